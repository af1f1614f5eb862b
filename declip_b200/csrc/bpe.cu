// declip_b200 — host-side byte-pair-encoding tokenizer (no device code): the text front end of the hot path
// (SURVEY.md §8f rank 1).  Re-implements prototype/model/utils/text_utils/simple_tokenizer.py:66-134 (OpenAI CLIP BPE
// with the extra <|mask|> token inserted before <|startoftext|> / <|endoftext|>, so the vocabulary is merges + 515) and
// the truncation / padding of TextTransformer.tokenize (text_encoder/text_transformer.py:144-170), multi-threaded over
// the captions of a batch.  The caller hands in text that is already cleaned and lower-cased
// (simple_tokenizer.py:53-63,126: html.unescape x2, strip, whitespace collapse, str.lower() — C-implemented one-liners
// in Python); everything after that — the pre-tokenisation regex, byte encoding, merges, id lookup — happens here.
#include <stdint.h>
#include <string.h>
#include <strings.h>

#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "bpe_unicode.h"
#include "internal.h"

namespace {

bool in_ranges(const uint32_t (*tab)[2], int n, uint32_t cp) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (cp < tab[mid][0]) hi = mid - 1;
    else if (cp > tab[mid][1]) lo = mid + 1;
    else return true;
  }
  return false;
}
inline bool is_letter(uint32_t cp) {
  if (cp < 128) return (cp >= 'a' && cp <= 'z') || (cp >= 'A' && cp <= 'Z');
  return in_ranges(DC_UNI_L, DC_UNI_L_COUNT, cp);
}
inline bool is_number(uint32_t cp) {
  if (cp < 128) return cp >= '0' && cp <= '9';
  return in_ranges(DC_UNI_N, DC_UNI_N_COUNT, cp);
}
// After whitespace_clean() the only whitespace left is U+0020; the remaining members of the regex module's \s (the
// Unicode White_Space property — NOT U+001C..U+001F, which Python's str.isspace() would add) are kept for callers
// that skip the cleaning step.
inline bool is_space(uint32_t cp) {
  return cp == ' ' || (cp >= 9 && cp <= 13) || cp == 0x85 || cp == 0xA0 || cp == 0x1680 ||
         (cp >= 0x2000 && cp <= 0x200A) || cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}

// Decodes one UTF-8 scalar at s[i] (malformed bytes decode as themselves, one at a time); returns its byte length.
inline int utf8_next(const unsigned char* s, size_t n, size_t i, uint32_t* cp) {
  const unsigned char c = s[i];
  if (c < 0x80) { *cp = c; return 1; }
  int len = (c >= 0xF0) ? 4 : (c >= 0xE0) ? 3 : (c >= 0xC0) ? 2 : 1;
  if (len == 1 || i + len > n) { *cp = c; return 1; }
  uint32_t v = c & (0xFF >> (len + 1));
  for (int k = 1; k < len; ++k) {
    if ((s[i + k] & 0xC0) != 0x80) { *cp = c; return 1; }
    v = (v << 6) | (s[i + k] & 0x3F);
  }
  *cp = v;
  return len;
}
inline void utf8_put(std::string& out, uint32_t cp) {
  if (cp < 0x80) out.push_back(static_cast<char>(cp));
  else if (cp < 0x800) { out.push_back(static_cast<char>(0xC0 | (cp >> 6))); out.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
  else { out.push_back(static_cast<char>(0xE0 | (cp >> 12))); out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
         out.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
}

struct Bpe {
  std::string byte_sym[256];                               // bytes_to_unicode(): byte -> printable stand-in (UTF-8)
  std::unordered_map<std::string, int> encoder;            // token string -> id
  std::unordered_map<std::string, int> ranks;              // "first\x01second" -> merge rank
  int sot = -1, eot = -1, mask = -1;
  std::shared_mutex mu;   // cache hits (the steady state) only take the shared side
  std::unordered_map<std::string, std::vector<int>> cache; // byte-encoded word -> ids

  // simple_tokenizer.py:16-37
  void build_bytes() {
    bool printable[256] = {false};
    for (int b = '!'; b <= '~'; ++b) printable[b] = true;
    for (int b = 0xA1; b <= 0xAC; ++b) printable[b] = true;
    for (int b = 0xAE; b <= 0xFF; ++b) printable[b] = true;
    int n = 0;
    for (int b = 0; b < 256; ++b) {
      std::string s;
      utf8_put(s, printable[b] ? static_cast<uint32_t>(b) : static_cast<uint32_t>(256 + n++));
      byte_sym[b] = s;
    }
  }
  // vocabulary order (simple_tokenizer.py:71-77): the 256 stand-ins in bytes_to_unicode() order, the same + "</w>", one
  // entry per merge, <|mask|>, <|startoftext|>, <|endoftext|>
  bool load(const char* text, size_t nbytes) {
    build_bytes();
    std::vector<std::string> order;
    for (int b = '!'; b <= '~'; ++b) order.push_back(byte_sym[b]);
    for (int b = 0xA1; b <= 0xAC; ++b) order.push_back(byte_sym[b]);
    for (int b = 0xAE; b <= 0xFF; ++b) order.push_back(byte_sym[b]);
    for (int b = 0; b < 256; ++b) {
      const bool pr = (b >= '!' && b <= '~') || (b >= 0xA1 && b <= 0xAC) || (b >= 0xAE);
      if (!pr) order.push_back(byte_sym[b]);
    }
    int id = 0;
    for (const auto& s : order) encoder.emplace(s, id++);
    for (const auto& s : order) encoder.emplace(s + "</w>", id++);
    // merges: lines 1 .. 49152-256-2 of the file (line 0 is the version header); split on whitespace like str.split()
    const size_t max_merges = 49152 - 256 - 2;
    size_t pos = 0, line = 0, taken = 0;
    while (pos <= nbytes && taken < max_merges) {
      size_t end = pos;
      while (end < nbytes && text[end] != '\n') ++end;
      if (line >= 1) {
        std::vector<std::string> parts;
        size_t i = pos;
        while (i < end) {
          while (i < end && (text[i] == ' ' || text[i] == '\t' || text[i] == '\r')) ++i;
          size_t j = i;
          while (j < end && text[j] != ' ' && text[j] != '\t' && text[j] != '\r') ++j;
          if (j > i) parts.emplace_back(text + i, j - i);
          i = j;
        }
        std::string joined;
        for (const auto& p : parts) joined += p;
        if (parts.size() == 2) ranks[parts[0] + '\x01' + parts[1]] = static_cast<int>(taken);   // dict(zip(...)): last wins
        // dict(zip(vocab, range(len(vocab)))): a later duplicate string overwrites the id of an earlier one
        encoder[joined] = id++;
        ++taken;
      }
      ++line;
      if (end >= nbytes) break;
      pos = end + 1;
    }
    encoder["<|mask|>"] = mask = id++;
    encoder["<|startoftext|>"] = sot = id++;
    encoder["<|endoftext|>"] = eot = id++;
    vocab = id;
    return true;
  }
  int vocab = 0;

  // simple_tokenizer.py:85-122 on the byte-encoded word `syms` (one stand-in per byte); appends the ids
  bool bpe_word(const std::string& key, std::vector<std::string>& w, std::vector<int>& out) {
    {
      std::shared_lock<std::shared_mutex> lk(mu);
      auto it = cache.find(key);
      if (it != cache.end()) { out.insert(out.end(), it->second.begin(), it->second.end()); return true; }
    }
    w.back() += "</w>";
    while (w.size() > 1) {
      int best = -1;
      size_t best_i = 0;
      for (size_t i = 0; i + 1 < w.size(); ++i) {
        auto it = ranks.find(w[i] + '\x01' + w[i + 1]);
        if (it != ranks.end() && (best < 0 || it->second < best)) { best = it->second; best_i = i; }
      }
      if (best < 0) break;
      const std::string first = w[best_i], second = w[best_i + 1];
      std::vector<std::string> nw;
      nw.reserve(w.size());
      for (size_t i = 0; i < w.size();) {
        if (i + 1 < w.size() && w[i] == first && w[i + 1] == second) { nw.push_back(first + second); i += 2; }
        else { nw.push_back(w[i]); ++i; }
      }
      w.swap(nw);
    }
    std::vector<int> ids;
    for (const auto& s : w) {
      auto it = encoder.find(s);
      if (it == encoder.end()) return false;
      ids.push_back(it->second);
    }
    out.insert(out.end(), ids.begin(), ids.end());
    std::unique_lock<std::shared_mutex> lk(mu);
    if (cache.size() < (1u << 20)) cache.emplace(key, std::move(ids));
    return true;
  }

  // the regex of simple_tokenizer.py:82 as a hand-written scanner, then bytes -> stand-ins -> merges -> ids
  bool encode(const char* text, std::vector<int>& out) {
    const unsigned char* s = reinterpret_cast<const unsigned char*>(text);
    const size_t n = strlen(text);
    static const char* kSpecial[2] = {"<|startoftext|>", "<|endoftext|>"};
    static const char* kContr[7] = {"'s", "'t", "'re", "'ve", "'m", "'ll", "'d"};
    size_t i = 0;
    std::vector<std::string> word;
    while (i < n) {
      size_t tok_end = 0;
      bool special = false;
      for (int k = 0; k < 2 && !tok_end; ++k) {
        const size_t L = strlen(kSpecial[k]);
        if (i + L <= n && strncasecmp(text + i, kSpecial[k], L) == 0) { tok_end = i + L; special = true; }
      }
      for (int k = 0; k < 7 && !tok_end; ++k) {
        const size_t L = strlen(kContr[k]);
        if (i + L <= n && strncasecmp(text + i, kContr[k], L) == 0) tok_end = i + L;
      }
      if (!tok_end) {
        uint32_t cp;
        const int len = utf8_next(s, n, i, &cp);
        if (is_letter(cp)) {
          size_t j = i + len;
          while (j < n) { uint32_t c2; const int l2 = utf8_next(s, n, j, &c2); if (!is_letter(c2)) break; j += l2; }
          tok_end = j;
        } else if (is_number(cp)) {
          tok_end = i + len;
        } else if (is_space(cp)) {
          i += len;
          continue;
        } else {
          size_t j = i + len;
          while (j < n) {
            uint32_t c2;
            const int l2 = utf8_next(s, n, j, &c2);
            if (is_space(c2) || is_letter(c2) || is_number(c2)) break;
            j += l2;
          }
          tok_end = j;
        }
      }
      if (special) {
        // findall returns the matched text; the bpe() cache maps only the exact lower-case literal to itself
        std::string lit(text + i, tok_end - i);
        auto it = encoder.find(lit);
        if (it != encoder.end() && (it->second == sot || it->second == eot)) { out.push_back(it->second); i = tok_end; continue; }
      }
      word.clear();
      std::string key;
      for (size_t b = i; b < tok_end; ++b) { word.push_back(byte_sym[s[b]]); key += byte_sym[s[b]]; }
      if (!bpe_word(key, word, out)) return false;
      i = tok_end;
    }
    return true;
  }
};

}  // namespace

struct dc_bpe { Bpe impl; };

extern "C" {

dc_bpe_t* dc_bpe_create(const char* merges_text, long long nbytes) {
  if (merges_text == nullptr || nbytes <= 0) { dc::set_error("bpe: empty merges text"); return nullptr; }
  dc_bpe* h = new dc_bpe();
  if (!h->impl.load(merges_text, static_cast<size_t>(nbytes))) { delete h; dc::set_error("bpe: cannot parse merges"); return nullptr; }
  return h;
}

void dc_bpe_destroy(dc_bpe_t* h) { delete h; }

int dc_bpe_vocab_size(const dc_bpe_t* h) { return h ? h->impl.vocab : 0; }

int dc_bpe_token_id(const dc_bpe_t* h, const char* token) {
  if (h == nullptr || token == nullptr) return -1;
  auto it = h->impl.encoder.find(token);
  return it == h->impl.encoder.end() ? -1 : it->second;
}

long long dc_bpe_encode(dc_bpe_t* h, const char* text, int* ids_out, long long capacity) {
  if (h == nullptr || text == nullptr) return dc::set_error("bpe: null argument");
  std::vector<int> ids;
  if (!h->impl.encode(text, ids)) return dc::set_error("bpe: symbol missing from the vocabulary");
  const long long n = static_cast<long long>(ids.size());
  for (long long i = 0; i < n && i < capacity; ++i) ids_out[i] = ids[i];
  return n;
}

// basic_clean + whitespace_clean + lower (simple_tokenizer.py:53-63,126) for a caption the caller has classified as
// printable ASCII without '&' (no entity to unescape, nothing for ftfy to repair): strip, collapse blanks, lower-case.
static void clean_printable_ascii(const char* t, std::string& out) {
  out.clear();
  bool pending_space = false;
  for (const char* p = t; *p; ++p) {
    const char c = *p;
    if (c == ' ') { pending_space = !out.empty(); continue; }
    if (pending_space) { out.push_back(' '); pending_space = false; }
    out.push_back((c >= 'A' && c <= 'Z') ? static_cast<char>(c + 32) : c);
  }
}

int dc_bpe_tokenize(dc_bpe_t* h, const char* const* texts, int n, int context_length, long long* ids, int* lengths,
                    int threads) {
  return dc_bpe_tokenize_ex(h, texts, nullptr, n, context_length, ids, lengths, threads);
}

int dc_bpe_tokenize_ex(dc_bpe_t* h, const char* const* texts, const unsigned char* raw_ascii, int n, int context_length,
                       long long* ids, int* lengths, int threads) {
  if (h == nullptr || texts == nullptr || ids == nullptr) return dc::set_error("bpe: null argument");
  if (context_length < 2) return dc::set_error("bpe: context_length must be at least 2");
  if (threads < 1) threads = 1;
  if (threads > n) threads = n > 0 ? n : 1;
  std::vector<int> status(threads, 0);
  auto work = [&](int t) {
    std::vector<int> tok;
    std::string cleaned;
    for (int i = t; i < n; i += threads) {
      tok.clear();
      tok.push_back(h->impl.sot);
      const char* text = texts[i];
      if (raw_ascii != nullptr && raw_ascii[i]) { clean_printable_ascii(text, cleaned); text = cleaned.c_str(); }
      if (!h->impl.encode(text, tok)) { status[t] = 1; return; }
      tok.push_back(h->impl.eot);
      long long* row = ids + static_cast<long long>(i) * context_length;
      int len = static_cast<int>(tok.size());
      if (len > context_length) {            // text_transformer.py:154-156: keep SOT, the first L-2 tokens, EOT
        tok[context_length - 1] = tok[len - 1];
        len = context_length;
      }
      for (int k = 0; k < len; ++k) row[k] = tok[k];
      for (int k = len; k < context_length; ++k) row[k] = 0;
      if (lengths != nullptr) lengths[i] = len;
    }
  };
  if (threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  for (int t = 0; t < threads; ++t)
    if (status[t]) return dc::set_error("bpe: symbol missing from the vocabulary");
  return 0;
}

}  // extern "C"
