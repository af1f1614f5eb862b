"""Generates tests/golden/bpe_small_merges.txt (a small but REAL merges table, learned here by a plain BPE trainer over an
embedded toy corpus — the official bpe_simple_vocab_16e6.txt.gz is a download the reference does not ship) and
tests/golden/bpe_tokenizer.json: the ids the UNMODIFIED reference SimpleTokenizer / TextTransformer.tokenize produce with
that table for a list of awkward captions (build container only).
    python tools/make_golden_bpe.py"""
import collections
import gzip
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import golden, ref_harness  # noqa: E402

CORPUS = """a photo of a dog . a photo of a cat . a photograph of the big red dog running in the park .
two dogs are playing with a ball on the grass . the quick brown fox jumps over the lazy dog .
a man is riding a bicycle down the street . a woman is holding an umbrella in the rain .
it's a beautiful day , isn't it ? we'll see what they've done . i'm sure she'd like it , they're here .
the cafe serves coffee and tea . a naive question about resume and fiance . 2023 was a good year , 42 is the answer .
children are playing in the playground . a plate of food with vegetables and rice . a close up of a flower .
the tokenizer splits words into subword units . photography of mountains , lakes and rivers at sunset .
an airplane is flying in the blue sky . people walking on the beach near the ocean waves .
""" * 3

TEXTS = [
    "a photo of a dog", "A Photo Of The BIG red dog!!!", "it's isn't we'll they've i'm she'd they're", "'sam 's rock'n'roll 'T",
    "2023 was 42 3.14 1,000 ٣٤ ½ x²", "  multiple   spaces\tand\nnewlines  ", "&amp; &lt;b&gt; &amp;amp; &quot;quoted&quot; &#39;x&#39;",
    "café naïve résumé fiancé", "日本語のテキスト と english mixed", "emoji 🙂🙂 test 👍🏽", "İstanbul ǅ ß STRASSE", "", "   ",
    "<|startoftext|> literal <|endoftext|> and <|mask|> here", "under_score-hyphen/slash\\back @user #tag $5 50% a+b=c",
    "don't can't won't o'clock y'all'd've", "...!!!???---", "a" * 40, "x", "the the the the the",
    "word " * 100, "photograph photography photographer photos", " nbsp emspace　ideographic",
    "tab\tseparated\x1fcontrol\x85nel", "Ünïcödé ÇÀPS ñandú ǆ", "1st 2nd 3rd 4th №5 Ⅷ", "a.b,c;d:e!f?g'h\"i", "ＦＵＬＬwidth １２３",
]


def train_bpe(corpus, n_merges):
    """Plain BPE over the byte-encoded corpus words (Sennrich et al.) — only used to obtain a realistic merges table."""
    from prototype.model.utils.text_utils.simple_tokenizer import bytes_to_unicode
    be = bytes_to_unicode()
    words = collections.Counter()
    for w in corpus.split():
        sym = tuple(be[b] for b in w.encode("utf-8"))
        words[sym[:-1] + (sym[-1] + "</w>",)] += 1
    merges = []
    for _ in range(n_merges):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        (a, b), cnt = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        if cnt < 2:
            break
        merges.append((a, b))
        nw = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and w[i] == a and w[i + 1] == b:
                    out.append(a + b); i += 2
                else:
                    out.append(w[i]); i += 1
            nw[tuple(out)] += c
        words = nw
    return merges


def main():
    ref_harness.setup()
    from prototype.model.utils.text_utils.simple_tokenizer import SimpleTokenizer
    merges = train_bpe(CORPUS, 400)
    txt = "#version: declip_b200 test merges (tools/make_golden_bpe.py)\n" + "\n".join("%s %s" % m for m in merges) + "\n"
    mpath = os.path.join(golden.GOLDEN_DIR, "bpe_small_merges.txt")
    open(mpath, "w", encoding="utf-8").write(txt)
    gz = os.path.join(tempfile.gettempdir(), "declip_b200_small_bpe.txt.gz")
    with gzip.open(gz, "wt", encoding="utf-8") as f:
        f.write(txt)
    tok = SimpleTokenizer(gz)
    out = {"generator": "tools/make_golden_bpe.py (reference SimpleTokenizer + TextTransformer.tokenize)",
           "merges": len(merges), "vocab": len(tok.encoder),
           "special": {k: tok.encoder[k] for k in ("<|mask|>", "<|startoftext|>", "<|endoftext|>")},
           "texts": TEXTS, "encode": [tok.encode(t) for t in TEXTS]}
    model = ref_harness.build_clip_vitb32(512, {"layers": 1}, {"transformer_layers": 1, "bpe_path": gz})
    for L in (77, 12):
        ids, lengths = model.encode_text.tokenize(TEXTS, context_length=L, return_length=True)
        out["tokenize_%d" % L] = {"ids": ids.tolist(), "lengths": lengths.tolist()}
    # the fake table of the model goldens (dummy merges that never apply): vocabulary size / special ids only
    fake = SimpleTokenizer(ref_harness._fake_bpe())
    out["fake_vocab"] = len(fake.encoder)
    out["fake_special"] = {k: fake.encoder[k] for k in ("<|mask|>", "<|startoftext|>", "<|endoftext|>")}
    out["fake_encode"] = [fake.encode(t) for t in TEXTS[:6]]
    path = os.path.join(golden.GOLDEN_DIR, "bpe_tokenizer.json")
    json.dump(out, open(path, "w"), ensure_ascii=True)
    print(mpath, len(merges), "merges; vocab", out["vocab"], out["special"], "->", path, os.path.getsize(path), "bytes")
    print("sample:", TEXTS[1], out["encode"][1])


if __name__ == "__main__":
    main()
