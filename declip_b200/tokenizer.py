"""Host-side text pipeline of the text tower: `SimpleTokenizer` with the interface of
prototype/model/utils/text_utils/simple_tokenizer.py (`encoder[...]` for the special tokens, `encode`, `decode`) backed by
the multi-threaded C++ BPE of the C-ABI library (csrc/bpe.cu), plus `tokenize()` = TextTransformer.tokenize
(text_encoder/text_transformer.py:144-170) for a whole batch in one call.

Cleaning (`basic_clean` + `whitespace_clean` + `lower`, simple_tokenizer.py:53-63,126) stays in Python: three C-implemented
string calls per caption.  `ftfy.fix_text` is applied when ftfy is importable, as in the reference; without it the text is
taken as is.  The BPE vocabulary (bpe_simple_vocab_16e6.txt.gz) is not shipped by the reference (docs/dataset_prepare.md:33-37):
pass its path."""
import ctypes
import gzip
import html
import os

import numpy as np
import torch

from . import _lib

try:
    import regex as _re
except ImportError:   # pragma: no cover - `regex` is what the reference uses
    import re as _re

try:
    import ftfy as _ftfy
    _fix_text = _ftfy.fix_text
except Exception:      # noqa: BLE001 - optional dependency, identity when absent or stubbed
    def _fix_text(s):
        return s

_WS = _re.compile(r"\s+")


def clean(text):
    """basic_clean + whitespace_clean + lower (simple_tokenizer.py:53-63,126)."""
    text = html.unescape(html.unescape(_fix_text(text))).strip()
    return _WS.sub(" ", text).strip().lower()


class _Encoder:
    """`tokenizer.encoder[token]` / `len(tokenizer.encoder)` as the reference's callers use them."""

    def __init__(self, lib, handle):
        self._lib, self._h = lib, handle

    def __getitem__(self, token):
        i = self._lib.dc_bpe_token_id(self._h, token.encode("utf-8"))
        if i < 0:
            raise KeyError(token)
        return i

    def __contains__(self, token):
        return self._lib.dc_bpe_token_id(self._h, token.encode("utf-8")) >= 0

    def __len__(self):
        return self._lib.dc_bpe_vocab_size(self._h)


class SimpleTokenizer:
    def __init__(self, bpe_path, threads=None):
        if bpe_path is None or not os.path.exists(bpe_path):
            raise FileNotFoundError("BPE vocabulary %r (bpe_simple_vocab_16e6.txt.gz is not shipped with the reference)" % (bpe_path,))
        opener = gzip.open if str(bpe_path).endswith(".gz") else open
        with opener(bpe_path, "rb") as f:
            data = f.read()
        self._lib = _lib.load()
        self._h = self._lib.dc_bpe_create(data, len(data))
        if not self._h:
            raise RuntimeError("dc_bpe_create failed: %s" % _lib.last_error())
        self.encoder = _Encoder(self._lib, self._h)
        self.threads = threads or min(16, os.cpu_count() or 1)
        self._decoder = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.dc_bpe_destroy(h)

    def encode(self, text):
        t = clean(text).encode("utf-8")
        cap = 4 * len(t) + 8
        buf = (ctypes.c_int * cap)()
        n = self._lib.dc_bpe_encode(self._h, t, buf, cap)
        if n < 0:
            raise RuntimeError("dc_bpe_encode failed: %s" % _lib.last_error())
        return list(buf[:n])

    def tokenize(self, texts, context_length=77, return_length=False):
        """List[str] -> LongTensor [n, context_length] (SOT + ids + EOT, truncated keeping EOT, zero padded)."""
        if isinstance(texts, str):
            texts = [texts]
        n = len(texts)
        ids = np.zeros((n, context_length), dtype=np.int64)
        lengths = np.zeros(n, dtype=np.int32)
        if n:
            enc = [clean(t).encode("utf-8") for t in texts]
            arr = (ctypes.c_char_p * n)(*enc)
            rc = self._lib.dc_bpe_tokenize(self._h, arr, n, context_length, ids.ctypes.data, lengths.ctypes.data, self.threads)
            _lib.check(rc, "dc_bpe_tokenize")
        out = torch.from_numpy(ids)
        if return_length:
            return out, torch.from_numpy(lengths.astype(np.int64))
        return out
