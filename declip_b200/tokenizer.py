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

    def tokenize(self, texts, context_length=77, return_length=False, pinned=None):
        """List[str] -> LongTensor [n, context_length] (SOT + ids + EOT, truncated keeping EOT, zero padded).

        Printable-ASCII captions without '&' (the common case) are cleaned inside the C++ library
        (dc_bpe_tokenize_ex); anything else goes through `clean()` here.  With `pinned` (default: whenever CUDA is
        available) the ids land in page-locked memory so the caller's `.to(device, non_blocking=True)` is a true
        asynchronous copy — a pageable source makes that copy synchronous and stalls the launch thread."""
        if isinstance(texts, str):
            texts = [texts]
        n = len(texts)
        if pinned is None:
            pinned = torch.cuda.is_available()
        out = torch.zeros((n, context_length), dtype=torch.int64, pin_memory=bool(pinned and n))
        lengths = np.zeros(n, dtype=np.int32)
        if n:
            raw = bytearray(n)
            enc = [None] * n
            for i, t in enumerate(texts):
                if t.isascii() and t.isprintable() and "&" not in t:
                    raw[i] = 1
                    enc[i] = t.encode("ascii")
                else:
                    enc[i] = clean(t).encode("utf-8")
            arr = (ctypes.c_char_p * n)(*enc)
            flags = (ctypes.c_ubyte * n).from_buffer(raw)
            rc = self._lib.dc_bpe_tokenize_ex(self._h, arr, flags, n, context_length, out.data_ptr(), lengths.ctypes.data,
                                              self.threads)
            _lib.check(rc, "dc_bpe_tokenize_ex")
        if return_length:
            return out, torch.from_numpy(lengths.astype(np.int64))
        return out


class CaptionPipeline:
    """Keeps host text work off the training thread: a worker thread tokenises batch i+1
    (the C++ BPE releases the GIL for the whole batch) — optionally after EDA augmentation for DeCLIP's second caption
    view — into pinned memory and uploads it on a copy stream while the device runs step i.  Iterating yields the
    input dict with `token_ids` (and `token_ids_aug`) already on the device; `captions` is kept for callers that log it.

        for batch in CaptionPipeline(loader, model.encode_text.tokenizer, device, eda=model.emd): model(batch, ...)
    """

    def __init__(self, batches, tokenizer, device, context_length=77, eda=None, depth=2):
        import queue
        import threading
        self._src, self._tok, self._dev, self._ctx, self._eda = iter(batches), tokenizer, torch.device(device), context_length, eda
        self._q = queue.Queue(maxsize=depth)
        self._stream = torch.cuda.Stream(device=self._dev)
        self._thread = threading.Thread(target=self._work, daemon=True)
        self._thread.start()

    @staticmethod
    def sample_captions(texts):
        return [t[0] if isinstance(t, (list, tuple)) else t for t in texts]          # clip.py:110-111

    def _work(self):
        try:
            torch.cuda.set_device(self._dev)
            for batch in self._src:
                texts = self.sample_captions(batch['captions'])
                out = dict(batch)
                with torch.cuda.stream(self._stream):
                    out['token_ids'] = self._tok.tokenize(texts, self._ctx, pinned=True).to(self._dev, non_blocking=True)
                    if self._eda is not None:
                        aug = self._eda.augment_batch(texts)
                        out['token_ids_aug'] = self._tok.tokenize(aug, self._ctx, pinned=True).to(self._dev, non_blocking=True)
                    if torch.is_tensor(out.get('images')) and not out['images'].is_cuda:
                        img = out['images'] if out['images'].is_pinned() else out['images'].pin_memory()
                        out['images'] = img.to(self._dev, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._stream)
                self._q.put((out, ev))
            self._q.put(None)
        except BaseException as e:   # noqa: BLE001 - re-raised on the consumer side
            self._q.put(e)

    def __iter__(self):
        return self

    def __next__(self):
        item = self._q.get()
        if item is None:
            raise StopIteration
        if isinstance(item, BaseException):
            raise item
        out, ev = item
        torch.cuda.current_stream(self._dev).wait_event(ev)
        for v in out.values():           # tensors produced on the copy stream, consumed on the compute stream
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(torch.cuda.current_stream(self._dev))
        return out
