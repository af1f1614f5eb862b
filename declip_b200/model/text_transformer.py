"""TextTransformer — mirror of prototype/model/text_encoder/text_transformer.py ('Transformer' branch).

forward accepts the reference's `List[str]` (host BPE tokenisation, text_transformer.py:144-180, when a
tokenizer is available) or — the fast path every benchmark and test uses — a pre-tokenised LongTensor
[B, 77] (SOT ... EOT, zero padded).  HF BERT/GPT2/Roberta branches are out of scope (hard-coded cluster
paths, text_transformer.py:51-102).
"""
import os

import torch
from torch import nn

from ..runtime import TowerRuntime, run_tower
from .base_transformer import LayerNorm, Transformer

VOCAB_SIZE = 49409   # simple_tokenizer.py:66-75: 49408 + <|mask|>


class TextTransformer(nn.Module):
    def __init__(self, embed_dim, context_length, transformer_width, transformer_heads, transformer_layers,
                 positional_embedding_flag, checkpoint, bpe_path=None, text_encode_type=None, text_model_utils=None):
        super().__init__()
        if text_encode_type != 'Transformer':
            raise NotImplementedError("declip_b200: only text_encode_type='Transformer' is on the hot path")
        if not positional_embedding_flag:
            raise NotImplementedError("declip_b200: positional_embedding_flag=False is not used by any config")
        self.context_length = context_length
        self.positional_embedding_flag = positional_embedding_flag
        self.text_encode_type = text_encode_type
        self.text_model_utils = text_model_utils or {}
        self.tokenizer = None
        self.bpe_path = bpe_path
        if bpe_path is not None and os.path.exists(str(bpe_path)):
            from ..tokenizer import SimpleTokenizer            # C++ BPE (csrc/bpe.cu), text_transformer.py:35-36
            self.tokenizer = SimpleTokenizer(bpe_path)
        self.transformer = Transformer(width=transformer_width, layers=transformer_layers, heads=transformer_heads,
                                       attn_mask=self.build_attention_mask(), checkpoint=checkpoint)
        self.vocab_size = VOCAB_SIZE
        self.token_embedding = nn.Embedding(self.vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.normal(mean=0, std=0.02, size=(context_length, transformer_width)))
        self.ln_final = LayerNorm(transformer_width)
        self.text_projection = nn.Linear(transformer_width, embed_dim)
        self.initialize_parameters()
        self._rt = TowerRuntime("text", self, transformer_layers, transformer_width, transformer_heads, context_length,
                                embed_dim, vocab=self.vocab_size)

    def initialize_parameters(self):
        # text_transformer.py:117-130
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection.weight, std=self.transformer.width ** -0.5)

    @property
    def dtype(self):
        return self.positional_embedding.dtype

    def build_attention_mask(self):
        # text_transformer.py:136-142 — kept for API parity; the kernel applies the causal mask analytically
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    def tokenize(self, texts, context_length=77, return_length=False, mask_type=None):
        """text_transformer.py:144-180.  Needs the BPE vocabulary file (not shipped with the reference)."""
        if self.tokenizer is None:
            raise RuntimeError("declip_b200: no BPE tokenizer loaded (bpe_path=%r). Pass a pre-tokenised LongTensor "
                               "[B,%d] instead of strings." % (self.bpe_path, self.context_length))
        if mask_type is not None:
            ids = self.tokenizer.tokenize(texts, context_length)
            from .text_utils import mask_tokens_batch
            return mask_tokens_batch(ids)
        return self.tokenizer.tokenize(texts, context_length, return_length=return_length)

    def forward(self, text, mask_type=None, return_dense=False):
        """text: List[str] (reference API) | LongTensor ids [B,77] | (masked_ids, labels) when already MLM-masked.
        Returns x [B,E]; with mask_type: (x, words_feat, labels); with return_dense: (x, words_feat)
        (text_transformer.py:183-274).  words_feat is bf16 [B,77,D] (ln_final of every token)."""
        labels = None
        if isinstance(text, (tuple, list)) and len(text) == 2 and torch.is_tensor(text[0]):
            ids, labels = text
        elif torch.is_tensor(text):
            ids = text
        else:
            ids = self.tokenize(text, context_length=self.context_length)
        if mask_type is not None and labels is None:
            if mask_type != 'MLM':
                raise NotImplementedError(mask_type)
            from .text_utils import mask_tokens_batch
            ids, labels = mask_tokens_batch(ids)                                   # text_transformer.py:154-162
        if ids.dim() != 2 or ids.shape[1] != self.context_length:
            raise ValueError("expected token ids [B,%d], got %s" % (self.context_length, tuple(ids.shape)))
        dev = self.positional_embedding.device
        if ids.device != dev:
            ids = ids.to(dev, non_blocking=True)                                   # text_transformer.py:188
        dense = mask_type is not None or return_dense
        if not dense:
            return run_tower(self._rt, ids)
        x, words = run_tower(self._rt, ids, dense=True)
        words = words.view(ids.shape[0], self.context_length, -1)
        if mask_type is not None:
            return x, words, labels
        return x, words


def text_transformers(**kwargs):
    default_kwargs = {'context_length': 77, 'transformer_width': 512, 'transformer_heads': 8, 'transformer_layers': 12,
                      'positional_embedding_flag': True, 'checkpoint': False}
    default_kwargs.update(**kwargs)
    return TextTransformer(**default_kwargs)
