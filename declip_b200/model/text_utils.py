"""Host-side text helpers mirroring prototype/model/utils/text_utils/mask_tokens.py, vectorised over the batch.

The reference masks one caption at a time in a Python loop inside TextTransformer.tokenize
(text_transformer.py:154-162, mask_tokens.py:5-29); the distribution here is identical (Bernoulli(0.15) over
non-special tokens; 80 % -> <|mask|>, 10 % -> random id, 10 % unchanged; labels -100 elsewhere) but the RNG stream
is torch's batched one, so parity tests inject pre-masked (ids, labels) instead of comparing random draws."""
import torch

SOT, EOT, MASK = 49407, 49408, 49406
VOCAB = 49409


def mask_tokens_batch(ids, mlm_probability=0.15, generator=None, vocab=VOCAB, mask_token=MASK, special=(SOT, EOT, MASK)):
    """ids int64 [B, L] (SOT ... EOT, zero padded) -> (masked_ids, labels) on the same device."""
    ids = ids.clone()
    dev = ids.device
    B, L = ids.shape
    lengths = ids.argmax(dim=1) + 1                                   # EOT is the highest id (text_transformer.py:203)
    pos = torch.arange(L, device=dev).unsqueeze(0)
    valid = pos < lengths.unsqueeze(1)
    for s in special:
        valid &= ids != s
    prob = torch.full((B, L), mlm_probability, device=dev) * valid
    masked = torch.bernoulli(prob, generator=generator).bool()
    labels = torch.where(masked, ids, torch.full_like(ids, -100))
    replaced = torch.bernoulli(torch.full((B, L), 0.8, device=dev), generator=generator).bool() & masked
    ids[replaced] = mask_token
    rnd = torch.bernoulli(torch.full((B, L), 0.5, device=dev), generator=generator).bool() & masked & ~replaced
    words = torch.randint(vocab, (B, L), device=dev, generator=generator)
    ids[rnd] = words[rnd]
    return ids, labels
