"""TEST INFRASTRUCTURE — fp32 restatements of the NT-Xent family (prototype/loss_functions/nt_xent.py:6-97),
pinned by tests/golden/nt_xent.pt (generated from the reference's own classes by tools/make_golden.py nt_xent)."""
import torch
import torch.nn.functional as F


def nt_xent(z_i, z_j, temperature=0.5):
    # nt_xent.py:28-44
    b = z_i.shape[0]
    p = torch.cat((z_i, z_j), dim=0)
    sim = F.cosine_similarity(p.unsqueeze(1), p.unsqueeze(0), dim=2) / temperature
    pos = torch.cat((torch.diag(sim, b), torch.diag(sim, -b))).reshape(2 * b, 1)
    mask = torch.ones(2 * b, 2 * b, dtype=torch.bool).fill_diagonal_(False)
    for i in range(b):
        mask[i, b + i] = False
        mask[b + i, i] = False
    logits = torch.cat((pos, sim[mask].reshape(2 * b, -1)), dim=1)
    return F.cross_entropy(logits, torch.zeros(2 * b, dtype=torch.long), reduction="sum") / (2 * b)


def nt_xent_gather(z_i, z_ib, z_j, z_jb, rank=0, temperature=0.1):
    # nt_xent.py:63-97
    bs, l_bs = z_i.shape[0], z_ib.shape[0]
    p0, p1 = torch.cat((z_i, z_j), dim=0), torch.cat((z_ib, z_jb), dim=0)
    sim = F.cosine_similarity(p0.unsqueeze(1), p1.unsqueeze(0), dim=2) / temperature
    ids = torch.arange(bs)
    labels = rank * bs + ids
    mp = torch.zeros(bs * 2, l_bs * 2, dtype=torch.bool)
    mp[ids + bs, labels] = True
    mp[ids, labels + l_bs] = True
    mn = torch.ones(bs * 2, l_bs * 2, dtype=torch.bool)
    mn[ids, labels] = False
    mn[ids + bs, labels] = False
    mn[ids, labels + l_bs] = False
    mn[ids + bs, labels + l_bs] = False
    logits = torch.cat((sim[mp].reshape(2 * bs, -1), sim[mn].reshape(2 * bs, -1)), dim=1)
    return F.cross_entropy(logits, torch.zeros(2 * bs, dtype=torch.long), reduction="sum") / (2 * bs)


def inputs(seed=0, b=8, n=24, d=64):
    g = torch.Generator().manual_seed(seed)
    z_i, z_j = torch.randn(b, d, generator=g), torch.randn(b, d, generator=g)
    z_ib, z_jb = torch.randn(n, d, generator=g), torch.randn(n, d, generator=g)
    rank = 1
    z_ib[rank * b:(rank + 1) * b] = z_i          # the gathered tensors contain this rank's rows
    z_jb[rank * b:(rank + 1) * b] = z_j
    return z_i, z_j, z_ib, z_jb, rank
