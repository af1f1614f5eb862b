"""Drop-in for prototype.loss_functions (loss.py): ClipInfoCELoss with the reference's call signature
`criterion(logits_per_image, logits_per_text) -> (loss, labels)`, computed by one fused CUDA kernel per
strip (log-softmax + NLL + top-1/top-5 in the forward, softmax - onehot in the backward)."""
import torch
from torch.nn.modules.loss import _Loss

from . import functions as F_


class ClipInfoCELoss(_Loss):
    def __init__(self):
        super().__init__()
        self.stats = {}

    def forward(self, logits_per_image, logits_per_text):
        bs, l_bs = logits_per_image.shape
        rank, _ = F_.dist_info()
        label0 = 0 if l_bs == bs else rank * bs                                    # loss.py:42-45
        labels = label0 + torch.arange(0, bs, dtype=torch.long, device=logits_per_image.device)
        loss = F_.ClipInfoCE.apply(logits_per_image, logits_per_text, label0, self.stats)
        return loss, labels

    def accuracy(self):
        """prec@1 / prec@5 (percent) of logits_per_image from the last forward — misc.py:415-428 — as device
        tensors (no host sync)."""
        rows = float(self.stats["rows"])
        return self.stats["top1_count"].float() * (100.0 / rows), self.stats["top5_count"].float() * (100.0 / rows)


class SimsiamLoss(torch.nn.Module):
    """loss_functions/loss.py:60-84: -0.5 * (D(p1, z2) + D(p2, z1)), D = mean cosine with stop-gradient on z."""

    def __init__(self, symmetry=True):
        super().__init__()
        self.symmetry = symmetry

    def forward(self, p1, z1, p2, z2, minimize_loss=False):
        if not self.symmetry or minimize_loss:
            raise NotImplementedError("declip_b200: only the symmetric SimSiam loss of the reference configs is built")
        return -0.5 * (F_.CosineMean.apply(p1, z2) + F_.CosineMean.apply(p2, z1))


class StripCE(torch.autograd.Function):
    """mean_r CE(logits[r], label0 + r) on one strip."""

    @staticmethod
    def forward(ctx, logits, label0):
        import ctypes
        from . import _lib, ops
        lib = ops.lib_for(logits)
        if logits.stride(1) != 1:
            logits = logits.contiguous()
        b, n = logits.shape
        acc = torch.zeros(1, device=logits.device, dtype=torch.float32)
        lse = torch.empty(b, device=logits.device, dtype=torch.float32)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = ctypes.c_void_p
        _lib.check(lib.dc_ce_strip_fwd(P(logits.data_ptr()), logits.stride(0), b, n, label0, None, None, P(acc.data_ptr()), None,
                                       None, P(lse.data_ptr()), st), "dc_ce_strip_fwd")
        ctx.save_for_backward(logits, lse)
        ctx.label0 = label0
        return acc[0] / b

    @staticmethod
    def backward(ctx, g):
        import ctypes
        from . import _lib, ops
        logits, lse = ctx.saved_tensors
        lib = ops.lib_for(logits)
        b, n = logits.shape
        g = g.contiguous().float().reshape(1)
        d = torch.empty(b, n, device=logits.device, dtype=torch.float32)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = ctypes.c_void_p
        _lib.check(lib.dc_ce_strip_bwd(P(logits.data_ptr()), logits.stride(0), b, n, ctx.label0, None, None, P(lse.data_ptr()),
                                       P(g.data_ptr()), 1.0 / b, P(d.data_ptr()), d.stride(0), 1, st), "dc_ce_strip_bwd")
        return d, None


class NTXentLoss(torch.nn.Module):
    """loss_functions/nt_xent_ConVIRT.py:4-86: local b x b image-text NT-Xent with soft targets = identity:
    alpha * CE(zi zj^T / T) + (1 - alpha) * CE(zj zi^T / T).  Evaluated every DeCLIP step by the reference solver
    (declip_solver.py:486-488) but only enters the loss for clip_simsiam_loss_weight.type == 'convirt'."""

    def __init__(self, batch_size, temperature=0.1, use_cosine_similarity=True, alpha_weight=0.75):
        super().__init__()
        self.batch_size = batch_size
        self.temperature = temperature
        self.alpha_weight = alpha_weight

    def forward(self, zis, zjs, norm=True, weights=1.0):
        if norm:
            zis = F_.L2Normalize.apply(zis, 1e-12)
            zjs = F_.L2Normalize.apply(zjs, 1e-12)
        ab, ba = F_.StripLogits.apply(None, 1.0 / self.temperature, False, False, ((0, 1), (1, 0)), zis, zjs)
        return self.alpha_weight * StripCE.apply(ab, 0) + (1 - self.alpha_weight) * StripCE.apply(ba, 0)


class NT_Xent(_Loss):
    """loss_functions/nt_xent.py:6-44 (SimCLR): 2b x 2b cosine / T; for every row the positive is its other view and
    the self-similarity is excluded — a row cross-entropy with one masked column, summed and divided by 2b."""

    def __init__(self, batch_size, temperature=0.5):
        super().__init__()
        self.batch_size = batch_size
        self.temperature = temperature

    def forward(self, z_i, z_j):
        b = self.batch_size
        p = torch.cat((z_i, z_j), dim=0)
        pn = F_.L2Normalize.apply(p, 1e-8)                                # nn.CosineSimilarity(dim=2), eps 1e-8
        sim = F_.MatmulNT.apply(pn, pn, 1.0 / self.temperature)
        r = torch.arange(2 * b, device=p.device)
        labels = (r + b) % (2 * b)                                        # nt_xent.py:31-35: the other view
        return F_.MaskedRowCE.apply(sim, labels.long(), r.int(), 2 * b)


class NT_Xent_gather(_Loss):
    """loss_functions/nt_xent.py:47-97 (SLIP): the 2b local rows against the 2N gathered columns; positives and the
    excluded self column follow the rank-offset labels of nt_xent.py:74-86."""

    def __init__(self, batch_size, temperature=0.1):
        super().__init__()
        self.batch_size = batch_size
        self.temperature = temperature

    def forward(self, z_i, z_ib, z_j, z_jb, temperature=None):
        bs, l_bs = z_i.shape[0], z_ib.shape[0]
        assert bs == self.batch_size
        rank, _ = F_.dist_info()
        p0 = F_.L2Normalize.apply(torch.cat((z_i, z_j), dim=0), 1e-8)
        p1 = F_.L2Normalize.apply(torch.cat((z_ib, z_jb), dim=0), 1e-8)
        sim = F_.MatmulNT.apply(p0, p1, 1.0 / self.temperature)          # the reference divides by self.temperature too
        ids = torch.arange(bs, device=z_i.device)
        lab = rank * bs + ids
        labels = torch.cat((lab + l_bs, lab))                             # positives: nt_xent.py:77-78
        skip = torch.cat((lab, lab + l_bs))                               # self columns: nt_xent.py:80,83
        return F_.MaskedRowCE.apply(sim, labels.long(), skip.int(), 2 * bs)
