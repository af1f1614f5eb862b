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
        head = getattr(logits_per_image, "_dc_head", None)
        if head is not None and getattr(logits_per_text, "_dc_head", None) is head:
            # fused head (csrc/head.cu): the row cross-entropies were reduced inside the kernel that formed the logits
            loss = head.parts.sum() / (2.0 * bs)
            self.stats["top1_count"], self.stats["top5_count"], self.stats["rows"] = head.top1_count, head.top5_count, bs
            return loss, labels
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


class DeclipCriterion(torch.nn.Module):
    """The DeCLIP solver's loss composition, prototype/solver/declip_solver.py:435-533, as one callable on the
    `return_dict=True` output of DECLIP.forward: four ClipInfoCELoss terms averaged (image_text_two_view) or two
    (only_image_two_view), MLM, nearest-neighbour text supervision, SimSiam, each divided by world size and combined
    with `clip_simsiam_loss_weight` (yfcc15m_vit_declip/config.yaml:28-32).  NTXentLoss is evaluated lazily: the
    reference computes it every step but it only enters the loss when the weight type is 'convirt'."""

    DEFAULT_WEIGHTS = dict(clip_loss=0.4, simsiam_loss=0.2, masking_language=0.2, nn_text=0.2)

    def __init__(self, weights=None, image_text_two_view=True, world_size=None):
        super().__init__()
        self.weights = dict(weights or self.DEFAULT_WEIGHTS)
        self.image_text_two_view = image_text_two_view
        self.world_size = world_size
        self.criterion = ClipInfoCELoss()
        self.simsiam_criterion = SimsiamLoss()

    def forward(self, out, curr_step=0, total_step=1):
        world = self.world_size if self.world_size is not None else F_.dist_info()[1]
        crit, w = self.criterion, self.weights
        li1, li2, lt1, lt2 = out['logits']
        clip_1, target = crit(li1, lt1)
        stats = dict(crit.stats)                        # prec@1/5 are logged on the first pair (declip_solver.py:535-536)
        clip_2, _ = crit(li2, lt2)
        if self.image_text_two_view:
            li1a, li2a, lt1a, lt2a = out['logits_aug']
            clip = (clip_1 + clip_2 + crit(li1a, lt1a)[0] + crit(li2a, lt2a)[0]) / 4
        else:
            clip = (clip_1 + clip_2) / 2
        clip = clip / world
        zero = torch.zeros_like(clip)
        mlm = out['text_self_supervised'] / world if 'text_self_supervised' in out else zero
        if 'nn_text_logits' in out:
            n1, n2, n1a, n2a = out['nn_text_logits']
            nn_text = (crit(n1, n1a)[0] + crit(n2, n2a)[0]) / 2 / world
        else:
            nn_text = zero
        p1, p2, z1, z2 = out['simsiam_features']
        simsiam = self.simsiam_criterion(p1, z1, p2, z2) / world
        parts = dict(clip=clip, mlm=mlm, nn=nn_text, simsiam=simsiam)
        kind = w.get('type', None)
        if not kind:
            loss = clip * w['clip_loss']
            if w.get('simsiam_loss', 0):
                loss = loss + simsiam * w['simsiam_loss']
            if w.get('masking_language', 0):
                loss = loss + mlm * w['masking_language']
            if w.get('nn_text', 0):
                loss = loss + nn_text * w['nn_text']
        elif kind == 'convirt':
            tf, f1, f2 = out['features']
            ntx = NTXentLoss(f1.shape[0])
            parts['nt_xent'] = (ntx(f1, tf) + ntx(f2, tf)) / world
            loss = (clip + parts['nt_xent']) / 2 * w['clip_loss'] + simsiam * w['simsiam_loss']
        elif kind == 'linear':
            cw = 0.2 + 0.8 * curr_step / total_step
            loss = clip * cw + simsiam * (1.0 - cw)
        elif kind == 'shift':
            loss = clip if curr_step % 2 == 0 else simsiam
        else:
            raise NotImplementedError("clip_simsiam_loss_weight.type %r" % kind)
        crit.stats = stats
        return loss, parts, target

    def accuracy(self):
        return self.criterion.accuracy()


class FilipCriterion(torch.nn.Module):
    """prototype/solver/filip_solver.py:436-520: ClipInfoCELoss on the global logits and on the token-wise
    `dense_logits`, weights clip_loss / clip_dense_loss (yfcc15m_vit_filip/config.yaml:33-35: 0.0 / 1.0)."""

    def __init__(self, weights=None, world_size=None):
        super().__init__()
        self.weights = dict(weights or dict(clip_loss=0.0, clip_dense_loss=1.0))
        self.world_size = world_size
        self.criterion = ClipInfoCELoss()

    def forward(self, out):
        world = self.world_size if self.world_size is not None else F_.dist_info()[1]
        dense, target = self.criterion(*out['dense_logits'])
        dense = dense / world
        loss = dense * self.weights['clip_dense_loss']
        parts = dict(dense=dense)
        if self.weights.get('clip_loss', 0):
            parts['clip'] = self.criterion(*out['logits'])[0] / world
            loss = loss + parts['clip'] * self.weights['clip_loss']
        return loss, parts, target
