"""Fused distributed contrastive head (csrc/head.cu, through the C ABI) against a plain PyTorch fp32 statement of
clip.py:129-146 + loss.py:40-50 + misc.py:415-428: loss, accuracy counts, d image_features, d text_features,
d logit_scale (clamp-on-.data semantics) — single rank, and W emulated ranks on one GPU (each rank's kernels are run in
turn on its slice of a gathered buffer, the 2b+2-float exchange vectors are assembled by hand) against autograd of the
GLOBAL symmetric InfoNCE, which is what the reference's AllGather + all-reduce computes."""
import ctypes
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
P = ctypes.c_void_p


def _cos(a, b):
    a, b = a.float().reshape(-1), b.float().reshape(-1)
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-20)).item()


def _reference(img, txt, ls, clamp=True, g=(1.0, 1.0)):
    """fp32 autograd statement on bf16-rounded normalised features (the kernel's operand precision)."""
    img = img.clone().requires_grad_(True)
    txt = txt.clone().requires_grad_(True)
    ls = ls.clone().requires_grad_(True)
    i_n = img / img.norm(dim=-1, keepdim=True)
    t_n = txt / (txt.norm(dim=-1, keepdim=True) + 1e-10)
    s = ls.exp()
    if clamp:
        s.data = torch.clamp(s.data, max=100)                    # clip.py:133-134
    li = s * i_n @ t_n.t()
    lt = s * t_n @ i_n.t()
    lab = torch.arange(img.shape[0], device=img.device)
    ce_i = torch.nn.functional.cross_entropy(li, lab, reduction="sum")
    ce_t = torch.nn.functional.cross_entropy(lt, lab, reduction="sum")
    (g[0] * ce_i + g[1] * ce_t).backward()
    rank = (li > li.diag().unsqueeze(1)).sum(1)
    return dict(ce=(ce_i.item(), ce_t.item()), d_img=img.grad, d_txt=txt.grad, dls=ls.grad, top1=(rank == 0).sum().item(),
                top5=(rank < 5).sum().item(), li=li.detach(), lt=lt.detach())


@pytest.mark.parametrize("b,e,ls0", [(8, 512, math.log(1 / 0.07)), (200, 512, 5.5), (512, 512, math.log(1 / 0.07)),
                                     (96, 768, 3.0), (64, 1024, 4.0), (300, 256, 2.0)])
def test_fused_head_single_rank(cuda_dev, b, e, ls0):
    from declip_b200 import functions as F_
    from declip_b200.loss_functions import ClipInfoCELoss
    torch.manual_seed(b + e)
    img = torch.randn(b, e, device=cuda_dev)
    # correlated pairs, weak enough that the softmax is not saturated at the given scale (label logit ~ 6 + noise)
    rho = min(0.6, 6.0 / math.exp(min(ls0, math.log(100.0))))
    txt = (rho * img + math.sqrt(1 - rho * rho) * torch.randn(b, e, device=cuda_dev)) * 3.0
    ls = torch.tensor([ls0], device=cuda_dev)
    ref = _reference(img, txt, ls, g=(0.5 / b, 0.5 / b))
    a, t, l = img.clone().requires_grad_(True), txt.clone().requires_grad_(True), ls.clone().requires_grad_(True)
    li, lt = F_.fused_clip_head(a, t, l, gather=False)
    crit = ClipInfoCELoss()
    loss, labels = crit(li, lt)
    loss.backward()
    torch.cuda.synchronize()
    want = 0.5 * (ref["ce"][0] + ref["ce"][1]) / b
    assert abs(loss.item() - want) <= 3e-3 * max(1.0, abs(want)), (loss.item(), want)
    assert li.shape == (b, b) and torch.equal(labels, torch.arange(b, device=cuda_dev))
    p1, p5 = crit.accuracy()
    slack = 100.0 * max(2, b // 64) / b + 1e-3                                      # near-ties may flip under bf16
    assert abs(p1.item() - 100.0 * ref["top1"] / b) <= slack and abs(p5.item() - 100.0 * ref["top5"] / b) <= slack
    assert _cos(a.grad, ref["d_img"]) > 0.995 and _cos(t.grad, ref["d_txt"]) > 0.995
    assert 0.97 < a.grad.norm().item() / ref["d_img"].norm().item() < 1.03
    assert abs(l.grad.item() - ref["dls"].item()) <= 0.03 * abs(ref["dls"].item()) + 1e-4, (l.grad.item(), ref["dls"].item())


def test_fused_head_compat_strips(cuda_dev):
    """strips[d] != NULL: the same kernel also stores the fp32 logit strips (compat mode)."""
    from declip_b200 import _lib, functions as F_, ops
    b, e = 136, 512
    torch.manual_seed(1)
    img, txt = torch.randn(b, e, device=cuda_dev), torch.randn(b, e, device=cuda_dev)
    ls = torch.tensor([2.5], device=cuda_dev)
    lib = ops.lib_for(img)
    L = F_.HeadLayout.get(lib, b, e)
    ws = torch.empty(L.total, device=cuda_dev)
    rows = torch.empty(b, 2 * e, device=cuda_dev, dtype=torch.bfloat16)
    feats, eps = (P * 2)(img.data_ptr(), txt.data_ptr()), (ctypes.c_float * 2)(0.0, 1e-10)
    st = P(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dc_head_prepare(feats, eps, 2, b, e, P(rows.data_ptr()), P(ws.data_ptr()), P(ls.data_ptr()), 100.0, st), "prep")
    strips = [torch.full((b, b), float("nan"), device=cuda_dev) for _ in range(2)]
    args = F_.head_args(b, b, e, 2 * e, 0, rows.data_ptr(), [rows.data_ptr()], ws.data_ptr(), strips=strips)
    _lib.check(lib.dc_head_forward(ctypes.byref(args), st), "fwd")
    torch.cuda.synchronize()
    ref = _reference(img, txt, ls)
    assert torch.allclose(strips[0], ref["li"], atol=3e-2) and torch.allclose(strips[1], ref["lt"], atol=3e-2)
    assert torch.allclose(strips[0], strips[1].t(), atol=1e-4)


@pytest.mark.parametrize("world,b,e", [(2, 64, 512), (4, 32, 512), (3, 160, 768), (8, 256, 512)])
def test_fused_head_emulated_ranks(cuda_dev, world, b, e):
    from declip_b200 import _lib, functions as F_, ops
    torch.manual_seed(world * 1000 + b)
    n = world * b
    img = torch.randn(n, e, device=cuda_dev)
    txt = (0.15 * img + torch.randn(n, e, device=cuda_dev)) * 2.0              # label logit ~ 5: per-row CE is O(1)
    ls = torch.tensor([3.5], device=cuda_dev)
    gs = [(0.5 / b / world * (1.0 + 0.1 * r), 0.5 / b / world * (1.0 - 0.05 * r)) for r in range(world)]   # per-rank upstream grads
    # ---- global reference: sum_r g_r[0] * sum CE(rows of rank r of li) + g_r[1] * (rows of lt)
    I, T, LS = img.clone().requires_grad_(True), txt.clone().requires_grad_(True), ls.clone().requires_grad_(True)
    i_n, t_n = I / I.norm(dim=-1, keepdim=True), T / (T.norm(dim=-1, keepdim=True) + 1e-10)
    s = LS.exp()
    s.data = torch.clamp(s.data, max=100)
    li, lt = s * i_n @ t_n.t(), s * t_n @ i_n.t()
    lab = torch.arange(n, device=cuda_dev)
    ce_i = torch.nn.functional.cross_entropy(li, lab, reduction="none").view(world, b).sum(1)
    ce_t = torch.nn.functional.cross_entropy(lt, lab, reduction="none").view(world, b).sum(1)
    total = sum(gs[r][0] * ce_i[r] + gs[r][1] * ce_t[r] for r in range(world))
    total.backward()
    # ---- emulated ranks through the C ABI
    lib = ops.lib_for(img)
    L = F_.HeadLayout.get(lib, b, e)
    st = P(torch.cuda.current_stream().cuda_stream)
    allb = torch.empty(n, 2 * e, device=cuda_dev, dtype=torch.bfloat16)
    wss = [torch.empty(L.total, device=cuda_dev) for _ in range(world)]
    eps = (ctypes.c_float * 2)(0.0, 1e-10)
    for r in range(world):            # every rank normalises its rows into the gather buffer ("all-gather")
        feats = (P * 2)(img[r * b:(r + 1) * b].data_ptr(), txt[r * b:(r + 1) * b].data_ptr())
        _lib.check(lib.dc_head_prepare(feats, eps, 2, b, e, P(allb[r * b:(r + 1) * b].data_ptr()), P(wss[r].data_ptr()),
                                       P(ls.data_ptr()), 100.0, st), "prep")
    argss = []
    for r in range(world):
        a = F_.head_args(b, n, e, 2 * e, r * b, allb[r * b:(r + 1) * b].data_ptr(), [allb.data_ptr()], wss[r].data_ptr())
        _lib.check(lib.dc_head_forward(ctypes.byref(a), st), "fwd")
        argss.append(a)
    torch.cuda.synchronize()
    for r in range(world):
        out = wss[r][L.out:L.out + 6]
        assert abs(out[0].item() - ce_i[r].item()) <= 3e-3 * ce_i[r].item() + 2e-3 * b, (r, out[0].item(), ce_i[r].item())
        assert abs(out[1].item() - ce_t[r].item()) <= 3e-3 * ce_t[r].item() + 2e-3 * b
    exch = torch.empty(world, 2 * b + 2, device=cuda_dev)
    for r in range(world):            # the 2b+2-float all-gather of the backward
        exch[r, :2 * b] = wss[r][L.lse:L.lse + 2 * b]
        exch[r, 2 * b], exch[r, 2 * b + 1] = gs[r]
    dls = 0.0
    for r in range(world):
        g = torch.tensor(gs[r], device=cuda_dev)
        d_img, d_txt = torch.empty(b, e, device=cuda_dev), torch.empty(b, e, device=cuda_dev)
        xraw = (P * 2)(img[r * b:(r + 1) * b].data_ptr(), txt[r * b:(r + 1) * b].data_ptr())
        dxo = (P * 2)(d_img.data_ptr(), d_txt.data_ptr())
        _lib.check(lib.dc_head_backward(ctypes.byref(argss[r]), P(g.data_ptr()), P(exch.data_ptr()), xraw, eps, dxo, st), "bwd")
        torch.cuda.synchronize()
        assert _cos(d_img, I.grad[r * b:(r + 1) * b]) > 0.995, r
        assert _cos(d_txt, T.grad[r * b:(r + 1) * b]) > 0.995, r
        assert 0.97 < d_txt.norm().item() / T.grad[r * b:(r + 1) * b].norm().item() < 1.03
        dls += wss[r][L.out + 10].item()            # summed over ranks by the gradient all-reduce
    assert abs(dls - LS.grad.item()) <= 0.03 * abs(LS.grad.item()) + 1e-4, (dls, LS.grad.item())


def test_fused_head_in_clip_model(cuda_dev):
    """clip_vitb32(fused_head=True): same loss and gradients as the compat (strip) path of the same model."""
    from declip_b200.loss_functions import ClipInfoCELoss
    from declip_b200.model import model_entry
    from oracle import synth
    res = []
    for fused in (False, True):
        cfg = dict(type='clip_vitb32', kwargs=dict(
            image_encode=dict(embed_dim=512, layers=1),
            text_encode=dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                             embed_dim=512, transformer_layers=1), clip=dict(use_allgather=False, fused_head=fused)))
        model = model_entry(cfg)
        model.load_state_dict(synth.clip_vit_state_dict(seed=5, v_layers=1, t_layers=1), strict=True)
        model = model.to(cuda_dev).train()
        images, ids = synth.synth_images(16, seed=5).to(cuda_dev), synth.synth_token_ids(16, seed=5).to(cuda_dev)
        crit = ClipInfoCELoss()
        loss, _ = crit(*model({"images": images, "captions": None, "token_ids": ids}))
        loss.backward()
        torch.cuda.synchronize()
        res.append((loss.item(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}, crit.accuracy()))
    (l0, g0, a0), (l1, g1, a1) = res
    assert abs(l0 - l1) < 2e-3
    assert set(g0) == set(g1)
    for k in g0:       # two bf16 evaluations of the same head: biases (sums with cancellation) are the noisiest
        assert _cos(g0[k], g1[k]) > (0.97 if k.endswith("bias") else 0.99), k
    assert abs(a0[0].item() - a1[0].item()) <= 100.0 / 16 + 1e-3
