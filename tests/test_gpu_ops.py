"""GPU: each C-ABI kernel against a plain PyTorch fp32 reference of the same op (bf16 I/O tolerance stated
per test).  Called through ctypes — the same path the product uses."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _cos(a, b):
    a, b = a.float().reshape(-1), b.float().reshape(-1)
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-20)).item()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("shape", [(128, 256, 64), (392, 768, 3072), (1000, 2304, 768), (8, 512, 768), (512, 4096, 512),
                                   (3072, 768, 1544)])
def test_gemm_majors(cuda_dev, a_mn, b_mn, shape):
    from declip_b200 import ops
    M, N, K = shape
    torch.manual_seed(0)
    a = (torch.randn((K, M) if a_mn else (M, K), device=cuda_dev) * 0.5).bfloat16()
    b = (torch.randn((K, N) if b_mn else (N, K), device=cuda_dev) * 0.5).bfloat16()
    want = (a.float().t() if a_mn else a.float()) @ (b.float() if b_mn else b.float().t())
    for bn in (0, 128, 256):
        got = ops.gemm(a, b, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), epilogue=ops.EPI_F32, block_n=bn)
        assert _rel(got, want) < 1e-5, (shape, a_mn, b_mn, bn)     # fp32 accumulate of exact bf16 products


@pytest.mark.parametrize("two_cta", [0, 1])
def test_gemm_1cta_and_2cta_variants(cuda_dev, two_cta):
    """Both GEMM kernels (gemm.cu 128x256 per CTA, gemm2.cu 256x256 per CTA pair) on cluster-tile-sized problems."""
    from declip_b200 import _lib, ops
    ops.lib_for(torch.zeros(1, device=cuda_dev))
    old = _lib.set_gemm_2cta(bool(two_cta))
    try:
        torch.manual_seed(7)
        for (a_mn, b_mn) in ((0, 0), (0, 1), (1, 1)):
            for (M, N, K) in ((256, 256, 64), (1000, 2304, 768), (3072, 768, 1544), (304, 520, 200)):
                a = (torch.randn((K, M) if a_mn else (M, K), device=cuda_dev) * 0.5).bfloat16()
                b = (torch.randn((K, N) if b_mn else (N, K), device=cuda_dev) * 0.5).bfloat16()
                want = (a.float().t() if a_mn else a.float()) @ (b.float() if b_mn else b.float().t())
                got = ops.gemm(a, b, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), epilogue=ops.EPI_F32)
                assert _rel(got, want) < 1e-5, (two_cta, a_mn, b_mn, M, N, K)
    finally:
        _lib.set_gemm_2cta(bool(old))


def test_gemm_epilogues(cuda_dev):
    from declip_b200 import ops
    torch.manual_seed(1)
    M, N, K = 1000, 768, 512
    a = (torch.randn(M, K, device=cuda_dev) * 0.5).bfloat16()
    b = (torch.randn(N, K, device=cuda_dev) * 0.5).bfloat16()
    bias = torch.randn(N, device=cuda_dev)
    aux = torch.randn(M, N, device=cuda_dev).bfloat16()
    want = a.float() @ b.float().t()
    assert _rel(ops.gemm(a, b, bias=bias), want + bias) < 4e-3            # bf16 output rounding
    h, u = ops.gemm(a, b, bias=bias, epilogue=ops.EPI_BF16_GELU)
    uu = want + bias
    assert _rel(u, uu) < 4e-3 and _rel(h, uu * torch.sigmoid(1.702 * uu)) < 4e-3
    assert _rel(ops.gemm(a, b, bias=bias, aux=aux, epilogue=ops.EPI_BF16_RESID), want + bias + aux.float()) < 4e-3
    x = aux.float()
    s = torch.sigmoid(1.702 * x)
    assert _rel(ops.gemm(a, b, aux=aux, epilogue=ops.EPI_BF16_DGELU), want * (s * (1 + 1.702 * x * (1 - s)))) < 5e-3
    cs = torch.ones(N, device=cuda_dev)
    o = ops.gemm(a, b, aux=aux, epilogue=ops.EPI_BF16_DGELU, colsum=cs)     # fused bias gradient of the output
    assert _rel(cs - 1, (want * (s * (1 + 1.702 * x * (1 - s)))).sum(0)) < 2e-3
    # ragged N / M through the TMA-store epilogue (clipping) and a strided output view
    big = torch.zeros(1000, 1024, device=cuda_dev, dtype=torch.bfloat16)
    ops.gemm(a, b[:520], bias=bias[:520].contiguous(), out=big[:, 8:528])
    assert _rel(big[:, 8:528], want[:, :520] + bias[:520]) < 4e-3
    assert big[:, :8].abs().max().item() == 0 and big[:, 528:].abs().max().item() == 0
    sc = torch.tensor([2.5], device=cuda_dev)
    assert _rel(ops.gemm(a, b, epilogue=ops.EPI_F32, alpha_dev=sc), 2.5 * want) < 1e-5
    acc = torch.ones(N, K, device=cuda_dev)
    dy = (torch.randn(4096, N, device=cuda_dev) * 0.5).bfloat16()
    xx = (torch.randn(4096, K, device=cuda_dev) * 0.5).bfloat16()
    for sp in (0, 1, 5):
        acc.fill_(1.0)
        ops.gemm(dy, xx, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC, out=acc, splits=sp)
        assert _rel(acc, dy.float().t() @ xx.float() + 1.0) < 1e-5


@pytest.mark.parametrize("width", [512, 768])
@pytest.mark.parametrize("rows", [1, 50, 3333])
def test_layernorm(cuda_dev, width, rows):
    from declip_b200 import ops
    torch.manual_seed(2)
    x = (torch.randn(rows, width, device=cuda_dev) * 2 + 0.3).bfloat16()
    g = 1 + 0.1 * torch.randn(width, device=cuda_dev)
    b = 0.1 * torch.randn(width, device=cuda_dev)
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    xr = x.float().requires_grad_(True)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (width,), gr, br, 1e-5)
    assert _rel(y, yr) < 4e-3
    assert torch.allclose(mean, xr.mean(1), atol=1e-4) and _rel(rstd, 1 / torch.sqrt(xr.var(1, unbiased=False) + 1e-5)) < 1e-4
    dy = torch.randn(rows, width, device=cuda_dev).bfloat16()
    dres = torch.randn(rows, width, device=cuda_dev).bfloat16()
    yr.backward(dy.float())
    dx, dg, db = ops.layernorm_bwd(dy, x, g, mean, rstd, dres)
    assert _rel(dx, xr.grad + dres.float()) < 5e-3
    assert _rel(dg, gr.grad) < 1e-3 and _rel(db, br.grad) < 1e-3
    dx2, _, _, dcol = ops.layernorm_bwd(dy, x, g, mean, rstd, None, with_colsum=True)
    assert _rel(dx2, xr.grad) < 5e-3
    assert _rel(dcol, xr.grad.sum(0)) < 2e-3 or (dcol - xr.grad.sum(0)).abs().max() < 2e-2   # fused bias gradient


def test_colsum(cuda_dev):
    from declip_b200 import ops
    x = torch.randn(2500, 2304, device=cuda_dev).bfloat16()
    out = torch.ones(2304, device=cuda_dev)
    ops.colsum(x, out)
    assert _rel(out, x.float().sum(0) + 1) < 1e-4
    # strided view (class-token rows of a [B, L*W] tensor)
    y = torch.randn(37, 50 * 768, device=cuda_dev).bfloat16()
    assert _rel(ops.colsum(y[:, :768]), y[:, :768].float().sum(0)) < 1e-4


def _ref_attention(qkv, B, L, H, causal):
    D = H * 64
    q, k, v = qkv.float().view(B, L, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    if causal:
        s = s + torch.full((L, L), float("-inf"), device=qkv.device).triu_(1)
    p = torch.softmax(s, -1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * L, D)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("L,H,causal", [(50, 12, False), (77, 8, True), (77, 8, False), (17, 4, True), (64, 2, False),
                                        (80, 2, True), (1, 2, True)])
def test_attention(cuda_dev, L, H, causal):
    from declip_b200 import ops
    torch.manual_seed(3)
    B, D = 5, H * 64
    qkv = torch.randn(B * L, 3 * D, device=cuda_dev).bfloat16()
    out, lse = ops.attention_fwd(qkv, B, L, H, causal)
    qr = qkv.float().requires_grad_(True)
    o_ref, lse_ref = _ref_attention(qr, B, L, H, causal)
    assert _rel(out, o_ref) < 1e-2                                   # P is rounded to bf16 before P V
    assert torch.allclose(lse.view(B, H, L), lse_ref, atol=2e-3)
    dout = torch.randn(B * L, D, device=cuda_dev).bfloat16()
    o_ref.backward(dout.float())
    dbias = torch.ones(3 * D, device=cuda_dev)
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, L, H, causal, dbias=dbias)
    assert _rel(dqkv, qr.grad) < 2e-2 and _cos(dqkv, qr.grad) > 0.9995
    assert _rel(dbias - 1, qr.grad.sum(0)) < 2e-2                      # fused in_proj_bias gradient


@pytest.mark.parametrize("B,L,H,causal", [(6, 50, 12, 0), (3, 77, 8, 1), (2, 128, 2, 1), (4, 50, 32, 0), (3, 17, 2, 0),
                                          (150, 50, 12, 0), (3, 50, 3, 0), (300, 50, 12, 0), (230, 77, 8, 1)])
def test_attention_tcgen05_core_against_fp32_and_the_mma_core(cuda_dev, B, L, H, causal):
    """The tcgen05 core (default) and the mma.sync core (fallback; also taken for odd head counts at L <= 64) against
    the fp32 restatement, at tile-boundary shapes: two pairs per tile (L <= 64), one pair (L > 64), L = 128 (full
    tile), more tiles than CTAs (B = 150; B = 300 / 230: >= 6 tiles per CTA also for the two-CTAs-per-SM forward), and a
    shape outside the tcgen05 envelope (3 heads at L = 50)."""
    from declip_b200 import _lib, ops
    torch.manual_seed(11)
    D = H * 64
    qkv = (torch.randn(B * L, 3 * D, device=cuda_dev) * 1.5).bfloat16()
    dout = torch.randn(B * L, D, device=cuda_dev).bfloat16()
    qr = qkv.float().requires_grad_(True)
    o_ref, lse_ref = _ref_attention(qr, B, L, H, causal)
    o_ref.backward(dout.float())
    try:
        for tc in (True, False):
            if not tc and L > 80:
                continue
            _lib.set_attention_tc(tc)
            out, lse = ops.attention_fwd(qkv, B, L, H, causal)
            dbias = torch.zeros(3 * D, device=cuda_dev)
            dqkv = ops.attention_bwd(qkv, out, dout, lse, B, L, H, causal, dbias=dbias)
            assert _cos(out, o_ref) > 0.9999 and torch.allclose(lse.view(B, H, L), lse_ref, atol=2e-3), tc
            assert _cos(dqkv, qr.grad) > 0.9995 and torch.isfinite(dqkv.float()).all(), tc
            ref_b = qr.grad.sum(0)
            # Q and V slices of the in_proj bias gradient; the K slice is zero in exact arithmetic (rounding noise in
            # both the reference and the kernels)
            for sl in (slice(0, D), slice(2 * D, 3 * D)):
                assert _cos(dbias[sl], ref_b[sl]) > 0.999, tc
            assert dbias[D:2 * D].abs().max() < 0.05 * dbias.abs().max(), tc
    finally:
        _lib.set_attention_tc(True)


def test_embeddings(cuda_dev):
    from declip_b200 import ops
    torch.manual_seed(4)
    img = torch.randn(3, 3, 224, 224, device=cuda_dev)
    p = ops.patchify(img, 32)
    ref = torch.nn.functional.unfold(img, kernel_size=32, stride=32).transpose(1, 2).reshape(3 * 49, 3072)
    assert torch.equal(p.float(), ref.bfloat16().float())
    # 6-channel two-view layout (DeCLIP): second view through a channel-offset view
    img6 = torch.randn(2, 6, 224, 224, device=cuda_dev)
    p2 = ops.patchify(img6[:, 3:], 32)
    ref2 = torch.nn.functional.unfold(img6[:, 3:], kernel_size=32, stride=32).transpose(1, 2).reshape(2 * 49, 3072)
    assert torch.equal(p2.float(), ref2.bfloat16().float())
    from oracle import synth
    ids = synth.synth_token_ids(9, seed=5).to(cuda_dev)
    table = torch.randn(49409, 512, device=cuda_dev)
    pos = torch.randn(77, 512, device=cuda_dev)
    x = ops.text_embed(ids, table, pos)
    assert _rel(x, (table[ids] + pos).reshape(-1, 512)) < 4e-3
    assert torch.equal(ops.eot_index(ids).long(), torch.arange(9, device=cuda_dev) * 77 + ids.argmax(1))
    dx = torch.randn(9 * 77, 512, device=cuda_dev).bfloat16()
    dt, dp = ops.text_embed_bwd(ids, dx, 49409)
    ref_t = torch.zeros(49409, 512, device=cuda_dev).index_add_(0, ids.reshape(-1), dx.float())
    assert _rel(dt, ref_t) < 1e-5 and _rel(dp, dx.float().view(9, 77, 512).sum(0)) < 1e-4


def test_head_functions(cuda_dev):
    """ClipLogits + ClipInfoCE against the oracle restatement (clip.py:129-141, loss.py:40-50), incl. the
    clamped-scale gradient semantics and top-1/top-5."""
    from declip_b200 import functions as F_
    from declip_b200.loss_functions import ClipInfoCELoss
    from oracle import clip_ref
    torch.manual_seed(5)
    for ls0 in (math.log(1 / 0.07), 5.5):
        fi = torch.randn(48, 512, device=cuda_dev, requires_grad=True)
        ft = torch.randn(48, 512, device=cuda_dev, requires_grad=True)
        ls = torch.tensor([ls0], device=cuda_dev, requires_grad=True)
        li, lt = F_.ClipLogits.apply(fi, ft, ls, False, True)
        crit = ClipInfoCELoss()
        loss, labels = crit(li, lt)
        loss.backward()
        fi_r, ft_r, ls_r = (t.detach().cpu().clone().requires_grad_(True) for t in (fi, ft, ls))
        li_r, lt_r, _, _ = clip_ref.clip_logits(fi_r, ft_r, ls_r)
        loss_r, labels_r = clip_ref.clip_info_ce(li_r, lt_r)
        loss_r.backward()
        assert _rel(li.cpu(), li_r) < 6e-3 and _rel(lt.cpu(), lt_r) < 6e-3       # bf16 features into the GEMM
        assert abs(loss.item() - loss_r.item()) < 2e-2
        assert torch.equal(labels.cpu(), labels_r)
        assert _cos(fi.grad.cpu(), fi_r.grad) > 0.995 and _cos(ft.grad.cpu(), ft_r.grad) > 0.995
        assert abs(ls.grad.item() - ls_r.grad.item()) < 0.05 * abs(ls_r.grad.item()) + 1e-3
        p1, p5 = crit.accuracy()
        r1, r5 = clip_ref.accuracy(li.detach().cpu(), labels_r)
        assert abs(p1.item() - r1.item()) < 1e-3 and abs(p5.item() - r5.item()) < 1e-3


def test_fused_adamw_matches_torch(cuda_dev):
    from declip_b200.optim import FusedAdamW
    torch.manual_seed(0)
    shapes = [(768, 3072), (3072,), (1,), (49409, 16), (50, 768), (7,)]
    ps = [torch.randn(s, device=cuda_dev).requires_grad_(True) for s in shapes]
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    groups = lambda xs: [dict(params=xs[:3], weight_decay=0.1), dict(params=xs[3:], weight_decay=0.0, lr=3e-4)]
    mine = FusedAdamW(groups(ps), lr=1e-3, betas=(0.9, 0.98), eps=1e-6)
    ref = torch.optim.AdamW(groups(qs), lr=1e-3, betas=(0.9, 0.98), eps=1e-6)
    for it in range(4):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        if it == 2:                                   # a scheduler changing the lr forces a table rebuild
            for o in (mine, ref):
                o.param_groups[0]["lr"] = 5e-4
        mine.step()
        ref.step()
    for p, q in zip(ps, qs):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), (p - q).abs().max()


def test_fused_adamw_state_dict_and_skipped_params(cuda_dev):
    """torch.optim.AdamW checkpoints store `step` as a tensor; a parameter with grad=None on some iteration keeps its
    own step count (torch semantics) instead of raising."""
    from declip_b200.optim import FusedAdamW
    torch.manual_seed(1)
    ps = [torch.randn(s, device=cuda_dev).requires_grad_(True) for s in [(64, 32), (32,), (5,)]]
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    ref = torch.optim.AdamW(qs, lr=1e-3, weight_decay=0.05)
    for _ in range(2):
        for q in qs:
            q.grad = torch.randn_like(q)
        ref.step()
    with torch.no_grad():
        for p, q in zip(ps, qs):
            p.copy_(q)
    mine = FusedAdamW(ps, lr=1e-3, weight_decay=0.05)
    import copy
    mine.load_state_dict(copy.deepcopy(ref.state_dict()))  # tensor-valued `step`; deepcopy: load_state_dict aliases same-device tensors
    for it in range(3):
        for i, (p, q) in enumerate(zip(ps, qs)):
            if it == 1 and i == 2:                         # the third parameter skips one iteration
                p.grad = q.grad = None
                continue
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        mine.step()
        ref.step()
    for p, q in zip(ps, qs):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), (p - q).abs().max()
    assert int(mine.state[ps[0]]["step"]) == 5 and int(mine.state[ps[2]]["step"]) == 4


def test_fused_adamw_rewrites_registered_shadow(cuda_dev):
    """The kernel refreshes a registered bf16 mirror (padded layouts included) and bumps the version counter."""
    from declip_b200.optim import FusedAdamW
    from declip_b200.runtime import shadow_current, weight_shadow
    torch.manual_seed(2)
    w = torch.nn.Parameter(torch.randn(49, 512, device=cuda_dev))
    sh = weight_shadow(w, pad_rows=56)
    assert sh.shape == (56, 512) and torch.equal(sh[:49], w.detach().bfloat16()) and not sh[49:].any()
    opt = FusedAdamW([w], lr=1e-2)
    v0 = w._version
    w.grad = torch.randn_like(w)
    opt.step()
    assert w._version == v0 + 1 and shadow_current(w)
    assert weight_shadow(w, pad_rows=56) is sh                       # no re-cast needed ...
    assert torch.equal(sh[:49], w.detach().bfloat16()) and not sh[49:].any()    # ... because the kernel wrote it
    with torch.no_grad():
        w.mul_(2.0)                                                    # any other in-place update re-casts lazily
    assert not shadow_current(w)
    assert torch.equal(weight_shadow(w, pad_rows=56)[:49], w.detach().bfloat16())


@pytest.mark.parametrize("M,groups,K", [(300, 37, 256), (128, 16, 64), (1000, 513, 256), (49 * 6, 6, 256)])
def test_gemm_groupmax16_epilogue(cuda_dev, M, groups, K):
    """DC_EPI_F32_GROUPMAX16 (FILIP, filip.py:93-104): max / arg-max over every 16 score columns straight out of TMEM."""
    from declip_b200 import ops
    torch.manual_seed(M + groups)
    a = torch.randn(M, K, device=cuda_dev).bfloat16()
    b = torch.randn(groups * 16, K, device=cuda_dev).bfloat16()
    s = torch.tensor([1.7], device=cuda_dev)
    mx, arg = ops.gemm(a, b, epilogue=ops.EPI_F32_GROUPMAX16, alpha_dev=s)
    ref = (1.7 * a.float() @ b.float().t()).view(M, groups, 16)
    rmx, rarg = ref.max(dim=-1)
    assert mx.shape == (M, groups) and arg.dtype == torch.uint8
    assert torch.allclose(mx, rmx, rtol=1e-5, atol=1e-4)
    same = arg.long() == rarg
    # a different index is only acceptable on an exact tie of the two candidates
    picked = ref.gather(-1, arg.long().unsqueeze(-1)).squeeze(-1)
    assert torch.allclose(picked[~same], rmx[~same], rtol=1e-6, atol=1e-5)
    assert same.float().mean().item() > 0.999
