"""GPU parity of the ModifiedResNet image tower (BASELINE configs[3]: clip_res50) against the golden vectors of the
reference's own clip_res50 and op-level checks of the convolution support kernels (through the C ABI)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.float().reshape(-1), b.float().reshape(-1)
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-20)).item()


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _nhwc(x):   # [B,C,H,W] -> [B*H*W, C] bf16
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1]).contiguous().bfloat16()


def _nchw(y, B, H, W):
    return y.float().view(B, H, W, -1).permute(0, 3, 1, 2)


def test_conv_bn_pool_ops(cuda_dev):
    from declip_b200 import functions_conv as C_
    torch.manual_seed(0)
    B, C, H, W = 3, 32, 14, 14
    x = torch.randn(B, C, H, W, device=cuda_dev)
    xh = _nhwc(x).requires_grad_(True)
    xr = xh.detach().float().view(B, H, W, C).permute(0, 3, 1, 2).requires_grad_(True)
    # 3x3 conv
    w3 = (torch.randn(64, C, 3, 3, device=cuda_dev) * 0.1).requires_grad_(True)
    y = C_.Conv3x3.apply(xh, w3, B, H, W)
    yr = torch.nn.functional.conv2d(xr, w3, padding=1)
    assert _rel(_nchw(y, B, H, W), yr) < 6e-3
    g = torch.randn_like(yr)
    y.backward(_nhwc(g))
    gw, gx = w3.grad.clone(), xh.grad.clone()
    w3.grad = None
    yr.backward(g)
    assert _cos(gw, w3.grad) > 0.999 and _cos(_nchw(gx, B, H, W), xr.grad) > 0.999
    # 1x1 conv + BN(+res, relu) + avgpool
    xh2 = _nhwc(x).requires_grad_(True)
    xr2 = xh2.detach().float().view(B, H, W, C).permute(0, 3, 1, 2).requires_grad_(True)
    w1 = (torch.randn(64, C, 1, 1, device=cuda_dev) * 0.2).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(64).to(cuda_dev).train()
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.normal_(0, 0.2)
    res = torch.randn(B, 64, H, W, device=cuda_dev)
    resh = _nhwc(res).requires_grad_(True)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    z = C_.Conv1x1.apply(xh2, w1)
    z = C_.BatchNorm2dNHWC.apply(z, resh, bn.weight, bn.bias, rm, rv, True, bn.eps, bn.momentum)
    z = C_.AvgPool2.apply(z, B, H, W)
    resr = resh.detach().float().view(B, H, W, 64).permute(0, 3, 1, 2).requires_grad_(True)
    zr = torch.nn.functional.avg_pool2d(torch.relu(bn(torch.nn.functional.conv2d(xr2, w1)) + resr), 2)
    assert _rel(_nchw(z, B, H // 2, W // 2), zr) < 1e-2
    g = torch.randn_like(zr)
    z.backward(_nhwc(g))
    gw1, gx2, gres, gg, gb = w1.grad.clone(), xh2.grad.clone(), resh.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone()
    w1.grad = None
    bn.zero_grad()
    zr.backward(g)
    assert _cos(gw1, w1.grad) > 0.995 and _cos(_nchw(gx2, B, H, W), xr2.grad) > 0.995
    assert _cos(_nchw(gres, B, H, W), resr.grad) > 0.999
    assert _cos(gg, bn.weight.grad) > 0.995 and _cos(gb, bn.bias.grad) > 0.995
    assert _rel(rm, bn.running_mean) < 2e-2 and _rel(rv, bn.running_var) < 2e-2
    # stem conv from the fp32 NCHW image
    img = torch.randn(2, 3, 224, 224, device=cuda_dev)
    ws = (torch.randn(32, 3, 3, 3, device=cuda_dev) * 0.2).requires_grad_(True)
    s = C_.StemConv.apply(img, ws)
    sr = torch.nn.functional.conv2d(img, ws, stride=2, padding=1)
    assert _rel(_nchw(s, 2, 112, 112), sr) < 6e-3
    g = torch.randn_like(sr)
    s.backward(_nhwc(g))
    gws = ws.grad.clone()
    ws.grad = None
    sr.backward(g)
    assert _cos(gws, ws.grad) > 0.999


def test_clip_res50_step_matches_reference_golden(cuda_dev):
    from declip_b200.loss_functions import ClipInfoCELoss
    from declip_b200.model import model_entry
    from oracle import golden
    g = golden.load("clip_res50_l1111_b4")
    c = g["case"]
    cfg = dict(type='clip_res50', kwargs=dict(
        image_encode=dict(embed_dim=c["embed_dim"], use_sync_bn=False, bn_group_size=1, layers=tuple(c["layers"])),
        text_encode=dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                         embed_dim=c["embed_dim"], transformer_layers=c["t_layers"]),
        clip=dict(use_allgather=False)))
    model = model_entry(cfg)
    sd, images, ids = golden.res_inputs(c)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda_dev).train()
    li, lt = model({"images": images.to(cuda_dev), "captions": None, "token_ids": ids.to(cuda_dev)})
    loss, _ = ClipInfoCELoss()(li, lt)
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - g["loss"]) <= 3e-2, (loss.item(), g["loss"])
    assert _cos(li.cpu(), g["logits_per_image"]) > 0.998
    params = dict(model.named_parameters())
    assert set(k for k, p in params.items() if p.grad is not None) == set(g["grads"])
    worst = []
    for k, ref in g["grads"].items():
        mine = params[k].grad.detach().float().reshape(-1).cpu()
        if ref["norm"] < 1e-7:
            continue
        nr = mine.norm().item() / (ref["norm"] + 1e-20)
        if (ref["sample"] != 0).sum().item() < 16:
            # the strided sample of a sparse gradient (token embedding: ~160 of 49409 rows are touched at batch 4) or a
            # scalar (logit_scale) holds too few non-zeros for a cosine; the full-tensor norm is compared instead
            lo, hi = (0.6, 1.5) if mine.numel() == 1 else (0.8, 1.2)   # a scalar gradient at batch 4 is the noisiest of all
            assert lo < nr < hi, (k, nr)
            continue
        worst.append((_cos(mine[golden.sample_index(mine.numel())], ref["sample"]), nr, k))
    worst.sort()
    txt = "\n".join("cos %.5f normratio %.4f %s" % w for w in worst[:12])
    # (1,1,1,1) network at batch 4: the noisiest parity case of the repo (BatchNorm over 4 samples; affine gradients of the
    # stem are sums with heavy cancellation over 50k positions; fp32 atomics make the value vary run to run: the stem
    # BatchNorm norms move +-10 %).  Same bounds as the full-depth golden (tests/parity_cases.py TOL["res"]; why bf16
    # activation storage cannot do better: profiles/r02_resnet_bf16_noise.md); block-level exactness is asserted by
    # test_bottleneck_block_isolated below.
    assert worst[0][0] > 0.6, txt
    assert worst[len(worst) // 10][0] > 0.8, txt
    assert all(0.8 < w[1] < 1.25 for w in worst), txt
    sdm = model.state_dict()
    for k, v in g["stats"].items():
        assert _rel(sdm[k].cpu(), v) < 5e-2, k


@pytest.mark.parametrize("kind", ["declip_res50", "filip_res50"])
def test_res50_wrappers_run(cuda_dev, kind):
    """declip_res50 / filip_res50 (declip.py:339-346, filip.py:146-153): the ResNet tower behind the DeCLIP / FILIP
    heads — forward dict + backward produce finite losses and gradients for every trainable parameter."""
    from declip_b200.loss_functions import ClipInfoCELoss, SimsiamLoss
    from declip_b200.model import model_entry
    from oracle import synth
    torch.manual_seed(0)
    E = 1024
    clip_kw = dict(use_allgather=True, text_mask_type='MLM', feature_dim=E)
    if kind == "declip_res50":
        clip_kw.update(return_nn_bank=True, nn_size=256)
    else:
        clip_kw.update(return_dense=True, select_topk=True, mask_rate=0.5, patch_number=14)
    cfg = dict(type=kind, kwargs=dict(
        image_encode=dict(embed_dim=E, use_sync_bn=False, bn_group_size=1, layers=(1, 1, 1, 1)),
        text_encode=dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                         embed_dim=E, transformer_layers=1), clip=clip_kw))
    model = model_entry(cfg).to(cuda_dev).train()
    B = 8
    ids = synth.synth_token_ids(B, seed=1)
    mlm = synth.synth_mlm(ids, seed=1)
    batch = {"images": torch.randn(B, 6, 224, 224, device=cuda_dev), "token_ids": ids.to(cuda_dev),
             "token_ids_aug": synth.synth_token_ids(B, seed=2).to(cuda_dev), "mlm": (mlm[0].to(cuda_dev), mlm[1])}
    out = model(batch, return_dict=True)
    crit = ClipInfoCELoss()
    if kind == "declip_res50":
        li1, li2, lt1, lt2 = out["logits"]
        p1, p2, z1, z2 = out["simsiam_features"]
        n1, n2, n1a, n2a = out["nn_text_logits"]
        loss = crit(li1, lt1)[0] + crit(li2, lt2)[0] + SimsiamLoss()(p1, z1, p2, z2) + out["text_self_supervised"] + \
            crit(n1, n1a)[0]
    else:
        loss = crit(*out["logits"])[0] + crit(*out["dense_logits"])[0]
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss).item()
    missing = [k for k, p in model.named_parameters() if p.requires_grad and p.grad is None]
    unused = {"text_label_predictor.weight", "text_label_predictor.bias", "visual.fc.weight", "visual.fc.bias"}
    assert set(missing) <= unused, missing
    for k, p in model.named_parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), k


@pytest.mark.parametrize("inplanes,planes,stride,H", [(64, 64, 1, 56), (256, 64, 1, 56), (256, 128, 2, 56), (512, 128, 1, 28),
                                                      (1024, 256, 1, 14), (1024, 512, 2, 14), (2048, 512, 1, 7)])
def test_bottleneck_block_isolated(cuda_dev, inplanes, planes, stride, H):
    """ONE Bottleneck (every variant of ModifiedResNet-50: projection / identity shortcut, stride 1 / 2, 56^2 .. 7^2) on
    the same input as an fp32 torch statement of modified_resnet.py:40-56: output, input gradient and every parameter
    gradient.  Run in isolation there is little depth for bf16 rounding to compound through (two ReLU gates: measured
    input-gradient cosine 0.997), so anything below ~0.99 here would be an indexing error (im2col / col2im / BatchNorm
    apply), not noise — which the deep-network goldens cannot tell apart (a randomly initialised 16-block network
    amplifies a 0.5 % perturbation ~1.25x per block)."""
    import torch.nn.functional as F
    from declip_b200.model.modified_resnet import Bottleneck
    torch.manual_seed(inplanes + planes + H)
    B = 4
    blk = Bottleneck(inplanes, planes, stride).to(cuda_dev).train()
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.2)
    x = torch.relu(torch.randn(B, inplanes, H, H, device=cuda_dev) + 0.3)
    xh = x.permute(0, 2, 3, 1).reshape(-1, inplanes).contiguous().bfloat16().requires_grad_(True)
    xr = xh.detach().float().view(B, H, H, inplanes).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    y, h2, w2 = blk.run(xh, B, H, H)
    g = torch.randn(B, planes * 4, h2, w2, device=cuda_dev)
    y.backward(g.permute(0, 2, 3, 1).reshape(-1, planes * 4).contiguous().bfloat16())
    mine = {k: p.grad.detach().clone() for k, p in blk.named_parameters()}
    # fp32 statement with the same parameters
    P = {k: p.detach().clone().requires_grad_(True) for k, p in blk.named_parameters()}

    def bn(t, n):
        return F.batch_norm(t, None, None, P[n + ".weight"], P[n + ".bias"], True, 0.1, 1e-5)
    o = F.relu(bn(F.conv2d(xr, P["conv1.weight"]), "bn1"))
    o = F.relu(bn(F.conv2d(o, P["conv2.weight"], padding=1), "bn2"))
    if stride > 1:
        o = F.avg_pool2d(o, stride)
    o = bn(F.conv2d(o, P["conv3.weight"]), "bn3")
    idn = xr
    if blk.downsample is not None:
        idn = bn(F.conv2d(F.avg_pool2d(xr, stride) if stride > 1 else xr, P["downsample.0.weight"]), "downsample.1")
    ref = F.relu(o + idn)
    ref.backward(g.bfloat16().float())
    torch.cuda.synchronize()
    assert _cos(_nchw(y, B, h2, w2), ref) > 0.9999
    assert _cos(_nchw(xh.grad, B, H, H), xr.grad) > 0.99
    for k, gref in P.items():
        assert _cos(mine[k], gref.grad) > 0.98, (k, _cos(mine[k], gref.grad))


@pytest.mark.parametrize("B,C,Cout,H", [(3, 64, 64, 56), (2, 128, 128, 28), (5, 256, 256, 14), (5, 512, 512, 7), (3, 64, 128, 56),
                                        (1, 128, 64, 28), (7, 512, 512, 7), (2, 32, 32, 112), (1, 32, 64, 112), (3, 64, 32, 20)])
def test_conv3x3_implicit_gemm(cuda_dev, B, C, Cout, H):
    """csrc/conv_igemm.cu: forward, input gradient and weight gradient through 4-D TMA boxes (padding = out-of-bounds
    zero fill; the weight gradient contracts over pixel boxes) — against torch conv2d in fp32 on the same bf16-rounded operands."""
    from declip_b200 import functions_conv as C_
    torch.manual_seed(B * 1000 + C + H)
    x = torch.randn(B, C, H, H, device=cuda_dev)
    xh = _nhwc(x).requires_grad_(True)
    xr = xh.detach().float().view(B, H, H, C).permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    w = (torch.randn(Cout, C, 3, 3, device=cuda_dev) * (9 * C) ** -0.5).requires_grad_(True)
    y = C_.Conv3x3.apply(xh, w, B, H, H)
    wr = w.detach().bfloat16().float().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr, padding=1)
    assert _rel(_nchw(y, B, H, H), yr) < 5e-3
    g = torch.randn_like(yr)
    y.backward(_nhwc(g))
    yr.backward(_nhwc(g).float().view(B, H, H, Cout).permute(0, 3, 1, 2))
    assert _rel(_nchw(xh.grad, B, H, H), xr.grad) < 6e-3
    assert _rel(w.grad, wr.grad) < 6e-3
