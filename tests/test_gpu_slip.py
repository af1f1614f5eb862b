"""GPU parity of the SLIP wrapper (CLIP on a base view + SimCLR through `predictor_sim` on the
image tower's pre-projection feature, slip.py:196-284) against the golden vectors of the reference's own SLIP module
(tools/make_golden.py slip_vitb32_l2_b8); loss = ClipInfoCELoss + NT_Xent_gather as in slip_solver.py:470-510."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.float().reshape(-1), b.float().reshape(-1)
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-20)).item()


def test_vit_return_feature_gradient(cuda_dev):
    """The pre-projection feature is a differentiable output of the tower: features == feature @ proj, and a loss on the
    feature alone reaches the tower parameters."""
    from declip_b200.model.visual_transformer import visual_transformer_B32
    torch.manual_seed(0)
    vit = visual_transformer_B32(embed_dim=512, layers=2).to(cuda_dev).train()
    x = torch.randn(4, 3, 224, 224, device=cuda_dev)
    feats, pre = vit(x, return_feature=True)
    assert pre.shape == (4, 768) and pre.dtype == torch.float32
    assert _cos(feats, pre @ vit.proj.detach().float()) > 0.9995
    feats2, dense, pre2 = vit(x, return_dense=True, return_feature=True)
    assert dense.shape == (4, 49, 768) and torch.equal(pre2, pre)
    w = torch.randn_like(pre)
    (pre * w).sum().backward()
    g = vit.ln_post.weight.grad
    assert g is not None and g.abs().sum().item() > 0 and vit.proj.grad.abs().sum().item() == 0
    # d/d(ln_post.bias) of sum(pre * w) is colsum(w)
    assert _cos(vit.ln_post.bias.grad, w.sum(0)) > 0.999


def test_slip_step_matches_reference_golden(cuda_dev):
    from declip_b200.loss_functions import ClipInfoCELoss, NT_Xent_gather
    from declip_b200.model import model_entry
    from oracle import golden
    g = golden.load("slip_vitb32_l2_b8")
    c = g["case"]
    model = model_entry(dict(type='slip_vitb32', kwargs=dict(
        image_encode=dict(embed_dim=c["embed_dim"], layers=c["v_layers"]),
        text_encode=dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                         embed_dim=c["embed_dim"], transformer_layers=c["t_layers"]),
        clip=dict(use_allgather=True, return_sim=True, feature_dim=768, sim_dim=256))))
    sd, images, ids = golden.slip_inputs(c)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda_dev).train()
    out = model({"images": images.to(cuda_dev), "captions": None, "token_ids": ids.to(cuda_dev)}, return_dict=True)
    clip_loss, _ = ClipInfoCELoss()(*out["logits"])
    s1, g1, s2, g2 = out["sim_features"]
    simclr = NT_Xent_gather(c["batch"])(s1, g1, s2, g2)
    loss = clip_loss + simclr
    loss.backward()
    torch.cuda.synchronize()
    assert abs(clip_loss.item() - g["parts"]["clip"]) <= 2e-2, (clip_loss.item(), g["parts"]["clip"])
    assert abs(simclr.item() - g["parts"]["simclr"]) <= 3e-2, (simclr.item(), g["parts"]["simclr"])
    for a, b in zip(out["logits"], g["logits"]):
        assert _cos(a.cpu(), b) > 0.999
    for a, b in zip(out["sim_features"], g["sim_features"]):
        assert _cos(a.cpu(), b) > 0.995          # behind two BatchNorms over a batch of 8
    for a, b in zip(out["features"], g["features"]):
        assert torch.nn.functional.cosine_similarity(a.cpu(), b, dim=1).min().item() > 0.999
    params = dict(model.named_parameters())
    assert set(k for k, p in params.items() if p.grad is not None) == set(g["grads"])
    worst = []
    for k, ref in g["grads"].items():
        mine = params[k].grad.detach().float().reshape(-1).cpu()
        if ref["norm"] < 1e-6:
            assert mine.abs().max().item() < 1e-4, k      # biases feeding a BatchNorm: exactly zero gradient
            continue
        worst.append((_cos(mine[golden.sample_index(mine.numel())], ref["sample"]), mine.norm().item() / (ref["norm"] + 1e-20), k))
    worst.sort()
    txt = "\n".join("cos %.5f normratio %.4f %s" % w for w in worst[:10])
    assert all(w[0] > (0.95 if "predictor_sim" in w[2] else 0.97) for w in worst), txt
    assert all(0.85 < w[1] < 1.15 for w in worst), txt
    sdm = model.state_dict()
    for k, v in g["stats"].items():
        rel = ((sdm[k].cpu().float() - v).norm() / (v.norm() + 1e-12)).item()
        assert rel < 3e-2, (k, rel)
