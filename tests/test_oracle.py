"""CPU: the oracle restatement (oracle/clip_ref.py) against the golden vectors produced by the
reference's own modules (tools/make_golden.py).  fp32 vs fp32, so the tolerance is tight."""
import pytest
import torch

from oracle import clip_ref, golden, synth


def _run(name):
    g = golden.load(name)
    c = g["case"]
    sd = synth.clip_vit_state_dict(seed=c["seed"], embed_dim=c["embed_dim"], v_layers=c["v_layers"],
                                   t_layers=c["t_layers"])
    images = synth.synth_images(c["batch"], seed=c["seed"])
    ids = synth.synth_token_ids(c["batch"], seed=c["seed"])
    return g, clip_ref.clip_step(sd, images, ids)


@pytest.mark.parametrize("name", ["clip_vitb32_l2_b8", "clip_vitb32_l12_b32"])
def test_restatement_matches_reference_golden(name):
    g, out = _run(name)
    assert abs(out["loss"].item() - g["loss"]) <= 2e-5
    torch.testing.assert_close(out["image_features"], g["image_features"], rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(out["text_features"], g["text_features"], rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(out["logits_per_image"], g["logits_per_image"], rtol=2e-4, atol=5e-4)
    torch.testing.assert_close(out["logits_per_text"], g["logits_per_text"], rtol=2e-4, atol=5e-4)
    assert torch.equal(out["labels"], g["labels"])
    # every trainable parameter of the reference has a gradient here, and they agree
    assert set(out["grads"]) == set(g["grads"])
    assert "visual.conv1.weight" not in g["grads"]          # frozen: visual_transformer.py:12,45-51
    for k, ref in g["grads"].items():
        mine = out["grads"][k].reshape(-1)
        samp = mine[golden.sample_index(mine.numel())]
        denom = ref["sample"].norm().item() + 1e-12
        assert (samp - ref["sample"]).norm().item() / denom <= 2e-3, k
        assert abs(mine.norm().item() - ref["norm"]) <= 2e-3 * ref["norm"] + 1e-9, k


def test_known_answers():
    # initial CLIP loss ~ ln(N) at random init
    g = golden.load("clip_vitb32_l12_b32")
    import math
    assert abs(g["loss"] - math.log(32)) < 0.5
    ids = synth.synth_token_ids(16, seed=3)
    assert (ids[:, 0] == synth.SOT).all()
    assert (ids.max(dim=1).values == synth.EOT).all()
    eot = ids.argmax(dim=1)
    for b in range(16):
        assert (ids[b, eot[b] + 1:] == 0).all()


def test_logit_scale_clamp_semantics():
    # clip.py:133-134: value clamped at 100 but gradient flows as exp(logit_scale)
    ls = torch.tensor([5.5], requires_grad=True)
    i = torch.nn.functional.normalize(torch.randn(4, 8), dim=-1)
    t = torch.nn.functional.normalize(torch.randn(4, 8), dim=-1)
    li, _, _, _ = clip_ref.clip_logits(i, t, ls)
    assert torch.allclose(li, 100.0 * (i / i.norm(dim=-1, keepdim=True)) @ (t / (t.norm(dim=-1, keepdim=True) + 1e-10)).t(),
                          atol=1e-4)
    li.sum().backward()
    raw = ((i / i.norm(dim=-1, keepdim=True)) @ (t / (t.norm(dim=-1, keepdim=True) + 1e-10)).t()).sum()
    assert abs(ls.grad.item() - (torch.exp(torch.tensor(5.5)) * raw).item()) < 1e-2


def test_declip_restatement_matches_reference_golden():
    """oracle/declip_ref.py (DECLIP.forward + NN bank + SimSiam + MLM + the solver's loss composition) against the
    golden vectors produced by the reference's own DECLIP module."""
    from oracle import declip_ref
    name = "declip_vitb32_l2_b8"
    g = golden.load(name)
    sd, images, mlm_ids, mlm_labels, ids_aug, bank = golden.declip_inputs(g["case"])
    res = declip_ref.declip_step(sd, images, mlm_ids, mlm_labels, ids_aug, bank)
    assert abs(res["loss"].item() - g["loss"]) <= 5e-5
    for k, v in g["parts"].items():
        assert abs(res["parts"][k].item() - v) <= 1e-4, k
    for key in ("logits", "logits_aug", "nn_text_logits", "simsiam_features", "features"):
        for a, b in zip(res["out"][key], g[key]):
            torch.testing.assert_close(a.detach(), b, rtol=3e-4, atol=1e-3)
    assert set(res["grads"]) == set(g["grads"])
    for k, ref in g["grads"].items():
        mine = res["grads"][k].reshape(-1)
        samp = mine[golden.sample_index(mine.numel())]
        assert (samp - ref["sample"]).norm().item() <= 3e-3 * (ref["sample"].norm().item() + 1e-9) + 1e-7, k
    for k, v in g["stats"].items():                                   # BatchNorm running statistics
        torch.testing.assert_close(res["stats"][k], v, rtol=1e-4, atol=1e-5)
    assert res["bank_ptr"] == g["bank_ptr"]
    assert abs(res["bank"].double().sum().item() - g["bank_checksum"]) < 1e-3   # FIFO enqueue of both text views


def _check_grads(res, g, tol=3e-3):
    assert set(res["grads"]) == set(g["grads"])
    for k, ref in g["grads"].items():
        mine = res["grads"][k].reshape(-1)
        samp = mine[golden.sample_index(mine.numel())]
        assert (samp - ref["sample"]).norm().item() <= tol * (ref["sample"].norm().item() + 1e-9) + 1e-7, k


def test_slip_restatement_matches_reference_golden():
    """oracle/slip_ref.py (SLIP.forward with the pre-projection feature + predictor_sim, ClipInfoCE + NT_Xent_gather)."""
    from oracle import slip_ref
    g = golden.load("slip_vitb32_l2_b8")
    sd, images, ids = golden.slip_inputs(g["case"])
    res = slip_ref.slip_step(sd, images, ids)
    assert abs(res["loss"].item() - g["loss"]) <= 5e-5
    for k, v in g["parts"].items():
        assert abs(res["parts"][k].item() - v) <= 1e-4, k
    for key in ("logits", "sim_features", "features"):
        for a, b in zip(res["out"][key], g[key]):
            torch.testing.assert_close(a.detach(), b, rtol=3e-4, atol=1e-3)
    _check_grads(res, g)
    for k, v in g["stats"].items():
        torch.testing.assert_close(res["stats"][k], v, rtol=1e-4, atol=1e-5)


def test_defilip_restatement_matches_reference_golden():
    """oracle/declip_ref.defilip_step (DeCLIP forward + token-wise logits on four view x caption combinations)."""
    from oracle import declip_ref
    g = golden.load("defilip_vitb32_l2_b8")
    sd, images, mlm_ids, mlm_labels, ids_aug, bank = golden.defilip_inputs(g["case"])
    res = declip_ref.defilip_step(sd, images, mlm_ids, mlm_labels, ids_aug, bank)
    assert abs(res["loss"].item() - g["loss"]) <= 5e-5
    for k, v in g["parts"].items():
        assert abs(res["parts"][k].item() - v) <= 1e-4, k
    for key in ("logits", "logits_aug", "nn_text_logits", "filip", "filip_aug"):
        for a, b in zip(res["out"][key], g[key]):
            torch.testing.assert_close(a.detach(), b, rtol=3e-4, atol=1e-3)
    _check_grads(res, g)
    assert res["bank_ptr"] == g["bank_ptr"]


def test_filip_restatement_matches_reference_golden():
    from oracle import filip_ref
    g = golden.load("filip_vitb32_l2_b8")
    sd, images, mlm_ids, mlm_labels = golden.filip_inputs(g["case"])
    res = filip_ref.filip_step(sd, images, mlm_ids)
    assert abs(res["loss"].item() - g["loss"]) <= 5e-5
    for key in ("logits", "dense_logits"):
        for a, b in zip(res["out"][key], g[key]):
            torch.testing.assert_close(a.detach(), b, rtol=3e-4, atol=1e-3)
    assert set(res["grads"]) == set(g["grads"])
    for k, ref in g["grads"].items():
        mine = res["grads"][k].reshape(-1)
        samp = mine[golden.sample_index(mine.numel())]
        assert (samp - ref["sample"]).norm().item() <= 3e-3 * (ref["sample"].norm().item() + 1e-9) + 1e-7, k


def test_resnet_restatement_matches_reference_golden():
    from oracle import resnet_ref
    g = golden.load("clip_res50_l1111_b4")
    sd, images, ids = golden.res_inputs(g["case"])
    res = resnet_ref.clip_res_step(sd, images, ids)
    assert abs(res["loss"].item() - g["loss"]) <= 5e-5
    torch.testing.assert_close(res["logits_per_image"], g["logits_per_image"], rtol=3e-4, atol=1e-3)
    assert set(res["grads"]) == set(g["grads"])
    for k, ref in g["grads"].items():
        mine = res["grads"][k].reshape(-1)
        samp = mine[golden.sample_index(mine.numel())]
        assert (samp - ref["sample"]).norm().item() <= 5e-3 * (ref["sample"].norm().item() + 1e-9) + 1e-6, k
    for k, v in g["stats"].items():
        torch.testing.assert_close(res["stats"][k], v, rtol=1e-4, atol=1e-5)


def test_nt_xent_restatements_match_reference_golden():
    from oracle import loss_ref
    g = golden.load("nt_xent")
    z_i, z_j, z_ib, z_jb, rank = loss_ref.inputs()
    a, b = z_i.clone().requires_grad_(True), z_j.clone().requires_grad_(True)
    l1 = loss_ref.nt_xent(a, b)
    l1.backward()
    assert abs(l1.item() - g["nt_xent"]) < 1e-5
    torch.testing.assert_close(a.grad, g["nt_xent_grad"], rtol=1e-4, atol=1e-6)
    c, d = z_i.clone().requires_grad_(True), z_j.clone().requires_grad_(True)
    l2 = loss_ref.nt_xent_gather(c, z_ib, d, z_jb, rank=rank)
    l2.backward()
    assert abs(l2.item() - g["nt_xent_gather"]) < 1e-5
    torch.testing.assert_close(c.grad, g["nt_xent_gather_grad_i"], rtol=1e-4, atol=1e-6)
