"""GPU parity of the DeFILIP wrapper (DeCLIP heads + FILIP token-wise logits on all four
(view, caption) combinations) against the golden vectors of the reference's own DEFILIP module
(tools/make_golden.py defilip_vitb32_l2_b8).  Tolerances as for DeCLIP / FILIP (bf16 storage, fp32 accumulate)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.float().reshape(-1), b.float().reshape(-1)
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-20)).item()


def test_defilip_step_matches_reference_golden(cuda_dev):
    from declip_b200.loss_functions import ClipInfoCELoss
    from declip_b200.model import model_entry
    from oracle import golden
    from oracle.declip_ref import DEFILIP_FILIP_WEIGHT
    from test_gpu_declip import _declip_loss
    g = golden.load("defilip_vitb32_l2_b8")
    c = g["case"]
    model = model_entry(dict(type='defilip_vitb32', kwargs=dict(
        image_encode=dict(embed_dim=c["embed_dim"], layers=c["v_layers"]),
        text_encode=dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                         embed_dim=c["embed_dim"], transformer_layers=c["t_layers"]),
        clip=dict(use_allgather=True, text_mask_type='MLM', return_nn_bank=True, feature_dim=c["embed_dim"],
                  nn_size=c["nn_size"], return_filip=True, dense_aug=True))))
    sd, images, mlm_ids, mlm_labels, ids_aug, bank = golden.defilip_inputs(c)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda_dev).train()
    model.nn_replacer_text.load_bank(bank, cuda_dev)
    batch = {"images": images.to(cuda_dev), "token_ids": mlm_ids.to(cuda_dev), "token_ids_aug": ids_aug.to(cuda_dev),
             "mlm": (mlm_ids.to(cuda_dev), mlm_labels)}
    out = model(batch, return_dict=True)
    loss, parts = _declip_loss(out)
    crit = ClipInfoCELoss()
    a = out["filip_aug"]
    filip = (crit(*out["filip"])[0] + crit(a[0], a[1])[0] + crit(a[2], a[3])[0] + crit(a[4], a[5])[0]) / 4
    loss = loss + filip * DEFILIP_FILIP_WEIGHT
    parts["filip"] = filip
    loss.backward()
    torch.cuda.synchronize()
    tol = dict(clip=2e-2, mlm=5e-2, nn=3e-2, simsiam=3e-3, nt_xent=3e-2, filip=2e-2)
    msg = {k: (parts[k].item(), g["parts"][k]) for k in parts}
    for k, v in g["parts"].items():
        assert abs(parts[k].item() - v) <= tol[k], msg
    assert abs(loss.item() - g["loss"]) <= 3e-2, (loss.item(), g["loss"])
    for key in ("logits", "logits_aug", "filip", "filip_aug"):
        for x, y in zip(out[key], g[key]):
            assert _cos(x.cpu(), y) > 0.999, key
    # nearest-neighbour logits: one column per neighbour.  The neighbour is an argmax over 1024 random bank entries whose
    # top-2 similarities can be closer than the bf16 noise of the query (exactness of the lookup itself is covered by
    # tests/test_gpu_declip.py::test_nn_bank_lookup_and_fifo), so a flipped neighbour may change at most 2 of 8 columns.
    for x, y in zip(out["nn_text_logits"], g["nn_text_logits"]):
        cols = torch.nn.functional.cosine_similarity(x.cpu().float(), y, dim=0)
        assert (cols > 0.999).sum().item() >= x.shape[1] - 2, cols
    params = dict(model.named_parameters())
    assert set(k for k, p in params.items() if p.grad is not None) == set(g["grads"])
    worst = []
    for k, ref in g["grads"].items():
        mine = params[k].grad.detach().float().reshape(-1).cpu()
        if ref["norm"] < 1e-6:
            assert mine.abs().max().item() < 1e-4, k        # biases feeding a BatchNorm: exactly zero gradient
            continue
        worst.append((_cos(mine[golden.sample_index(mine.numel())], ref["sample"]), mine.norm().item() / (ref["norm"] + 1e-20), k))
    worst.sort()
    txt = "\n".join("cos %.5f normratio %.4f %s" % w for w in worst[:10])
    assert all(w[0] > (0.95 if ("projector" in w[2] or "predictor" in w[2]) else 0.97) for w in worst), txt
    assert all(0.85 < w[1] < 1.15 for w in worst), txt
    assert model.nn_replacer_text.bank_ptr == g["bank_ptr"]
