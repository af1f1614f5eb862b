"""TEST INFRASTRUCTURE — CPU/fp32 restatement (oracle) of SLIP.forward (prototype/model/slip.py:196-284: CLIP on the base
view, `predictor_sim` on the pre-projection feature of two augmented views) and the solver's loss
(prototype/solver/slip_solver.py:470-510: ClipInfoCELoss + NT_Xent_gather, weights 1 / 1), world size 1.
Pinned by tests/test_oracle.py against the golden generated from the reference's own SLIP module."""
import torch
import torch.nn.functional as F

from . import clip_ref, loss_ref


def predictor_sim(x, sd, stats, p="predictor_sim."):
    # slip.py:49-108 with out_bn=False: Linear-BN-ReLU, Linear-BN-ReLU, Linear (BN = per-rank nn.BatchNorm1d, linklink/nn.py:4-5)
    def bn(y, name):
        return F.batch_norm(y, stats[name + ".running_mean"], stats[name + ".running_var"], sd[name + ".weight"],
                            sd[name + ".bias"], True, 0.1, 1e-5)
    x = F.relu(bn(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"]), p + "bn1"))
    x = F.relu(bn(F.linear(x, sd[p + "linear2.weight"], sd[p + "linear2.bias"]), p + "bn2"))
    return F.linear(x, sd[p + "linear3.weight"], sd[p + "linear3.bias"])


def slip_forward(params, stats, images9, ids):
    base, v1, v2 = torch.split(images9, [3, 3, 3], dim=1)                                   # slip.py:241
    tf = clip_ref.encode_text(ids, params, prefix="text_encoder.")                          # :244
    fi = clip_ref.encode_image(base, params)                                                # :246
    _, feat1 = clip_ref.encode_image(v1, params, return_feature=True)                       # :247-248, 230-237
    _, feat2 = clip_ref.encode_image(v2, params, return_feature=True)
    s1, s2 = predictor_sim(feat1, params, stats), predictor_sim(feat2, params, stats)
    fi = fi / fi.norm(dim=-1, keepdim=True)                                                 # :251-252
    tf = tf / (tf.norm(dim=-1, keepdim=True) + 1e-10)
    s = params["logit_scale"].exp()                                                         # :258 (no clamp)
    return {"logits": (s * fi @ tf.t(), s * tf @ fi.t()), "sim_features": (s1, s1, s2, s2),  # world 1: gathered == local
            "features": (tf, fi)}


def slip_step(sd, images9, ids):
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k != "visual.conv1.weight" and
                                                   "running_" not in k) for k, v in sd.items()}
    stats = {k: v.detach().clone() for k, v in sd.items() if "running_" in k}
    out = slip_forward(params, stats, images9, ids)
    clip = clip_ref.clip_info_ce(*out["logits"])[0]
    s1, g1, s2, g2 = out["sim_features"]
    simclr = loss_ref.nt_xent_gather(s1, g1, s2, g2, rank=0, temperature=0.1)
    loss = clip + simclr
    loss.backward()
    return {"loss": loss.detach(), "parts": {"clip": clip.detach(), "simclr": simclr.detach()}, "out": out,
            "grads": {k: p.grad for k, p in params.items() if p.grad is not None}, "stats": stats}
