"""TEST INFRASTRUCTURE — CPU/fp32 restatement (oracle) of FILIP.forward / get_weighted_dense_logits
(prototype/model/filip.py:71-142) and the solver's loss (prototype/solver/filip_solver.py:436-520), world size 1.
Pinned by tests/test_oracle.py against golden vectors generated from the reference's own FILIP module."""
import torch
import torch.nn.functional as F

from . import clip_ref

# yfcc15m_vit_filip/config.yaml:33-35 trains with clip_loss 0.0 / clip_dense_loss 1.0; the golden case uses 1.0 / 1.0 so
# that both heads carry gradient.
LOSS_WEIGHTS = dict(clip_loss=1.0, clip_dense_loss=1.0)


def weighted_dense_logits(d1, d2, logit_scale_dense, top_k=16):
    # filip.py:71-106 (select_topk=True, world size 1: all_gather is the identity)
    d1 = d1 / d1.norm(dim=-1, keepdim=True)
    d2 = d2 / d2.norm(dim=-1, keepdim=True)
    s = logit_scale_dense.exp()
    cross = torch.matmul(d1, d2.permute(0, 2, 1))
    _, id1 = torch.topk(cross.sum(dim=2), dim=1, k=top_k)
    _, id2 = torch.topk(cross.sum(dim=1), dim=1, k=top_k)
    bs, n1 = d1.shape[:2]
    n2 = d2.shape[1]
    sel1 = d1.reshape(bs * n1, -1)[id1 + (torch.arange(bs) * n1)[:, None]].reshape(bs, top_k, -1)
    sel2 = d2.reshape(bs * n2, -1)[id2 + (torch.arange(bs) * n2)[:, None]].reshape(bs, top_k, -1)

    def get_logits(a, sel):
        i, j, k = a.shape
        l, m, k = sel.shape
        return (s * a.reshape(-1, k) @ sel.reshape(-1, k).t()).reshape(i, j, l, m).permute(0, 2, 1, 3)
    l1 = get_logits(d1, sel2).max(dim=-1)[0].mean(dim=-1)
    l2 = get_logits(d2, sel1).max(dim=-1)[0].mean(dim=-1)
    return l1, l2, (id1, id2)


def filip_forward(params, images6, mlm_ids):
    im1, _ = torch.split(images6, [3, 3], dim=1)                                         # filip.py:112
    tf, words = clip_ref.encode_text(mlm_ids, params, return_dense=True)                 # :116
    f1, dense = clip_ref.encode_image(im1, params, return_dense=True)                    # :119
    s = params["logit_scale"].exp()                                                      # :121 (no clamp)
    f1 = f1 / f1.norm(dim=-1, keepdim=True)
    tf = tf / (tf.norm(dim=-1, keepdim=True) + 1e-10)
    li, lt = s * f1 @ tf.t(), s * tf @ f1.t()                                            # :126-129
    d1 = F.linear(dense, params["image_mapping.weight"], params["image_mapping.bias"])   # :133
    d2 = F.linear(words, params["text_mapping.weight"], params["text_mapping.bias"])     # :134
    l1, l2, ids = weighted_dense_logits(d1, d2, params["logit_scale_dense"])
    return {"logits": (li, lt), "dense_logits": (l1, l2), "topk": ids}


def filip_loss(out, weights=LOSS_WEIGHTS, world=1):
    ce = lambda a, b: clip_ref.clip_info_ce(a, b)[0]
    clip_loss = ce(*out["logits"]) / world
    dense = ce(*out["dense_logits"]) / world
    return clip_loss * weights["clip_loss"] + dense * weights["clip_dense_loss"], dict(clip=clip_loss, dense=dense)


def filip_step(sd, images6, mlm_ids):
    params = {k: v.detach().clone().requires_grad_(k != "visual.conv1.weight") for k, v in sd.items()}
    out = filip_forward(params, images6, mlm_ids)
    loss, parts = filip_loss(out)
    loss.backward()
    return {"loss": loss.detach(), "parts": {k: v.detach() for k, v in parts.items()}, "out": out,
            "grads": {k: p.grad for k, p in params.items() if p.grad is not None}}
