"""TEST INFRASTRUCTURE — CPU/fp32 restatement (oracle) of DECLIP.forward (prototype/model/declip.py:196-336), the
nearest-neighbour bank (utils/nnclr_modules/{memory_bank,nn_memory_bank}.py), SimsiamLoss / NTXentLoss
(loss_functions/loss.py:52-84, nt_xent_ConVIRT.py) and the solver's loss composition
(prototype/solver/declip_solver.py:435-517), world size 1.  Pinned by tests/test_oracle.py against golden vectors
generated from the reference's own DECLIP module (tools/make_golden.py)."""
import torch
import torch.nn.functional as F

from . import clip_ref

LOSS_WEIGHTS = dict(clip_loss=0.4, simsiam_loss=0.2, masking_language=0.2, nn_text=0.2)   # yfcc15m_vit_declip/config.yaml:28-32


def _bn(x, sd, p, stats):
    """nn.BatchNorm1d in training mode: batch statistics; running stats updated in `stats` (functional copy)."""
    rm, rv = stats[p + ".running_mean"], stats[p + ".running_var"]
    return F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], True, 0.1, 1e-5)


def projector(x, sd, stats, p="projector."):
    # declip.py:65-90
    x = F.relu(_bn(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"]), sd, p + "bn1", stats))
    x = F.relu(_bn(F.linear(x, sd[p + "linear2.weight"], sd[p + "linear2.bias"]), sd, p + "bn2", stats))
    return _bn(F.linear(x, sd[p + "linear3.weight"], sd[p + "linear3.bias"]), sd, p + "bn3", stats)


def predictor(x, sd, stats, p="predictor."):
    # declip.py:118-130
    x = F.relu(_bn(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"]), sd, p + "bn1", stats))
    return F.linear(x, sd[p + "layer2.weight"], sd[p + "layer2.bias"])


class Bank:
    """memory_bank.py:9-124 + nn_memory_bank.py:42-65, reference layout [dim, size]."""

    def __init__(self, bank, ptr=0):
        self.bank, self.ptr, self.size = bank.clone(), int(ptr), bank.shape[1]

    def __call__(self, output, update=False):
        bank = self.bank.clone()
        if update:
            bs = output.shape[0]
            if self.ptr + bs >= self.size:
                self.bank[:, self.ptr:] = output[:self.size - self.ptr].T.detach()
                self.ptr = 0
            else:
                self.bank[:, self.ptr:self.ptr + bs] = output.T.detach()
                self.ptr += bs
        bank = bank.t()
        sim = F.normalize(output, dim=1) @ F.normalize(bank, dim=1).t()
        idx = sim.topk(1, dim=1).indices[:, 0]
        return bank.index_select(0, idx), idx


def simsiam_loss(p1, z1, p2, z2):
    # loss.py:52-84
    def D(p, z):
        z = z.detach()
        return ((p / p.norm(dim=-1, keepdim=True)) * (z / z.norm(dim=-1, keepdim=True))).sum(dim=1).mean()
    return -0.5 * (D(p1, z2) + D(p2, z1))


def nt_xent(zis, zjs, temperature=0.1, alpha=0.75):
    # nt_xent_ConVIRT.py:28-86
    zis, zjs = F.normalize(zis, p=2, dim=1), F.normalize(zjs, p=2, dim=1)
    n = zis.shape[0]
    labels = torch.eye(n)
    ab, ba = zis @ zjs.t() / temperature, zjs @ zis.t() / temperature
    soft = lambda t, l: -(t * F.log_softmax(l, dim=1)).sum() / l.shape[0]
    return alpha * soft(labels, ab) + (1 - alpha) * soft(labels, ba)


def declip_forward(params, stats, images6, mlm_ids, mlm_labels, ids_aug, bank):
    """DECLIP.forward(return_dict=True) with text_mask_type='MLM', return_nn_bank=True, world size 1."""
    im1, im2 = torch.split(images6, [3, 3], dim=1)                                         # declip.py:199
    tf, words = clip_ref.encode_text(mlm_ids, params, return_dense=True)                   # :215 (masked caption)
    tfa = clip_ref.encode_text(ids_aug, params)                                            # :216
    f1 = clip_ref.encode_image(im1, params)                                                # :231-232
    f2 = clip_ref.encode_image(im2, params)
    z1, z2 = projector(f1, params, stats), projector(f2, params, stats)                    # :238-241
    p1, p2 = predictor(z1, params, stats), predictor(z2, params, stats)
    f1 = f1 / f1.norm(dim=-1, keepdim=True)                                                # :245-248
    f2 = f2 / f2.norm(dim=-1, keepdim=True)
    tf = tf / (tf.norm(dim=-1, keepdim=True) + 1e-10)
    tfa = tfa / (tfa.norm(dim=-1, keepdim=True) + 1e-10)
    s = params["logit_scale"].exp()
    s.data = torch.clamp(s.data, max=100)                                                  # :251-252
    out = {"logits": (s * f1 @ tf.t(), s * f2 @ tf.t(), s * tf @ f1.t(), s * tf @ f2.t()),           # :271-279
           "logits_aug": (s * f1 @ tfa.t(), s * f2 @ tfa.t(), s * tfa @ f1.t(), s * tfa @ f2.t()),
           "simsiam_features": (p1, p2, z1, z2), "features": (tf, f1, f2)}
    nn_t, idx_t = bank(tf.detach().float(), update=False)                                  # :282-288
    nn_t = nn_t / (nn_t.norm(dim=-1, keepdim=True) + 1e-10)
    nn_ta, idx_ta = bank(tfa.detach().float(), update=True)
    nn_ta = nn_ta / (nn_ta.norm(dim=-1, keepdim=True) + 1e-10)
    bank(tf.detach().float(), update=True)
    out["nn_text_logits"] = (s * f1 @ nn_t.t(), s * f2 @ nn_t.t(), s * f1 @ nn_ta.t(), s * f2 @ nn_ta.t())   # :293-300
    out["nn_index"] = (idx_t, idx_ta)
    pred = F.linear(words, params["text_label_predictor.weight"], params["text_label_predictor.bias"])    # :329
    m = mlm_labels != -100
    out["text_self_supervised"] = F.cross_entropy(pred[m], mlm_labels[m])                  # :331-333
    return out


def declip_loss(out, weights=LOSS_WEIGHTS, world=1):
    """declip_solver.py:435-517 (image_text_two_view, default weighting type)."""
    ce = lambda a, b: clip_ref.clip_info_ce(a, b)[0]
    li1, li2, lt1, lt2 = out["logits"]
    li1a, li2a, lt1a, lt2a = out["logits_aug"]
    clip_loss = (ce(li1, lt1) + ce(li2, lt2) + ce(li1a, lt1a) + ce(li2a, lt2a)) / 4 / world
    mlm = out["text_self_supervised"] / world
    n1, n2, n1a, n2a = out["nn_text_logits"]
    nn_loss = (ce(n1, n1a) + ce(n2, n2a)) / 2 / world                                      # :474-479
    p1, p2, z1, z2 = out["simsiam_features"]
    ss = simsiam_loss(p1, z1, p2, z2) / world
    tf, f1, f2 = out["features"]
    ntx = (nt_xent(f1, tf) + nt_xent(f2, tf)) / world                                      # :486-488 (logged only)
    loss = clip_loss * weights["clip_loss"] + ss * weights["simsiam_loss"] + mlm * weights["masking_language"] + \
        nn_loss * weights["nn_text"]
    return loss, dict(clip=clip_loss, mlm=mlm, nn=nn_loss, simsiam=ss, nt_xent=ntx)


def declip_step(sd, images6, mlm_ids, mlm_labels, ids_aug, bank_dim_by_size):
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k != "visual.conv1.weight" and
                                                   "running_" not in k) for k, v in sd.items()}
    stats = {k: v.detach().clone() for k, v in sd.items() if "running_" in k}
    bank = Bank(bank_dim_by_size)
    out = declip_forward(params, stats, images6, mlm_ids, mlm_labels, ids_aug, bank)
    loss, parts = declip_loss(out)
    loss.backward()
    return {"loss": loss.detach(), "parts": {k: v.detach() for k, v in parts.items()}, "out": out,
            "grads": {k: p.grad for k, p in params.items() if p.grad is not None}, "stats": stats,
            "bank": bank.bank, "bank_ptr": bank.ptr}


# yfcc15m_vit_defilip/config.yaml: the DeCLIP weights plus `filip: 0.2` (defilip_solver.py:462-480,541-542)
DEFILIP_FILIP_WEIGHT = 0.2


def defilip_loss(out, weights=LOSS_WEIGHTS, filip_weight=DEFILIP_FILIP_WEIGHT, world=1):
    """declip_loss + the FILIP term: mean of the (up to four) dense InfoNCE losses."""
    ce = lambda a, b: clip_ref.clip_info_ce(a, b)[0]
    loss, parts = declip_loss(out, weights, world)
    f = ce(*out["filip"])
    if "filip_aug" in out:
        a = out["filip_aug"]
        f = (f + ce(a[0], a[1]) + ce(a[2], a[3]) + ce(a[4], a[5])) / 4
    f = f / world
    parts = dict(parts, filip=f)
    return loss + f * filip_weight, parts


def defilip_step(sd, images6, mlm_ids, mlm_labels, ids_aug, bank_dim_by_size):
    """DEFILIP.forward (defilip.py:269-431: the DeCLIP forward with dense image / word tokens, `return_filip`, `dense_aug`)
    + the solver's loss (defilip_loss) + backward, world size 1."""
    from . import filip_ref
    params = {k: v.detach().clone().requires_grad_(v.is_floating_point() and k != "visual.conv1.weight" and
                                                   "running_" not in k) for k, v in sd.items()}
    stats = {k: v.detach().clone() for k, v in sd.items() if "running_" in k}
    bank = Bank(bank_dim_by_size)
    out = declip_forward(params, stats, images6, mlm_ids, mlm_labels, ids_aug, bank)
    # token-wise late interaction on the four (view, caption) combinations               defilip.py:311-313,331-342
    im1, im2 = torch.split(images6, [3, 3], dim=1)
    _, d1 = clip_ref.encode_image(im1, params, return_dense=True)
    _, d2 = clip_ref.encode_image(im2, params, return_dense=True)
    _, w1 = clip_ref.encode_text(mlm_ids, params, return_dense=True)
    _, w2 = clip_ref.encode_text(ids_aug, params, return_dense=True)
    img = lambda t: F.linear(t, params["image_mapping.weight"], params["image_mapping.bias"])
    txt = lambda t: F.linear(t, params["text_mapping.weight"], params["text_mapping.bias"])
    dense = lambda a, b: filip_ref.weighted_dense_logits(a, b, params["logit_scale_dense"])[:2]
    out["filip"] = dense(img(d1), txt(w1))
    out["filip_aug"] = (*dense(img(d2), txt(w1)), *dense(img(d1), txt(w2)), *dense(img(d2), txt(w2)))
    loss, parts = defilip_loss(out)
    loss.backward()
    return {"loss": loss.detach(), "parts": {k: v.detach() for k, v in parts.items()}, "out": out,
            "grads": {k: p.grad for k, p in params.items() if p.grad is not None}, "bank_ptr": bank.ptr}
