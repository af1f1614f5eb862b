"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference modules from /root/reference (build container
only; the path does not exist on the GPU box) so golden vectors can be generated and the restatement in
`oracle/clip_ref.py` can be pinned.  Recipe: SURVEY.md §8c / Appendix D.

Nothing here is copied from the reference: the modules are imported where they lie.
"""
import gzip
import os
import sys
import tempfile
import types

import torch

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/build_ref.py


def _resolve_root():
    env = os.environ.get("DECLIP_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/prototype"):
        return "/root/reference"
    return _STAGED          # GPU box: the unmodified modules staged by oracle/build_ref.py


REF_ROOT = _resolve_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "prototype"))


_ready = False
_bpe_path = None


def _fake_bpe():
    """The real BPE merges file is a Google-Drive download, not in the repo (docs/dataset_prepare.md:33-37).
    1 header + 48894 dummy merges gives len(tokenizer.encoder) == 49409 (simple_tokenizer.py:66-75)."""
    global _bpe_path
    if _bpe_path is None:
        path = os.path.join(tempfile.gettempdir(), "declip_b200_fake_bpe.txt.gz")
        if not os.path.exists(path):
            with gzip.open(path, "wt") as f:
                f.write("#version: fake\n")
                for i in range(49152 - 256 - 2):
                    f.write("a%d b%d\n" % (i, i))
        _bpe_path = path
    return _bpe_path


def setup(force_cpu=False):
    global _ready
    if force_cpu and torch.cuda.is_available():
        # CPU arm on a GPU box: the reference's hard-coded .cuda() calls must stay on the host
        torch.Tensor.cuda = lambda self, *a, **k: self
    if _ready:
        return
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name in ("ipdb", "timm", "ftfy", "textaugment"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["ftfy"].fix_text = lambda s: s

    class _EDA:   # identity text augmentation keeps DeCLIP parity deterministic (declip.py:203-212)
        def synonym_replacement(self, s): return s
        def random_swap(self, s): return s
        def random_deletion(self, s): return s
    sys.modules["textaugment"].EDA = _EDA
    if not torch.cuda.is_available():
        # hard-coded .cuda() calls: text_transformer.py:188, loss.py:43,45
        torch.Tensor.cuda = lambda self, *a, **k: self
    _ready = True


def build_clip_vitb32(embed_dim=512, image_kwargs=None, text_kwargs=None, use_allgather=False):
    """prototype.model.model_entry({'type': 'clip_vitb32', ...}) — model/__init__.py:15-21, clip.py:158-165."""
    setup()
    from prototype.model import model_entry
    ie = dict(embed_dim=embed_dim)
    ie.update(image_kwargs or {})
    te = dict(bpe_path=_fake_bpe(), text_encode_type="Transformer", text_model_utils=dict(random=False, freeze=False),
              embed_dim=embed_dim)
    te.update(text_kwargs or {})
    cfg = dict(type="clip_vitb32", kwargs=dict(image_encode=ie, text_encode=te, clip=dict(use_allgather=use_allgather)))
    return model_entry(cfg)


def set_token_ids(model, ids, labels=None):
    """Bypass the host BPE tokeniser with fixed ids (text_transformer.py:144-180)."""
    def _tok(texts, context_length=77, return_length=False, mask_type=None):
        if mask_type is not None:
            return ids, labels
        return ids
    model.encode_text.tokenize = _tok


def clip_loss_fn():
    setup()
    from prototype.loss_functions import ClipInfoCELoss
    return ClipInfoCELoss()


def reference_clip_step(sd, images, ids, embed_dim=512, v_layers=12, t_layers=12, backward=True):
    """Reference forward + ClipInfoCELoss + backward on the given state_dict / inputs (CPU fp32)."""
    model = build_clip_vitb32(embed_dim, dict(layers=v_layers), dict(transformer_layers=t_layers)).train()
    missing, unexpected = model.load_state_dict(sd, strict=True), None
    set_token_ids(model, ids)
    B = images.shape[0]
    li, lt = model({"images": images, "captions": [["x"]] * B})
    loss, labels = clip_loss_fn()(li, lt)
    out = {"loss": loss.detach(), "logits_per_image": li.detach(), "logits_per_text": lt.detach(), "labels": labels}
    with torch.no_grad():
        out["image_features"] = model.encode_image(images)
        out["text_features"] = model.encode_text(["x"] * B)
    if backward:
        loss.backward()
        out["grads"] = {k: p.grad.detach() for k, p in model.named_parameters() if p.grad is not None}
    return out, model


def _ensure_pg():
    """The reference AllGather calls torch.distributed even at world size 1 (clip.py:36): a 1-rank gloo group."""
    import torch.distributed as dist
    if not dist.is_initialized():
        import tempfile as _tf
        store = dist.FileStore(os.path.join(_tf.mkdtemp(), "pg"), 1)
        dist.init_process_group("gloo", store=store, rank=0, world_size=1)


def reference_declip_step(sd, images6, mlm_ids, mlm_labels, ids_aug, bank_dim_by_size, embed_dim=512, v_layers=12,
                          t_layers=12, weights=None):
    """Reference declip_vitb32 (MLM + NN bank) forward, the solver's loss composition (declip_solver.py:435-517,
    restated by oracle.declip_ref.declip_loss on the reference's OWN outputs) and backward.  CPU fp32."""
    setup()
    _ensure_pg()
    from prototype.model import model_entry
    from . import declip_ref
    nn_size = bank_dim_by_size.shape[1]
    cfg = dict(type="declip_vitb32", kwargs=dict(
        image_encode=dict(embed_dim=embed_dim, layers=v_layers),
        text_encode=dict(bpe_path=_fake_bpe(), text_encode_type="Transformer", text_model_utils=dict(random=False, freeze=False),
                         embed_dim=embed_dim, transformer_layers=t_layers),
        clip=dict(use_allgather=True, text_mask_type="MLM", return_nn_bank=True, feature_dim=embed_dim, nn_size=nn_size)))
    model = model_entry(cfg).train()
    model.load_state_dict(sd, strict=True)
    model.nn_replacer_text.bank = bank_dim_by_size.clone()
    model.nn_replacer_text.bank_ptr = torch.LongTensor([0])

    def _tok(texts, context_length=77, return_length=False, mask_type=None):
        return (mlm_ids.clone(), mlm_labels.clone()) if mask_type is not None else ids_aug
    model.encode_text.tokenize = _tok
    B = images6.shape[0]
    out = model({"images": images6, "captions": [["x"]] * B}, return_dict=True)
    loss, parts = declip_ref.declip_loss(out, weights or declip_ref.LOSS_WEIGHTS)
    loss.backward()
    res = {"loss": loss.detach(), "parts": {k: v.detach() for k, v in parts.items()},
           "out": out, "grads": {k: p.grad.detach() for k, p in model.named_parameters() if p.grad is not None},
           "stats": {k: v.detach().clone() for k, v in model.state_dict().items() if "running_" in k},
           "bank": model.nn_replacer_text.bank.clone(), "bank_ptr": int(model.nn_replacer_text.bank_ptr)}
    return res, model


def reference_defilip_step(sd, images6, mlm_ids, mlm_labels, ids_aug, bank_dim_by_size, embed_dim=512, v_layers=12,
                           t_layers=12):
    """Reference defilip_vitb32 (DeCLIP heads + return_filip + dense_aug) forward, the solver's loss composition
    (defilip_solver.py:435-545 restated by oracle.declip_ref.defilip_loss on the reference's OWN outputs), backward."""
    setup()
    _ensure_pg()
    from prototype.model import model_entry
    from . import declip_ref
    nn_size = bank_dim_by_size.shape[1]
    cfg = dict(type="defilip_vitb32", kwargs=dict(
        image_encode=dict(embed_dim=embed_dim, layers=v_layers),
        text_encode=dict(bpe_path=_fake_bpe(), text_encode_type="Transformer", text_model_utils=dict(random=False, freeze=False),
                         embed_dim=embed_dim, transformer_layers=t_layers),
        clip=dict(use_allgather=True, text_mask_type="MLM", return_nn_bank=True, feature_dim=embed_dim, nn_size=nn_size,
                  return_filip=True, dense_aug=True)))
    model = model_entry(cfg).train()
    model.load_state_dict(sd, strict=True)
    model.nn_replacer_text.bank = bank_dim_by_size.clone()
    model.nn_replacer_text.bank_ptr = torch.LongTensor([0])

    def _tok(texts, context_length=77, return_length=False, mask_type=None):
        return (mlm_ids.clone(), mlm_labels.clone()) if mask_type is not None else ids_aug
    model.encode_text.tokenize = _tok
    B = images6.shape[0]
    out = model({"images": images6, "captions": [["x"]] * B}, return_dict=True)
    loss, parts = declip_ref.defilip_loss(out)
    loss.backward()
    return {"loss": loss.detach(), "parts": {k: v.detach() for k, v in parts.items()}, "out": out,
            "grads": {k: p.grad.detach() for k, p in model.named_parameters() if p.grad is not None},
            "stats": {k: v.detach().clone() for k, v in model.state_dict().items() if "running_" in k},
            "bank": model.nn_replacer_text.bank.clone(), "bank_ptr": int(model.nn_replacer_text.bank_ptr)}, model


def reference_slip_step(sd, images9, ids, embed_dim=512, v_layers=12, t_layers=12, sim_dim=256):
    """Reference slip_vitb32 (return_sim) forward + the solver's loss (slip_solver.py:470-510: ClipInfoCELoss on the base
    view + NT_Xent_gather on the two augmented views' sim features, weights 1 / 1) + backward.  CPU fp32, world 1."""
    setup()
    _ensure_pg()
    os.environ.setdefault("SLURM_PROCID", "0")
    os.environ.setdefault("SLURM_NTASKS", "1")
    from prototype.loss_functions import NT_Xent_gather
    from prototype.model import model_entry
    cfg = dict(type="slip_vitb32", kwargs=dict(
        image_encode=dict(embed_dim=embed_dim, layers=v_layers),
        text_encode=dict(bpe_path=_fake_bpe(), text_encode_type="Transformer", text_model_utils=dict(random=False, freeze=False),
                         embed_dim=embed_dim, transformer_layers=t_layers),
        clip=dict(use_allgather=True, return_sim=True, feature_dim=768, sim_dim=sim_dim)))
    model = model_entry(cfg).train()
    model.load_state_dict(sd, strict=True)

    def _tok(texts, context_length=77, return_length=False, mask_type=None):
        return ids
    model.text_encoder.tokenize = _tok
    B = images9.shape[0]
    out = model({"images": images9, "captions": [["x"]] * B}, return_dict=True)
    clip_loss, _ = clip_loss_fn()(*out["logits"])
    s1, g1, s2, g2 = out["sim_features"]
    simclr = NT_Xent_gather(B)(s1, g1, s2, g2)
    loss = clip_loss + simclr
    loss.backward()
    return {"loss": loss.detach(), "parts": {"clip": clip_loss.detach(), "simclr": simclr.detach()}, "out": out,
            "grads": {k: p.grad.detach() for k, p in model.named_parameters() if p.grad is not None},
            "stats": {k: v.detach().clone() for k, v in model.state_dict().items() if "running_" in k}}, model


def reference_filip_step(sd, images6, mlm_ids, mlm_labels, embed_dim=768, v_layers=12, t_layers=12, weights=None):
    """Reference filip_vitb32 (return_dense, select_topk, MLM tokenisation) forward + the solver's loss + backward."""
    setup()
    _ensure_pg()
    from prototype.model import model_entry
    from . import filip_ref
    cfg = dict(type="filip_vitb32", kwargs=dict(
        image_encode=dict(embed_dim=embed_dim, layers=v_layers),
        text_encode=dict(bpe_path=_fake_bpe(), text_encode_type="Transformer", text_model_utils=dict(random=False, freeze=False),
                         embed_dim=embed_dim, transformer_layers=t_layers),
        clip=dict(use_allgather=True, text_mask_type="MLM", return_dense=True, select_topk=True, feature_dim=embed_dim,
                  mask_rate=0.5, patch_number=14)))
    model = model_entry(cfg).train()
    model.load_state_dict(sd, strict=True)

    def _tok(texts, context_length=77, return_length=False, mask_type=None):
        return (mlm_ids.clone(), mlm_labels.clone()) if mask_type is not None else mlm_ids
    model.encode_text.tokenize = _tok
    B = images6.shape[0]
    out = model({"images": images6, "captions": [["x"]] * B}, return_dict=True)
    loss, parts = filip_ref.filip_loss(out, weights or filip_ref.LOSS_WEIGHTS)
    loss.backward()
    return {"loss": loss.detach(), "parts": {k: v.detach() for k, v in parts.items()}, "out": out,
            "grads": {k: p.grad.detach() for k, p in model.named_parameters() if p.grad is not None}}, model


def reference_clip_res_step(sd, images, ids, embed_dim=1024, layers=(3, 4, 6, 3), t_layers=12):
    """Reference clip_res50 (use_sync_bn False) forward + ClipInfoCELoss + backward, CPU fp32."""
    setup()
    from prototype.model import model_entry
    cfg = dict(type="clip_res50", kwargs=dict(
        image_encode=dict(embed_dim=embed_dim, use_sync_bn=False, bn_group_size=1, layers=tuple(layers)),
        text_encode=dict(bpe_path=_fake_bpe(), text_encode_type="Transformer", text_model_utils=dict(random=False, freeze=False),
                         embed_dim=embed_dim, transformer_layers=t_layers),
        clip=dict(use_allgather=False)))
    model = model_entry(cfg).train()
    model.load_state_dict(sd, strict=True)
    set_token_ids(model, ids)
    B = images.shape[0]
    li, lt = model({"images": images, "captions": [["x"]] * B})
    loss, labels = clip_loss_fn()(li, lt)
    loss.backward()
    with torch.no_grad():
        model.eval()
    return {"loss": loss.detach(), "logits_per_image": li.detach(),
            "grads": {k: p.grad.detach() for k, p in model.named_parameters() if p.grad is not None},
            "stats": {k: v.detach().clone() for k, v in model.state_dict().items() if "running_" in k}}, model


class Stepper:
    """One persistent reference model + synthetic batch for timing (bench.py --impl reference / cpu_baseline): the
    UNMODIFIED reference modules, built once; `step()` = zero_grad + forward + the solver's loss + backward, CPU fp32.
    config in {'clip', 'declip', 'filip', 'res50'} (BASELINE configs[1..4] at a bounded sample batch)."""

    def __init__(self, config, batch, seed=0):
        from . import golden, synth
        setup(force_cpu=True)
        from prototype.model import model_entry
        self.config, self.batch = config, batch
        tcfg = dict(bpe_path=_fake_bpe(), text_encode_type="Transformer", text_model_utils=dict(random=False, freeze=False))
        B = batch
        if config == "clip":
            sd = synth.clip_vit_state_dict(seed=seed)
            cfg = dict(type="clip_vitb32", kwargs=dict(image_encode=dict(embed_dim=512), text_encode=dict(embed_dim=512, **tcfg),
                                                       clip=dict(use_allgather=False)))
            images, ids = synth.synth_images(B, seed=seed), synth.synth_token_ids(B, seed=seed)
            tok = lambda texts, context_length=77, return_length=False, mask_type=None: ids
        elif config == "res50":
            sd = synth.clip_res_state_dict(seed=seed)
            cfg = dict(type="clip_res50", kwargs=dict(image_encode=dict(embed_dim=1024, use_sync_bn=False, bn_group_size=1),
                                                      text_encode=dict(embed_dim=1024, **tcfg), clip=dict(use_allgather=False)))
            images, ids = synth.synth_images(B, seed=seed), synth.synth_token_ids(B, seed=seed)
            tok = lambda texts, context_length=77, return_length=False, mask_type=None: ids
        elif config == "declip":
            _ensure_pg()
            c = dict(batch=B, v_layers=12, t_layers=12, embed_dim=512, seed=seed, nn_size=65536)
            sd, images, mlm_ids, mlm_labels, ids_aug, bank = golden.declip_inputs(c)
            cfg = dict(type="declip_vitb32", kwargs=dict(
                image_encode=dict(embed_dim=512), text_encode=dict(embed_dim=512, **tcfg),
                clip=dict(use_allgather=True, text_mask_type="MLM", return_nn_bank=True, feature_dim=512, nn_size=65536)))
            tok = lambda texts, context_length=77, return_length=False, mask_type=None: (
                (mlm_ids.clone(), mlm_labels.clone()) if mask_type is not None else ids_aug)
            self._bank = bank
        elif config == "filip":
            _ensure_pg()
            c = dict(batch=B, v_layers=12, t_layers=12, embed_dim=768, seed=seed)
            sd, images, mlm_ids, mlm_labels = golden.filip_inputs(c)
            cfg = dict(type="filip_vitb32", kwargs=dict(
                image_encode=dict(embed_dim=768), text_encode=dict(embed_dim=768, **tcfg),
                clip=dict(use_allgather=True, text_mask_type="MLM", return_dense=True, select_topk=True, feature_dim=768,
                          mask_rate=0.5, patch_number=14)))
            tok = lambda texts, context_length=77, return_length=False, mask_type=None: (
                (mlm_ids.clone(), mlm_labels.clone()) if mask_type is not None else mlm_ids)
        else:
            raise ValueError(config)
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):          # the reference prints banners while building
            self.model = model_entry(cfg).train()
        self.model.load_state_dict(sd, strict=True)
        self.model.encode_text.tokenize = tok
        if config == "declip":
            self.model.nn_replacer_text.bank = self._bank.clone()
            self.model.nn_replacer_text.bank_ptr = torch.LongTensor([0])
        self.images = images
        self.crit = clip_loss_fn()

    def step(self):
        from . import declip_ref, filip_ref
        self.model.zero_grad(set_to_none=True)
        batch = {"images": self.images, "captions": [["x"]] * self.batch}
        if self.config in ("clip", "res50"):
            loss, _ = self.crit(*self.model(batch))
        elif self.config == "declip":
            loss, _ = declip_ref.declip_loss(self.model(batch, return_dict=True), declip_ref.LOSS_WEIGHTS)
        else:
            loss, _ = filip_ref.filip_loss(self.model(batch, return_dict=True), filip_ref.LOSS_WEIGHTS)
        loss.backward()
        return loss.item()
