"""Generates tests/golden/zeroshot_{vit,res}.pt with the UNMODIFIED reference models in eval() mode (build container only):
prompt-ensemble classifier + image logits exactly as prototype/solver/clip_solver.py:675-737 computes them.
    python tools/make_golden_zeroshot.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import golden, ref_harness, synth  # noqa: E402

LABELS, PROMPTS, IMAGES = 6, 3, 8


def run(model, images, ids):
    model.eval()
    with torch.no_grad():
        rows = []
        for i in range(LABELS):
            ref_harness.set_token_ids(model, ids[i * PROMPTS:(i + 1) * PROMPTS])
            f = model.encode_text(["x"] * PROMPTS)
            f = f / f.norm(dim=-1, keepdim=True)
            f = f.mean(dim=0)
            rows.append(f / f.norm())
        cls = torch.stack(rows, 0)
        img = model.encode_image(images)
        img = img / img.norm(dim=-1, keepdim=True)
        logits = img @ cls.t()
    return {"classifier": cls, "image_features": img, "logits": logits, "preds": logits.argmax(1)}


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    ref_harness.setup()
    from prototype.model import model_entry
    # ViT
    c = golden.CASES["clip_vitb32_l2_b8"]
    sd = synth.clip_vit_state_dict(seed=c["seed"], embed_dim=c["embed_dim"], v_layers=c["v_layers"], t_layers=c["t_layers"])
    model = ref_harness.build_clip_vitb32(c["embed_dim"], {"layers": c["v_layers"]}, {"transformer_layers": c["t_layers"]})
    model.load_state_dict(sd, strict=True)
    images = synth.synth_images(IMAGES, seed=11)
    ids = synth.synth_token_ids(LABELS * PROMPTS, seed=12)
    out = run(model, images, ids)
    out.update(generator="tools/make_golden_zeroshot.py (reference clip_vitb32 eval)", case="clip_vitb32_l2_b8",
               image_seed=11, ids_seed=12, labels=LABELS, prompts=PROMPTS)
    torch.save(out, golden.path("zeroshot_vit"))
    print("vit preds", out["preds"].tolist(), "logit range", out["logits"].min().item(), out["logits"].max().item())
    # ResNet: running statistics = the ones the training golden left behind after its step
    c = golden.RES_CASES["clip_res50_l1111_b4"]
    sd, _, _ = golden.res_inputs(c)
    sd.update(golden.load("clip_res50_l1111_b4")["stats"])
    model = model_entry(dict(type="clip_res50", kwargs=dict(
        image_encode=dict(embed_dim=c["embed_dim"], use_sync_bn=False, bn_group_size=1, layers=tuple(c["layers"])),
        text_encode=dict(bpe_path=ref_harness._fake_bpe(), text_encode_type="Transformer",
                         text_model_utils=dict(random=False, freeze=False), embed_dim=c["embed_dim"],
                         transformer_layers=c["t_layers"]),
        clip=dict(use_allgather=False))))
    model.load_state_dict(sd, strict=True)
    out = run(model, images, ids)
    out.update(generator="tools/make_golden_zeroshot.py (reference clip_res50 eval, BN running stats from the training golden)",
               case="clip_res50_l1111_b4", image_seed=11, ids_seed=12, labels=LABELS, prompts=PROMPTS)
    torch.save(out, golden.path("zeroshot_res"))
    print("res preds", out["preds"].tolist(), "logit range", out["logits"].min().item(), out["logits"].max().item())


if __name__ == "__main__":
    main()
