"""Generates tests/golden/*.pt by executing the UNMODIFIED reference modules (imported from
/root/reference — build container only) on the synthetic weights/inputs of oracle/synth.py.

    python tools/make_golden.py            # all cases
Committed beside the vectors so the provenance of every golden number is reproducible.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import golden, ref_harness, synth  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    os.makedirs(golden.GOLDEN_DIR, exist_ok=True)
    names = sys.argv[1:] or (list(golden.CASES) + list(golden.DECLIP_CASES) + list(golden.FILIP_CASES) +
                             list(golden.RES_CASES) + list(golden.DEFILIP_CASES) + list(golden.SLIP_CASES))
    if "nt_xent" in names or not sys.argv[1:]:
        from oracle import loss_ref
        ref_harness.setup()
        os.environ["SLURM_PROCID"], os.environ["SLURM_NTASKS"] = "1", "3"     # link.get_rank() reads SLURM env
        from prototype.loss_functions import NT_Xent, NT_Xent_gather
        z_i, z_j, z_ib, z_jb, rank = loss_ref.inputs()
        a, b = z_i.clone().requires_grad_(True), z_j.clone().requires_grad_(True)
        l1 = NT_Xent(z_i.shape[0], 0.5)(a, b)
        l1.backward()
        c, d = z_i.clone().requires_grad_(True), z_j.clone().requires_grad_(True)
        l2 = NT_Xent_gather(z_i.shape[0], 0.1)(c, z_ib, d, z_jb)
        l2.backward()
        del os.environ["SLURM_PROCID"], os.environ["SLURM_NTASKS"]
        torch.save({"generator": "tools/make_golden.py nt_xent (reference NT_Xent / NT_Xent_gather, CPU fp32)",
                    "nt_xent": l1.item(), "nt_xent_grad": a.grad.clone(), "nt_xent_gather": l2.item(),
                    "nt_xent_gather_grad_i": c.grad.clone(), "nt_xent_gather_grad_j": d.grad.clone(), "rank": rank},
                   golden.path("nt_xent"))
        print("nt_xent: %.6f  nt_xent_gather: %.6f" % (l1.item(), l2.item()))
    names = [n for n in names if n != "nt_xent"]
    for name in [n for n in names if n in golden.RES_CASES]:
        c = golden.RES_CASES[name]
        t0 = time.time()
        sd, images, ids = golden.res_inputs(c)
        res, model = ref_harness.reference_clip_res_step(sd, images, ids, c["embed_dim"], c["layers"], c["t_layers"])
        blob = {"case": c, "torch": torch.__version__,
                "generator": "tools/make_golden.py via oracle/ref_harness.reference_clip_res_step (reference clip_res50, CPU fp32)",
                "loss": res["loss"].item(), "logits_per_image": res["logits_per_image"].clone(),
                "grads": golden.summarise_grads(res["grads"]), "stats": res["stats"]}
        torch.save(blob, golden.path(name))
        print("%s: loss %.6f, %d grads, %.1fs -> %.1f KB" % (name, blob["loss"], len(blob["grads"]), time.time() - t0,
              os.path.getsize(golden.path(name)) / 1024))
    for name in [n for n in names if n in golden.SLIP_CASES]:
        c = golden.SLIP_CASES[name]
        t0 = time.time()
        sd, images, ids = golden.slip_inputs(c)
        res, model = ref_harness.reference_slip_step(sd, images, ids, c["embed_dim"], c["v_layers"], c["t_layers"])
        o = res["out"]
        blob = {"case": c, "torch": torch.__version__,
                "generator": "tools/make_golden.py via oracle/ref_harness.reference_slip_step (reference SLIP, CPU fp32)",
                "loss": res["loss"].item(), "parts": {k: v.item() for k, v in res["parts"].items()},
                "logits": [t.detach().clone() for t in o["logits"]],
                "sim_features": [t.detach().clone() for t in o["sim_features"]],
                "features": [t.detach().clone() for t in o["features"]],
                "grads": golden.summarise_grads(res["grads"]), "stats": res["stats"]}
        torch.save(blob, golden.path(name))
        print("%s: loss %.6f parts %s, %d grads, %.1fs -> %.1f KB" % (name, blob["loss"], {k: round(v, 4) for k, v in
              blob["parts"].items()}, len(blob["grads"]), time.time() - t0, os.path.getsize(golden.path(name)) / 1024))
    for name in [n for n in names if n in golden.DEFILIP_CASES]:
        c = golden.DEFILIP_CASES[name]
        t0 = time.time()
        sd, images, mlm_ids, mlm_labels, ids_aug, bank = golden.defilip_inputs(c)
        res, model = ref_harness.reference_defilip_step(sd, images, mlm_ids, mlm_labels, ids_aug, bank, c["embed_dim"],
                                                        c["v_layers"], c["t_layers"])
        o = res["out"]
        blob = {"case": c, "torch": torch.__version__,
                "generator": "tools/make_golden.py via oracle/ref_harness.reference_defilip_step (reference DEFILIP, CPU fp32)",
                "loss": res["loss"].item(), "parts": {k: v.item() for k, v in res["parts"].items()},
                "logits": [t.detach().clone() for t in o["logits"]], "logits_aug": [t.detach().clone() for t in o["logits_aug"]],
                "filip": [t.detach().clone() for t in o["filip"]], "filip_aug": [t.detach().clone() for t in o["filip_aug"]],
                "nn_text_logits": [t.detach().clone() for t in o["nn_text_logits"]],
                "grads": golden.summarise_grads(res["grads"]), "bank_ptr": res["bank_ptr"]}
        torch.save(blob, golden.path(name))
        print("%s: loss %.6f parts %s, %d grads, %.1fs -> %.1f KB" % (name, blob["loss"], {k: round(v, 4) for k, v in
              blob["parts"].items()}, len(blob["grads"]), time.time() - t0, os.path.getsize(golden.path(name)) / 1024))
    for name in [n for n in names if n in golden.FILIP_CASES]:
        c = golden.FILIP_CASES[name]
        t0 = time.time()
        sd, images, mlm_ids, mlm_labels = golden.filip_inputs(c)
        res, model = ref_harness.reference_filip_step(sd, images, mlm_ids, mlm_labels, c["embed_dim"], c["v_layers"],
                                                      c["t_layers"])
        blob = {"case": c, "torch": torch.__version__,
                "generator": "tools/make_golden.py via oracle/ref_harness.reference_filip_step (reference FILIP, CPU fp32)",
                "loss": res["loss"].item(), "parts": {k: v.item() for k, v in res["parts"].items()},
                "logits": [t.detach().clone() for t in res["out"]["logits"]],
                "dense_logits": [t.detach().clone() for t in res["out"]["dense_logits"]],
                "grads": golden.summarise_grads(res["grads"])}
        torch.save(blob, golden.path(name))
        print("%s: loss %.6f parts %s, %d grads, %.1fs -> %.1f KB" % (name, blob["loss"], {k: round(v, 4) for k, v in
              blob["parts"].items()}, len(blob["grads"]), time.time() - t0, os.path.getsize(golden.path(name)) / 1024))
    for name in [n for n in names if n in golden.DECLIP_CASES]:
        c = golden.DECLIP_CASES[name]
        t0 = time.time()
        sd, images, mlm_ids, mlm_labels, ids_aug, bank = golden.declip_inputs(c)
        res, model = ref_harness.reference_declip_step(sd, images, mlm_ids, mlm_labels, ids_aug, bank, c["embed_dim"],
                                                       c["v_layers"], c["t_layers"])
        o = res["out"]
        blob = {
            "case": c, "torch": torch.__version__,
            "generator": "tools/make_golden.py via oracle/ref_harness.reference_declip_step (reference DECLIP, CPU fp32)",
            "loss": res["loss"].item(), "parts": {k: v.item() for k, v in res["parts"].items()},
            "logits": [t.detach().clone() for t in o["logits"]], "logits_aug": [t.detach().clone() for t in o["logits_aug"]],
            "nn_text_logits": [t.detach().clone() for t in o["nn_text_logits"]],
            "simsiam_features": [t.detach().clone() for t in o["simsiam_features"]],
            "features": [t.detach().clone() for t in o["features"]],
            "text_self_supervised": o["text_self_supervised"].item(),
            "grads": golden.summarise_grads(res["grads"]), "stats": res["stats"],
            "bank_ptr": res["bank_ptr"], "bank_checksum": res["bank"].double().sum().item(),
            "bank_tail": res["bank"][:, :2 * c["batch"]].clone(),
        }
        torch.save(blob, golden.path(name))
        print("%s: loss %.6f parts %s, %d grads, %.1fs -> %.1f KB" % (name, blob["loss"], {k: round(v, 4) for k, v in
              blob["parts"].items()}, len(blob["grads"]), time.time() - t0, os.path.getsize(golden.path(name)) / 1024))
    for name in [n for n in names if n in golden.CASES]:
        c = golden.CASES[name]
        t0 = time.time()
        sd = synth.clip_vit_state_dict(seed=c["seed"], embed_dim=c["embed_dim"], v_layers=c["v_layers"],
                                       t_layers=c["t_layers"])
        images = synth.synth_images(c["batch"], seed=c["seed"])
        ids = synth.synth_token_ids(c["batch"], seed=c["seed"])
        out, model = ref_harness.reference_clip_step(sd, images, ids, c["embed_dim"], c["v_layers"], c["t_layers"])
        blob = {
            "case": c, "torch": torch.__version__,
            "generator": "tools/make_golden.py via oracle/ref_harness.py (reference @ /root/reference, CPU fp32)",
            "loss": out["loss"].item(),
            "logits_per_image": out["logits_per_image"].clone(), "logits_per_text": out["logits_per_text"].clone(),
            "image_features": out["image_features"].clone(), "text_features": out["text_features"].clone(),
            "labels": out["labels"].clone(),
            "grads": golden.summarise_grads(out["grads"]),
            "param_names": [k for k, _ in model.named_parameters()],
            "trainable": [k for k, p in model.named_parameters() if p.requires_grad],
        }
        torch.save(blob, golden.path(name))
        print("%s: loss %.6f, %d grads, %.1fs -> %s (%.1f KB)" % (name, blob["loss"], len(blob["grads"]), time.time() - t0,
              golden.path(name), os.path.getsize(golden.path(name)) / 1024))


if __name__ == "__main__":
    main()
