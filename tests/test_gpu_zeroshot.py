"""Zero-shot evaluation path (clip_solver.py:675-737) against goldens produced by the reference
models in eval() mode (tools/make_golden_zeroshot.py): prompt-ensemble classifier, unit image features, logits, top-1.
Stated tolerance (bf16 towers vs fp32 reference): classifier / feature cosine >= 0.999 per row, |d logit| <= 3e-3,
identical top-1 wherever the reference's top-1 margin exceeds 6e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(model, g, dev):
    from declip_b200 import zero_shot
    from oracle import synth
    images = synth.synth_images(8, seed=g["image_seed"]).to(dev)
    ids = synth.synth_token_ids(g["labels"] * g["prompts"], seed=g["ids_seed"]).to(dev)
    model.train()                                   # the helpers must switch to eval() themselves and restore the mode
    cls = zero_shot.build_classifier(model, ids, g["labels"])
    logits, preds, scores = zero_shot.classify(model, images, cls, ensemble_matrix=torch.eye(g["labels"]))
    torch.cuda.synchronize()
    assert model.training
    cos_c = torch.nn.functional.cosine_similarity(cls.cpu(), g["classifier"], dim=1)
    assert cos_c.min().item() > 0.999, cos_c
    assert torch.allclose(cls.norm(dim=1).cpu(), torch.ones(g["labels"]), atol=1e-4)
    err = (logits.cpu() - g["logits"]).abs().max().item()
    assert err <= 3e-3, err
    top2 = g["logits"].topk(2, dim=1).values
    sure = (top2[:, 0] - top2[:, 1]) > 6e-3
    assert torch.equal(preds.cpu()[sure], g["preds"][sure])
    assert torch.allclose(scores.sum(1).cpu(), torch.ones(8), atol=1e-4)


def test_zero_shot_vit(cuda_dev):
    from declip_b200.model import model_entry
    from oracle import golden, synth
    g = golden.load("zeroshot_vit")
    c = golden.CASES[g["case"]]
    model = model_entry(dict(type='clip_vitb32', kwargs=dict(
        image_encode=dict(embed_dim=c["embed_dim"], layers=c["v_layers"]),
        text_encode=dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                         embed_dim=c["embed_dim"], transformer_layers=c["t_layers"]),
        clip=dict(use_allgather=False))))
    model.load_state_dict(synth.clip_vit_state_dict(seed=c["seed"], embed_dim=c["embed_dim"], v_layers=c["v_layers"],
                                                    t_layers=c["t_layers"]), strict=True)
    _check(model.to(cuda_dev), g, cuda_dev)


def test_zero_shot_res50_eval_batchnorm(cuda_dev):
    from declip_b200.model import model_entry
    from oracle import golden
    g = golden.load("zeroshot_res")
    c = golden.RES_CASES[g["case"]]
    model = model_entry(dict(type='clip_res50', kwargs=dict(
        image_encode=dict(embed_dim=c["embed_dim"], use_sync_bn=False, bn_group_size=1, layers=tuple(c["layers"])),
        text_encode=dict(bpe_path=None, text_encode_type='Transformer', text_model_utils=dict(random=False, freeze=False),
                         embed_dim=c["embed_dim"], transformer_layers=c["t_layers"]),
        clip=dict(use_allgather=False))))
    sd, _, _ = golden.res_inputs(c)
    sd.update(golden.load(g["case"])["stats"])      # BatchNorm running statistics after the golden training step
    model.load_state_dict(sd, strict=True)
    _check(model.to(cuda_dev), g, cuda_dev)
