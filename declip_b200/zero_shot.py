"""Zero-shot classification with a trained CLIP-family model — the evaluation path of the reference solvers
(prototype/solver/clip_solver.py:675-737): prompt-ensemble text classifier, image features, logits, top-1 and the
label-ensemble scores.  Forward-only: `model.eval()` + `torch.no_grad()`; ModifiedResNet towers use their BatchNorm
running statistics (`dc_bn2d_fwd(training=0)`)."""
import torch

from . import functions as F_


def _unit(x):
    return x / x.norm(dim=-1, keepdim=True)


@torch.no_grad()
def build_classifier(model, label_texts, label_num):
    """label_texts: List[str] of length label_num * prompts in CLASS-MAJOR order (all prompts of class 0 first — the
    layout `dataset.get_label_texts()` returns), or the same as pre-tokenised LongTensor ids [label_num * prompts, 77].
    Returns fp32 [label_num, embed_dim]: per class, mean of the unit-norm prompt embeddings, re-normalised
    (clip_solver.py:693-700)."""
    was_training = model.training
    model.eval()
    try:
        n = len(label_texts)
        if n % label_num != 0:
            raise ValueError("len(label_texts) = %d is not a multiple of label_num = %d" % (n, label_num))
        prompts = n // label_num
        rows = []
        for i in range(label_num):
            feats = model.encode_text(label_texts[i * prompts:(i + 1) * prompts]).float()
            rows.append(_unit(_unit(feats).mean(dim=0)))
        return torch.stack(rows, dim=0)
    finally:
        model.train(was_training)


@torch.no_grad()
def classify(model, images, classifier, ensemble_matrix=None, return_dense=False):
    """images fp32 [B,3,R,R] -> (logits [B,label_num] fp32, preds [B] int64, scores or None)   (clip_solver.py:705-720).
    logits = unit(image features) @ classifier^T through the fp32-accurate split-bf16 tcgen05 GEMM."""
    was_training = model.training
    model.eval()
    try:
        feats = model.encode_image(images, return_dense=True)[0] if return_dense else model.encode_image(images)
        feats = _unit(feats.float())
        n = classifier.shape[0]
        pad = (-n) % 8                                      # the GEMM wants the class count in multiples of 8
        w = torch.cat([classifier, classifier.new_zeros(pad, classifier.shape[1])]) if pad else classifier
        logits = F_.LinearF32.apply(feats.contiguous(), w.float().contiguous(), None)[:, :n]
        preds = logits.argmax(dim=1)
        scores = None
        if ensemble_matrix is not None:
            scores = torch.softmax(logits, dim=1) @ ensemble_matrix.to(logits)
        return logits, preds, scores
    finally:
        model.train(was_training)
