"""Drop-in for prototype.loss_functions (loss.py): ClipInfoCELoss with the reference's call signature
`criterion(logits_per_image, logits_per_text) -> (loss, labels)`, computed by one fused CUDA kernel per
strip (log-softmax + NLL + top-1/top-5 in the forward, softmax - onehot in the backward)."""
import torch
from torch.nn.modules.loss import _Loss

from . import functions as F_


class ClipInfoCELoss(_Loss):
    def __init__(self):
        super().__init__()
        self.stats = {}

    def forward(self, logits_per_image, logits_per_text):
        bs, l_bs = logits_per_image.shape
        rank, _ = F_.dist_info()
        label0 = 0 if l_bs == bs else rank * bs                                    # loss.py:42-45
        labels = label0 + torch.arange(0, bs, dtype=torch.long, device=logits_per_image.device)
        loss = F_.ClipInfoCE.apply(logits_per_image, logits_per_text, label0, self.stats)
        return loss, labels

    def accuracy(self):
        """prec@1 / prec@5 (percent) of logits_per_image from the last forward — misc.py:415-428 — as device
        tensors (no host sync)."""
        rows = float(self.stats["rows"])
        return self.stats["top1_count"].float() * (100.0 / rows), self.stats["top5_count"].float() * (100.0 / rows)
