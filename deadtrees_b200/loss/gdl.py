"""``GeneralizedDiceLoss`` with the reference's signature (``deadtrees/loss/gdl.py:10-27``)."""
import torch

from .. import ops


class GeneralizedDiceLoss(torch.nn.Module):
    def forward(self, inp, targ):
        s = ops.prob_loss_partials(inp, targ).sum(dim=0)  # (K, 4): batch-global sums per class
        count = s[:, 2].round().long()
        w = 1.0 / ((count ** 2).float() + 1e-9)           # int64 count**2, then float32 (gdl.py:15)
        numerator = torch.sum(w * s[:, 0].float())
        denominator = torch.sum(w * (s[:, 2] + s[:, 1]).float())
        dice = 2.0 * (numerator + 1e-9) / (denominator + 1e-9)
        return 1.0 - dice
