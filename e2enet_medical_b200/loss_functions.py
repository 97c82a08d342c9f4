"""Drop-in mirrors of the reference's deep-supervision loss modules (the first
"next" row after the hot path): `DC_and_CE_loss` (e2enet/training/loss_functions/dice_loss.py:302-359,
with `SoftDiceLoss` :155-190 and `RobustCrossEntropyLoss` crossentropy.py:4-11) and
`MultipleOutputLoss2` (deep_supervision.py:18-43).  Same constructor signatures and the same value /
gradient; the ~10 voxel-sized passes per output collapse into one fused statistics kernel forward and
one backward (ops.SoftmaxStats), the dice / CE formulas act on (B, C) tensors."""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class SoftDiceLoss(nn.Module):
    def __init__(self, apply_nonlin=None, batch_dice=False, do_bg=True, smooth=1.):
        super().__init__()
        self.do_bg, self.batch_dice, self.apply_nonlin, self.smooth = do_bg, batch_dice, apply_nonlin, smooth

    def from_stats(self, sp, tp, sy):
        """tp / fp / fn sums -> -mean dice (reference :172-190)"""
        fp, fn = sp - tp, sy - tp
        if self.batch_dice:
            tp, fp, fn = tp.sum(0), fp.sum(0), fn.sum(0)
        dc = (2 * tp + self.smooth) / (2 * tp + fp + fn + self.smooth + 1e-8)
        if not self.do_bg:
            dc = dc[1:] if self.batch_dice else dc[:, 1:]
        return -dc.mean()


class DC_and_CE_loss(nn.Module):
    def __init__(self, soft_dice_kwargs, ce_kwargs, aggregate="sum", square_dice=False, weight_ce=1, weight_dice=1,
                 log_dice=False, ignore_label=None):
        super().__init__()
        if square_dice or ignore_label is not None or ce_kwargs:
            raise NotImplementedError("the fused E2ENet loss implements the trainer's configuration: "
                                      "DC_and_CE_loss({'batch_dice': ..., 'smooth': 1e-5, 'do_bg': False}, {}) "
                                      "(nnUNetTrainer_simple.py:100)")
        if aggregate != "sum":
            raise NotImplementedError("nah son")
        self.log_dice, self.weight_dice, self.weight_ce, self.aggregate = log_dice, weight_dice, weight_ce, aggregate
        self.ignore_label = None
        self.dc = SoftDiceLoss(apply_nonlin=None, **soft_dice_kwargs)

    def forward(self, net_output, target):
        if not self.log_dice and self.dc.apply_nonlin is None:
            # the trainer's configuration: statistics -> loss -> gradient coefficients entirely in 4 + 1 kernel launches
            return ops.DiceCELoss.apply(net_output, target, self.dc.smooth, self.dc.do_bg, self.dc.batch_dice,
                                        self.weight_ce, self.weight_dice)
        sp, tp, sy, ce_sum = ops.SoftmaxStats.apply(net_output, target)
        n_vox = net_output.shape[0] * net_output[0, 0].numel()
        dc_loss = self.dc.from_stats(sp, tp, sy) if self.weight_dice != 0 else 0
        if self.log_dice:
            dc_loss = -torch.log(-dc_loss)
        ce_loss = ce_sum / n_vox if self.weight_ce != 0 else 0
        return self.weight_ce * ce_loss + self.weight_dice * dc_loss


class MultipleOutputLoss2(nn.Module):
    def __init__(self, loss, weight_factors=None):
        super().__init__()
        self.weight_factors = weight_factors
        self.loss = loss

    def forward(self, x, y):
        assert isinstance(x, (tuple, list)), "x must be either tuple or list"
        assert isinstance(y, (tuple, list)), "y must be either tuple or list"
        weights = [1] * len(x) if self.weight_factors is None else self.weight_factors
        l = weights[0] * self.loss(x[0], y[0])
        for i in range(1, len(x)):
            if weights[i] != 0:
                l += weights[i] * self.loss(x[i], y[i])
        return l
