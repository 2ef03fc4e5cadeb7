"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

Restatement of the boundary-method criteria of /root/reference/src/training/losses.py: ``dice_loss`` (:40-68) and
``ce_dice`` (:71-96) in plain PyTorch (fp32, autograd).  PINNED: tests/golden/ce_dice_*.npz hold loss values and
gradients produced by the REAL reference functions (imported from /root/reference by tests/golden/make_golden.py);
tests/test_oracle_losses.py checks this restatement against them."""
import torch
import torch.nn.functional as F


def dice_loss(y_pred, y_true):
    smooth = 1.
    gt = y_true.contiguous().view(-1)
    pred = y_pred.contiguous().view(-1)
    pred_gt = torch.sum(gt * pred)
    return 1 - (2. * pred_gt + smooth) / (torch.sum(gt ** 2) + torch.sum(pred ** 2) + smooth)


def ce_dice(y_pred, y_true, num_classes=3):
    """y_pred [N,3,H,W] logits, y_true [N,H,W] int64"""
    one_hot = F.one_hot(y_true, num_classes).float().permute(0, 3, 1, 2)
    soft = F.softmax(y_pred, dim=1)
    ce = F.cross_entropy(y_pred, y_true)
    dice = 0
    for index in range(1, num_classes):
        dice = dice + index * dice_loss(soft[:, index], one_hot[:, index])
    return ce + 0.5 * dice


def boundary_loss(y_pred, y_true, kind):
    return ce_dice(y_pred, y_true) if kind == 'ce_dice' else F.cross_entropy(y_pred, y_true)
