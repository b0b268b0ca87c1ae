"""Oracle: forward values of the reference's training losses (depthnet/losses.py), restated with torch-CPU ops in
fp64.  TEST INFRASTRUCTURE ONLY.

  idepth_loss            losses.py:31-48   (IdepthLoss.forward)
  idepth_prob_loss       losses.py:51-74   (IdepthwithProbLoss.forward)
  idepth_loss_234        losses.py:7-28    (IdepthLoss_234.forward)
  surface_normal_loss    losses.py:77-122
"""
import numpy as np
import torch
import torch.nn.functional as F


def _mask(pred, gt):
    return (gt > 0.0) & torch.isfinite(gt) & torch.isfinite(pred) & (pred > 0.0)


def idepth_loss(pred, gt, log=False):
    m = _mask(pred, gt)
    p, g = pred[m].double(), gt[m].double()
    return float((torch.log10(p) - torch.log10(g)).abs().mean() if log else (p - g).abs().mean())


def idepth_prob_loss(pred, gt, prob, log=False):
    m = _mask(pred, gt)
    p, g, w = pred[m].double(), gt[m].double(), prob[m].double()
    diff = 10 * (torch.log10(p) - torch.log10(g)).abs() if log else (p - g).abs()
    return float((w * diff).mean())


def idepth_loss_234(disps, gt):
    _, d2, d3, d4 = disps
    tot = 0.0
    for d in (d2, d3, d4):
        g = F.interpolate(gt, size=d.shape[2:4])      # default mode: nearest
        tot += float((d.double() - g.double()).abs().mean())
    return 0.1 * tot / 3.0


def surface_normal_loss(prediction, surface_normal, valid_region, probability_map=None):
    ok = torch.isfinite(surface_normal.sum(1, keepdim=True)) & torch.isfinite(prediction.sum(1, keepdim=True)) & valid_region
    m = ok[:, 0]
    p = prediction.permute(0, 2, 3, 1)[m].double()
    g = surface_normal.permute(0, 2, 3, 1)[m].double()
    sim = (p * g).sum(1) / (p.norm(dim=1).clamp_min(1e-8) * g.norm(dim=1).clamp_min(1e-8))
    if probability_map is None:
        loss = (1 - sim).mean()
    else:
        w = probability_map[:, 0][m].double()
        loss = ((1 - sim) * w).sum() / w.sum()
    angle = torch.acos(sim.clamp(-1, 1)).mean() / np.pi * 180
    return float(loss), float(angle)
