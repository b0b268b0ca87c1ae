"""Generates tests/golden/losses.npz by calling the UNMODIFIED reference depthnet/losses.py on seeded synthetic tensors.
Build-container only (needs /root/reference).

    python -m oracle.make_golden_losses
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("CNM_REFERENCE", "/root/reference")


def main():
    spec = importlib.util.spec_from_file_location("ref_losses", os.path.join(REF, "depthnet", "losses.py"))
    L = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(L)
    g = torch.Generator().manual_seed(77)
    B, H, W = 2, 32, 48
    gt = torch.rand(B, 1, H, W, generator=g) * 2 + 0.1
    gt[0, 0, :3] = 0.0
    gt[1, 0, 5, 5:9] = float("nan")
    gt[1, 0, 7, 7] = float("inf")
    pred = gt.nan_to_num(1.0, 1.0) * (1 + 0.1 * torch.randn(B, 1, H, W, generator=g)) + 0.01
    pred[0, 0, 10, :4] = -0.5
    pred[1, 0, 11, 11] = float("nan")
    prob = torch.rand(B, 1, H, W, generator=g)
    disps = [pred] + [torch.rand(B, 1, H // f, W // f, generator=g) * 2 for f in (2, 4, 8)]
    gt_pos = gt.nan_to_num(1.0, 1.0).clamp_min(0.05)
    n_pred = torch.randn(B, 3, H, W, generator=g)
    n_gt = torch.randn(B, 3, H, W, generator=g)
    n_gt[0, :, 3, 3] = float("nan")
    n_pred[1, 1, 4, 4] = float("inf")
    valid = torch.rand(B, 1, H, W, generator=g) > 0.2
    out = {
        "idepth": float(L.IdepthLoss()(pred, gt)), "idepth_log": float(L.IdepthLoss()(pred, gt, log=True)),
        "idepth_prob": float(L.IdepthwithProbLoss()(pred, gt, prob)),
        "idepth_prob_log": float(L.IdepthwithProbLoss()(pred, gt, prob, log=True)),
        "idepth_234": float(L.IdepthLoss_234()(disps, gt_pos)),
    }
    loss, ang = L.surface_normal_loss(n_pred, n_gt, valid)
    out["normal"], out["normal_angle"] = float(loss), float(ang)
    loss, ang = L.surface_normal_loss(n_pred, n_gt, valid, prob)
    out["normal_prob"], out["normal_prob_angle"] = float(loss), float(ang)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "losses.npz"), gt=gt.numpy(), pred=pred.numpy(), prob=prob.numpy(),
                        disp2=disps[1].numpy(), disp3=disps[2].numpy(), disp4=disps[3].numpy(), gt_pos=gt_pos.numpy(),
                        n_pred=n_pred.numpy(), n_gt=n_gt.numpy(), valid=valid.numpy(),
                        **{k: np.float64(v) for k, v in out.items()})
    print(out)


if __name__ == "__main__":
    main()
