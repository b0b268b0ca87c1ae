"""Generates tests/golden/depth_metrics.npz by calling the UNMODIFIED reference functions (utils/metric.py) the way
eval.py:1030-1047 (cal_metrics) does, on seeded synthetic depth maps.  Build-container only (needs /root/reference).

    python -m oracle.make_golden_metrics
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from cnmnet_b200 import synth  # noqa: E402

REF = os.environ.get("CNM_REFERENCE", "/root/reference")


def main():
    spec = importlib.util.spec_from_file_location("ref_metric", os.path.join(REF, "utils", "metric.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    rng = np.random.default_rng(2024)
    B, H, W = 4, 48, 64
    gt = synth.depth_maps(rng, B, H, W).astype(np.float64)          # has zero / >= 10 m patches: invalid pixels
    gt[1, 5:9, 5:9] = np.inf
    gt[3] = 0.0                                                      # an image without a single valid pixel
    pred = (gt * (1.0 + 0.08 * rng.standard_normal(gt.shape)) + 0.02 * rng.standard_normal(gt.shape)).astype(np.float32)
    pred[~np.isfinite(pred)] = 3.0
    pred[0, :4] = 0.01                                               # below MIN_DEPTH: clamped
    pred[0, -4:] = 30.0                                              # above MAX_DEPTH: clamped
    MIN_DEPTH, MAX_DEPTH = 0.3, 8.0                                  # eval.py:1009-1010
    rows = []
    for b in range(B):
        g, p = gt[b].copy(), pred[b].astype(np.float64)
        p[p < MIN_DEPTH] = MIN_DEPTH
        p[p > MAX_DEPTH] = MAX_DEPTH
        mask = m.compute_valid_depth_mask(g, min_thred=MIN_DEPTH, max_thred=MAX_DEPTH)
        g, p = g[mask], p[mask]
        rows.append([float(mask.sum()), m.l1(g, p), m.abs_relative(depth_gt=g, depth_pred=p), m.rmse(g, p),
                     m.scale_invariant(g, p), m.sq_relative(p, g), m.rmse_log(g, p), m.ratio_threshold(g, p, 1.25),
                     m.ratio_threshold(g, p, 1.25 * 1.25), m.ratio_threshold(g, p, 1.25 * 1.25 * 1.25)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "depth_metrics.npz"), gt=gt, pred=pred,
                        metrics=np.array(rows, dtype=np.float64))
    print(np.array(rows))


if __name__ == "__main__":
    main()
