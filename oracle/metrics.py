"""Oracle: the reference's depth-evaluation metrics (utils/metric.py:149-362 as called by eval.py:1009-1060,
``cal_metrics``), restated in numpy.  TEST INFRASTRUCTURE ONLY.

Per image: pred is clamped to [MIN_DEPTH, MAX_DEPTH] (eval.py:1031-1032), the valid mask comes from the ground truth
alone (``compute_valid_depth_mask(gt, min, max)``, utils/metric.py:149-162: min < gt < max and finite), and the nine
numbers of eval.py:1037-1047 are computed on the masked pixels with the reference's argument order
(``l1(gt, pred)``, ``abs_relative(depth_gt=gt, depth_pred=pred)``, ``sq_relative(pred, gt)``, log differences as
log(gt) - log(pred)).  The sequence-level result is the plain mean over images (eval.py:1060-1069).
"""
import numpy as np

MIN_DEPTH, MAX_DEPTH = 0.3, 8.0          # eval.py:1009-1010
NAMES = ("l1", "abs_relative", "rmse", "scale_invariant", "sq_relative", "rmse_log", "a1", "a2", "a3")
THRESHOLDS = (1.25, 1.25 * 1.25, 1.25 * 1.25 * 1.25)


def image_metrics(pred, gt, min_depth=MIN_DEPTH, max_depth=MAX_DEPTH):
    """pred, gt [H,W] -> dict of the nine metrics + 'num_valid' (NaN metrics when no pixel is valid)."""
    pred = np.asarray(pred, dtype=np.float64).copy()
    gt = np.asarray(gt, dtype=np.float64)
    pred[pred < min_depth] = min_depth
    pred[pred > max_depth] = max_depth
    valid = (gt < max_depth) & (gt > min_depth) & np.isfinite(gt)
    g, p = gt[valid], pred[valid]
    n = float(g.size)
    out = {"num_valid": int(g.size)}
    if n == 0:
        out.update({k: float("nan") for k in NAMES})
        return out
    diff = g - p
    logd = np.log(g) - np.log(p)
    out["l1"] = np.sum(np.abs(diff)) / n                                   # metric.py:165-183
    out["abs_relative"] = np.sum(np.abs(p - g) / g) / n                    # :276-294
    out["rmse"] = np.sqrt(np.sum(np.square(diff)) / n)                     # :232-250
    out["scale_invariant"] = np.sqrt(np.sum(np.square(logd)) / n - np.square(np.sum(logd)) / np.square(n))  # :253-273
    out["sq_relative"] = np.sum(np.square(p - g) / g) / n                  # :320-338
    out["rmse_log"] = np.sqrt(np.sum(np.square(logd)) / n)                 # :209-229
    for name, thr in zip(("a1", "a2", "a3"), THRESHOLDS):                  # :341-362
        out[name] = float(np.sum(np.abs(logd) < np.log(thr))) / n
    return out


def sequence_metrics(preds, gts, min_depth=MIN_DEPTH, max_depth=MAX_DEPTH):
    """[B,H,W] x2 -> (per-image list of dicts, dict of means over images) like eval.py:1060-1069."""
    per = [image_metrics(p, g, min_depth, max_depth) for p, g in zip(preds, gts)]
    return per, {k: float(np.mean([m[k] for m in per])) for k in NAMES}
