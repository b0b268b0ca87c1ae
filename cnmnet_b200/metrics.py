"""Depth-evaluation metrics of the reference's ``cal_metrics`` (eval.py:1009-1069, utils/metric.py:149-362) on the GPU:
one fused masked reduction per image (csrc/metrics.cu) instead of nine numpy passes over masked copies."""
import torch

from . import _lib

MIN_DEPTH, MAX_DEPTH = 0.3, 8.0      # eval.py:1009-1010
NAMES = ("l1", "abs_relative", "rmse", "scale_invariant", "sq_relative", "rmse_log", "a1", "a2", "a3")


def depth_metrics(pred_depth, gt_depth, min_depth=MIN_DEPTH, max_depth=MAX_DEPTH):
    """pred_depth, gt_depth [B,H,W] CUDA tensors -> (per_image [B,10] float64 tensor on the device with columns
    num_valid + NAMES, dict of means over images like eval.py:1060-1069).  Prediction and ground truth must have the
    same size (the reference resizes the prediction to the ground truth with cv2 first, eval.py:1030)."""
    if pred_depth.shape != gt_depth.shape or pred_depth.dim() != 3:
        raise ValueError("expected pred and gt depth maps [B,H,W] of equal shape")
    if not pred_depth.is_cuda:
        raise RuntimeError("cnmnet_b200 runs on CUDA tensors only (no CPU path)")
    lib = _lib.load()
    B, H, W = pred_depth.shape
    p, g = pred_depth.float().contiguous(), gt_depth.float().contiguous()
    out = torch.empty((B, 10), dtype=torch.float64, device=p.device)
    ws = torch.empty((lib.cnm_depth_metrics_workspace_bytes(B, H, W),), dtype=torch.uint8, device=p.device)
    _lib.check(lib.cnm_depth_metrics(p.data_ptr(), g.data_ptr(), out.data_ptr(), ws.data_ptr(), B, H, W, float(min_depth),
                                     float(max_depth), _lib.stream_ptr()), "depth_metrics")
    means = out[:, 1:].mean(0).tolist()       # plain mean over images: NaN if any image has no valid pixel, as np.mean
    return out, dict(zip(NAMES, means))
