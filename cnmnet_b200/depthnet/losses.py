"""Forward values of the reference's training losses (depthnet/losses.py) on the GPU: same class / function names and
arguments, one fused masked reduction each (csrc/losses.cu).  Forward only -- this package implements the inference
path; the values serve validation-loss reporting and as the first step toward training."""
import torch
import torch.nn as nn

from .. import _lib


def _prep(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cnmnet_b200 runs on CUDA tensors only (no CPU path)")
    return [None if t is None else t.float().contiguous() for t in tensors]


def _workspace(lib, device):
    return torch.empty((lib.cnm_loss_workspace_bytes(),), dtype=torch.uint8, device=device)


def _idepth(pred, gt, prob, log):
    if pred.shape != gt.shape or (prob is not None and prob.shape != gt.shape):
        raise ValueError("idepth_pred, idepth_gt (and prob_map) must have the same shape")
    lib = _lib.load()
    p, g, w = _prep(pred, gt, prob)
    out = torch.empty((2,), dtype=torch.float64, device=p.device)
    ws = _workspace(lib, p.device)
    _lib.check(lib.cnm_idepth_loss(p.data_ptr(), g.data_ptr(), _lib.ptr(w), out.data_ptr(), ws.data_ptr(), p.numel(),
                                   1 if log else 0, _lib.stream_ptr()), "idepth_loss")
    return out[0].float()


class IdepthLoss(nn.Module):
    """losses.py:31-48: L1 (or L1 of log10) over pixels with gt > 0, pred > 0, both finite."""

    def forward(self, idepth_pred, idepth_groud_truth, log=False):
        return _idepth(idepth_pred, idepth_groud_truth, None, log)


class IdepthwithProbLoss(nn.Module):
    """losses.py:51-74: mean of prob_map * |diff| over the same mask (log mode: 10 * |log10 diff|)."""

    def forward(self, idepth_pred, idepth_gt, prob_map, log=False):
        return _idepth(idepth_pred, idepth_gt, prob_map, log)


class IdepthLoss_234(nn.Module):
    """losses.py:7-28: 0.1 * mean over disp2..4 of mean|disp_k - nearest-downsampled ground truth|."""

    def forward(self, idepth_preds, idepth_ground_truth):
        _, disp2, disp3, disp4 = idepth_preds
        B, _, H, W = idepth_ground_truth.shape
        for k, d in enumerate((disp2, disp3, disp4)):
            if tuple(d.shape) != (B, 1, H >> (k + 1), W >> (k + 1)):
                raise ValueError("disp%d must be [B,1,H/%d,W/%d]" % (k + 2, 2 << k, 2 << k))
        lib = _lib.load()
        d2, d3, d4, g = _prep(disp2, disp3, disp4, idepth_ground_truth)
        out = torch.empty((1,), dtype=torch.float64, device=g.device)
        ws = _workspace(lib, g.device)
        _lib.check(lib.cnm_idepth_loss_234(d2.data_ptr(), d3.data_ptr(), d4.data_ptr(), g.data_ptr(), out.data_ptr(),
                                           ws.data_ptr(), B, H, W, _lib.stream_ptr()), "idepth_loss_234")
        return out[0].float()


def surface_normal_loss(prediction, surface_normal, valid_region, probability_map=None):
    """losses.py:77-122 -> (loss, mean angle in degrees).  prediction / surface_normal [B,3,H,W], valid_region
    [B,1,H,W] bool, probability_map [B,1,H,W] or None."""
    B, C, H, W = prediction.shape
    if C != 3 or surface_normal.shape != prediction.shape:
        raise ValueError("expected prediction and surface_normal [B,3,H,W]")
    lib = _lib.load()
    p, g, w = _prep(prediction, surface_normal, probability_map)
    v = valid_region.to(torch.uint8).contiguous()
    out = torch.empty((3,), dtype=torch.float64, device=p.device)
    ws = _workspace(lib, p.device)
    _lib.check(lib.cnm_surface_normal_loss(p.data_ptr(), g.data_ptr(), v.data_ptr(), _lib.ptr(w), out.data_ptr(),
                                           ws.data_ptr(), B, H, W, _lib.stream_ptr()), "surface_normal_loss")
    return out[0].float(), out[1].float()
