"""pixel2cam and the depth-driven inverse warp with the reference's signatures
(reference: depthnet/inverse_warp.py:18-118)."""
import torch

from .. import _lib


def check_sizes(input, input_name, expected):
    """Same contract as inverse_warp.py:18-24: AssertionError on a rank / fixed-size mismatch."""
    condition = [input.ndimension() == len(expected)]
    for i, size in enumerate(expected):
        if size.isdigit():
            condition.append(input.size(i) == int(size))
    assert all(condition), "wrong size for {}, expected {}, got  {}".format(input_name, "x".join(expected),
                                                                            list(input.size()))


def pixel2cam(depth, intrinsics_inv):
    """depth [B,H,W], Kinv [B,3,3] -> camera-frame points [B,3,H,W]   (inverse_warp.py:27-43)."""
    check_sizes(depth, "depth", "BHW")
    check_sizes(intrinsics_inv, "intrinsics_inv", "B33")
    if not depth.is_cuda:
        raise RuntimeError("cnmnet_b200 runs on CUDA tensors only (no CPU path)")
    B, H, W = depth.shape
    d = depth.float().contiguous()
    Ki = intrinsics_inv.float().contiguous()
    scratch = torch.empty((B, 3, H, W), dtype=torch.float32, device=d.device)
    points = torch.empty((B, 3, H, W), dtype=torch.float32, device=d.device)
    lib = _lib.load()
    # k_size = 1: the plane fit degenerates, only the point cloud output is used
    _lib.check(lib.cnm_depth2normal(d.data_ptr(), Ki.data_ptr(), scratch.data_ptr(), points.data_ptr(), B, H, W, 1,
                                    _lib.stream_ptr()), "pixel2cam")
    return points


def inverse_warp(feat, depth, pose, intrinsics, intrinsics_inv, padding_mode="zeros"):
    """feat [B,C,H,W], depth [B,H,W], pose [B,3,4], K/Kinv [B,3,3] -> warped [B,C,H,W]   (inverse_warp.py:79-118)."""
    check_sizes(depth, "depth", "BHW")
    check_sizes(pose, "pose", "B34")
    check_sizes(intrinsics, "intrinsics", "B33")
    check_sizes(intrinsics_inv, "intrinsics", "B33")
    assert intrinsics_inv.size() == intrinsics.size()
    if padding_mode not in ("zeros", "border"):
        raise ValueError("padding_mode must be 'zeros' or 'border'")
    if not feat.is_cuda:
        raise RuntimeError("cnmnet_b200 runs on CUDA tensors only (no CPU path)")
    B, Cc, H, W = feat.shape
    if tuple(depth.shape) != (B, H, W):
        # the reference builds the sampling grid from depth's own size (inverse_warp.py:101-110); the kernel indexes
        # depth with feat's H and W, so a depth map at another resolution must be resampled by the caller
        raise ValueError("inverse_warp: depth %s must be [B,H,W] = %s of feat" % (tuple(depth.shape), (B, H, W)))
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=feat.device)
    lib = _lib.load()
    args = [t.float().contiguous() for t in (feat, depth, pose, intrinsics, intrinsics_inv)]
    _lib.check(lib.cnm_inverse_warp(*[a.data_ptr() for a in args], out.data_ptr(), B, Cc, H, W,
                                    0 if padding_mode == "zeros" else 1, _lib.stream_ptr()), "inverse_warp")
    return out
