"""Oracle: plane-sweep cost volume (numpy restatement).  TEST INFRASTRUCTURE ONLY.

Follows the reference arithmetic step by step:
  * camera pre-processing   depthnet/depth_util.py:24-56   (process_camera_parameters)
  * pixel grid (x-major)    depthnet/depth_util.py:13-21   (get_pixel_coordinates)
  * plane loop + warp       depthnet/depthNet_model.py:185-224 (depthNet.getVolume)
  * F.grid_sample(bilinear, zeros, align_corners=False) as pinned by torch 2.11
    (the arithmetic lives in the third-party dependency PyTorch).

``dtype=np.float64`` is the exact restatement used as the parity target for the CUDA
kernel (inputs are the fp32 tensors promoted to fp64); ``dtype=np.float32`` mimics the
reference's own rounding points and is only used to measure its noise floor.
"""
import numpy as np


def idepth_range(idepth_scale):
    """depthNet_model.py:186-191.  Any other scale is an UnboundLocalError in the reference."""
    if idepth_scale == 2.0:
        return 0.02, 2.0
    if idepth_scale == 3.0:
        return 0.1, 3.0
    raise ValueError("reference getVolume only defines idepth_scale 2.0 and 3.0")


def plane_depths(num_planes, idepth_min, idepth_max):
    """z_d = 1 / (idmin + d * (idmax - idmin) / (D - 1)) in Python double (depthNet_model.py:193-194,209)."""
    step = (idepth_max - idepth_min) / (num_planes - 1.0)
    return np.array([1.0 / (idepth_min + d * step) for d in range(num_planes)], dtype=np.float64)


def homography(left_cam, right_cam, dtype=np.float64):
    """Per pair: Hm = K_s R K_r^-1 [B,3,3], kt = K_s t [B,3]   (depth_util.py:32-52)."""
    left_cam = np.asarray(left_cam, dtype=dtype)
    right_cam = np.asarray(right_cam, dtype=dtype)
    E_r, K_r = left_cam[:, 0], left_cam[:, 1, :3, :3]
    E_s, K_s = right_cam[:, 0], right_cam[:, 1, :3, :3]
    T = E_s @ np.linalg.inv(E_r)
    R, t = T[:, :3, :3], T[:, :3, 3]
    Hm = K_s @ (R @ np.linalg.inv(K_r))
    kt = np.einsum("bij,bj->bi", K_s, t)
    return Hm.astype(dtype), kt.astype(dtype)


def cost_volume(left_image, right_image, left_cam, right_cam, num_planes=64,
                idepth_min=0.1, idepth_max=3.0, dtype=np.float64, Hm=None, kt=None):
    """[B,C,H,W] x2 + cams -> cost [B,D,H,W].  cost[b,d,y,x] = sum_c |warp_d(right)[b,c,y,x] - left[b,c,y,x]|."""
    left = np.asarray(left_image, dtype=dtype)
    right = np.asarray(right_image, dtype=dtype)
    B, C, H, W = left.shape
    if Hm is None:
        Hm, kt = homography(left_cam, right_cam, dtype)
    Hm = np.asarray(Hm, dtype=dtype)
    kt = np.asarray(kt, dtype=dtype)
    ys, xs = np.meshgrid(np.arange(H, dtype=dtype), np.arange(W, dtype=dtype), indexing="ij")
    zs = plane_depths(num_planes, idepth_min, idepth_max)
    out = np.empty((B, num_planes, H, W), dtype=dtype)
    half_w, half_h = dtype(W / 2.0), dtype(H / 2.0)
    eps = dtype(1e-6)
    for b in range(B):
        # a = Hm . [x, y, 1]^T   (KRKiUV, depth_util.py:43)
        a = [Hm[b, i, 0] * xs + Hm[b, i, 1] * ys + Hm[b, i, 2] for i in range(3)]
        for d in range(num_planes):
            z = dtype(zs[d])
            px, py, pz = a[0] * z + kt[b, 0], a[1] * z + kt[b, 1], a[2] * z + kt[b, 2]
            with np.errstate(divide="ignore", invalid="ignore"):
                u = px / (pz + eps)
                v = py / (pz + eps)
                # normalise (depthNet_model.py:213) then grid_sample's un-normalise (align_corners=False)
                gx = (u - half_w) / half_w
                gy = (v - half_h) / half_h
                ix = ((gx + 1) * W - 1) / 2
                iy = ((gy + 1) * H - 1) / 2
            x0f, y0f = np.floor(ix), np.floor(iy)
            wx1, wy1 = ix - x0f, iy - y0f
            wx0, wy0 = 1 - wx1, 1 - wy1
            warped = np.zeros((C, H, W), dtype=dtype)
            for dy, wy in ((0, wy0), (1, wy1)):
                for dx, wx in ((0, wx0), (1, wx1)):
                    xt, yt = x0f + dx, y0f + dy
                    ok = (xt >= 0) & (xt <= W - 1) & (yt >= 0) & (yt <= H - 1)  # NaN -> False
                    xi = np.where(ok, xt, 0).astype(np.int64)
                    yi = np.where(ok, yt, 0).astype(np.int64)
                    w = np.where(ok, wx * wy, 0).astype(dtype)
                    warped += right[b][:, yi, xi] * w[None]
            out[b, d] = np.abs(warped - left[b]).sum(axis=0)
    return out
