"""Seeded synthetic inputs for the CNMNet hot path.

ScanNet / 7-Scenes frames and the pretrained checkpoint are not available offline, so
tests, smoke and bench all run on these deterministic stand-ins.  Everything is generated
with numpy's PCG64 ``Generator`` on the host (bit-reproducible for a given numpy build) and
returned as CPU torch tensors; callers move them to the device.

Camera tensor layout follows the reference (scannet/preprocess.py:29-46, eval.py:136-146):
``cam[B, 2, 4, 4]`` with ``cam[:, 0]`` the 4x4 world->camera extrinsic and
``cam[:, 1, :3, :3]`` the 3x3 intrinsic K.
"""
import math

import numpy as np
import torch

# 7-Scenes intrinsics at 640x480 (reference eval.py:48-50): fx = fy = 585, cx = 320, cy = 240.
_FX640, _CX640, _CY480 = 585.0, 320.0, 240.0


def intrinsics(height, width):
    """K scaled to the image the way reference eval.py:80-85,148-159 scales it."""
    sx, sy = width / 640.0, height / 480.0
    K = np.eye(3, dtype=np.float64)
    K[0, 0], K[1, 1] = _FX640 * sx, _FX640 * sy
    K[0, 2], K[1, 2] = _CX640 * sx, _CY480 * sy
    return K


def _rodrigues(rvec):
    th = float(np.linalg.norm(rvec))
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * Kx + (1 - math.cos(th)) * (Kx @ Kx)


def smooth_images(rng, batch, height, width, channels=3, noise=0.05):
    """Band-limited images: 8 random 2-D sinusoids per channel (periods 4..64 px), std ~ 1."""
    yy, xx = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64), indexing="ij")
    out = np.zeros((batch, channels, height, width), dtype=np.float64)
    for b in range(batch):
        for c in range(channels):
            acc = np.zeros((height, width))
            for _ in range(8):
                period = rng.uniform(4.0, 64.0)
                ang = rng.uniform(0.0, 2 * math.pi)
                ph = rng.uniform(0.0, 2 * math.pi)
                fx, fy = math.cos(ang) / period, math.sin(ang) / period
                acc += np.sin(2 * math.pi * (fx * xx + fy * yy) + ph)
            out[b, c] = acc / 2.0  # 8 unit sinusoids: std = sqrt(8 / 2) = 2
    out += noise * rng.standard_normal(out.shape)
    return out.astype(np.float32)


def cameras(rng, batch, height, width, rot_sigma=0.05, trans_sigma=0.15):
    """(ref_cam, src_cam) pair: random rigid ref pose, source = small motion applied to it."""
    K = intrinsics(height, width)
    ref = np.zeros((batch, 2, 4, 4), dtype=np.float64)
    src = np.zeros((batch, 2, 4, 4), dtype=np.float64)
    for b in range(batch):
        E_r = np.eye(4)
        E_r[:3, :3] = _rodrigues(rng.normal(0.0, 0.6, 3))
        E_r[:3, 3] = rng.normal(0.0, 1.0, 3)
        M = np.eye(4)
        M[:3, :3] = _rodrigues(rng.normal(0.0, rot_sigma, 3))
        M[:3, 3] = rng.normal(0.0, trans_sigma, 3)
        ref[b, 0], src[b, 0] = E_r, M @ E_r
        ref[b, 1, :3, :3] = K
        src[b, 1, :3, :3] = K
        ref[b, 1, 3, 3] = src[b, 1, 3, 3] = 1.0
    return ref.astype(np.float32), src.astype(np.float32)


def depth_maps(rng, batch, height, width, with_invalid=True):
    """Smooth slanted-plane depth + ripples + 2 cm noise, with invalid pixels for the mask path."""
    yy, xx = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64), indexing="ij")
    out = np.zeros((batch, height, width), dtype=np.float64)
    for b in range(batch):
        d = 2.0 + 0.002 * xx + 0.003 * yy + 0.05 * np.sin(xx / 7.0 + b) * np.cos(yy / 5.0)
        d += 0.02 * rng.standard_normal(d.shape)
        if with_invalid and height >= 24 and width >= 24:
            d[5:15, 7:17] = 0.0  # hole (invalid: depth <= 0)
            d[height - 3, width - 4] = 12.5  # beyond the 10 m validity bound
            d[2, 3] = 10.0  # exactly the (exclusive) bound
            d[height // 2, width // 2] = -1.0
        out[b] = d
    return out.astype(np.float32)


def plane_segments(rng, batch, height, width, max_planes=20):
    """Disjoint plane-instance masks [B, 20, H, W] (uint8) + planes_num [B] (reference
    scannet/dataloader_batch.py:180-194 builds them disjoint).  Vertical stripes of random width."""
    segs = np.zeros((batch, max_planes, height, width), dtype=np.uint8)
    nums = np.zeros((batch,), dtype=np.int64)
    for b in range(batch):
        n = int(rng.integers(3, 7))
        cuts = np.sort(rng.choice(np.arange(4, width - 4), size=n, replace=False))
        x0 = 0
        for i, x1 in enumerate(cuts):
            y0, y1 = int(rng.integers(0, height // 3)), int(rng.integers(2 * height // 3, height))
            segs[b, i, y0:y1, x0:x1] = 1
            x0 = int(x1) + int(rng.integers(0, 3))
        nums[b] = n
    return segs, nums


def frame_tuples(seed, batch, height, width, views=2):
    """One batch of frame tuples: ref image, V source images, ref cam, V source cams (CPU tensors)."""
    rng = np.random.default_rng(seed)
    ref = smooth_images(rng, batch, height, width)
    srcs = np.stack([smooth_images(rng, batch, height, width) for _ in range(views)], axis=1)
    ref_cam = None
    src_cams = []
    for v in range(views):
        rc, sc = cameras(rng, batch, height, width)
        if ref_cam is None:
            ref_cam = rc
        else:  # keep one ref pose; re-express this view's motion relative to it
            for b in range(batch):
                M = sc[b, 0].astype(np.float64) @ np.linalg.inv(rc[b, 0].astype(np.float64))
                sc[b, 0] = (M @ ref_cam[b, 0].astype(np.float64)).astype(np.float32)
        src_cams.append(sc)
    src_cams = np.stack(src_cams, axis=1)
    return (torch.from_numpy(ref), torch.from_numpy(srcs), torch.from_numpy(ref_cam), torch.from_numpy(src_cams))


def kaiming_state(shapes, seed, head_scale=1.0):
    """Deterministic stand-in for the reference's random init (depthNet_model.py:163-182):
    conv weights ~ N(0, 2 / fan_out), BN gamma 1 / beta 0 / mean 0 / var 1, biases 0.
    ``shapes`` maps state_dict key -> shape.  ``head_scale`` scales the 1-channel heads
    (x0.02 un-saturates the sigmoids so depth parity is not vacuous)."""
    rng = np.random.default_rng(seed)
    state = {}
    for key, shape in shapes.items():
        if key.endswith("num_batches_tracked"):
            state[key] = torch.zeros((), dtype=torch.long)
        elif key.endswith("running_var"):
            state[key] = torch.from_numpy((1.0 + 0.2 * rng.random(shape)).astype(np.float32))
        elif key.endswith("running_mean"):
            state[key] = torch.from_numpy((0.05 * rng.standard_normal(shape)).astype(np.float32))
        elif len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            w = rng.standard_normal(shape) * math.sqrt(2.0 / fan_out)
            if shape[0] == 1:
                w = w * head_scale
            state[key] = torch.from_numpy(w.astype(np.float32))
        elif key.endswith(".weight"):  # BN gamma
            state[key] = torch.from_numpy((1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32))
        else:  # BN beta / conv bias
            state[key] = torch.from_numpy((0.05 * rng.standard_normal(shape)).astype(np.float32))
    return state
