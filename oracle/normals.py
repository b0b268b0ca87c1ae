"""Oracle: depth -> point cloud -> k x k plane-fit normals, and the Combined Normal Map.
TEST INFRASTRUCTURE ONLY.

Restates (in numpy, fp64 by default):
  * pixel2cam                       depthnet/inverse_warp.py:27-43
  * Depth2normal.forward            depthnet/depth_util.py:149-203,240
  * CNM branch of Depth2normal      depthnet/depth_util.py:205-238
  * get_normal_by_planes            depthnet/depth_util.py:243-278
"""
import numpy as np


def pixel2cam(depth, intrinsics_inv, dtype=np.float64):
    """P[b,:,y,x] = Kinv[b] . [x, y, 1]^T * depth[b,y,x]  -> [B,3,H,W]."""
    depth = np.asarray(depth, dtype=dtype)
    Kinv = np.asarray(intrinsics_inv, dtype=dtype)
    B, H, W = depth.shape
    ys, xs = np.meshgrid(np.arange(H, dtype=dtype), np.arange(W, dtype=dtype), indexing="ij")
    pix = np.stack([xs, ys, np.ones_like(xs)], axis=0).reshape(3, -1)
    rays = np.einsum("bij,jn->bin", Kinv, pix).reshape(B, 3, H, W)
    with np.errstate(invalid="ignore"):
        return rays * depth[:, None]


def depth2normal(depth, intrinsics_inv, k_size=9, dtype=np.float64, return_det=False):
    """-> (normals [B,3,H,W], points [B,3,H,W]).  See SURVEY.md Appendix A.2."""
    depth = np.asarray(depth)
    B, H, W = depth.shape
    points = pixel2cam(depth, intrinsics_inv, dtype)
    with np.errstate(invalid="ignore"):
        valid = (depth > 0) & (depth < 10.0)  # NaN / inf -> invalid (depth_util.py:162)
    q = np.where(valid[:, None], points, 0.0).astype(dtype)  # depth_util.py:176-177
    r = k_size // 2
    qp = np.pad(q, ((0, 0), (0, 0), (r, r), (r, r)))  # Unfold zero padding (depth_util.py:165)
    M = np.zeros((B, H, W, 3, 3), dtype=dtype)
    s = np.zeros((B, H, W, 3), dtype=dtype)
    for j in range(k_size):
        for i in range(k_size):
            w = qp[:, :, j:j + H, i:i + W]  # [B,3,H,W]
            wv = np.moveaxis(w, 1, -1)  # [B,H,W,3]
            M += wv[..., :, None] * wv[..., None, :]
            s += wv
    det = np.linalg.det(M)
    bad = np.isnan(det) | (det < 1e-5)  # depth_util.py:185-187
    M = np.where(bad[..., None, None], np.eye(3, dtype=dtype), M)
    n = np.einsum("...ij,...j->...i", np.linalg.inv(M), s)
    n = n / (np.linalg.norm(n, axis=-1, keepdims=True) + 1e-5)  # depth_util.py:201
    normals = np.moveaxis(n, -1, 1)
    if return_det:
        return normals, points, det
    return normals, points


def _replace_by_plane_mean(normal_bhw3, instance_segs, planes_num, with_loss, dtype):
    """Sequential per-plane replace-by-mean (depth_util.py:217-236 / :261-276)."""
    out = np.array(normal_bhw3, dtype=dtype, copy=True)
    segs = np.asarray(instance_segs) != 0
    B, H, W, _ = out.shape
    loss = dtype(0.0)
    for b in range(len(planes_num)):
        cur = out[b]
        for i in range(int(planes_num[b])):
            m = segs[b, i]
            cnt = dtype(m.sum())
            with np.errstate(invalid="ignore", divide="ignore"):
                mean = (cur * m[..., None]).reshape(-1, 3).sum(axis=0) / cnt  # NOT renormalised
            if with_loss:
                orig = np.where(m[..., None], cur, 0.0).reshape(-1, 3)
                # F.cosine_similarity(eps=1e-8): x.y / (max(|x|, eps) * max(|y|, eps))
                nm = max(float(np.linalg.norm(mean)), 1e-8)
                no = np.maximum(np.linalg.norm(orig, axis=1), 1e-8)
                cos = (orig @ mean) / (nm * no)
                loss = loss + np.mean(1.0 - cos)
            cur = np.where(m[..., None], mean[None, None, :], cur)
        out[b] = cur
    return out, loss


def depth2normal_cnm(depth, intrinsics_inv, instance_segs, planes_num, k_size=9, dtype=np.float64):
    """Depth2normal.forward with planes -> (normals [B,3,H,W], loss, points)."""
    normals, points = depth2normal(depth, intrinsics_inv, k_size, dtype)
    out, loss = _replace_by_plane_mean(np.moveaxis(normals, 1, -1), instance_segs, planes_num, True, dtype)
    return np.moveaxis(out, -1, 1), loss, points


def get_normal_by_planes(gt_normal, instance_segs, planes_num, dtype=np.float64):
    """[B,3,H,W] -> [B,3,H,W] with each plane's pixels replaced by the plane-mean normal."""
    n = np.moveaxis(np.asarray(gt_normal, dtype=dtype), 1, -1)
    out, _ = _replace_by_plane_mean(n, instance_segs, planes_num, False, dtype)
    return np.moveaxis(out, -1, 1)
