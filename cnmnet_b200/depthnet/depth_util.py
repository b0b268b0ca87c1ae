"""Camera pre-processing, depth->normal and the Combined Normal Map with the reference's
signatures (reference: depthnet/depth_util.py:13-56, 140-278), executed by csrc/normals.cu and
csrc/cost_volume.cu."""
import torch
import torch.nn as nn

from .. import _lib


def get_pixel_coordinates(height, width, device="cuda"):
    """Homogeneous pixel grid [3, W*H], x-major flattening (flat = x*H + y) like depth_util.py:13-21."""
    xs = torch.arange(width, dtype=torch.float32, device=device).view(width, 1).expand(width, height)
    ys = torch.arange(height, dtype=torch.float32, device=device).view(1, height).expand(width, height)
    return torch.stack([xs, ys, torch.ones_like(xs)], 0).reshape(3, -1)


def process_camera_parameters(left_cam, right_cam, pixel_coordinates):
    """cams [B,2,4,4] -> (KRKiUV [B,3,W*H], KT [B,3,1])   (depth_util.py:24-56).

    The 3x3 homography and K_s t come from cnm_camera_setup (fp64 on device); expanding them over the
    pixel grid is kept only for API compatibility -- the fused cost-volume kernel never needs it."""
    if not left_cam.is_cuda:
        raise RuntimeError("cnmnet_b200 runs on CUDA tensors only (no CPU path)")
    B = left_cam.shape[0]
    lib = _lib.load()
    hk = torch.empty((B, 12), dtype=torch.float32, device=left_cam.device)
    _lib.check(lib.cnm_camera_setup(_lib.ptr(left_cam.float().contiguous()), _lib.ptr(right_cam.float().contiguous()),
                                    hk.data_ptr(), B, 1, _lib.stream_ptr()), "camera_setup")
    Hm, kt = hk[:, :9].reshape(B, 3, 3), hk[:, 9:].reshape(B, 3, 1)
    return torch.matmul(Hm, pixel_coordinates.to(hk.device)), kt.contiguous()


def _planes_num_tensor(planes_num, device):
    if torch.is_tensor(planes_num):
        return planes_num.to(device=device, dtype=torch.int32).contiguous()
    return torch.tensor([int(p) for p in planes_num], dtype=torch.int32, device=device)


class Depth2normal(nn.Module):
    """Reference: depth_util.py:140 -- ``Depth2normal(k_size=9)``."""

    def __init__(self, k_size=9):
        super().__init__()
        self.k_size = k_size

    def forward(self, depth, intrinsic_inv, instance_segs=None, planes_num=None, need_points=True):
        """depth [B,H,W], intrinsic_inv [B,3,3] -> (normals [B,3,H,W], points [B,3,H,W]); with planes:
        (normals, loss, points)   (depth_util.py:149-240)."""
        if depth.dim() != 3 or tuple(intrinsic_inv.shape) != (depth.shape[0], 3, 3):
            raise ValueError("expected depth [B,H,W] and intrinsic_inv [B,3,3]")
        if not depth.is_cuda:
            raise RuntimeError("cnmnet_b200 runs on CUDA tensors only (no CPU path)")
        B, H, W = depth.shape
        lib, st = _lib.load(), _lib.stream_ptr()
        d = depth.float().contiguous()
        Ki = intrinsic_inv.float().contiguous()
        normals = torch.empty((B, 3, H, W), dtype=torch.float32, device=d.device)
        points = torch.empty((B, 3, H, W), dtype=torch.float32, device=d.device) if need_points else None
        _lib.check(lib.cnm_depth2normal(d.data_ptr(), Ki.data_ptr(), normals.data_ptr(), _lib.ptr(points), B, H, W,
                                        int(self.k_size), st), "depth2normal")
        if planes_num is None:
            return normals, points
        segs = instance_segs.to(torch.uint8).contiguous()
        pn = _planes_num_tensor(planes_num, d.device)
        loss = torch.zeros((), dtype=torch.float64, device=d.device)
        ws = _cnm_workspace(lib, B, segs.shape[1], H, W, d.device)
        _lib.check(lib.cnm_plane_regularize(normals.data_ptr(), segs.data_ptr(), pn.data_ptr(), loss.data_ptr(),
                                            ws.data_ptr(), B, segs.shape[1], H, W, st), "plane_regularize")
        return normals, loss.float(), points


def _cnm_workspace(lib, B, P, H, W, device):
    nbytes = lib.cnm_plane_regularize_workspace_bytes(B, P, H, W)
    if nbytes <= 0:
        raise ValueError("cnm_plane_regularize supports 1..32 plane slots, got %d" % P)
    return torch.empty((nbytes,), dtype=torch.uint8, device=device)


def get_normal_by_planes(gt_normal, instance_segs, planes_num):
    """[B,3,H,W], segs [B,P,H,W], planes_num [B] -> [B,3,H,W]   (depth_util.py:243-278)."""
    if not gt_normal.is_cuda:
        raise RuntimeError("cnmnet_b200 runs on CUDA tensors only (no CPU path)")
    B, _, H, W = gt_normal.shape
    out = gt_normal.float().clone().contiguous()
    segs = instance_segs.to(torch.uint8).contiguous()
    pn = _planes_num_tensor(planes_num, out.device)
    lib = _lib.load()
    ws = _cnm_workspace(lib, B, segs.shape[1], H, W, out.device)
    _lib.check(lib.cnm_plane_regularize(out.data_ptr(), segs.data_ptr(), pn.data_ptr(), None, ws.data_ptr(), B,
                                        segs.shape[1], H, W, _lib.stream_ptr()), "plane_regularize")
    return out
