"""Oracle: DepthNet / DepthRefineNet forward as a functional torch-CPU restatement, the
refine-input aggregation, idepth->depth and the depth-driven inverse warp.
TEST INFRASTRUCTURE ONLY.

The convolution / batch-norm / upsample arithmetic lives in the reference's third-party
dependency PyTorch (pinned here: torch 2.11), so the oracle calls the same functional ops
on CPU; what is restated is the wiring:
  * layer factories         depthnet/depthNet_model.py:19-112
  * depthNet.forward        depthnet/depthNet_model.py:226-263
  * DepthRefineNet.forward  depthnet/depthNet_model.py:331-370
  * view-group means        eval.py:659-663, 920-929
  * idepth -> depth         eval.py:269, 452, 934
  * inverse_warp            depthnet/inverse_warp.py:46-118
Everything takes a plain ``state_dict`` (reference key names) so it runs on weights
produced by the reference modules, and a dtype so fp64 runs are possible.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import cost_volume as _cv

BN_EPS = 1e-5

# (name, kind, cin, cout, k).  kind: down = conv(s1)+BN+ReLU, conv(s2)+BN+ReLU; up = bilinear x2 + conv+BN+ReLU;
# conv = conv+BN+ReLU; head = conv3x3(+bias) + sigmoid.  Order = registration order in the reference ctor.
DEPTHNET_LAYERS = [
    ("conv1", "down", 67, 128, 7), ("conv2", "down", 128, 256, 5), ("conv3", "down", 256, 512, 3),
    ("conv4", "down", 512, 512, 3), ("conv5", "down", 512, 512, 3),
    ("upconv5", "up", 512, 512, 3), ("iconv5", "conv", 1024, 512, 3),
    ("upconv4", "up", 512, 512, 3), ("iconv4", "conv", 1024, 512, 3), ("disp4", "head", 512, 1, 3),
    ("upconv3", "up", 512, 256, 3), ("iconv3", "conv", 513, 256, 3), ("disp3", "head", 256, 1, 3),
    ("upconv2", "up", 256, 128, 3), ("iconv2", "conv", 257, 128, 3), ("disp2", "head", 128, 1, 3),
    ("upconv1", "up", 128, 64, 3), ("iconv1", "conv", 65, 64, 3), ("disp1", "head", 64, 1, 3),
]
REFINENET_LAYERS = [
    ("conv1", "down", 67, 128, 3), ("conv2", "down", 128, 256, 3), ("conv3", "down", 256, 512, 3),
    ("upconv3_depth", "up", 512, 256, 3), ("iconv3_depth", "conv", 512, 256, 3),
    ("upconv2_depth", "up", 256, 128, 3), ("iconv2_depth", "conv", 256, 128, 3),
    ("upconv1_depth", "up", 128, 64, 3), ("iconv1_depth", "conv", 64, 64, 3), ("disp_refine", "head", 64, 1, 3),
    ("upconv3_prob", "up", 512, 256, 3), ("iconv3_prob", "conv", 512, 256, 3),
    ("upconv2_prob", "up", 256, 128, 3), ("iconv2_prob", "conv", 256, 128, 3),
    ("upconv1_prob", "up", 128, 64, 3), ("iconv1_prob", "conv", 64, 64, 3), ("prob", "head", 64, 1, 3),
]


def depthnet_layers(planes=64):
    """DEPTHNET_LAYERS with conv1 widened to 3 + planes inputs.  planes = 64 is the reference (depthNet_model.py:129);
    other values are the extension BASELINE config 5 names (128 planes) -- no reference implementation exists for them,
    the wiring below is unchanged."""
    return [("conv1", "down", 3 + planes, 128, 7)] + DEPTHNET_LAYERS[1:]


def state_shapes(layers):
    """state_dict key -> shape, with the reference's nn.Sequential indices."""
    shapes = OrderedDict()

    def conv_bn(prefix, ci, cs, co, k):
        shapes["%s.%d.weight" % (prefix, ci)] = (co, cs, k, k)
        for nm, shp in (("weight", (co,)), ("bias", (co,)), ("running_mean", (co,)), ("running_var", (co,)),
                        ("num_batches_tracked", ())):
            shapes["%s.%d.%s" % (prefix, ci + 1, nm)] = shp

    for name, kind, cin, cout, k in layers:
        if kind == "down":
            conv_bn(name, 0, cin, cout, k)
            conv_bn(name, 3, cout, cout, k)
        elif kind == "conv":
            conv_bn(name, 0, cin, cout, k)
        elif kind == "up":
            conv_bn(name, 1, cin, cout, k)
        elif kind == "head":
            shapes[name + ".0.weight"] = (1, cin, 3, 3)
            shapes[name + ".0.bias"] = (1,)
    return shapes


def _cbr(sd, prefix, ci, x, k, stride):
    w = sd["%s.%d.weight" % (prefix, ci)].to(x.dtype)
    x = F.conv2d(x, w, None, stride=stride, padding=(k - 1) // 2)
    p = "%s.%d." % (prefix, ci + 1)
    x = F.batch_norm(x, sd[p + "running_mean"].to(x.dtype), sd[p + "running_var"].to(x.dtype),
                     sd[p + "weight"].to(x.dtype), sd[p + "bias"].to(x.dtype), False, 0.0, BN_EPS)
    return F.relu(x)


def _down(sd, name, x, k):
    return _cbr(sd, name, 3, _cbr(sd, name, 0, x, k, 1), k, 2)


def _up(sd, name, x):
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)  # nn.Upsample(2,'bilinear')
    return _cbr(sd, name, 1, x, 3, 1)


def _head_logit(sd, name, x):
    return F.conv2d(x, sd[name + ".0.weight"].to(x.dtype), sd[name + ".0.bias"].to(x.dtype), padding=1)


def _nearest2(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")  # F.upsample default mode


def depthnet_trunk(sd, x, idepth_scale=3.0, return_logits=False):
    """x = cat(left_image, volume) [B,67,H,W] -> ([disp1..disp4], iconv1)   (depthNet_model.py:235-263)."""
    k = {n: kk for n, _, _, _, kk in DEPTHNET_LAYERS}
    conv1 = _down(sd, "conv1", x, k["conv1"])
    conv2 = _down(sd, "conv2", conv1, k["conv2"])
    conv3 = _down(sd, "conv3", conv2, 3)
    conv4 = _down(sd, "conv4", conv3, 3)
    conv5 = _down(sd, "conv5", conv4, 3)
    iconv5 = _cbr(sd, "iconv5", 0, torch.cat((_up(sd, "upconv5", conv5), conv4), 1), 3, 1)
    iconv4 = _cbr(sd, "iconv4", 0, torch.cat((_up(sd, "upconv4", iconv5), conv3), 1), 3, 1)
    l4 = _head_logit(sd, "disp4", iconv4)
    disp4 = idepth_scale * torch.sigmoid(l4)
    iconv3 = _cbr(sd, "iconv3", 0, torch.cat((_up(sd, "upconv3", iconv4), conv2, _nearest2(disp4)), 1), 3, 1)
    l3 = _head_logit(sd, "disp3", iconv3)
    disp3 = idepth_scale * torch.sigmoid(l3)
    iconv2 = _cbr(sd, "iconv2", 0, torch.cat((_up(sd, "upconv2", iconv3), conv1, _nearest2(disp3)), 1), 3, 1)
    l2 = _head_logit(sd, "disp2", iconv2)
    disp2 = idepth_scale * torch.sigmoid(l2)
    iconv1 = _cbr(sd, "iconv1", 0, torch.cat((_up(sd, "upconv1", iconv2), _nearest2(disp2)), 1), 3, 1)
    l1 = _head_logit(sd, "disp1", iconv1)
    disp1 = idepth_scale * torch.sigmoid(l1)
    if return_logits:
        return [disp1, disp2, disp3, disp4], iconv1, [l1, l2, l3, l4]
    return [disp1, disp2, disp3, disp4], iconv1


def depthnet_forward(sd, left_image, right_image, left_cam, right_cam, idepth_scale=3.0,
                     dtype=torch.float32, return_logits=False, volume=None, num_planes=64):
    """depthNet.forward (depthNet_model.py:226-263): cost volume (fp64 oracle, cast) + trunk."""
    if volume is None:
        idmin, idmax = _cv.idepth_range(float(idepth_scale))
        np_dt = np.float64 if dtype == torch.float64 else np.float32
        vol = _cv.cost_volume(left_image.numpy(), right_image.numpy(), left_cam.numpy(), right_cam.numpy(),
                              num_planes, idmin, idmax, dtype=np.float64).astype(np_dt)
        volume = torch.from_numpy(vol)
    x = torch.cat((left_image.to(dtype), volume.to(dtype)), 1)
    return depthnet_trunk(sd, x, idepth_scale, return_logits)


def refine_input(idepth01, idepth02, iconv01, iconv02):
    """x = cat(id1, id2, |id1 - id2|, iconv1 + iconv2)   (depthNet_model.py:332-333)."""
    return torch.cat((idepth01, idepth02, torch.abs(idepth01 - idepth02), iconv01 + iconv02), 1)


def group_mean(tensors):
    """View-group mean with the reference's op order: (x0+x2)/2 (eval.py:660-663), (x0+x2+x4)/3. (eval.py:920-929)."""
    acc = tensors[0]
    for t in tensors[1:]:
        acc = acc + t
    return acc / float(len(tensors)) if len(tensors) > 1 else acc


def refinenet_forward(sd, idepth01, idepth02, iconv01, iconv02, idepth_scale=2.0, return_volume=False,
                      return_logits=False):
    """DepthRefineNet.forward (depthNet_model.py:331-370)."""
    x = refine_input(idepth01, idepth02, iconv01, iconv02)
    conv1 = _down(sd, "conv1", x, 3)
    conv2 = _down(sd, "conv2", conv1, 3)
    conv3 = _down(sd, "conv3", conv2, 3)
    outs = {}
    for br, head in (("depth", "disp_refine"), ("prob", "prob")):
        i3 = _cbr(sd, "iconv3_" + br, 0, torch.cat((_up(sd, "upconv3_" + br, conv3), conv2), 1), 3, 1)
        i2 = _cbr(sd, "iconv2_" + br, 0, torch.cat((_up(sd, "upconv2_" + br, i3), conv1), 1), 3, 1)
        i1 = _cbr(sd, "iconv1_" + br, 0, _up(sd, "upconv1_" + br, i2), 3, 1)
        outs[br] = (i1, _head_logit(sd, head, i1))
    disp_refined = idepth_scale * torch.sigmoid(outs["depth"][1])
    prob_map = torch.sigmoid(outs["prob"][1])
    res = [disp_refined, prob_map]
    if return_volume:
        res.append(outs["depth"][0])
    if return_logits:
        res.append((outs["depth"][1], outs["prob"][1]))
    return tuple(res)


def idepth_to_depth(idepth):
    """eval.py:269 / 452 / 934: depth = 1 / idepth  ([B,1,H,W] -> [B,H,W])."""
    return torch.div(1.0, idepth).squeeze(1)


def inverse_warp(feat, depth, pose, intrinsics, intrinsics_inv, padding_mode="zeros"):
    """Depth-driven warp (inverse_warp.py:46-118), torch restatement in the input dtype."""
    B, C, H, W = feat.shape
    dt = feat.dtype
    ys, xs = torch.meshgrid(torch.arange(H, dtype=dt), torch.arange(W, dtype=dt), indexing="ij")
    pix = torch.stack((xs, ys, torch.ones_like(xs)), 0).reshape(1, 3, -1).expand(B, 3, -1)
    cam = intrinsics_inv.bmm(pix) * depth.reshape(B, 1, -1)
    proj = intrinsics.bmm(pose)  # [B,3,4]
    p = proj[:, :, :3].bmm(cam) + proj[:, :, 3:]
    X, Y, Z = p[:, 0], p[:, 1], p[:, 2].clamp(min=1e-3)
    xn = 2 * (X / Z) / (W - 1) - 1
    yn = 2 * (Y / Z) / (H - 1) - 1
    if padding_mode == "zeros":
        xn = torch.where((xn > 1) | (xn < -1), torch.full_like(xn, 2.0), xn)
        yn = torch.where((yn > 1) | (yn < -1), torch.full_like(yn, 2.0), yn)
    grid = torch.stack((xn, yn), 2).reshape(B, H, W, 2)
    return F.grid_sample(feat, grid, mode="bilinear", padding_mode=padding_mode, align_corners=False)


class _RefModules(torch.nn.Module):
    """Module tree with the reference's registration order, used only to reproduce its
    random init stream (depthNet_model.py:127-182, 273-329): nn.Conv2d default init consumes
    the RNG at construction, then kaiming_normal_(fan_out) re-draws in modules() order."""

    def __init__(self, layers):
        super().__init__()
        nn = torch.nn
        for name, kind, cin, cout, k in layers:
            pad = (k - 1) // 2
            if kind == "down":
                seq = nn.Sequential(nn.Conv2d(cin, cout, k, padding=pad, stride=1, bias=False), nn.BatchNorm2d(cout),
                                    nn.ReLU(), nn.Conv2d(cout, cout, k, padding=pad, stride=2, bias=False),
                                    nn.BatchNorm2d(cout), nn.ReLU())
            elif kind == "conv":
                seq = nn.Sequential(nn.Conv2d(cin, cout, k, padding=pad, bias=False), nn.BatchNorm2d(cout), nn.ReLU())
            elif kind == "up":
                seq = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear"),
                                    nn.Conv2d(cin, cout, k, padding=pad, bias=False), nn.BatchNorm2d(cout), nn.ReLU())
            else:
                seq = nn.Sequential(nn.Conv2d(cin, 1, 3, padding=1), nn.Sigmoid())
            setattr(self, name, seq)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)


def reference_init_state(layers, seed):
    """state_dict equal to what the reference constructor produces under torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    return _RefModules(layers).state_dict()


def cost_volume_torch(left_image, right_image, left_cam, right_cam, idepth_scale=3.0, num_planes=64):
    """fp32 torch-CPU port of the reference's camera pre-processing + getVolume, op for op
    (depth_util.py:13-56, depthNet_model.py:185-224).  Used for the timed CPU baseline (it runs at the
    reference's own speed: 64 x {mul, add, div, grid_sample, abs, sum}); parity checks use the fp64
    numpy restatement in oracle/cost_volume.py."""
    B, C, H, W = left_image.shape
    idmin, idmax = _cv.idepth_range(float(idepth_scale))
    pix = np.indices([W, H]).astype(np.float32)
    pix = np.concatenate((pix, np.ones([1, W, H], dtype=np.float32)), axis=0).reshape(3, -1)
    pix = torch.from_numpy(pix)
    left_E, left_K = left_cam[:, 0], left_cam[:, 1, :3, :3]
    right_E, right_K = right_cam[:, 0], right_cam[:, 1, :3, :3]
    r2l = torch.matmul(right_E, left_E.inverse())
    R, t = r2l[:, :3, :3], r2l[:, :3, 3]
    KRKiUV = torch.matmul(torch.matmul(right_K, torch.matmul(R, left_K.inverse())), pix)
    KT = torch.matmul(right_K, t.unsqueeze(-1))
    step = (idmax - idmin) / (num_planes - 1.0)
    norm = torch.tensor([W / 2.0, H / 2.0], dtype=torch.float32).view(1, 2, 1)
    out = torch.empty((B, num_planes, H, W), dtype=torch.float32)
    for d in range(num_planes):
        z = 1.0 / (idmin + d * step)
        tr = KRKiUV * z + KT
        uv = tr[:, 0:2, :] / (tr[:, 2:3, :] + 1e-6)
        uv = ((uv - norm) / norm).view(B, 2, W, H).permute(0, 3, 2, 1)
        warped = F.grid_sample(right_image, uv, mode="bilinear", padding_mode="zeros", align_corners=False)
        out[:, d] = torch.sum(torch.abs(warped - left_image), dim=1)
    return out


def depth2normal_torch(depth, intrinsic_inv, k_size=9):
    """fp32 torch-CPU port of Depth2normal.forward without planes (depth_util.py:149-203), op for op:
    Unfold -> mask -> bmm -> det -> inverse -> matmul -> normalise.  Timed CPU baseline only."""
    b, h, w = depth.shape
    ys, xs = torch.meshgrid(torch.arange(h, dtype=depth.dtype), torch.arange(w, dtype=depth.dtype), indexing="ij")
    pix = torch.stack((xs, ys, torch.ones_like(xs)), 0).reshape(1, 3, -1).expand(b, 3, -1)
    points = intrinsic_inv.bmm(pix).view(b, 3, h, w) * depth.unsqueeze(1)
    valid = ((depth > 0) & (depth < 10.0)).to(depth.dtype).unsqueeze(1)
    kk = k_size * k_size
    unf = torch.nn.Unfold(kernel_size=(k_size, k_size), padding=k_size // 2, stride=(1, 1))
    A = unf(points).view(-1, 3, kk, h, w).permute(0, 3, 4, 2, 1)
    vm = unf(valid).view(-1, 1, kk, h, w).permute(0, 3, 4, 2, 1).repeat(1, 1, 1, 1, 3) > 0.5
    A = torch.where(vm, A, torch.zeros_like(A))
    At = torch.transpose(A, 3, 4).reshape(-1, 3, kk)
    A = A.reshape(-1, kk, 3)
    M = torch.bmm(At, A)
    det = M.det()
    bad = torch.isnan(det) | (det < 1e-5)
    M = torch.where(bad.view(-1, 1, 1), torch.eye(3, dtype=depth.dtype).expand_as(M), M)
    n = torch.matmul(torch.matmul(torch.inverse(M), At), torch.ones(At.shape[0], kk, 1, dtype=depth.dtype)).squeeze(-1)
    n = n / (torch.norm(n, dim=1, keepdim=True) + 1e-5)
    return n.view(b, h, w, 3).permute(0, 3, 1, 2), points


def reference_pipeline(dsd, rsd, ref, srcs, ref_cam, src_cams, idepth_scale=3.0, k_size=9, num_planes=64):
    """The reference's eval_refine call sequence (eval.py:440-455) on the fp32 torch-CPU ports above:
    V x depthNet (with its own cost volume), refine, 1/idepth, Depth2normal.  -> depth, normal, prob."""
    V = srcs.shape[1]
    ids, ics = [], []
    with torch.no_grad():
        for v in range(V):
            vol = cost_volume_torch(ref, srcs[:, v], ref_cam, src_cams[:, v], idepth_scale, num_planes)
            d, ic = depthnet_trunk(dsd, torch.cat((ref, vol), 1), idepth_scale)
            ids.append(d[0])
            ics.append(ic)
        idr, prob = refinenet_forward(rsd, group_mean(ids[0::2]), group_mean(ids[1::2]), group_mean(ics[0::2]),
                                      group_mean(ics[1::2]), idepth_scale)
        depth = idepth_to_depth(idr)
        Kinv = torch.inverse(ref_cam[:, 1, :3, :3])
        normal, _ = depth2normal_torch(depth, Kinv, k_size)
    return depth, normal, prob
