"""depthNet and DepthRefineNet with the reference's constructor / forward / state_dict contract
(reference: depthnet/depthNet_model.py:124-263 and :268-370), executed by the sm_100a kernels.

The nn.Module tree below exists only to hold parameters under the reference's key names
("conv1.0.weight", "conv1.1.running_mean", "disp1.0.bias", ...) and to reproduce its random
init; forward() never calls these modules -- it folds/packs the weights into an execution plan
(cnmnet_b200/engine.py) and runs CUDA kernels through the C ABI.
"""
import torch
import torch.nn as nn

from .. import _lib
from ..engine import DepthNetPlan, RefinePlan, parse_dtype

_DEPTHNET = [
    ("conv1", "down", 67, 128, 7), ("conv2", "down", 128, 256, 5), ("conv3", "down", 256, 512, 3),
    ("conv4", "down", 512, 512, 3), ("conv5", "down", 512, 512, 3),
    ("upconv5", "up", 512, 512, 3), ("iconv5", "conv", 1024, 512, 3),
    ("upconv4", "up", 512, 512, 3), ("iconv4", "conv", 1024, 512, 3), ("disp4", "head", 512, 1, 3),
    ("upconv3", "up", 512, 256, 3), ("iconv3", "conv", 513, 256, 3), ("disp3", "head", 256, 1, 3),
    ("upconv2", "up", 256, 128, 3), ("iconv2", "conv", 257, 128, 3), ("disp2", "head", 128, 1, 3),
    ("upconv1", "up", 128, 64, 3), ("iconv1", "conv", 65, 64, 3), ("disp1", "head", 64, 1, 3),
]
_REFINE = [("conv1", "down", 67, 128, 3), ("conv2", "down", 128, 256, 3), ("conv3", "down", 256, 512, 3)]
for _br, _head in (("depth", "disp_refine"), ("prob", "prob")):
    _REFINE += [("upconv3_" + _br, "up", 512, 256, 3), ("iconv3_" + _br, "conv", 512, 256, 3),
                ("upconv2_" + _br, "up", 256, 128, 3), ("iconv2_" + _br, "conv", 256, 128, 3),
                ("upconv1_" + _br, "up", 128, 64, 3), ("iconv1_" + _br, "conv", 64, 64, 3), (_head, "head", 64, 1, 3)]


def _block(kind, cin, cout, k):
    pad = (k - 1) // 2
    if kind == "down":
        return nn.Sequential(nn.Conv2d(cin, cout, k, padding=pad, stride=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(),
                             nn.Conv2d(cout, cout, k, padding=pad, stride=2, bias=False), nn.BatchNorm2d(cout), nn.ReLU())
    if kind == "conv":
        return nn.Sequential(nn.Conv2d(cin, cout, k, padding=pad, bias=False), nn.BatchNorm2d(cout), nn.ReLU())
    if kind == "up":
        return nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear"),
                             nn.Conv2d(cin, cout, k, padding=pad, bias=False), nn.BatchNorm2d(cout), nn.ReLU())
    return nn.Sequential(nn.Conv2d(cin, 1, 3, padding=1), nn.Sigmoid())


class _KernelNet(nn.Module):
    """Parameter holder + plan cache shared by both networks."""

    def _build(self, layers):
        for name, kind, cin, cout, k in layers:
            setattr(self, name, _block(kind, cin, cout, k))
        # same init as the reference (depthNet_model.py:163-182): kaiming_normal(fan_out), BN 1/0, bias 0
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self._plans = {}
        self._packed = {}

    def _weights_version(self):
        return tuple(t._version for t in self.state_dict(keep_vars=True).values()) + (next(self.parameters()).data_ptr(),)

    def _get_plan(self, key, factory):
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = factory()
        ver = self._weights_version()
        if self._packed.get(key) != ver:
            with torch.no_grad():
                plan.load_state({k: v.detach() for k, v in self.state_dict().items()})
            self._packed[key] = ver
            # some packed values are launch ARGUMENTS (head bias, the f16f8 accumulator scale): a CUDA graph captured
            # before this point has the old ones baked in -- holders of a graph compare this counter (CNMPipeline)
            plan.weights_epoch = getattr(plan, "weights_epoch", 0) + 1
        return plan

    def load_reference_checkpoint(self, state_dict):
        """Accepts the reference's checkpoints, with or without DataParallel's 'module.' prefix (eval.py:184-198)."""
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        return self.load_state_dict(sd)

    def train(self, mode=True):
        if mode:
            raise RuntimeError("cnmnet_b200 implements the inference path only (BatchNorm is folded); use .eval()")
        return super().train(False)


def _check_image_pair(left_image, right_image):
    if left_image.dim() != 4 or left_image.shape[1] != 3 or left_image.shape != right_image.shape:
        raise ValueError("expected left/right images [B,3,H,W] of equal shape, got %s and %s"
                         % (tuple(left_image.shape), tuple(right_image.shape)))
    if not left_image.is_cuda:
        raise RuntimeError("cnmnet_b200 runs on CUDA tensors only (no CPU path)")


class depthNet(_KernelNet):
    """Reference: depthnet/depthNet_model.py:124 -- ``depthNet(idepth_scale=3)``.

    Extra keyword arguments (not in the reference): ``precision`` 'f16f8' (default: the mode that meets the 1e-3
    final-depth tolerance at the best speed) | 'bf16x3' | 'tf32' | 'bf16' (speed mode, misses the tolerance) selects the
    tensor-core operand type, ``backend`` 'tc' (tcgen05, product) | 'simt' (CUDA-core validation twin), ``planes`` the
    number of inverse-depth hypotheses.  The reference hard-codes 64 planes (getVolume loops ``range(64)``,
    conv1 takes 3 + 64 channels, depthNet_model.py:129,193); ``planes=128`` is the variant BASELINE config 5 names:
    the same sweep sampled at 128 planes over the same inverse-depth range, conv1.0 widened to 3 + 128 inputs."""

    def __init__(self, idepth_scale=3, precision="f16f8", backend="tc", planes=64):
        super().__init__()
        if planes < 64 or planes % 64:
            raise ValueError("planes must be a positive multiple of 64 (one 128-byte operand row per 64 planes)")
        self.idepth_scale = idepth_scale
        self.precision = parse_dtype(precision)
        self.backend = backend
        self.planes = int(planes)
        self._build([("conv1", "down", 3 + self.planes, 128, 7)] + _DEPTHNET[1:])
        super().train(False)

    def plan(self, N, Bref, H, W, want_logits=False, lane=0):
        """`lane` distinguishes independent buffer sets of the same shape (plans that run concurrently on two streams)."""
        dev = next(self.parameters()).device
        key = (N, Bref, H, W, self.precision, self.backend, str(dev), want_logits, lane)
        return self._get_plan(key, lambda: DepthNetPlan(N, Bref, H, W, float(self.idepth_scale), self.precision, dev,
                                                        self.backend, planes=self.planes, want_logits=want_logits,
                                                        lane=lane))

    def getVolume(self, left_image, right_image, KRKiUV_T, KT_T):
        """[B,3,H,W] x2, KRKiUV_T [B,3,W*H] (x-major), KT_T [B,3,1] -> cost volume [B,planes (64),H,W] fp32
        (depthNet_model.py:185-224).  The 3x3 homography is recovered from three columns of KRKiUV_T: the origin and
        the two FARTHEST pixels along x and y, differenced in fp64, so that the fp32 rounding of the tabulated values
        is divided by W-1 / H-1 instead of being multiplied by the pixel coordinate."""
        _check_image_pair(left_image, right_image)
        B, _, H, W = left_image.shape
        s = float(self.idepth_scale)
        if s == 2.0:
            idmin, idmax = 0.02, 2.0
        elif s == 3.0:
            idmin, idmax = 0.1, 3.0
        else:
            raise ValueError("idepth_scale must be 2.0 or 3.0")
        a00 = KRKiUV_T[:, :, 0].double()                                            # Hm.[0,0,1]
        col0 = (KRKiUV_T[:, :, (W - 1) * H].double() - a00) / max(W - 1, 1)         # flat index x*H + y
        col1 = (KRKiUV_T[:, :, H - 1].double() - a00) / max(H - 1, 1)
        hk = torch.cat([torch.stack([col0, col1, a00], dim=2).reshape(B, 9), KT_T.reshape(B, 3).double()],
                       1).float().contiguous()
        lib = _lib.load()
        st = _lib.stream_ptr()
        left = left_image.float().contiguous()
        right = right_image.float().contiguous()
        out = torch.empty((B, self.planes, H, W), dtype=torch.float32, device=left.device)
        from .. import engine
        if engine.PLANE_SWEEP and W % 4 == 0 and self.planes <= 128 and right.data_ptr() % 16 == 0:
            _lib.check(lib.cnm_plane_sweep(_lib.ptr(left), _lib.ptr(right), hk.data_ptr(), out.data_ptr(), B, 1, H, W,
                                           self.planes, idmin, idmax, _lib.VOLUME_NCHW_F32, 0, st), "plane_sweep")
            return out
        # widths that TMA cannot address (row pitch not a multiple of 16 bytes): repack + L1-gather kernel
        tex = torch.empty((lib.cnm_packed_source_bytes(B, H, W),), dtype=torch.uint8, device=left.device)
        _lib.check(lib.cnm_pack_source(_lib.ptr(right), tex.data_ptr(), B, H, W, st), "pack_source")
        _lib.check(lib.cnm_cost_volume(_lib.ptr(left), tex.data_ptr(), hk.data_ptr(), out.data_ptr(), B, 1, H, W, self.planes, idmin,
                                       idmax, _lib.VOLUME_NCHW_F32, 0, st), "cost_volume")
        return out

    def forward(self, left_image, right_image, left_cam, right_cam):
        """-> ([disp1, disp2, disp3, disp4], iconv1)   (depthNet_model.py:226-263)."""
        _check_image_pair(left_image, right_image)
        B, _, H, W = left_image.shape
        plan = self.plan(B, B, H, W)
        plan.bind(left_image.float().contiguous(), right_image.float().contiguous(),
                  left_cam.float().contiguous(), right_cam.float().contiguous())
        plan.run()
        disps = [plan.disp1.clone(), plan.disp2.clone(), plan.disp3.clone(), plan.disp4.clone()]
        return disps, plan.iconv1_nchw()


class DepthRefineNet(_KernelNet):
    """Reference: depthnet/depthNet_model.py:268 -- ``DepthRefineNet(base_channels_num=32, idepth_scale=2)``
    (the first positional argument is unused there too)."""

    def __init__(self, base_channels_num=32, idepth_scale=2, precision="f16f8", backend="tc"):
        super().__init__()
        self.base_channels_num = base_channels_num
        self.idepth_scale = idepth_scale
        self.precision = parse_dtype(precision)
        self.backend = backend
        self._build(_REFINE)
        super().train(False)

    def plan(self, B, H, W, group=1, nchw_input=False, want_logits=False, lane=0):
        dev = next(self.parameters()).device
        key = (B, H, W, group, nchw_input, self.precision, self.backend, str(dev), want_logits, lane)
        return self._get_plan(key, lambda: RefinePlan(B, H, W, float(self.idepth_scale), self.precision, dev,
                                                      self.backend, group, nchw_input, want_logits, lane=lane))

    def forward(self, idepth01, idepth02, iconv01, iconv02, ReturnVolume=False):
        """-> (disp_refined, prob_map[, iconv1_depth])   (depthNet_model.py:331-370)."""
        if idepth01.dim() != 4 or idepth01.shape[1] != 1 or iconv01.shape[1] != 64:
            raise ValueError("expected idepth [B,1,H,W] and iconv [B,64,H,W]")
        if not idepth01.is_cuda:
            raise RuntimeError("cnmnet_b200 runs on CUDA tensors only (no CPU path)")
        B, _, H, W = idepth01.shape
        plan = self.plan(B, H, W, 1, True)
        idepth = torch.stack([idepth01.reshape(B, H, W), idepth02.reshape(B, H, W)]).float().contiguous()
        iconv = torch.stack([iconv01, iconv02]).float().contiguous()
        plan.bind(idepth, iconv)
        plan.run()
        res = (plan.out["depth"].clone(), plan.out["prob"].clone())
        if ReturnVolume:
            res = res + (plan.iconv1_depth_nchw(),)
        return res
