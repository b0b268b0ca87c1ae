"""Shared helpers for the -m gpu parity tests (all compute calls go through the C ABI)."""
import numpy as np
import torch

from cnmnet_b200 import _lib, engine


def dev():
    return torch.device("cuda", 0)


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def to_act(x_nchw, dtype):
    """NCHW fp32 (CPU) -> act tensor on the GPU through cnm_nchw_to_act."""
    lib = _lib.load()
    x = x_nchw.float().contiguous().to(dev())
    N, C, H, W = x.shape
    out = torch.empty((N, H, W, C * engine.phys_factor(dtype)), dtype=engine.TORCH_DTYPE[dtype], device=dev())
    _lib.check(lib.cnm_nchw_to_act(x.data_ptr(), out.data_ptr(), N, C, H, W, C, 0, dtype, _lib.stream_ptr()))
    return out


def from_act(a, dtype):
    lib = _lib.load()
    N, H, W, C = a.shape
    C //= engine.phys_factor(dtype)
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=a.device)
    _lib.check(lib.cnm_act_to_nchw(a.data_ptr(), out.data_ptr(), N, C, H, W, C, 0, dtype, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return out.cpu()


def quantize(x, dtype):
    """What storing x in an act tensor does to it."""
    if dtype == _lib.CNM_BF16:
        return x.to(torch.bfloat16).float()
    if dtype == _lib.CNM_BF16X3:
        hi = x.float().to(torch.bfloat16).float()
        return hi + (x.float() - hi).to(torch.bfloat16).float()
    if dtype == _lib.CNM_F16F8:
        hi, lo8, _ = f16f8_parts(x)
        return hi + lo8 / engine.F8_LO_SCALE
    return engine.round_tf32(x.float())


def e4m3(t):
    return t.float().clamp(-448.0, 448.0).to(torch.float8_e4m3fn).float()


def f16f8_parts(x):
    """CNM_F16F8 storage of x (include/cnmnet_b200.h): fp16 hi, e4m3 lo8 = (x - hi) * 2^7, e4m3 hi8 = hi * 2^-4."""
    x = x.float()
    hi = x.clamp(-65504.0, 65504.0).to(torch.float16).float()
    return hi, e4m3((x - hi) * engine.F8_LO_SCALE), e4m3(hi * engine.F8_HI_SCALE)


def act_empty(shape, dtype):
    """Uninitialised act tensor for a LOGICAL [N,H,W,C] shape (physical channels per the mode)."""
    N, H, W, C = shape
    return torch.empty((N, H, W, C * engine.phys_factor(dtype)), dtype=engine.TORCH_DTYPE[dtype], device=dev())
