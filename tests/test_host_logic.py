"""CPU tests of the host-side logic: K-step programs + weight packing (emulated in torch), BN folding,
module / state_dict contract, the C-ABI export list, sharding helpers.  No GPU, no kernel calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cnmnet_b200
from cnmnet_b200 import _lib, engine, synth
from cnmnet_b200.dist import shard_range
from oracle import nets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def emulate_steps(srcs, steps, wp, bias, stride, ch, relu):
    """Reference semantics of the device K-step loop (include/cnmnet_b200.h), NHWC float tensors."""
    N, H, W, _ = srcs[0].shape
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    out = torch.zeros((N, Ho, Wo, wp.shape[0]), dtype=torch.float64)
    for i, (si, sub, dx, dy, c0) in enumerate(steps):
        x = srcs[si].double()
        if stride == 2:
            x = x[:, (sub >> 1)::2, (sub & 1)::2]
        hs, ws = x.shape[1:3]
        a = torch.zeros((N, Ho, Wo, ch), dtype=torch.float64)
        ys, xs = torch.arange(Ho) + dy, torch.arange(Wo) + dx
        vy, vx = (ys >= 0) & (ys < hs), (xs >= 0) & (xs < ws)
        a[:, vy.nonzero().flatten()[:, None], vx.nonzero().flatten()[None, :]] = \
            x[:, ys[vy][:, None], xs[vx][None, :], c0:c0 + ch]
        out += a @ wp[:, i * ch:(i + 1) * ch].double().T
    out += bias.double()
    return F.relu(out) if relu else out


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("k,stride", [(3, 1), (3, 2), (5, 1), (5, 2), (7, 1), (7, 2)])
def test_act_steps_reproduce_conv(k, stride):
    torch.manual_seed(k * 10 + stride)
    ch, cin, cout, N, H, W = 32, 64, 8, 2, 10, 12
    x = torch.randn(N, cin, H, W)
    w = torch.randn(cout, cin, k, k)
    b = torch.randn(cout)
    src = engine.act_src(nhwc(x), 0)
    steps, cols = engine.make_steps([src], k, stride, ch)
    wp = engine.pack_weights(w, cols, ch, _lib.CNM_TF32)
    got = emulate_steps([nhwc(x)], steps, wp, b, stride, ch, True)
    want = nhwc(F.relu(F.conv2d(x.double(), engine.round_tf32(w).double(), b.double(), stride=stride,
                                padding=(k - 1) // 2)))
    assert torch.allclose(got, want, atol=1e-9)


def test_concat_and_side_sources_reproduce_conv():
    """iconv3-style layer: cat(upconv [64], skip [32], nearest-upsampled 1-channel disparity)."""
    torch.manual_seed(3)
    ch, N, h, w = 32, 2, 5, 6
    up, skip, disp = torch.randn(N, 64, 2 * h, 2 * w), torch.randn(N, 32, 2 * h, 2 * w), torch.randn(N, 1, h, w)
    wt, b = torch.randn(16, 97, 3, 3), torch.randn(16)
    udisp = F.interpolate(disp, scale_factor=2, mode="nearest")
    want = nhwc(F.conv2d(torch.cat((up, skip, udisp), 1).double(), engine.round_tf32(wt).double(), b.double(), padding=1))
    # side tensor: channel 3r+s = tap (r,s) of the upsampled disparity (what cnm_pack_disp_taps writes)
    taps = torch.zeros(N, 2 * h, 2 * w, ch)
    pd = F.pad(udisp, (1, 1, 1, 1))
    for r in range(3):
        for s in range(3):
            taps[..., 3 * r + s] = pd[:, 0, r:r + 2 * h, s:s + 2 * w]
    srcs = [engine.act_src(nhwc(up), 0), engine.act_src(nhwc(skip), 64), engine.Src(taps, "taps3x3", 96, ch, 0, 1)]
    steps, cols = engine.make_steps(srcs, 3, 1, ch)
    wp = engine.pack_weights(wt, cols, ch, _lib.CNM_TF32)
    got = emulate_steps([s.buf for s in srcs], steps, wp, b, 1, ch, False)
    assert torch.allclose(got, want, atol=1e-9)


def test_rowtaps_source_reproduces_conv1():
    """conv1.0-style layer: 3 image channels as a 7-tap row-packed side tensor + cost channels."""
    torch.manual_seed(4)
    ch, N, H, W, k = 32, 1, 9, 11, 7
    img, cost = torch.randn(N, 3, H, W), torch.randn(N, 32, H, W)
    wt, b = torch.randn(8, 35, k, k), torch.randn(8)
    want = nhwc(F.conv2d(torch.cat((img, cost), 1).double(), engine.round_tf32(wt).double(), b.double(), padding=3))
    taps = torch.zeros(N, H, W, ch)
    pi = F.pad(img, (3, 3, 0, 0))
    for s in range(k):
        for c in range(3):
            taps[..., s * 3 + c] = pi[:, c, :, s:s + W]
    srcs = [engine.act_src(nhwc(cost), 3), engine.Src(taps, "rowtaps", 0, ch, 0, 3)]
    steps, cols = engine.make_steps(srcs, k, 1, ch)
    wp = engine.pack_weights(wt, cols, ch, _lib.CNM_TF32)
    got = emulate_steps([s.buf for s in srcs], steps, wp, b, 1, ch, False)
    assert torch.allclose(got, want, atol=1e-9)


def test_halo_groups_reproduce_conv():
    """Row-halo K loop (csrc/conv_halo.cu): groups + group-ordered weights == conv2d, incl. side sources."""
    torch.manual_seed(8)
    ch, N, H, W = 32, 1, 9, 11
    for k in (3, 5, 7):
        x, extra = torch.randn(N, 64, H, W), torch.randn(N, 3, H, W)
        wt, b = torch.randn(8, 67, k, k), torch.randn(8)
        want = nhwc(F.conv2d(torch.cat((extra, x), 1).double(), engine.round_tf32(wt).double(), b.double(), padding=k // 2))
        taps = torch.zeros(N, H, W, ch)
        if k == 3:      # 3 extra channels as 27 packed 3x3 taps (refine conv1.0)
            pe = F.pad(extra, (1, 1, 1, 1))
            for c in range(3):
                for r in range(3):
                    for s_ in range(3):
                        taps[..., 9 * c + 3 * r + s_] = pe[:, c, r:r + H, s_:s_ + W]
            side = engine.Src(taps, "taps3x3", 0, ch, 0, 3)
        elif k == 7:    # 3 extra channels as 7 row taps (DepthNet conv1.0)
            pe = F.pad(extra, (3, 3, 0, 0))
            for s_ in range(7):
                for c in range(3):
                    taps[..., s_ * 3 + c] = pe[:, c, :, s_:s_ + W]
            side = engine.Src(taps, "rowtaps", 0, ch, 0, 3)
        else:
            side = None
            wt[:, :3] = 0
            want = nhwc(F.conv2d(torch.cat((extra, x), 1).double(), engine.round_tf32(wt).double(), b.double(), padding=2))
        srcs = [engine.act_src(nhwc(x), 3)] + ([side] if side is not None else [])
        groups, cols, info = engine.make_groups(srcs, k, ch)
        steps = [(m // 4, 0, dx, t - info[m][0], c0) for (m, c0, dx, nt) in groups for t in range(nt)]
        assert len(steps) == len(cols)
        wp = engine.pack_weights(wt, cols, ch, _lib.CNM_TF32)
        got = emulate_steps([s_.buf for s_ in srcs], steps, wp, b, 1, ch, False)
        assert torch.allclose(got, want, atol=1e-9), k


@pytest.mark.parametrize("k", [3, 5, 7])
def test_halo_groups_stride2_reproduce_conv(k):
    torch.manual_seed(k)
    ch, cin, cout, N, H, W = 32, 64, 8, 1, 12, 14
    x = torch.randn(N, cin, H, W)
    w, b = torch.randn(cout, cin, k, k), torch.randn(cout)
    groups, cols, info = engine.make_groups([engine.act_src(nhwc(x), 0)], k, ch, stride=2)
    steps = [(m // 4, m % 4, dx, t - info[m][0], c0) for (m, c0, dx, nt) in groups for t in range(nt)]
    assert len(steps) == len(cols) == k * k * (cin // ch)
    assert all(info[m][1] == 16 + max(nt for (mm, _, _, nt) in groups if mm == m) - 1 for m in {g[0] for g in groups})
    wp = engine.pack_weights(w, cols, ch, _lib.CNM_TF32)
    got = emulate_steps([nhwc(x)], steps, wp, b, 2, ch, True)
    want = nhwc(F.relu(F.conv2d(x.double(), engine.round_tf32(w).double(), b.double(), stride=2, padding=k // 2)))
    assert torch.allclose(got, want, atol=1e-9)


def _split_phys(x_nhwc):
    """Logical NHWC float tensor -> CNM_BF16X3 physical layout [.., (hi 64 | lo 64) per 64-channel chunk]."""
    hi = x_nhwc.to(torch.bfloat16).float()
    lo = (x_nhwc - hi).to(torch.bfloat16).float()
    N, H, W, C = x_nhwc.shape
    return torch.stack((hi.view(N, H, W, C // 64, 64), lo.view(N, H, W, C // 64, 64)), 4).reshape(N, H, W, 2 * C)


@pytest.mark.parametrize("halo", [False, True])
def test_split_bf16_programs_reproduce_conv(halo):
    """bf16x3: three bf16 passes per logical chunk (hi*Whi + lo*Whi + hi*Wlo) on the hi/lo physical layout."""
    torch.manual_seed(21)
    ch, N, H, W, k = 64, 1, 7, 9, 3
    x, disp = torch.randn(N, 128, H, W), torch.randn(N, 1, H, W)
    wt, b = torch.randn(8, 129, k, k) / 30, torch.randn(8)
    want = nhwc(F.conv2d(torch.cat((x, disp), 1).double(), wt.double(), b.double(), padding=1))
    taps = torch.zeros(N, H, W, ch)
    pd = F.pad(disp, (1, 1, 1, 1))
    for r in range(3):
        for s_ in range(3):
            taps[..., 3 * r + s_] = pd[:, 0, r:r + H, s_:s_ + W]
    xp, tp = _split_phys(nhwc(x)), _split_phys(taps)
    srcs = [engine.act_src(xp, 0, phys=2), engine.Src(tp, "taps3x3", 128, ch, 0, 1)]
    if halo:
        groups, cols, info = engine.make_groups(srcs, k, ch, 1, split=True)
        steps = [(m // 4, 0, dx, t - info[m][0], c0) for (m, c0, dx, nt) in groups for t in range(nt)]
    else:
        steps, cols = engine.make_steps(srcs, k, 1, ch, split=True)
    assert len(steps) == len(cols) == 3 * (9 * 2 + 1)
    wp = engine.pack_weights(wt, cols, ch, _lib.CNM_BF16X3).float()
    got = emulate_steps([xp, tp], steps, wp, b, 1, ch, False)
    # 16 mantissa bits per operand, the lo*lo term dropped: ~1e-5 relative
    assert (got - want).abs().max() < 3e-5 * want.abs().max()
    bf = emulate_steps([nhwc(x).to(torch.bfloat16).float(), taps.to(torch.bfloat16).float()],
                       *engine.make_steps([engine.act_src(nhwc(x), 0), engine.Src(taps, "taps3x3", 128, ch, 0, 1)], k, 1, ch)[:1],
                       engine.pack_weights(wt, engine.make_steps([engine.act_src(nhwc(x), 0),
                                                                  engine.Src(taps, "taps3x3", 128, ch, 0, 1)], k, 1, ch)[1],
                                           ch, _lib.CNM_BF16).float(), b, 1, ch, False)
    assert (bf - want).abs().max() > 50 * (got - want).abs().max()   # plain bf16 is >50x worse on the same layer


def _f16f8_phys(x_nhwc):
    """Logical NHWC float tensor -> CNM_F16F8 physical BYTES [.., 256 per 64-channel chunk]: 128 B of fp16 hi, then
    64 B of e4m3 lo8 = (x - hi) * 2^7 and 64 B of e4m3 hi8 = hi * 2^-4 (include/cnmnet_b200.h)."""
    N, H, W, C = x_nhwc.shape
    hi = x_nhwc.to(torch.float16)
    lo8 = ((x_nhwc - hi.float()) * engine.F8_LO_SCALE).clamp(-448, 448).to(torch.float8_e4m3fn)
    hi8 = (hi.float() * engine.F8_HI_SCALE).clamp(-448, 448).to(torch.float8_e4m3fn)
    parts = (hi.view(N, H, W, C // 64, 64).contiguous().view(torch.uint8),
             lo8.view(N, H, W, C // 64, 64).contiguous().view(torch.uint8),
             hi8.view(N, H, W, C // 64, 64).contiguous().view(torch.uint8))
    return torch.cat(parts, 4).reshape(N, H, W, C * 4)


def emulate_f16f8_steps(srcs_bytes, steps, wp_bytes, bias, acc_exp):
    """The CNM_F16F8 K loop: the first half of the steps multiply a 128-byte e4m3 row with 128 e4m3 weight columns, the
    second half a 128-byte fp16 row (64 ch) with 64 fp16 weight columns; one accumulator, scaled by 2^-acc_exp before the
    bias."""
    N, H, W, _ = srcs_bytes[0].shape
    out = torch.zeros((N, H, W, wp_bytes.shape[0]), dtype=torch.float64)
    for i, (si, sub, dx, dy, c0) in enumerate(steps):
        f8 = i < len(steps) // 2                                     # corrections first, then the fp16 rows
        assert sub == 0 and (c0 % 128 == 64) == f8
        x = srcs_bytes[si]
        a = torch.zeros((N, H, W, 128), dtype=torch.uint8)
        ys, xs = torch.arange(H) + dy, torch.arange(W) + dx
        vy, vx = (ys >= 0) & (ys < H), (xs >= 0) & (xs < W)
        a[:, vy.nonzero().flatten()[:, None], vx.nonzero().flatten()[None, :]] = \
            x[:, ys[vy][:, None], xs[vx][None, :], 2 * c0:2 * c0 + 128]
        wrow = wp_bytes[:, i * 128:(i + 1) * 128].contiguous()
        if f8:
            out += a.view(torch.float8_e4m3fn).double() @ wrow.view(torch.float8_e4m3fn).double().T
        else:
            out += a.view(torch.float16).double() @ wrow.view(torch.float16).double().T
    return out * 2.0 ** -acc_exp + bias.double()


@pytest.mark.parametrize("halo", [False, True])
def test_f16f8_programs_reproduce_conv(halo):
    """f16f8: an fp16 pass and one fp8 correction pass per logical chunk on the fp16-row / fp8-row physical layout."""
    torch.manual_seed(22)
    ch, N, H, W, k = 64, 1, 7, 9, 3
    x, disp = torch.randn(N, 128, H, W), torch.randn(N, 1, H, W)
    wt, b = torch.randn(8, 129, k, k) / 30, torch.randn(8)
    want = nhwc(F.conv2d(torch.cat((x, disp), 1).double(), wt.double(), b.double(), padding=1))
    taps = torch.zeros(N, H, W, ch)
    pd = F.pad(disp, (1, 1, 1, 1))
    for r in range(3):
        for s_ in range(3):
            taps[..., 3 * r + s_] = pd[:, 0, r:r + H, s_:s_ + W]
    xb, tb = _f16f8_phys(nhwc(x)), _f16f8_phys(taps)
    # shapes seen by the program builder: 2-byte physical channels
    xp, tp = torch.empty(N, H, W, 256, dtype=torch.float16), torch.empty(N, H, W, 128, dtype=torch.float16)
    srcs = [engine.act_src(xp, 0, phys=2), engine.Src(tp, "taps3x3", 128, ch, 0, 1)]
    passes = engine.split_passes(_lib.CNM_F16F8)
    if halo:
        groups, gcols, info = engine.make_groups(srcs, k, ch, 1, split=passes)
        assert all((c0 % 128 == 64) == (i < len(groups) // 2) for i, (m, c0, dx, nt) in enumerate(groups))
        steps = [(m // 4, 0, dx, t - info[m][0], c0) for (m, c0, dx, nt) in groups for t in range(nt)]
    else:
        steps, gcols = engine.make_steps(srcs, k, 1, ch, split=passes)
    assert len(steps) == len(gcols) == 2 * (9 * 2 + 1)
    e = engine.f16f8_weight_exp(wt)
    assert 224.0 < float(wt.abs().max()) * 2.0 ** (e - 7) <= 448.0
    wp = engine.pack_weights(wt, gcols, ch, _lib.CNM_F16F8)
    assert wp.dtype == torch.float16 and wp.shape == (8, len(gcols) * 64)
    got = emulate_f16f8_steps([xb, tb], steps, wp.view(torch.uint8), b, e)
    assert (got - want).abs().max() < 3e-5 * want.abs().max()     # same accuracy class as split bf16


def test_fold_bn_matches_batchnorm():
    sd = synth.kaiming_state(nets.state_shapes([("c", "conv", 8, 4, 3)]), seed=5)
    x = torch.randn(2, 8, 6, 6)
    want = F.batch_norm(F.conv2d(x, sd["c.0.weight"], padding=1), sd["c.1.running_mean"], sd["c.1.running_var"],
                        sd["c.1.weight"], sd["c.1.bias"], False, 0.0, 1e-5)
    wf, bf = engine.fold_bn(sd, "c", 0)
    assert torch.allclose(F.conv2d(x, wf, bf, padding=1), want, atol=1e-5)


def test_round_tf32():
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -10, -3.14159, 1e-30, 65504.0])
    r = engine.round_tf32(x)
    assert r[0] == 1.0 and r[1] == 1.0 + 2 ** -10 and r[2] == 1.0 + 2 ** -10
    assert torch.all((r.view(torch.int32) & 0x1FFF) == 0)
    assert torch.allclose(r, x, rtol=2 ** -11)


def test_module_contract_matches_reference_keys():
    d = cnmnet_b200.depthNet(idepth_scale=3.0)
    r = cnmnet_b200.DepthRefineNet(idepth_scale=3.0)
    for mod, layers in ((d, nets.DEPTHNET_LAYERS), (r, nets.REFINENET_LAYERS)):
        shapes = nets.state_shapes(layers)
        sd = mod.state_dict()
        assert list(sd.keys()) == list(shapes.keys())
        assert all(tuple(sd[k].shape) == tuple(v) for k, v in shapes.items())
    assert sum(p.numel() for p in d.parameters()) == 33898500
    assert sum(p.numel() for p in r.parameters()) == 10776066
    # same init stream as the reference constructor
    torch.manual_seed(0)
    d0 = cnmnet_b200.depthNet(idepth_scale=3.0).state_dict()
    ref = nets.reference_init_state(nets.DEPTHNET_LAYERS, 0)
    assert all(torch.equal(d0[k], ref[k]) for k in ref)
    # DataParallel-prefixed checkpoints load (eval.py:184-198)
    d.load_reference_checkpoint({"module." + k: v for k, v in ref.items()})
    assert torch.equal(d.state_dict()["conv1.0.weight"], ref["conv1.0.weight"])
    # reference defaults (depthNet_model.py:127, 273)
    assert cnmnet_b200.DepthRefineNet().idepth_scale == 2 and cnmnet_b200.depthNet().idepth_scale == 3


def test_errors_without_gpu_are_loud():
    d = cnmnet_b200.depthNet()
    with pytest.raises(RuntimeError):
        d(torch.zeros(1, 3, 32, 32), torch.zeros(1, 3, 32, 32), torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(RuntimeError):
        cnmnet_b200.Depth2normal(9)(torch.ones(1, 8, 8), torch.eye(3)[None])
    with pytest.raises(RuntimeError):
        d.train()
    with pytest.raises(AssertionError):  # reference check_sizes contract (inverse_warp.py:18-24)
        cnmnet_b200.inverse_warp(torch.zeros(1, 2, 4, 4), torch.zeros(1, 4, 4), torch.zeros(1, 3, 5), torch.zeros(1, 3, 3),
                                 torch.zeros(1, 3, 3))


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "cnmnet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cnm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES.keys()), declared ^ set(_lib.SIGNATURES.keys())
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.cnm_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.cnm_version()
    assert _lib.load().cnm_packed_source_bytes(2, 192, 256) == 2 * 195 * 260 * 16


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.ConvSrc) == 40 and ctypes.sizeof(_lib.ConvStep) == 32 and ctypes.sizeof(_lib.ConvDesc) == 112


def test_shard_range_partitions():
    for total in (0, 1, 5, 8, 37):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_video_reference_tuples_follow_the_eval_loops():
    """eval.py:408-415 (3 views), 581-592 (5 views), 817-830 (7 views): source offsets, order and frame stride."""
    from cnmnet_b200 import video
    t3 = video.reference_tuples(100, 3)
    assert t3[0] == (12, (22, 2)) and t3[-1] == (87, (97, 77)) and all(i % 3 == 0 for i, _ in t3)
    assert [i for i, _ in t3] == [i for i in range(10, 90) if i % 3 == 0]          # the reference's own index list
    t5 = video.reference_tuples(100, 5)
    assert t5[0] == (12, (22, 2, 17, 7))
    t7 = video.reference_tuples(100, 7)
    assert t7[0] == (27, (37, 17, 32, 22, 47, 7)) and all(i % 9 == 0 for i, _ in t7)
    assert all(0 <= s < 100 for _, srcs in t5 + t7 for s in srcs)                   # no wrap-around (unlike eval.py:581)
    assert video.reference_tuples(15, 3) == []
    with pytest.raises(ValueError):
        video.reference_tuples(50, offsets=(5,))


def test_video_batches_gather_and_pad():
    from cnmnet_b200 import video
    T, H, W = 40, 4, 6
    frames = torch.arange(T, dtype=torch.float32).view(T, 1, 1, 1).expand(T, 3, H, W).contiguous()
    cams = torch.arange(T, dtype=torch.float32).view(T, 1, 1, 1).expand(T, 2, 4, 4).contiguous() * 10
    vs = video.VideoDepthStream(None, batch=4, views=3)
    idx, gen = vs.batches(frames, cams)
    want = [i for i in range(10, 30) if i % 3 == 0]                                  # 12, 15, ..., 27
    assert [i for g in idx for i in g] == want and [len(g) for g in idx] == [4, 2]
    got = [tuple(t.clone() for t in b) for b in gen]
    assert len(got) == 2
    ref, srcs, rc, sc = got[0]
    assert ref.shape == (4, 3, H, W) and srcs.shape == (4, 2, 3, H, W) and sc.shape == (4, 2, 2, 4, 4)
    assert ref[:, 0, 0, 0].tolist() == [12, 15, 18, 21]
    assert srcs[:, 0, 0, 0, 0].tolist() == [22, 25, 28, 31] and srcs[:, 1, 0, 0, 0].tolist() == [2, 5, 8, 11]
    assert rc[:, 0, 0, 0].tolist() == [120, 150, 180, 210] and sc[:, 1, 0, 0, 0].tolist() == [20, 50, 80, 110]
    ref, srcs, _, _ = got[1]
    assert ref[:, 0, 0, 0].tolist() == [24, 27, 27, 27]                              # padded with the last tuple


def test_correction_steps_come_first():
    """K order of the two-row modes (engine._corrections_first): every correction step before every leading-term step,
    for the per-tap step list and group-wise for the row-halo list; one-row modes keep their order."""
    x = torch.empty(1, 8, 8, 256, dtype=torch.float16)
    t = torch.empty(1, 8, 8, 128, dtype=torch.float16)
    srcs = [engine.act_src(x, 0, phys=2), engine.Src(t, "taps3x3", 128, 64, 0, 1)]
    for passes, main in ((engine.split_passes(_lib.CNM_F16F8), ("hi", "h16")), (engine.split_passes(_lib.CNM_BF16X3), ("hi", "hi"))):
        steps, cols = engine.make_steps(srcs, 3, 1, 64, split=passes)
        is_main = [(c[5], c[4]) == main for c in cols]
        n_main = sum(is_main)
        assert n_main == len(cols) // len(passes) and all(is_main[-n_main:]) and not any(is_main[:-n_main])
        groups, gcols, _ = engine.make_groups(srcs, 3, 64, 1, split=passes)
        g_main = []
        pos = 0
        for g in groups:
            kinds = {(c[5], c[4]) == main for c in gcols[pos:pos + g[3]]}
            assert len(kinds) == 1                                    # a group is all-leading or all-correction
            g_main.append(kinds.pop())
            pos += g[3]
        k = sum(g_main)
        assert pos == len(gcols) and all(g_main[-k:]) and not any(g_main[:-k])
        if passes is engine.split_passes(_lib.CNM_F16F8):
            assert k == len(groups) // 2 and all(c0 % 128 == 64 for (_, c0, _, _) in groups[:k])   # fp8 rows = first half
    steps, cols = engine.make_steps([engine.act_src(torch.empty(1, 8, 8, 128), 0)], 3, 1, 64)
    assert [c[2:4] for c in cols[:2]] == [(0, 0), (0, 0)] and len(cols) == 18   # unchanged: tap-major, chunk inner


def test_plan_arena_recycles_released_buffers(monkeypatch):
    """_Plan.buf / release: a released buffer backs the next request that fits (best fit), `fresh` never reuses, live
    buffers never alias, PLAN_REUSE=0 disables recycling."""
    plan = object.__new__(engine._Plan)
    plan.dtype, plan.device, plan.pf = _lib.CNM_F16F8, torch.device("cpu"), 2
    plan._storage, plan._free, plan.bytes_allocated = {}, [], 0
    a = plan.buf(2, 8, 8, 64)                       # act tensor: 2 physical channels per logical one
    assert a.shape == (2, 8, 8, 128) and a.dtype == torch.float16   # (CUDA allocations are 512-byte aligned; CPU ones only 64)
    b = plan.buf(2, 4, 4, 64)
    first = plan.bytes_allocated
    assert a.data_ptr() != b.data_ptr()
    plan.release(a)
    c = plan.buf(1, 8, 8, 64)                       # fits into a's storage
    assert c.data_ptr() == a.data_ptr() and plan.bytes_allocated == first
    d = plan.buf(1, 8, 8, 64, fresh=True)           # never recycled
    assert d.data_ptr() not in (a.data_ptr(), b.data_ptr()) and plan.bytes_allocated > first
    plan.release(b, c)
    e = plan.buf(2, 2, 2, 12, dtype=torch.float32)  # best fit: the smaller free block (b's)
    assert e.data_ptr() == b.data_ptr() and e.dtype == torch.float32
    plan.release(e)
    plan.release(e)                                 # double release is harmless
    assert len(plan._free) == 2
    monkeypatch.setattr(engine, "PLAN_REUSE", False)
    f = plan.buf(1, 2, 2, 64)
    assert f.data_ptr() not in (a.data_ptr(), b.data_ptr())


@pytest.mark.parametrize("halo", [False, True])
def test_128_plane_conv1_program(halo):
    """depthNet(planes=128): conv1.0 = 7x7 over cat(image [3], cost volume [128]) -- four cost chunks + the row-tap side
    tensor of the image, as a per-tap step list and as row-halo groups."""
    torch.manual_seed(128)
    ch, N, H, W, k, D = 32, 1, 9, 10, 7, 128
    img, cost = torch.randn(N, 3, H, W), torch.randn(N, D, H, W)
    wt, b = torch.randn(8, 3 + D, k, k) / 20, torch.randn(8)
    want = nhwc(F.conv2d(torch.cat((img, cost), 1).double(), engine.round_tf32(wt).double(), b.double(), padding=3))
    taps = torch.zeros(N, H, W, ch)
    pi = F.pad(img, (3, 3, 0, 0))
    for s in range(k):
        for c in range(3):
            taps[..., s * 3 + c] = pi[:, c, :, s:s + W]
    srcs = [engine.act_src(nhwc(cost), 3), engine.Src(taps, "rowtaps", 0, ch, 0, 3)]
    if halo:
        groups, cols, info = engine.make_groups(srcs, k, ch)
        steps = [(m // 4, 0, dx, t - info[m][0], c0) for (m, c0, dx, nt) in groups for t in range(nt)]
    else:
        steps, cols = engine.make_steps(srcs, k, 1, ch)
    assert len(steps) == len(cols) == 49 * (D // ch) + 7
    wp = engine.pack_weights(wt, cols, ch, _lib.CNM_TF32)
    got = emulate_steps([s_.buf for s_ in srcs], steps, wp, b, 1, ch, False)
    assert torch.allclose(got, want, atol=1e-9)


def test_folded_upconv_weights_equal_upsample_then_conv():
    """engine.fold_upconv_weights: Upsample(x2, bilinear, align_corners=False) + conv3x3(pad 1) == the sum over nine
    zero-extended operands (x, border rows / columns, corners) of 3x3 convolutions with the folded weights, phase by phase."""
    import torch.nn.functional as F
    from cnmnet_b200 import engine
    torch.manual_seed(0)
    for h, w in ((5, 6), (2, 2), (3, 7)):
        cout, cin = 3, 2
        wt = torch.randn(cout, cin, 3, 3, dtype=torch.float64)
        x = torch.randn(2, cin, h, w, dtype=torch.float64)
        want = F.conv2d(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), wt, padding=1)
        big = engine.fold_upconv_weights(wt).double().reshape(4, cout, 9, cin, 3, 3)
        srcs = [x]
        for rows, cols in ((slice(0, 1), slice(None)), (slice(h - 1, h), slice(None)), (slice(None), slice(0, 1)),
                           (slice(None), slice(w - 1, w)), (slice(0, 1), slice(0, 1)), (slice(0, 1), slice(w - 1, w)),
                           (slice(h - 1, h), slice(0, 1)), (slice(h - 1, h), slice(w - 1, w))):
            v = torch.zeros_like(x)
            v[:, :, rows, cols] = x[:, :, rows, cols]
            srcs.append(v)
        got = torch.zeros_like(want)
        for ph in range(4):
            got[:, :, (ph >> 1)::2, (ph & 1)::2] = sum(F.conv2d(srcs[s_], big[ph][:, s_], padding=1) for s_ in range(9))
        assert (got - want).abs().max() < 1e-6 * want.abs().max()


def test_dense_rowtaps_source_reproduces_conv1():
    """conv1.0 with the image taps packed three kernel rows per K-step (kind 'rowtaps3': side tensor with one apron row above
    and below, what cnm_pack_row_taps3 writes): per-tap steps and row-halo groups (taps 3 rows apart) == conv2d."""
    torch.manual_seed(14)
    ch, N, H, W, k = 64, 1, 9, 11, 7
    img, cost = torch.randn(N, 3, H, W), torch.randn(N, 64, H, W)
    wt, b = torch.randn(8, 67, k, k), torch.randn(8)
    want = nhwc(F.conv2d(torch.cat((img, cost), 1).double(), engine.round_tf32(wt).double(), b.double(), padding=3))
    taps = torch.zeros(N, H + 2, W, ch)
    pi = F.pad(img, (3, 3, 3, 3))            # image row y at padded row y + 3
    for j in range(H + 2):                   # tensor row j is centred on image row j - 1
        for q in range(3):
            for s in range(k):
                for c in range(3):
                    taps[:, j, :, (q * k + s) * 3 + c] = pi[:, c, j - 2 + q + 3, s:s + W]
    srcs = [engine.act_src(nhwc(cost), 3), engine.Src(taps, "rowtaps3", 0, ch, 0, 3)]
    steps, cols = engine.make_steps(srcs, k, 1, ch)
    assert len(steps) == 49 + 3
    wp = engine.pack_weights(wt, cols, ch, _lib.CNM_TF32)
    got = emulate_steps([s.buf for s in srcs], steps, wp, b, 1, ch, False)
    assert torch.allclose(got, want, atol=1e-9)
    groups, gcols, info = engine.make_groups(srcs, k, ch)
    gsteps = []
    for (m, c0, dx, nt) in groups:
        taps_n, tstride = engine.group_taps(nt)
        gsteps += [(m // 4, 0, dx, t * tstride - info[m][0], c0) for t in range(taps_n)]
    assert len(gsteps) == len(gcols) == 52 and info[4] == [1, 22]
    wp = engine.pack_weights(wt, gcols, ch, _lib.CNM_TF32)
    got = emulate_steps([s.buf for s in srcs], gsteps, wp, b, 1, ch, False)
    assert torch.allclose(got, want, atol=1e-9)


def test_folded_upconv_program_reproduces_operator():
    """engine.FoldedUpconvOp's device program, emulated on the CPU exactly as csrc/conv_halo.cu (FOLD) walks it: per 16x16
    block (pairs: the union class of two blocks), every group whose border class the tile has, its slab = a window of the
    source VIEW (zeros outside the view's extent), `taps` row windows against consecutive weight steps, phase scatter.
    Must equal Upsample(x2, bilinear) + conv3x3 with the (bf16-rounded) folded weights."""
    torch.manual_seed(21)
    N, h, w, cin, cout = 2, 40, 48, 64, 64
    dt = _lib.CNM_BF16
    x = torch.randn(N, h, w, cin).to(torch.bfloat16)
    out = torch.zeros(N, 2 * h, 2 * w, cout, dtype=torch.bfloat16)
    op = engine.FoldedUpconvOp("t", x, cout, N, h, w, dt, "cpu", out)
    wt, b = torch.randn(cout, cin, 3, 3) / 24.0, torch.randn(cout)
    op.load(wt, b)
    v = op.variants[op.backend]
    groups = v["groups_dev"].tolist()
    pads = list(v["pads"])
    wp = op.wp.double()                                  # [4 cout, steps * 64]
    xs = x.double()
    srcs = op.srcs_arr
    pix = x.shape[3] * x.element_size()

    def view_window(m, n, y_lo, x_lo, rows, cols):
        """rows x cols window of source m starting at full-image coordinates (y_lo, x_lo); zeros outside the view."""
        s = srcs[m]
        off = (s.ptr - x.data_ptr()) // pix                # first pixel of the view inside image 0
        vy0, vx0 = off // w, off % w
        assert (vx0, vy0) == (s.x_off, s.y_off) and (m == 0 or (s.pitch_w, s.pitch_hw) == (w, h * w))
        o = torch.zeros(rows, cols, cin, dtype=torch.float64)
        for r in range(rows):
            for c in range(cols):
                vy, vx = y_lo + r - s.y_off, x_lo + c - s.x_off    # coordinates inside the view
                if 0 <= vy < s.H and 0 <= vx < s.W:
                    o[r, c] = xs[n, vy + vy0, vx + vx0]
        return o

    def block_class(ty, tx):
        return (1 if ty == 0 else 0) | (2 if ty * 16 + 16 >= h else 0) | (4 if tx == 0 else 0) | (8 if tx * 16 + 16 >= w else 0)

    tiles_x, tiles_y = (w + 15) // 16, (h + 15) // 16
    got = torch.zeros(N, 2 * h, 2 * w, cout, dtype=torch.float64)
    pairs = op.backend == "halo2f"
    blocks = [(n, ty, tx) for n in range(N) for ty in range(tiles_y) for tx in range(tiles_x)]
    executed = set()
    for bi, (n, ty, tx) in enumerate(blocks):
        cls = block_class(ty, tx)
        if pairs:
            pn, pty, ptx = blocks[bi ^ 1]
            cls |= block_class(pty, ptx)
        y0, x0 = ty * 16, tx * 16
        acc = torch.zeros(16, 16, 4 * cout, dtype=torch.float64)
        for m, c0, dx, nf in groups:
            taps, flags, first = nf & 15, (nf >> 4) & 15, (nf >> 8) & 0xFFFFF
            if flags & ~cls:
                continue
            executed.add(flags)
            pad, rows = pads[2 * m], pads[2 * m + 1]
            slab = view_window(m, n, y0 - pad, x0 + dx, rows, 16)          # one TMA box {64, 16, rows}
            for t in range(taps):
                a = slab[t:t + 16, :, c0:c0 + 64]
                acc += a @ wp[:, (first + t) * 64:(first + t + 1) * 64].T
        for ph in range(4):
            blk = acc[:, :, ph * cout:(ph + 1) * cout] + b.double()
            ys = [2 * (y0 + r) + (ph >> 1) for r in range(16) if y0 + r < h]
            xs_ = [2 * (x0 + c) + (ph & 1) for c in range(16) if x0 + c < w]
            got[n][ys[0]:ys[-1] + 1:2, xs_[0]:xs_[-1] + 1:2] = blk[:len(ys), :len(xs_)]
    assert executed == {0, 1, 2, 4, 8, 5, 9, 6, 10}          # every border class ran somewhere
    big = engine.fold_upconv_weights(wt).to(torch.bfloat16).double()       # what pack_weights stored
    xn = xs.permute(0, 3, 1, 2)
    want = torch.zeros(N, cout, 2 * h, 2 * w, dtype=torch.float64)
    zs = [xn]
    for rows, cols in ((slice(0, 1), slice(None)), (slice(h - 1, h), slice(None)), (slice(None), slice(0, 1)),
                       (slice(None), slice(w - 1, w)), (slice(0, 1), slice(0, 1)), (slice(0, 1), slice(w - 1, w)),
                       (slice(h - 1, h), slice(0, 1)), (slice(h - 1, h), slice(w - 1, w))):
        z = torch.zeros_like(xn)
        z[:, :, rows, cols] = xn[:, :, rows, cols]
        zs.append(z)
    for ph in range(4):
        y = sum(F.conv2d(zs[s_], big[ph * cout:(ph + 1) * cout, s_ * cin:(s_ + 1) * cin], padding=1) for s_ in range(9))
        want[:, :, (ph >> 1)::2, (ph & 1)::2] = y + b.double().view(1, -1, 1, 1)
    assert torch.allclose(got.permute(0, 3, 1, 2), want, atol=1e-9)
    # ... and that is the operator: Upsample + conv3x3 with the unrounded weights, to bf16 weight rounding
    exact = F.conv2d(F.interpolate(xn, scale_factor=2, mode="bilinear", align_corners=False), wt.double(), b.double(), padding=1)
    assert (want - exact).abs().max() < 2e-2 * exact.abs().max()


def _balanced_segments(tiles, S, G):
    """Python mirror of TcWalk / HaloWalk (csrc/conv_tc.cu, conv_halo.cu) in balanced mode: CTA c -> [(tile, k0, k1)]."""
    T = tiles * S
    out = []
    for c in range(G):
        pos, end, segs = T * c // G, T * (c + 1) // G, []
        while pos < end:
            tile = pos // S
            k0 = pos - tile * S
            k1 = min(S, k0 + end - pos)
            segs.append((tile, k0, k1))
            pos += k1 - k0
        out.append(segs)
    return out


@pytest.mark.parametrize("tiles,S,G", [(120, 144, 148), (240, 72, 74), (960, 98, 74), (20, 144, 148), (74, 36, 74), (1, 300, 148)])
def test_balanced_schedule_partition(tiles, S, G):
    """The stream-K partition the balanced kernels walk: every (tile, K-step) exactly once; a CTA has at most one contributor
    piece (k0 > 0) and it comes first, at most one head piece (k0 == 0, k1 < S) and it comes last; the contributors of a
    tile are the CTAs right after its head CTA, in order -- what the head CTA's epilogue assumes when it gathers."""
    segs = _balanced_segments(tiles, S, G)
    cover = {}
    for c, ss in enumerate(segs):
        kinds = ["contrib" if k0 > 0 else ("head" if k1 < S else "whole") for (_, k0, k1) in ss]
        assert kinds.count("contrib") <= 1 and kinds.count("head") <= 1
        if "contrib" in kinds:
            assert kinds[0] == "contrib"
        if "head" in kinds:
            assert kinds[-1] == "head"
        for (t, k0, k1) in ss:
            assert 0 <= k0 < k1 <= S
            cover.setdefault(t, []).append((k0, k1, c))
    assert sorted(cover) == list(range(tiles))
    T = tiles * S
    for t, pieces in cover.items():
        pieces.sort()
        assert pieces[0][0] == 0 and pieces[-1][1] == S and all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
        head = pieces[0][2]
        assert [p_[2] for p_ in pieces] == list(range(head, head + len(pieces)))
        # the kernel's contributor count: CTAs after the head whose range starts before the tile ends
        n = 0
        while T * (head + 1 + n) // G < (t + 1) * S:
            n += 1
        assert n == len(pieces) - 1
