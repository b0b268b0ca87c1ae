"""GPU parity: implicit-GEMM convolution (tcgen05 kernel and its SIMT twin) vs torch CPU conv2d in fp64
on the same quantised operands, through the C ABI."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cnmnet_b200 import _lib, engine
from gpu_util import dev, e4m3, f16f8_parts, from_act, quantize, rel_l2, to_act

pytestmark = pytest.mark.gpu

DTYPES = [_lib.CNM_BF16, _lib.CNM_TF32, _lib.CNM_BF16X3, _lib.CNM_F16F8]
# operands identical to the oracle's; what differs is fp32 accumulation order and the output rounding of the mode
MAX_TOL = {_lib.CNM_BF16: 6e-3, _lib.CNM_TF32: 1e-3, _lib.CNM_BF16X3: 5e-5, _lib.CNM_F16F8: 6e-5}
L2_TOL = {_lib.CNM_BF16: 3e-3, _lib.CNM_TF32: 3e-4, _lib.CNM_BF16X3: 2e-5, _lib.CNM_F16F8: 2e-5}


def run_conv(backend, dtype, xs, wchs, w, b, k, stride, relu=True):
    N, _, H, W = xs[0].shape
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    acts = [to_act(x, dtype) for x in xs]
    srcs = [engine.act_src(a, wch, phys=engine.phys_factor(dtype)) for a, wch in zip(acts, wchs)]
    op = engine.ConvOp("t", srcs, k, stride, w.shape[0], N, Ho, Wo, dtype, dev(), relu, backend=backend)
    op.load(w, b)
    op(_lib.stream_ptr())
    torch.cuda.synchronize()
    return from_act(op.out, dtype)


def oracle_conv(dtype, xs, w, b, k, stride, relu=True):
    if dtype == _lib.CNM_F16F8:
        # the two tensor-core passes on the stored parts: hi * W16 + [lo8 | hi8] * [Whi8 | Wlo8], scaled by 2^-e
        hi, lo8, hi8 = (t.double() for t in f16f8_parts(torch.cat(xs, 1)))
        e = engine.f16f8_weight_exp(w)
        ws = w * (2.0 ** e)
        w16 = ws.to(torch.float16).float()
        whi8, wlo8 = e4m3(ws / engine.F8_LO_SCALE), e4m3((ws - w16) / engine.F8_HI_SCALE)
        conv = lambda a, ww: F.conv2d(a, ww.double(), None, stride=stride, padding=(k - 1) // 2)  # noqa: E731
        y = (conv(hi, w16) + conv(lo8, whi8) + conv(hi8, wlo8)) * (2.0 ** -e) + b.double().view(1, -1, 1, 1)
        return (F.relu(y) if relu else y).float()
    x = torch.cat([quantize(t, dtype) for t in xs], 1).double()
    wq = quantize(w, dtype).double()
    y = F.conv2d(x, wq, b.double(), stride=stride, padding=(k - 1) // 2)
    return (F.relu(y) if relu else y).float()


CASES = [
    # (N, H, W, [Cin per source], Cout, k, stride)
    (1, 8, 16, [64], 64, 1, 1),       # plain GEMM, one tile
    (1, 8, 16, [64], 64, 3, 1),
    (2, 16, 32, [128], 128, 3, 1),
    (2, 16, 32, [128], 128, 3, 2),
    (1, 24, 32, [64], 128, 7, 1),
    (1, 24, 32, [64], 64, 7, 2),
    (1, 16, 16, [64], 256, 5, 1),
    (1, 16, 16, [64], 64, 5, 2),
    (3, 6, 8, [128], 512, 3, 1),      # tiny layer: batch packed into the pixel tile
    (5, 12, 16, [64, 64], 256, 3, 1), # concatenated sources, ragged batch tile
    (1, 20, 40, [64], 64, 3, 1),      # partial tiles (640x480 pyramid sizes)
    (2, 96, 128, [64], 128, 3, 1),    # enough tiles for BN=128 and several tiles per CTA
    (1, 48, 64, [128], 512, 3, 1),    # two BN=256 column tiles
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("backend", ["simt", "tc"])
def test_conv_matches_oracle(backend, dtype, case):
    N, H, W, cins, cout, k, stride = case
    g = torch.Generator().manual_seed(N * 1000 + H * 7 + W + cout + k * 13 + stride)
    xs = [torch.randn(N, c, H, W, generator=g) for c in cins]
    w = torch.randn(cout, sum(cins), k, k, generator=g) / np.sqrt(sum(cins) * k * k)
    b = torch.randn(cout, generator=g)
    wchs = [sum(cins[:i]) for i in range(len(cins))]
    got = run_conv(backend, dtype, xs, wchs, w, b, k, stride)
    want = oracle_conv(dtype, xs, w, b, k, stride)
    assert got.shape == want.shape
    # operands are identical; only fp32 accumulation order and the output rounding differ
    err = (got - quantize(want, dtype)).abs().max().item()
    assert err < MAX_TOL[dtype] * max(1.0, want.abs().max().item()), (err, rel_l2(got, want))
    assert rel_l2(got, want) < L2_TOL[dtype]


HALO_CASES = [
    (1, 16, 16, [64], 64, 3, 1),
    (2, 32, 48, [128], 128, 3, 1),
    (1, 24, 40, [64], 128, 7, 1),      # partial tiles in both directions, deepest halo
    (1, 32, 32, [64], 256, 5, 1),      # BN = 128 (too few tiles for 256)
    (2, 16, 32, [64, 64], 256, 3, 1),  # concatenated sources
    (3, 48, 64, [64], 512, 3, 1),
    (16, 48, 64, [64], 256, 3, 1),     # >= 120 tiles of BN = 256: single-buffered accumulators, several tiles per CTA
    (10, 192, 256, [64], 64, 3, 1),    # C4-sized Cout = 64 layer: many tiles per CTA
    (2, 32, 64, [64], 128, 3, 2),      # stride 2: slabs of the four parity sub-images
    (1, 64, 96, [128], 128, 7, 2),
    (2, 40, 72, [64], 256, 5, 2),      # partial tiles, odd sub-image extents
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", HALO_CASES)
@pytest.mark.parametrize("backend", ["halo", "halo2", "halo2h"])
def test_halo_conv_matches_oracle(backend, dtype, case):
    N, H, W, cins, cout, k, stride = case
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    if backend == "halo2" and (N * ((Ho + 15) // 16) * ((Wo + 15) // 16)) % 2:
        pytest.skip("the CTA-pair kernel needs an even number of 16x16 pixel blocks")
    if backend == "halo2h" and (dtype == _lib.CNM_TF32 or (N * ((Ho + 7) // 8) * ((Wo + 15) // 16)) % 2):
        pytest.skip("16x8 blocks: 2-byte dtypes, even number of blocks")
    g = torch.Generator().manual_seed(N * 1000 + H * 7 + W + cout + k * 13)
    xs = [torch.randn(N, c, H, W, generator=g) for c in cins]
    w = torch.randn(cout, sum(cins), k, k, generator=g) / np.sqrt(sum(cins) * k * k)
    b = torch.randn(cout, generator=g)
    wchs = [sum(cins[:i]) for i in range(len(cins))]
    got = run_conv(backend, dtype, xs, wchs, w, b, k, stride)
    want = oracle_conv(dtype, xs, w, b, k, stride)
    err = (got - quantize(want, dtype)).abs().max().item()
    assert err < MAX_TOL[dtype] * max(1.0, want.abs().max().item()), (err, rel_l2(got, want))
    assert rel_l2(got, want) < L2_TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv_no_relu_negative_outputs(dtype):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 64, 16, 16, generator=g)
    w = torch.randn(64, 64, 3, 3, generator=g) / 24.0
    b = torch.randn(64, generator=g)
    got = run_conv("tc", dtype, [x], [0], w, b, 3, 1, relu=False)
    want = oracle_conv(dtype, [x], w, b, 3, 1, relu=False)
    assert (want < 0).any() and rel_l2(got, want) < L2_TOL[dtype]


@pytest.mark.parametrize("backend", ["tc", "halo2"])
def test_f16f8_accuracy_class(backend):
    """CNM_F16F8 against the UNQUANTISED fp64 convolution: the fp16 + fp8-correction scheme must stay in the accuracy
    class of CNM_BF16X3 (operand error ~2^-17), two orders below single-pass bf16."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 128, 32, 32, generator=g).abs()
    w = torch.randn(128, 128, 3, 3, generator=g) / np.sqrt(128 * 9)
    b = torch.randn(128, generator=g)
    want = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1)).float()
    errs = {}
    for dtype in (_lib.CNM_F16F8, _lib.CNM_BF16X3, _lib.CNM_BF16):
        errs[dtype] = rel_l2(run_conv(backend, dtype, [x], [0], w, b, 3, 1), want)
    assert errs[_lib.CNM_F16F8] < 3e-5 and errs[_lib.CNM_F16F8] < errs[_lib.CNM_BF16] / 50, errs


@pytest.mark.parametrize("dtype", [_lib.CNM_BF16, _lib.CNM_BF16X3, _lib.CNM_F16F8])
@pytest.mark.parametrize("backend,cout,bn", [("tc", 64, 0), ("halo", 64, 0), ("halo2", 64, 0), ("halo2", 128, 128),
                                             ("halo", 256, 128), ("halo2", 256, 256), ("tc", 256, 128), ("tc", 512, 256),
                                             ("tc", 128, 64)])
def test_fused_head_matches_separate_head(backend, cout, bn, dtype):
    """cnm_conv_desc.head_w / head_partial + cnm_head_finish: the 1-channel 3x3 head computed from the fp32 accumulators in
    the convolution epilogue (one partial slice per 64-channel chunk, any output-channel tile) equals conv3x3(+bias) ->
    scale * sigmoid on the (unrounded) layer output; partial tiles."""
    N, H, W, cin, k = 2, 24, 40, 64, 3
    g = torch.Generator().manual_seed(5 + cout)
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / np.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g)
    hw = torch.randn(1, cout, 3, 3, generator=g) * 0.05
    act = to_act(x, dtype)
    op = engine.ConvOp("t", [engine.act_src(act, 0, phys=engine.phys_factor(dtype))], k, 1, cout, N, H, W, dtype, dev(), True,
                       backend=backend)
    op.tunable = False
    op.select(backend, bn)
    wt = hw[0].permute(1, 2, 0).reshape(9, cout).contiguous().to(dev())
    S = cout // 64
    partial = torch.full((S * N, H, W, 12), float("nan"), device=dev())
    op.fuse_head(wt, partial)
    op.load(w, b)
    op(_lib.stream_ptr())
    out = torch.empty((N, 1, H, W), device=dev())
    logits = torch.empty((N, 1, H, W), device=dev())
    lib = _lib.load()
    _lib.check(lib.cnm_head_finish(partial.data_ptr(), S, 0.25, 3.0, out.data_ptr(), logits.data_ptr(), N, H, W,
                                   _lib.stream_ptr()))
    torch.cuda.synchronize()
    want_act = oracle_conv(dtype, [x], w, b, k, 1)                         # fp32 layer output before storage rounding
    assert rel_l2(from_act(op.out, dtype), quantize(want_act, dtype)) < L2_TOL[dtype]   # the act output is still written
    wl = F.conv2d(want_act.double(), hw.double(), torch.tensor([0.25], dtype=torch.float64), padding=1)
    assert not torch.isnan(partial).any()
    assert (logits.cpu() - wl.float()).abs().max() < 2e-5 * max(1.0, wl.abs().max().item())
    assert (out.cpu() - 3.0 * torch.sigmoid(wl).float()).abs().max() < 3e-5   # fp32 sums over up to 8 slices of 576 terms


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case,ksplit", [((3, 6, 8, [128], 512, 3, 1), 4), ((1, 12, 16, [64, 64], 256, 3, 1), 3),
                                         ((2, 24, 32, [128], 128, 3, 2), 2), ((5, 24, 32, [512], 256, 3, 1), 8)])
def test_split_k_matches_oracle(dtype, case, ksplit):
    """cnm_conv_desc.ksplit: K ranges on different CTAs + the fixed-order finishing kernel; deterministic across runs."""
    N, H, W, cins, cout, k, stride = case
    g = torch.Generator().manual_seed(cout + ksplit)
    xs = [torch.randn(N, c, H, W, generator=g) for c in cins]
    w = torch.randn(cout, sum(cins), k, k, generator=g) / np.sqrt(sum(cins) * k * k)
    b = torch.randn(cout, generator=g)
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    acts = [to_act(x, dtype) for x in xs]
    srcs = [engine.act_src(a, sum(cins[:i]), phys=engine.phys_factor(dtype)) for i, a in enumerate(acts)]
    op = engine.ConvOp("t", srcs, k, stride, cout, N, Ho, Wo, dtype, dev(), True, backend="tc")
    op.tunable = False
    op.load(w, b)
    op.select("tc", 128 if cout % 128 == 0 else 64, ksplit)
    outs = []
    for _ in range(2):
        op.out.zero_()
        op(_lib.stream_ptr())
        torch.cuda.synchronize()
        outs.append(from_act(op.out, dtype))
    assert torch.equal(outs[0], outs[1])
    want = oracle_conv(dtype, xs, w, b, k, stride)
    err = (outs[0] - quantize(want, dtype)).abs().max().item()
    assert err < MAX_TOL[dtype] * max(1.0, want.abs().max().item()), (err, rel_l2(outs[0], want))
    assert rel_l2(outs[0], want) < L2_TOL[dtype]


MC_DTYPES = [d for d in DTYPES if d != _lib.CNM_TF32]


@pytest.mark.parametrize("dtype", MC_DTYPES)
@pytest.mark.parametrize("case,bn", [((5, 48, 64, [128], 256, 3, 1), 256),    # refine 48x64 level: 60 pairs, one per CTA pair
                                     ((10, 48, 64, [64], 256, 3, 1), 256),    # 120 pairs: two accumulator stages in use
                                     ((5, 24, 32, [256], 512, 3, 1), 128),
                                     ((2, 40, 72, [64], 256, 5, 2), 256),     # stride 2, partial blocks
                                     ((10, 192, 256, [64], 64, 3, 1), 64)])
def test_half_block_halo_matches_oracle_and_full_blocks(dtype, case, bn):
    """cnm_conv_desc.tile_h = 8 (16x8 blocks on CTA pairs) against the oracle, and bit-identical to the 16x16-block
    CTA-pair kernel where that runs: a pixel's K-steps come in the same order in both."""
    N, H, W, cins, cout, k, stride = case
    g = torch.Generator().manual_seed(cout + bn + N)
    xs = [torch.randn(N, c, H, W, generator=g) for c in cins]
    w = torch.randn(cout, sum(cins), k, k, generator=g) / np.sqrt(sum(cins) * k * k)
    b = torch.randn(cout, generator=g)
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    acts = [to_act(x, dtype) for x in xs]
    srcs = [engine.act_src(a, sum(cins[:i]), phys=engine.phys_factor(dtype)) for i, a in enumerate(acts)]
    op = engine.ConvOp("t", srcs, k, stride, cout, N, Ho, Wo, dtype, dev(), True, backend="halo2h")
    op.tunable = False
    op.load(w, b)
    outs = {}
    for be in ("halo2h", "halo2"):
        if be not in op.variants:
            continue
        op.select(be, bn)
        op.out.fill_(7)
        op(_lib.stream_ptr())
        torch.cuda.synchronize()
        outs[be] = op.out.clone().view(torch.uint8)
        assert rel_l2(from_act(op.out, dtype), oracle_conv(dtype, xs, w, b, k, stride)) < L2_TOL[dtype]
    if "halo2" in outs:
        diff = int((outs["halo2h"] != outs["halo2"]).sum())
        assert diff == 0, "%d of %d output bytes differ between 16x8 and 16x16 blocks" % (diff, outs["halo2"].numel())


@pytest.mark.parametrize("dtype", MC_DTYPES)
@pytest.mark.parametrize("case,bn,ksplit", [
    ((5, 24, 32, [256], 512, 3, 1), 256, 1),      # the 24x32 level: 30 pixel tiles x 2 column tiles
    ((5, 48, 64, [128], 256, 3, 1), 256, 1),      # the 48x64 level of the refine network: 120 tiles
    ((3, 24, 32, [128], 256, 3, 2), 128, 1),      # stride 2 (parity sub-images), 5 pixel tiles: dummy tiles in the cluster
    ((5, 12, 16, [64, 64], 256, 3, 1), 128, 2),   # ragged batch tile + split-K
    ((2, 96, 128, [64], 128, 3, 1), 128, 1),      # more work items than clusters: several per CTA
    ((1, 8, 16, [64], 128, 1, 1), 128, 1),        # one pixel tile: three of four CTAs of a cluster are dummies
])
def test_multicast_clusters_are_bit_identical_to_single_ctas(dtype, case, bn, ksplit):
    """cnm_conv_desc.cluster = 2 / 4: the weight tiles arrive by TMA multicast from the cluster's CTAs; every tile still
    accumulates its K-steps in the same order, so the output bits are those of the single-CTA launch."""
    N, H, W, cins, cout, k, stride = case
    g = torch.Generator().manual_seed(cout + bn + ksplit)
    xs = [torch.randn(N, c, H, W, generator=g) for c in cins]
    w = torch.randn(cout, sum(cins), k, k, generator=g) / np.sqrt(sum(cins) * k * k)
    b = torch.randn(cout, generator=g)
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    acts = [to_act(x, dtype) for x in xs]
    srcs = [engine.act_src(a, sum(cins[:i]), phys=engine.phys_factor(dtype)) for i, a in enumerate(acts)]
    op = engine.ConvOp("t", srcs, k, stride, cout, N, Ho, Wo, dtype, dev(), True, backend="tc")
    op.tunable = False
    op.load(w, b)
    outs = {}
    for be in ("tc", "tcm2", "tcm4"):
        assert be in op.variants
        op.select(be, bn, ksplit)
        assert op.desc.cluster == {"tc": 0, "tcm2": 2, "tcm4": 4}[be]
        op.out.fill_(7)
        op(_lib.stream_ptr())
        torch.cuda.synchronize()
        outs[be] = op.out.clone().view(torch.uint8)
    want = oracle_conv(dtype, xs, w, b, k, stride)
    assert rel_l2(from_act(op.out, dtype), want) < L2_TOL[dtype]
    for be in ("tcm2", "tcm4"):
        diff = int((outs[be] != outs["tc"]).sum())
        assert diff == 0, "%s: %d of %d output bytes differ from the single-CTA kernel" % (be, diff, outs["tc"].numel())


@pytest.mark.parametrize("dtype", MC_DTYPES)
@pytest.mark.parametrize("case,bn", [
    ((5, 48, 64, [128], 256, 3, 1), 256),          # 120 tiles: every CTA holds 0.81 of a tile (head + contributor pieces)
    ((5, 48, 64, [128, 128], 256, 3, 1), 128),     # 240 tiles on 148 CTAs, two sources
    ((10, 96, 128, [64], 128, 3, 1), 128),         # 240 pair tiles on 74 pairs: whole tiles between the two pieces
    ((5, 96, 128, [64, 64], 128, 3, 1), 64),       # column tile 64, two column tiles per pixel block
    ((2, 40, 72, [64], 128, 5, 2), 128),           # stride 2 (parity sub-images), partial blocks
    ((10, 12, 16, [128], 256, 3, 1), 256),         # 20 tiles: up to eight contributors per tile
    ((10, 48, 64, [64, 64], 256, 3, 1), 128),      # + a 1-channel side tensor below: groups of unequal tap counts
])
def test_balanced_schedule_matches_static_schedule(dtype, case, bn):
    """cnm_conv_desc.balance (stream-K): equal K ranges per SM, contributor pieces summed in a fixed order by the tile's
    head CTA.  Against the oracle, against the static schedule of the same kernel (same products, another summation tree:
    equal to fp32 rounding of the accumulator), bit-identical across runs, and the flags end up zeroed."""
    N, H, W, cins, cout, k, stride = case
    g = torch.Generator().manual_seed(cout + bn + N + H)
    xs = [torch.randn(N, c, H, W, generator=g) for c in cins]
    w = torch.randn(cout, sum(cins), k, k, generator=g) / np.sqrt(sum(cins) * k * k)
    b = torch.randn(cout, generator=g)
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    acts = [to_act(x, dtype) for x in xs]
    srcs = [engine.act_src(a, sum(cins[:i]), phys=engine.phys_factor(dtype)) for i, a in enumerate(acts)]
    op = engine.ConvOp("t", srcs, k, stride, cout, N, Ho, Wo, dtype, dev(), True, backend="tc")
    op.tunable = False
    op.load(w, b)
    want = oracle_conv(dtype, xs, w, b, k, stride)
    ran = 0
    for static, bal in (("tc", "tcb"), ("halo", "halob"), ("halo2", "halo2b")):
        if bal not in op.variants or (bn > 128 and bal != "tcb"):
            continue
        op.select(static, bn)
        op.out.fill_(7)
        op(_lib.stream_ptr())
        torch.cuda.synchronize()
        ref = from_act(op.out, dtype)
        outs = []
        for _ in range(3):
            op.select(bal, bn)
            assert op.desc.balance == 1
            op.out.fill_(7)
            op(_lib.stream_ptr())
            torch.cuda.synchronize()
            outs.append(op.out.clone().view(torch.uint8))
        ran += 1
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), bal
        got = from_act(op.out, dtype)
        assert rel_l2(got, want) < L2_TOL[dtype], bal
        scale = max(1.0, want.abs().max().item())
        assert (got - ref).abs().max().item() <= MAX_TOL[dtype] * scale, bal   # one storage-rounding flip at most
        ws = engine.balance_workspace(dev())
        assert int(ws[:256].view(torch.int32).abs().sum()) == 0, "%s left flags set" % bal
    assert ran >= 1


def test_balanced_schedule_with_fused_head_and_side_tensor():
    """The balanced schedule on a layer of the iconv kind: two act sources + a 9-tap side tensor (a group with one tap) and
    a fused 1-channel head computed from the summed accumulators."""
    # (the side tensor case is the last one of test_balanced_schedule_matches_static_schedule's list)
    dtype = _lib.CNM_F16F8
    N, H, W, cin, cout, k = 5, 48, 64, 64, 64, 3
    g = torch.Generator().manual_seed(77)
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / np.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g)
    hw = torch.randn(1, cout, 3, 3, generator=g) * 0.05
    act = to_act(x, dtype)
    wt = hw[0].permute(1, 2, 0).reshape(9, cout).contiguous().to(dev())
    res = {}
    for be in ("halo2", "halo2b", "tc", "tcb"):
        op = engine.ConvOp("t", [engine.act_src(act, 0, phys=engine.phys_factor(dtype))], k, 1, cout, N, H, W, dtype, dev(),
                           True, backend=be)
        op.tunable = False
        op.select(be, 64)
        partial = torch.full((N, H, W, 12), float("nan"), device=dev())
        op.fuse_head(wt, partial)
        op.load(w, b)
        op(_lib.stream_ptr())
        torch.cuda.synchronize()
        assert not torch.isnan(partial).any()
        res[be] = (from_act(op.out, dtype), partial.clone())
    for a, bb in (("halo2", "halo2b"), ("tc", "tcb")):
        assert (res[a][0] - res[bb][0]).abs().max().item() < 3e-4   # a flip of the fp8 residual byte at most
        assert (res[a][1] - res[bb][1]).abs().max().item() < 1e-5


FOLD_CASES = [
    # (N, h, w, Cin, Cout) of the LOW-RES input; output [N, 2h, 2w, Cout]
    (2, 32, 48, 64, 64),      # 2 x 3 blocks: every tile is a border tile, corners included
    (2, 96, 128, 128, 64),    # the C4 upconv1 shape (fewer images): interior tiles skip the border groups
    (1, 48, 48, 64, 128),     # 9 blocks: odd count -> single-CTA kernel (halof); two phases per column tile
    (2, 40, 56, 64, 64),      # partial blocks on the bottom / right edges
    (1, 16, 16, 64, 64),      # one block holding all four corners
]


def _zero_extended_views(x):
    """The nine operands of a folded upconv as full-size tensors: x, then its first / last row, first / last column and
    four corner pixels with zeros everywhere else (what the kernel reads through the views' TMA bounds)."""
    h, w = x.shape[2:]
    out = [x]
    for rows, cols in ((slice(0, 1), slice(None)), (slice(h - 1, h), slice(None)), (slice(None), slice(0, 1)),
                       (slice(None), slice(w - 1, w)), (slice(0, 1), slice(0, 1)), (slice(0, 1), slice(w - 1, w)),
                       (slice(h - 1, h), slice(0, 1)), (slice(h - 1, h), slice(w - 1, w))):
        v = torch.zeros_like(x)
        v[:, :, rows, cols] = x[:, :, rows, cols]
        out.append(v)
    return out


@pytest.mark.parametrize("dtype", MC_DTYPES)
@pytest.mark.parametrize("case", FOLD_CASES)
def test_folded_upconv_matches_oracle(dtype, case):
    """cnm_conv_desc.fold (engine.FoldedUpconvOp): Upsample(x2, bilinear) + conv3x3 + bias + ReLU as one convolution on the
    low-res tensor.  (1) Kernel mechanics -- views, border classes, phase stores -- against the fp64 convolution of the same
    quantised operands (the nine zero-extended sources x the folded weights), at the tolerance of every other convolution
    test; (2) the operator itself against F.interpolate + F.conv2d in fp64 with the unrounded weights."""
    N, h, w, cin, cout = case
    g = torch.Generator().manual_seed(N + h * 3 + w + cin + cout)
    x = torch.randn(N, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / np.sqrt(cin * 9)
    b = torch.randn(cout, generator=g)
    act = to_act(x, dtype)
    pf = engine.phys_factor(dtype)
    out = torch.full((N, 2 * h, 2 * w, cout * pf), 7, dtype=engine.TORCH_DTYPE[dtype], device=dev())
    op = engine.FoldedUpconvOp("t", act, cout, N, h, w, dtype, dev(), out)
    assert op.backend == ("halo2f" if (N * ((h + 15) // 16) * ((w + 15) // 16)) % 2 == 0 else "halof")
    op.load(wt, b)
    op(_lib.stream_ptr())
    torch.cuda.synchronize()
    got = from_act(out, dtype)
    big = engine.fold_upconv_weights(wt)
    y = oracle_conv(dtype, _zero_extended_views(x), big, b.repeat(4), 3, 1)          # [N, 4 cout, h, w], phase-major
    want = torch.empty(N, cout, 2 * h, 2 * w)
    for ph in range(4):
        want[:, :, (ph >> 1)::2, (ph & 1)::2] = y[:, ph * cout:(ph + 1) * cout]
    err = (got - quantize(want, dtype)).abs()
    assert rel_l2(got, want) < L2_TOL[dtype], (rel_l2(got, want), np.unravel_index(int(err.argmax()), err.shape))
    assert err.max().item() < MAX_TOL[dtype] * max(1.0, want.abs().max().item())
    xq = quantize(x, dtype).double()
    exact = F.relu(F.conv2d(F.interpolate(xq, scale_factor=2, mode="bilinear", align_corners=False), wt.double(), b.double(),
                            padding=1)).float()
    assert rel_l2(got, exact) < {_lib.CNM_BF16: 6e-3, _lib.CNM_BF16X3: 5e-5, _lib.CNM_F16F8: 5e-5}[dtype]
    # the border ring and the interior separately (a wrong border class would hide in a global norm)
    ring = torch.ones(2 * h, 2 * w, dtype=torch.bool)
    ring[2:-2, 2:-2] = False
    assert rel_l2(got[..., ring], exact[..., ring]) < {_lib.CNM_BF16: 8e-3, _lib.CNM_BF16X3: 8e-5, _lib.CNM_F16F8: 8e-5}[dtype]


@pytest.mark.parametrize("dtype", MC_DTYPES)
@pytest.mark.parametrize("backend,case,bn", [
    ("halo", (2, 32, 48, [64], 64), 64),            # iconv1 kind: one source
    ("halo2", (2, 32, 48, [64, 64], 128), 128),     # iconv2 kind: two sources, CTA pairs
    ("halo2", (2, 32, 48, [64], 64), 64),
    ("halo", (1, 40, 56, [64, 64], 256), 0),        # library-chosen tile (256 -> 128 for side layers), partial blocks
    ("tc", (3, 8, 12, [64], 128), 128),             # small map: per-tap kernel, images packed into one tile
    ("tc", (2, 32, 48, [64], 256), 256),
])
def test_side_disparity_input_matches_oracle(dtype, backend, case, bn):
    """cnm_conv_desc.side_disp / side_w (ConvOp.attach_side_disp): the nearest-upsampled 1-channel disparity as an input
    channel of a 3x3 layer, added in the epilogue in fp32 -- against the fp64 convolution of the quantised operands plus the
    fp64 convolution of the unrounded upsampled disparity, before bias-ReLU."""
    N, H, W, cins, cout = case
    g = torch.Generator().manual_seed(N + H + W + cout + sum(cins))
    xs = [torch.randn(N, c, H, W, generator=g) for c in cins]
    disp = torch.rand(N, 1, H // 2, W // 2, generator=g) * 3.0
    ctot = sum(cins)
    w = torch.randn(cout, ctot + 1, 3, 3, generator=g) / np.sqrt((ctot + 1) * 9)
    w[:, ctot] *= 0.5                                     # the largest weight (-> f16f8 scale) stays among the act channels
    b = torch.randn(cout, generator=g)
    acts = [to_act(x, dtype) for x in xs]
    srcs = [engine.act_src(a, sum(cins[:i]), phys=engine.phys_factor(dtype)) for i, a in enumerate(acts)]
    op = engine.ConvOp("t", srcs, 3, 1, cout, N, H, W, dtype, dev(), True, backend=backend)
    op.tunable = False
    dd = disp.to(dev())
    op.attach_side_disp(dd, ctot)
    op.select(backend, bn)
    op.load(w, b)
    op.out.fill_(7)
    op(_lib.stream_ptr())
    torch.cuda.synchronize()
    got = from_act(op.out, dtype)
    # oracle: the layer without the disparity channel, pre-activation, on the quantised operands of the mode ...
    wmain = w.clone()
    wmain[:, ctot] = 0                                    # keeps the mode's weight scale (acc_exp) identical
    pre = oracle_conv(dtype, xs + [torch.zeros(N, 1, H, W)], wmain, b, 3, 1, relu=False).double()
    # ... plus the disparity channel exactly
    ud = F.interpolate(disp.double(), scale_factor=2, mode="nearest")
    want = F.relu(pre + F.conv2d(ud, w[:, ctot:ctot + 1].double(), padding=1)).float()
    assert rel_l2(got, want) < L2_TOL[dtype], rel_l2(got, want)
    assert (got - quantize(want, dtype)).abs().max().item() < MAX_TOL[dtype] * max(1.0, want.abs().max().item())


def test_conv_rejects_bad_arguments():
    lib = _lib.load()
    a = torch.zeros((1, 8, 8, 48), dtype=torch.bfloat16, device=dev())  # 48 channels: not chunk aligned
    desc = _lib.ConvDesc(1, 8, 8, 64, 64, 0, 1, 1, 1, 9, _lib.CNM_BF16, 0)
    srcs = (_lib.ConvSrc * 1)(_lib.ConvSrc(a.data_ptr(), 8, 8, 48, 0))
    import ctypes as C
    rc = lib.cnm_conv2d_tc(C.byref(desc), srcs, a.data_ptr(), a.data_ptr(), None, a.data_ptr(), None)
    assert rc == -2 and b"channels" in lib.cnm_last_error()
    rc = lib.cnm_conv2d_tc(None, srcs, a.data_ptr(), a.data_ptr(), None, a.data_ptr(), None)
    assert rc == -1
