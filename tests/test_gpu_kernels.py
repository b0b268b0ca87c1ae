"""GPU parity of the non-conv kernels through the C ABI / the reference-shaped Python API:
cost volume (K1), depth->normal (K3), CNM (K3b), refine aggregation (K4), upsample, heads, packers,
inverse warp.  Oracles: oracle/ (numpy fp64) and the committed reference outputs in tests/golden/."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cnmnet_b200
from cnmnet_b200 import _lib, engine, synth
from gpu_util import act_empty, dev, from_act, quantize, rel_l2, to_act

ALL_DTYPES = [_lib.CNM_BF16, _lib.CNM_TF32, _lib.CNM_BF16X3, _lib.CNM_F16F8]
from oracle import cost_volume as ocv
from oracle import nets as onets
from oracle import normals as onrm

pytestmark = pytest.mark.gpu


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


# ------------------------------------------------------------------ K1 cost volume
@pytest.mark.parametrize("tag", ["s3", "s2"])
def test_cost_volume_golden(golden_dir, tag):
    g = _g(golden_dir, "cost_volume_%s.npz" % tag)
    net = cnmnet_b200.depthNet(idepth_scale=float(g["idepth_scale"])).to(dev())
    H, W = g["left"].shape[2:]
    pix = cnmnet_b200.get_pixel_coordinates(H, W)
    KRKiUV, KT = cnmnet_b200.process_camera_parameters(_cuda(g["left_cam"]), _cuda(g["right_cam"]), pix)
    assert np.allclose(KT.cpu().numpy(), g["KT"], rtol=2e-4, atol=2e-4)
    assert np.allclose(KRKiUV.cpu().numpy()[:, :, ::97], g["KRKiUV"], rtol=2e-4, atol=2e-3)
    vol = net.getVolume(_cuda(g["left"]), _cuda(g["right"]), KRKiUV, KT).cpu().numpy()
    idmin, idmax = ocv.idepth_range(float(g["idepth_scale"]))
    exact = ocv.cost_volume(g["left"], g["right"], g["left_cam"], g["right_cam"], 64, idmin, idmax, np.float64)
    ref_err = rel_l2(g["volume"], exact)          # the reference's own fp32 noise on these inputs
    assert rel_l2(vol, exact) <= max(1e-5, 2 * ref_err), (rel_l2(vol, exact), ref_err)
    assert rel_l2(vol, g["volume"]) < 2e-5


def _volume(kernel, refd, src_vm, hk, out, B, V, H, W, D, layout, dtype, idmin=0.1, idmax=3.0):
    """One cost-volume launch through the C ABI: 'sweep' = cnm_plane_sweep (TMA-staged, reads the sources in place),
    'gather' = cnm_pack_source + cnm_cost_volume (first-generation kernel, taps through L1)."""
    lib, st = _lib.load(), _lib.stream_ptr()
    if kernel == "sweep":
        _lib.check(lib.cnm_plane_sweep(refd.data_ptr(), src_vm.data_ptr(), hk.data_ptr(), out.data_ptr(), B, V, H, W, D,
                                       idmin, idmax, layout, dtype, st))
    else:
        tex = torch.empty((lib.cnm_packed_source_bytes(V * B, H, W),), dtype=torch.uint8, device=dev())
        _lib.check(lib.cnm_pack_source(src_vm.data_ptr(), tex.data_ptr(), V * B, H, W, st))
        _lib.check(lib.cnm_cost_volume(refd.data_ptr(), tex.data_ptr(), hk.data_ptr(), out.data_ptr(), B, V, H, W, D,
                                       idmin, idmax, layout, dtype, st))
    torch.cuda.synchronize()


@pytest.mark.parametrize("kernel", ["sweep", "gather"])
def test_cost_volume_c2_size_vs_oracle(kernel):
    """BASELINE config 2: 256x192, 64 planes, 2 source views, batch 5, fp32."""
    B, V, H, W = 5, 2, 192, 256
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(1234, B, H, W, V)
    lib, st = _lib.load(), _lib.stream_ptr()
    src_vm = srcs.transpose(0, 1).contiguous().to(dev())        # [V,B,3,H,W]
    cam_vm = src_cams.transpose(0, 1).contiguous().to(dev())
    refd, rcd = ref.to(dev()), ref_cam.to(dev())
    hk = torch.empty((V * B, 12), dtype=torch.float32, device=dev())
    out = torch.empty((V * B, 64, H, W), dtype=torch.float32, device=dev())
    _lib.check(lib.cnm_camera_setup(rcd.data_ptr(), cam_vm.data_ptr(), hk.data_ptr(), B, V, st))
    _volume(kernel, refd, src_vm, hk, out, B, V, H, W, 64, _lib.VOLUME_NCHW_F32, 0)
    got = out.cpu().numpy().reshape(V, B, 64, H, W)
    for v in range(V):
        exact = ocv.cost_volume(ref.numpy(), srcs[:, v].numpy(), ref_cam.numpy(), src_cams[:, v].numpy(), 64, 0.1, 3.0)
        assert rel_l2(got[v], exact) < 1e-5
        rel = np.abs(got[v] - exact) / np.maximum(np.abs(exact), 1e-3)
        assert (rel > 1e-5).mean() < 0.02          # SURVEY.md 8d gate: <= 2 % of elements off by > 1e-5 relative
    # the act layouts carry the same numbers (bf16-rounded / exact fp32)
    for dtype in ALL_DTYPES:
        act = act_empty((V * B, H, W, 64), dtype)
        _volume(kernel, refd, src_vm, hk, act, B, V, H, W, 64, _lib.VOLUME_ACT, dtype)
        back = from_act(act, dtype)
        want = out.cpu()
        assert torch.equal(back, quantize(want, dtype))


@pytest.mark.parametrize("case", ["wide_baseline", "behind_camera", "rotation", "planes128", "partial_tiles", "scale2"])
def test_plane_sweep_footprint_cases(case):
    """The TMA-staged kernel against the fp64 oracle where its box logic is exercised: footprints that move too fast
    for 8 planes per box (finer levels), sample rays that cross the source camera plane (p_z changes sign: no box, the
    global-memory path), strong rotation (boxes much taller than the tile), 128 planes (16 groups), image sizes that
    leave partial tiles, and idepth_scale 2 (planes out to 50 m)."""
    rng = np.random.default_rng(len(case))
    B, V, H, W, D, idr = 2, 1, 64, 96, 64, (0.1, 3.0)
    rot, trans = 0.05, 0.15
    if case == "wide_baseline":
        trans = 1.0
    elif case == "behind_camera":
        trans = 3.0
    elif case == "rotation":
        rot = 0.6
    elif case == "planes128":
        D = 128
    elif case == "partial_tiles":
        H, W = 44, 84
    elif case == "scale2":
        idr = (0.02, 2.0)
    left, right = synth.smooth_images(rng, B, H, W), synth.smooth_images(rng, B, H, W)
    lcam, rcam = synth.cameras(rng, B, H, W, rot_sigma=rot, trans_sigma=trans)
    lib, st = _lib.load(), _lib.stream_ptr()
    hk = torch.empty((B, 12), dtype=torch.float32, device=dev())
    _lib.check(lib.cnm_camera_setup(_cuda(lcam).data_ptr(), _cuda(rcam).data_ptr(), hk.data_ptr(), B, 1, st))
    outs = {}
    for kernel in ("sweep", "gather"):
        out = torch.full((B, D, H, W), float("nan"), dtype=torch.float32, device=dev())
        _volume(kernel, _cuda(left), _cuda(right), hk, out, B, 1, H, W, D, _lib.VOLUME_NCHW_F32, 0, *idr)
        outs[kernel] = out.cpu().numpy()
    exact = ocv.cost_volume(left, right, lcam, rcam, D, idr[0], idr[1])
    assert np.isfinite(outs["sweep"]).all()
    # Sample positions are amplified where p_z is small (wide baselines, rays near the source camera plane): both
    # kernels are gated by what the first-generation kernel -- the round-1 parity reference -- scores on the same inputs
    err_s, err_g = rel_l2(outs["sweep"], exact), rel_l2(outs["gather"], exact)
    assert err_s < max(1e-5, 2.0 * err_g), (case, err_s, err_g)
    big = np.abs(outs["sweep"] - exact) > 1e-3 * np.maximum(np.abs(exact), 1.0)
    big_g = np.abs(outs["gather"] - exact) > 1e-3 * np.maximum(np.abs(exact), 1.0)
    assert big.mean() <= max(1e-4, 2.0 * big_g.mean()), (case, big.mean(), big_g.mean())


@pytest.mark.parametrize("H,W", [(24, 40), (33, 70), (8, 8)])
def test_cost_volume_ragged_shapes_and_identity_pose(H, W):
    rng = np.random.default_rng(H * W)
    left, right = synth.smooth_images(rng, 2, H, W), synth.smooth_images(rng, 2, H, W)
    lcam, rcam = synth.cameras(rng, 2, H, W)
    rcam[1] = lcam[1]                                 # identity motion for batch item 1
    net = cnmnet_b200.depthNet(idepth_scale=3.0).to(dev())
    pix = cnmnet_b200.get_pixel_coordinates(H, W)
    KRKiUV, KT = cnmnet_b200.process_camera_parameters(_cuda(lcam), _cuda(rcam), pix)
    vol = net.getVolume(_cuda(left), _cuda(right), KRKiUV, KT).cpu().numpy()
    exact = ocv.cost_volume(left, right, lcam, rcam, 64, 0.1, 3.0)
    assert rel_l2(vol, exact) < 1e-5


# ------------------------------------------------------------------ K3 normals, K3b CNM
@pytest.mark.parametrize("k", [9, 5])
def test_normals_golden_and_oracle(golden_dir, k):
    g = _g(golden_dir, "normals_k%d.npz" % k)
    d2n = cnmnet_b200.Depth2normal(k)
    normals, points = d2n(_cuda(g["depth"]), _cuda(g["Kinv"]))
    n, p = normals.cpu().numpy(), points.cpu().numpy()
    n64, p64, det = onrm.depth2normal(g["depth"], g["Kinv"], k, np.float64, return_det=True)
    stable = np.broadcast_to((np.abs(det - 1e-5) > 1e-6)[:, None], n.shape)
    assert np.isfinite(n).all()
    assert np.abs(n - n64)[stable].max() < 1e-5        # fp64 kernel vs fp64 oracle
    fin = np.isfinite(g["normals"])
    assert np.abs(n - g["normals"])[fin & stable].mean() < 2e-3   # vs the reference's fp32 output (its own noise)
    pf = np.isfinite(p64)
    assert np.allclose(p[pf], p64[pf], rtol=1e-6, atol=1e-6)
    assert (np.isnan(p) == np.isnan(g["points"])).all()


def test_normals_c3_size_vs_oracle():
    B, H, W = 5, 192, 256
    rng = np.random.default_rng(77)
    depth = synth.depth_maps(rng, B, H, W)
    Kinv = np.broadcast_to(np.linalg.inv(synth.intrinsics(H, W)).astype(np.float32), (B, 3, 3)).copy()
    normals, _ = cnmnet_b200.Depth2normal(9)(_cuda(depth), _cuda(Kinv))
    n64, _, det = onrm.depth2normal(depth, Kinv, 9, np.float64, return_det=True)
    stable = np.broadcast_to((np.abs(det - 1e-5) > 1e-6)[:, None], n64.shape)
    assert np.abs(normals.cpu().numpy() - n64)[stable].max() < 1e-5


def test_normals_all_invalid_and_tiny_images():
    d2n = cnmnet_b200.Depth2normal(9)
    z = torch.zeros((1, 7, 5), device=dev())
    n, p = d2n(z, torch.eye(3, device=dev())[None])
    assert torch.count_nonzero(n) == 0 and torch.count_nonzero(p) == 0
    d = torch.full((1, 3, 3), 2.0, device=dev())
    n, _ = d2n(d, torch.eye(3, device=dev())[None])
    n64, _ = onrm.depth2normal(d.cpu().numpy(), np.eye(3)[None], 9)
    assert np.allclose(n.cpu().numpy(), n64, atol=1e-5)


def test_cnm_golden_and_oracle(golden_dir):
    g = _g(golden_dir, "normals_k9.npz")
    out = cnmnet_b200.get_normal_by_planes(_cuda(g["gt_like"]), _cuda(g["segs"]), [int(v) for v in g["planes_num"]])
    assert np.allclose(out.cpu().numpy(), g["normals_by_planes"], rtol=1e-4, atol=1e-5)
    want = onrm.get_normal_by_planes(g["gt_like"], g["segs"], g["planes_num"])
    assert rel_l2(out.cpu().numpy(), want) < 1e-6
    n, loss, p = cnmnet_b200.Depth2normal(9)(_cuda(g["depth"]), _cuda(g["Kinv"]), _cuda(g["segs"]),
                                             torch.from_numpy(g["planes_num"]))
    n64, loss64, _ = onrm.depth2normal_cnm(g["depth"], g["Kinv"], g["segs"], g["planes_num"])
    ok = np.isfinite(n64)
    assert np.abs(n.cpu().numpy() - n64)[ok].max() < 1e-5
    assert abs(float(loss) - float(loss64)) < 1e-5 * max(1.0, abs(float(loss64)))
    # overlapping masks: later planes see earlier replacements (sequential semantics of the reference loop)
    segs = np.zeros((1, 20, 16, 16), np.uint8)
    segs[0, 0, :, :10] = 1
    segs[0, 1, :, 6:] = 1
    x = np.random.default_rng(3).standard_normal((1, 3, 16, 16)).astype(np.float32)
    got = cnmnet_b200.get_normal_by_planes(_cuda(x), _cuda(segs), [2]).cpu().numpy()
    assert rel_l2(got, onrm.get_normal_by_planes(x, segs, [2])) < 1e-6


def test_cnm_overlapping_chains_and_loss(golden_dir):
    """Masks that overlap in chains (plane i sees what planes < i wrote): the parallel accumulate / solve / apply
    formulation must reproduce the reference's sequential loop, loss term included."""
    g = _g(golden_dir, "normals_k9.npz")
    depth, Kinv = g["depth"], g["Kinv"]
    B, H, W = depth.shape
    rng = np.random.default_rng(17)

    def boxes(counts):
        segs = np.zeros((len(counts), 20, H, W), np.uint8)
        for b, n in enumerate(counts):
            for i in range(n):
                y0, x0 = int(rng.integers(0, H // 2)), int(rng.integers(0, W // 2))
                segs[b, i, y0:y0 + H // 2, x0:x0 + W // 2] = 1       # half-image boxes: heavy overlap
        return segs

    counts = [5, 3][:B] + [2] * max(0, B - 2)
    segs, pn = boxes(counts), np.array(counts)
    n, loss, _ = cnmnet_b200.Depth2normal(9)(_cuda(depth), _cuda(Kinv), _cuda(segs), torch.from_numpy(pn))
    n64, loss64, _ = onrm.depth2normal_cnm(depth, Kinv, segs, pn)
    ok = np.isfinite(n64)
    assert np.abs(n.cpu().numpy() - n64)[ok].max() < 1e-5
    assert abs(float(loss) - float(loss64)) < 1e-5 * max(1.0, abs(float(loss64)))
    counts3 = [6, 3, 0]
    segs3 = boxes(counts3)
    x = rng.standard_normal((3, 3, H, W)).astype(np.float32)
    got = cnmnet_b200.get_normal_by_planes(_cuda(x), _cuda(segs3), counts3).cpu().numpy()
    assert rel_l2(got, onrm.get_normal_by_planes(x, segs3, np.array(counts3))) < 1e-6
    assert np.array_equal(got[2], x[2])                               # no planes: untouched


def test_cnm_empty_mask_below_planes_num_gives_nan_loss(golden_dir):
    """A plane slot below planes_num with an empty mask: the reference's mean is 0/0, cosine_similarity is NaN at every
    pixel and the loss is NaN (depth_util.py:222-231); the oracle's sequential loop reproduces that."""
    g = _g(golden_dir, "normals_k9.npz")
    depth, Kinv = g["depth"][:1], g["Kinv"][:1]
    _, H, W = depth.shape
    segs = np.zeros((1, 20, H, W), np.uint8)
    segs[0, 0, : H // 2] = 1                     # slot 1 stays empty, slot 2 is used
    segs[0, 2, H // 2:] = 1
    pn = np.array([3])
    n, loss, _ = cnmnet_b200.Depth2normal(9)(_cuda(depth), _cuda(Kinv), _cuda(segs), torch.from_numpy(pn))
    n64, loss64, _ = onrm.depth2normal_cnm(depth, Kinv, segs, pn)
    assert np.isnan(float(loss64)) and np.isnan(float(loss))
    ok = np.isfinite(n64)
    assert np.abs(n.cpu().numpy() - n64)[ok].max() < 1e-5


# ------------------------------------------------------------------ K4 aggregation and the small kernels
@pytest.mark.parametrize("dtype", ALL_DTYPES)
@pytest.mark.parametrize("group", [1, 2, 3])
def test_refine_aggregate(dtype, group):
    B, H, W, V = 2, 12, 20, 2 * group
    g = torch.Generator().manual_seed(group)
    idepth = torch.rand(V, B, H, W, generator=g) * 2 + 0.2
    iconv = torch.rand(V * B, 64, H, W, generator=g) * 3
    lib, st, ch = _lib.load(), _lib.stream_ptr(), engine.chunk_channels(dtype)
    ic_act = to_act(iconv, dtype)
    sum_act = act_empty((B, H, W, 64), dtype)
    taps = act_empty((B, H, W, ch), dtype)
    id_out = torch.empty((2, B, H, W), dtype=torch.float32, device=dev())
    idd = idepth.to(dev())
    _lib.check(lib.cnm_refine_aggregate(idd.data_ptr(), ic_act.data_ptr(), sum_act.data_ptr(), taps.data_ptr(),
                                        id_out.data_ptr(), B, H, W, group, dtype, st))
    icq = quantize(iconv, dtype).view(V, B, 64, H, W)
    m = [onets.group_mean([idepth[v] for v in range(gi, V, 2)]) for gi in range(2)]
    mi = [onets.group_mean([icq[v] for v in range(gi, V, 2)]) for gi in range(2)]
    assert torch.equal(id_out.cpu(), torch.stack(m))                       # bit-exact group means
    want_sum = quantize(mi[0] + mi[1], dtype)
    assert torch.equal(from_act(sum_act, dtype), want_sum)
    x3 = torch.stack([m[0], m[1], (m[0] - m[1]).abs()], 1)                 # [B,3,H,W]
    pd = F.pad(x3, (1, 1, 1, 1))
    got = from_act(taps, dtype)
    for c in range(3):
        for r in range(3):
            for s in range(3):
                assert torch.equal(got[:, 9 * c + 3 * r + s], quantize(pd[:, c, r:r + H, s:s + W], dtype))
    assert torch.count_nonzero(got[:, 27:]) == 0


@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_upsample_head_packers(dtype):
    lib, st, ch = _lib.load(), _lib.stream_ptr(), engine.chunk_channels(dtype)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 64, 6, 10, generator=g)
    xa = to_act(x, dtype)
    up = act_empty((2, 12, 20, 64), dtype)
    _lib.check(lib.cnm_upsample2x(xa.data_ptr(), up.data_ptr(), 2, 6, 10, 64, dtype, st))
    want = F.interpolate(quantize(x, dtype), scale_factor=2, mode="bilinear", align_corners=False)
    tol = {_lib.CNM_BF16: 2e-2, _lib.CNM_TF32: 2e-3, _lib.CNM_BF16X3: 1e-4, _lib.CNM_F16F8: 1e-4}[dtype]   # one output rounding
    assert (from_act(up, dtype) - want).abs().max() < tol
    # head: conv3x3(C->1)+bias -> scale*sigmoid
    w = torch.randn(1, 64, 3, 3, generator=g) * 0.05
    wv = w[0].permute(1, 2, 0).reshape(-1).contiguous().to(dev())
    out = torch.empty((2, 1, 6, 10), device=dev())
    logit = torch.empty((2, 1, 6, 10), device=dev())
    _lib.check(lib.cnm_head3x3(xa.data_ptr(), wv.data_ptr(), 0.25, 3.0, out.data_ptr(), logit.data_ptr(), 2, 6, 10, 64,
                               dtype, st))
    wl = F.conv2d(quantize(x, dtype).double(), w.double(), torch.tensor([0.25], dtype=torch.float64), padding=1)
    assert (logit.cpu() - wl.float()).abs().max() < 1e-4
    assert (out.cpu() - 3.0 * torch.sigmoid(wl).float()).abs().max() < 1e-5
    # the same head in two passes (per-pixel tap dot products, then the 3x3 gather): what the pipeline runs
    part = torch.full((2, 6, 10, 12), float("nan"), device=dev())
    wt = w[0].permute(1, 2, 0).reshape(9, 64).contiguous().to(dev())
    out2, logit2 = torch.empty_like(out), torch.empty_like(logit)
    _lib.check(lib.cnm_head_partial(xa.data_ptr(), wt.data_ptr(), part.data_ptr(), 2, 6, 10, 64, dtype, st))
    _lib.check(lib.cnm_head_finish(part.data_ptr(), 1, 0.25, 3.0, out2.data_ptr(), logit2.data_ptr(), 2, 6, 10, st))
    assert (logit2.cpu() - wl.float()).abs().max() < 1e-4 and (out2.cpu() - out.cpu()).abs().max() < 1e-5
    # disparity taps: nearest x2 then 3x3 taps
    disp = torch.rand(2, 1, 6, 10, generator=g)
    t = act_empty((2, 12, 20, ch), dtype)
    dd = disp.to(dev())
    _lib.check(lib.cnm_pack_disp_taps(dd.data_ptr(), t.data_ptr(), 2, 6, 10, dtype, st))
    pu = F.pad(F.interpolate(disp, scale_factor=2, mode="nearest"), (1, 1, 1, 1))
    got = from_act(t, dtype)
    for r in range(3):
        for s in range(3):
            assert torch.equal(got[:, 3 * r + s], quantize(pu[:, 0, r:r + 12, s:s + 20], dtype))
    # image row taps, one ref image feeding two views
    img = torch.randn(2, 3, 6, 10, generator=g)
    t = act_empty((4, 6, 10, ch), dtype)
    im = img.to(dev())
    _lib.check(lib.cnm_pack_row_taps(im.data_ptr(), t.data_ptr(), 4, 2, 3, 6, 10, 7, dtype, st))
    pi = F.pad(img, (3, 3, 0, 0))
    got = from_act(t, dtype)
    for n in range(4):
        for s in range(7):
            for c in range(3):
                assert torch.equal(got[n, s * 3 + c], quantize(pi[n % 2, c, :, s:s + 10], dtype))
    if dtype != _lib.CNM_TF32:
        # the denser form: three kernel rows per chunk, one apron row above and below (row j centred on image row j - 1)
        t3 = act_empty((4, 8, 10, ch), dtype)
        _lib.check(lib.cnm_pack_row_taps3(im.data_ptr(), t3.data_ptr(), 4, 2, 3, 6, 10, 7, dtype, st))
        pi3 = F.pad(img, (3, 3, 3, 3))
        got = from_act(t3, dtype)
        assert got.shape == (4, ch, 8, 10)
        for n in range(4):
            for j in range(8):
                for q in range(3):
                    for s in range(7):
                        for c in range(3):
                            assert torch.equal(got[n, (q * 7 + s) * 3 + c, j], quantize(pi3[n % 2, c, j - 2 + q + 3, s:s + 10], dtype))
        assert float(got[:, 63].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", ALL_DTYPES)
@pytest.mark.parametrize("C,H,W", [(128, 12, 20), (256, 9, 16), (512, 6, 8)])
def test_wide_heads_both_forms(dtype, C, H, W):
    """disp2 / disp3 / disp4-sized heads: the one-kernel form (cnm_head3x3) and the two-pass form agree with torch."""
    lib, st = _lib.load(), _lib.stream_ptr()
    g = torch.Generator().manual_seed(C)
    x = torch.randn(3, C, H, W, generator=g).abs()
    w = torch.randn(1, C, 3, 3, generator=g) * 0.02
    xa = to_act(x, dtype)
    wl = F.conv2d(quantize(x, dtype).double(), w.double(), torch.tensor([-0.1], dtype=torch.float64), padding=1).float()
    wv = w[0].permute(1, 2, 0).reshape(-1).contiguous().to(dev())
    out, logit = torch.empty((3, 1, H, W), device=dev()), torch.empty((3, 1, H, W), device=dev())
    _lib.check(lib.cnm_head3x3(xa.data_ptr(), wv.data_ptr(), -0.1, 2.0, out.data_ptr(), logit.data_ptr(), 3, H, W, C, dtype, st))
    part = torch.empty((3, H, W, 12), device=dev())
    out2, logit2 = torch.empty_like(out), torch.empty_like(logit)
    _lib.check(lib.cnm_head_partial(xa.data_ptr(), wv.data_ptr(), part.data_ptr(), 3, H, W, C, dtype, st))
    _lib.check(lib.cnm_head_finish(part.data_ptr(), 1, -0.1, 2.0, out2.data_ptr(), logit2.data_ptr(), 3, H, W, st))
    for lg, o in ((logit, out), (logit2, out2)):
        assert (lg.cpu() - wl).abs().max() < 2e-4
        assert (o.cpu() - 2.0 * torch.sigmoid(wl)).abs().max() < 5e-5


def test_depth_metrics_golden_and_oracle(golden_dir):
    """cnm_depth_metrics vs the reference's functions (golden) and vs the oracle on a larger random case."""
    from oracle import metrics as omet
    g = _g(golden_dir, "depth_metrics.npz")
    per, means = cnmnet_b200.depth_metrics(_cuda(g["pred"]), _cuda(g["gt"].astype(np.float32)))
    per = per.cpu().numpy()
    want = g["metrics"]
    assert np.array_equal(per[:, 0], want[:, 0])
    ok = want[:, 0] > 0
    assert np.allclose(per[ok, 1:7], want[ok, 1:7], rtol=2e-6)          # gt is fp32 on the device, fp64 in the fixture
    assert np.abs(per[ok, 7:] - want[ok, 7:]).max() <= 2.0 / want[ok, 0].min()   # threshold counts: +-2 pixels
    assert np.isnan(per[~ok, 1:]).all() and np.isnan(means["l1"])
    rng = np.random.default_rng(9)
    gt = synth.depth_maps(rng, 3, 192, 256).astype(np.float32)
    pred = (gt * (1 + 0.05 * rng.standard_normal(gt.shape)) + 0.01).astype(np.float32)
    per, means = cnmnet_b200.depth_metrics(_cuda(pred), _cuda(gt))
    operi, omeans = omet.sequence_metrics(pred, gt)
    for k in omet.NAMES[:6]:
        assert abs(means[k] - omeans[k]) < 1e-6 * max(1.0, abs(omeans[k])), k
    for k in omet.NAMES[6:]:
        assert abs(means[k] - omeans[k]) < 1e-4, k


def test_losses_golden_and_oracle(golden_dir):
    """cnmnet_b200.depthnet.losses (forward values) vs the reference module's outputs (golden) and the fp64 oracle."""
    from cnmnet_b200.depthnet import losses as L
    from oracle import losses as olos
    g = _g(golden_dir, "losses.npz")
    t = {k: _cuda(g[k]) for k in ("gt", "pred", "prob", "disp2", "disp3", "disp4", "gt_pos", "n_pred", "n_gt", "valid")}
    got = {"idepth": L.IdepthLoss()(t["pred"], t["gt"]), "idepth_log": L.IdepthLoss()(t["pred"], t["gt"], log=True),
           "idepth_prob": L.IdepthwithProbLoss()(t["pred"], t["gt"], t["prob"]),
           "idepth_prob_log": L.IdepthwithProbLoss()(t["pred"], t["gt"], t["prob"], log=True),
           "idepth_234": L.IdepthLoss_234()([t["pred"], t["disp2"], t["disp3"], t["disp4"]], t["gt_pos"])}
    got["normal"], got["normal_angle"] = L.surface_normal_loss(t["n_pred"], t["n_gt"], t["valid"])
    got["normal_prob"], got["normal_prob_angle"] = L.surface_normal_loss(t["n_pred"], t["n_gt"], t["valid"], t["prob"])
    for k, v in got.items():
        assert abs(float(v) - float(g[k])) < 5e-6 * max(1.0, abs(float(g[k]))), (k, float(v), float(g[k]))
    # a C4-sized case against the oracle
    gen = torch.Generator().manual_seed(3)
    gt = torch.rand(5, 1, 192, 256, generator=gen) * 2 + 0.05
    pred = gt * (1 + 0.05 * torch.randn(5, 1, 192, 256, generator=gen))
    prob = torch.rand(5, 1, 192, 256, generator=gen)
    assert abs(float(L.IdepthwithProbLoss()(pred.to(dev()), gt.to(dev()), prob.to(dev()))) -
               olos.idepth_prob_loss(pred, gt, prob)) < 1e-6
    with pytest.raises(RuntimeError):
        L.IdepthLoss()(pred, gt)                                      # CPU tensors: no CPU path


def test_reciprocal_and_inverse_warp(golden_dir):
    lib, st = _lib.load(), _lib.stream_ptr()
    x = torch.rand(1000, device=dev()) + 0.1
    y = torch.empty_like(x)
    _lib.check(lib.cnm_reciprocal(x.data_ptr(), y.data_ptr(), 1000, st))
    assert torch.equal(y, torch.div(1.0, x))
    g = _g(golden_dir, "inverse_warp.npz")
    out = cnmnet_b200.inverse_warp(_cuda(g["feat"]), _cuda(g["depth"]), _cuda(g["pose"]), _cuda(g["K"]), _cuda(g["Kinv"]))
    assert np.allclose(out.cpu().numpy(), g["warped"], rtol=1e-4, atol=2e-4)
    pts = cnmnet_b200.pixel2cam(_cuda(g["depth"]), _cuda(g["Kinv"]))
    assert np.allclose(pts.cpu().numpy(), onrm.pixel2cam(g["depth"], g["Kinv"]), rtol=1e-5, atol=1e-5)
    cam = torch.from_numpy(synth.cameras(np.random.default_rng(1), 3, 192, 256)[0]).to(dev())
    Kinv = torch.empty((3, 3, 3), device=dev())
    _lib.check(lib.cnm_intrinsics_inverse(cam.data_ptr(), Kinv.data_ptr(), 3, st))
    assert np.allclose(Kinv.cpu().numpy(), np.linalg.inv(cam[:, 1, :3, :3].cpu().numpy().astype(np.float64)), rtol=1e-6)
