"""GPU parity of the networks and the end-to-end pipeline against the reference's outputs (tests/golden/)
and the CPU oracle, through the reference-shaped module API.

Tolerances (north_star): final depth within 1e-3 relative in the TF32 tensor-core mode; the BF16 speed mode
is reported against a looser bound (pure-bf16 operands give ~1.4e-3 mean / 7e-3 max)."""
import os

import numpy as np
import pytest
import torch

import cnmnet_b200
from cnmnet_b200 import synth
from gpu_util import dev, rel_l2
from oracle import nets as onets
from oracle import normals as onrm

pytestmark = pytest.mark.gpu

# (precision, backend) -> (iconv rel-L2, depth rel-err p99, max, mean).  Measured on B200 (round 1): tf32 p99 ~6e-4,
# max 1.1-1.6e-3, mean 2-3e-4; bf16 p99 ~1e-2, max 1.5e-2, mean 3.7e-3.  The tensor-core kernel and its fp32-FMA
# SIMT twin give the same numbers, i.e. the error is the operand rounding (10 / 7 mantissa bits) of the mode.
TOL = {("tf32", "tc"): (1.5e-3, 1e-3, 2.5e-3, 4e-4), ("tf32", "simt"): (1.5e-3, 1e-3, 2.5e-3, 4e-4),
       ("bf16", "tc"): (1.5e-2, 1.5e-2, 3e-2, 5e-3), ("bf16", "simt"): (1.5e-2, 1.5e-2, 3e-2, 5e-3),
       # split bf16 (hi + lo, three tensor-core passes): 16 operand mantissa bits -- the mode that meets the north-star
       # 1e-3 on the final depth with two orders of magnitude to spare
       # (measured: iconv rel-L2 9e-5 on the tensor cores / < 5e-5 with fp32 FMAs, depth p99 < 1e-4)
       ("bf16x3", "tc"): (2e-4, 1e-4, 1e-3, 5e-5), ("bf16x3", "simt"): (2e-4, 1e-4, 1e-3, 5e-5),
       # fp16 + fp8 corrections (one kind::f16 pass + one kind::f8f6f4 pass): operand error ~2^-17 like bf16x3
       # (tools/emulate_modes.py predicts ~2x the bf16x3 error), at 2/3 of its tensor time
       ("f16f8", "tc"): (3e-4, 2e-4, 1e-3, 8e-5), ("f16f8", "simt"): (3e-4, 2e-4, 1e-3, 8e-5)}
MODES = list(TOL.keys())


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def _state(layers, g):
    return synth.kaiming_state(onets.state_shapes(layers), int(g["weight_seed"]), float(g["head_scale"]))


@pytest.mark.parametrize("precision,backend", MODES)
def test_depthnet_forward_golden(golden_dir, precision, backend):
    g = _g(golden_dir, "depthnet_fwd.npz")
    net = cnmnet_b200.depthNet(idepth_scale=3.0, precision=precision, backend=backend)
    net.load_state_dict(_state(onets.DEPTHNET_LAYERS, g))
    net = net.to(dev()).eval()
    disps, iconv1 = net(_cuda(g["left"]), _cuda(g["right"]), _cuda(g["left_cam"]), _cuda(g["right_cam"]))
    t_ic, t_p99, t_max, t_mean = TOL[(precision, backend)]
    assert iconv1.shape == g["iconv1"].shape
    assert rel_l2(iconv1.cpu().numpy(), g["iconv1"]) < t_ic
    for i, k in enumerate(["disp1", "disp2", "disp3", "disp4"]):
        got = disps[i].cpu().numpy()
        assert got.shape == g[k].shape
        rel = np.abs(1.0 / got - 1.0 / g[k]) * g[k]          # relative depth error, depth = 1/idepth
        # the north-star tolerance is on the FINAL depth (disp1); the coarser side outputs get 1.5x
        f = 1.0 if k == "disp1" else 1.5
        assert np.percentile(rel, 99) < f * t_p99 and rel.max() < f * t_max and rel.mean() < f * t_mean, \
            (k, np.percentile(rel, 99), rel.max(), rel.mean())


@pytest.mark.parametrize("precision,backend", MODES)
def test_refinenet_forward_golden(golden_dir, precision, backend):
    g = _g(golden_dir, "refinenet_fwd.npz")
    net = cnmnet_b200.DepthRefineNet(idepth_scale=3.0, precision=precision, backend=backend)
    net.load_state_dict(_state(onets.REFINENET_LAYERS, g))
    net = net.to(dev()).eval()
    dr, pm, vol = net(_cuda(g["idepth01"]), _cuda(g["idepth02"]), _cuda(g["iconv01"]), _cuda(g["iconv02"]),
                      ReturnVolume=True)
    t_ic, t_p99, t_max, t_mean = TOL[(precision, backend)]
    assert rel_l2(vol.cpu().numpy(), g["iconv1_depth"]) < t_ic
    rel = np.abs(1.0 / dr.cpu().numpy() - 1.0 / g["disp_refined"]) * g["disp_refined"]
    assert np.percentile(rel, 99) < t_p99 and rel.max() < t_max and rel.mean() < t_mean, (rel.max(), rel.mean())
    assert np.abs(pm.cpu().numpy() - g["prob_map"]).max() < 10 * t_max
    dr2, pm2 = net(_cuda(g["idepth01"]), _cuda(g["idepth02"]), _cuda(g["iconv01"]), _cuda(g["iconv02"]))
    assert torch.equal(dr2, dr) and torch.equal(pm2, pm)     # deterministic, and the 2-tuple return form


def _oracle_pipeline(dsd, rsd, ref, srcs, ref_cam, src_cams, scale, k_size, planes=64):
    B, V = srcs.shape[:2]
    ids, ics = [], []
    with torch.no_grad():
        for v in range(V):
            d, ic = onets.depthnet_forward(dsd, ref, srcs[:, v], ref_cam, src_cams[:, v], scale, num_planes=planes)
            ids.append(d[0])
            ics.append(ic)
        id1, id2 = onets.group_mean(ids[0::2]), onets.group_mean(ids[1::2])
        ic1, ic2 = onets.group_mean(ics[0::2]), onets.group_mean(ics[1::2])
        idr, prob = onets.refinenet_forward(rsd, id1, id2, ic1, ic2, scale)
        depth = onets.idepth_to_depth(idr)
    Kinv = np.linalg.inv(ref_cam[:, 1, :3, :3].numpy().astype(np.float64))
    normal, _ = onrm.depth2normal(depth.numpy(), Kinv, k_size)
    return depth.numpy(), normal, prob.numpy()


@pytest.mark.parametrize("precision,V", [("tf32", 2), ("bf16", 2), ("tf32", 4), ("bf16x3", 2), ("f16f8", 2), ("f16f8", 4)])
def test_pipeline_matches_oracle(precision, V):
    B, H, W = 2, 64, 96
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(4321 + V, B, H, W, V)
    dsd = synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 21, 0.02)
    rsd = synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), 22, 0.02)
    dnet = cnmnet_b200.depthNet(idepth_scale=3.0, precision=precision)
    rnet = cnmnet_b200.DepthRefineNet(idepth_scale=3.0, precision=precision)
    dnet.load_state_dict(dsd)
    rnet.load_state_dict(rsd)
    dnet, rnet = dnet.to(dev()), rnet.to(dev())
    want_d, want_n, want_p = _oracle_pipeline(dsd, rsd, ref, srcs, ref_cam, src_cams, 3.0, 9)
    outs = {}
    for graph in (False, True):
        pipe = cnmnet_b200.CNMPipeline(dnet, rnet, k_size=9, use_graph=graph)
        d, n, p = pipe(ref.to(dev()), srcs.to(dev()), ref_cam.to(dev()), src_cams.to(dev()))
        d2, n2, p2 = pipe(ref.to(dev()), srcs.to(dev()), ref_cam.to(dev()), src_cams.to(dev()))   # replay
        assert torch.equal(d, d2) and torch.equal(n, n2) and torch.equal(p, p2)
        outs[graph] = (d.cpu().numpy(), n.cpu().numpy(), p.cpu().numpy())
    for a, b in zip(outs[False], outs[True]):
        assert np.array_equal(a, b)                          # the CUDA graph replays the same kernels
    d, n, p = outs[True]
    rel = np.abs(d - want_d) / np.abs(want_d)
    _, t_p99, t_max, t_mean = TOL[(precision, "tc")]
    assert np.percentile(rel, 99) < t_p99 and rel.max() < t_max and rel.mean() < t_mean, \
        (np.percentile(rel, 99), rel.max(), rel.mean())
    assert np.abs(p - want_p).max() < 10 * t_max
    # normals amplify depth noise (9x9 plane fit on a ~1e-3 perturbed surface): compare by angle
    cos = np.clip((n * want_n).sum(1), -1, 1)
    ang = np.degrees(np.arccos(cos))
    assert np.median(ang) < {"tf32": 1.0, "bf16": 8.0, "bf16x3": 0.5, "f16f8": 0.5}[precision], np.median(ang)


@pytest.mark.parametrize("precision", ["bf16x3", "f16f8"])
@pytest.mark.parametrize("H,W,B", [(96, 160, 3), (480, 640, 1)])
def test_pipeline_odd_pyramid_sizes(H, W, B, precision):
    """Image sizes whose pyramid levels are not multiples of the 16x16 / 16x8 pixel tiles (96x160 -> 48x80 ... 3x5;
    640x480 = BASELINE config 5's resolution -> 240x320 ... 15x20): partial tiles, TMA clipping, odd batch."""
    V = 2
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(500 + H, B, H, W, V)
    dsd = synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 51, 0.02)
    rsd = synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), 52, 0.02)
    dnet = cnmnet_b200.depthNet(idepth_scale=3.0, precision=precision)
    rnet = cnmnet_b200.DepthRefineNet(idepth_scale=3.0, precision=precision)
    dnet.load_state_dict(dsd)
    rnet.load_state_dict(rsd)
    pipe = cnmnet_b200.CNMPipeline(dnet.to(dev()), rnet.to(dev()), k_size=9)
    d, n, p = pipe(ref.to(dev()), srcs.to(dev()), ref_cam.to(dev()), src_cams.to(dev()))
    want_d, want_n, want_p = _oracle_pipeline(dsd, rsd, ref, srcs, ref_cam, src_cams, 3.0, 9)
    rel = np.abs(d.cpu().numpy() - want_d) / np.abs(want_d)
    # these weights (seed 51/52) are ~8x more sensitive than the others in this file (tools/acc_probe.py: tf32 p99 7e-3):
    # the north-star 1e-3 holds in the accuracy modes with a 2x margin
    assert rel.max() < 5e-4 and rel.mean() < 1e-4, (rel.max(), rel.mean())
    assert np.abs(p.cpu().numpy() - want_p).max() < 2e-3


@pytest.mark.parametrize("precision", ["f16f8", "bf16x3"])
def test_pipeline_128_planes(precision):
    """BASELINE config 5's plane count: depthNet(planes=128) = the same sweep at 128 hypotheses, conv1.0 with 3 + 128
    input channels (an extension: the reference hard-codes 64).  Checked against the oracle wiring with the same
    widened conv1 and the fp64 cost volume at 128 planes; the NCHW getVolume() result is checked too."""
    B, H, W, V, D = 1, 64, 96, 2, 128
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(777, B, H, W, V)
    dsd = synth.kaiming_state(onets.state_shapes(onets.depthnet_layers(D)), 61, 0.02)
    rsd = synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), 62, 0.02)
    assert tuple(dsd["conv1.0.weight"].shape) == (128, 3 + D, 7, 7)
    dnet = cnmnet_b200.depthNet(idepth_scale=3.0, precision=precision, planes=D)
    rnet = cnmnet_b200.DepthRefineNet(idepth_scale=3.0, precision=precision)
    dnet.load_state_dict(dsd)
    rnet.load_state_dict(rsd)
    dnet, rnet = dnet.to(dev()), rnet.to(dev())
    pix = cnmnet_b200.get_pixel_coordinates(H, W)
    KRKiUV, KT = cnmnet_b200.process_camera_parameters(ref_cam.to(dev()), src_cams[:, 0].to(dev()), pix)
    vol = dnet.getVolume(ref.to(dev()), srcs[:, 0].to(dev()), KRKiUV, KT).cpu().numpy()
    from oracle import cost_volume as ocv
    exact = ocv.cost_volume(ref.numpy(), srcs[:, 0].numpy(), ref_cam.numpy(), src_cams[:, 0].numpy(), D, 0.1, 3.0)
    # SURVEY.md 8d gates (the reference's own fp32 port scores rel-L2 2.4e-6 / 1.5 % on this pair)
    assert vol.shape == (B, D, H, W) and rel_l2(vol, exact) < 1e-5
    assert (np.abs(vol - exact) / np.maximum(np.abs(exact), 1e-3) > 1e-5).mean() < 0.02
    pipe = cnmnet_b200.CNMPipeline(dnet, rnet, k_size=9)
    d, n, p = pipe(ref.to(dev()), srcs.to(dev()), ref_cam.to(dev()), src_cams.to(dev()))
    want_d, want_n, want_p = _oracle_pipeline(dsd, rsd, ref, srcs, ref_cam, src_cams, 3.0, 9, planes=D)
    rel = np.abs(d.cpu().numpy() - want_d) / np.abs(want_d)
    # These weights are the most error-sensitive set in the suite (CPU emulation of the operand roundings,
    # tools/emulate_modes.py: tf32 mean 1.6e-3 / max 1e-2; f16f8 mean 5.7e-5 / max 2.9e-4, which the fp32-FMA SIMT twin
    # reproduces on the GPU).  With the correction steps accumulated first (engine._corrections_first) the tensor cores
    # stay close to that: tools/acc_floor.py measures f16f8 mean 9.3e-5 / max 3.9e-4, bf16x3 7.5e-5 / 2.2e-4 (the
    # interleaved K order gave 4.4e-4 / 1.1e-3 and 6.2e-4 / 1.4e-3).  The figure moves with the kernel variant a layer runs
    # on (K order of the per-tap vs row-halo programs): untuned default variants measure mean 2.1e-4 / max 6.8e-4.
    assert np.percentile(rel, 99) < 7e-4 and rel.max() < 1e-3 and rel.mean() < 3e-4, \
        (np.percentile(rel, 99), rel.max(), rel.mean())


def test_pipeline_with_halo_kernel_forced(monkeypatch):
    """Every eligible stride-1 layer through csrc/conv_halo.cu (incl. the side-tensor sources) vs the per-tap kernel."""
    from cnmnet_b200 import engine
    B, H, W, V = 1, 64, 64, 2
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(99, B, H, W, V)
    dsd = synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 41, 0.02)
    rsd = synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), 42, 0.02)
    outs = {}
    for mode in ("off", "force"):
        monkeypatch.setattr(engine, "HALO_MODE", mode)
        dnet = cnmnet_b200.depthNet(idepth_scale=3.0, precision="tf32")
        rnet = cnmnet_b200.DepthRefineNet(idepth_scale=3.0, precision="tf32")
        dnet.load_state_dict(dsd)
        rnet.load_state_dict(rsd)
        pipe = cnmnet_b200.CNMPipeline(dnet.to(dev()), rnet.to(dev()), use_graph=False)
        st = pipe.static_state(B, V, H, W, dev())
        kinds = {op.backend for op in list(st["dplan"].convs.values()) + list(st["rplan"].convs.values())}
        assert ("halo" in kinds) == (mode == "force")
        d, n, p = pipe(ref.to(dev()), srcs.to(dev()), ref_cam.to(dev()), src_cams.to(dev()))
        outs[mode] = (d.cpu().numpy(), p.cpu().numpy())
    # Same operands, different K order.  fp32 accumulation-order noise (~4e-6) flips the tf32 rounding of ~1 % of the
    # activations per layer (one flip = 4.9e-4 relative) and the flips compound through ~20 layers: measured rel-L2
    # between the two kernels grows from 4e-5 (conv1.0) to 8e-4 (last layers) -- tools/compare_halo.py.  So the two
    # paths agree like two tf32 evaluations do (same size as the distance to the fp32 oracle), not bit-wise.
    rel = np.abs(outs["force"][0] - outs["off"][0]) / np.abs(outs["off"][0])
    assert rel.mean() < 2e-3 and np.percentile(rel, 99) < 8e-3, (rel.mean(), rel.max())
    assert np.abs(outs["force"][1] - outs["off"][1]).mean() < 2e-3


def test_infer_stream_matches_single_calls():
    """The double-buffered streaming API returns, batch by batch, what the blocking call returns."""
    B, H, W, V = 2, 32, 64, 2
    dnet = cnmnet_b200.depthNet(idepth_scale=3.0).to(dev())
    rnet = cnmnet_b200.DepthRefineNet(idepth_scale=3.0).to(dev())
    pipe = cnmnet_b200.CNMPipeline(dnet, rnet, use_graph=True)
    batches = [tuple(t.pin_memory() for t in synth.frame_tuples(100 + i, B, H, W, V)) for i in range(5)]
    want = [tuple(o.cpu().clone() for o in pipe(*[t.to(dev()) for t in b])) for b in batches]
    for lanes in (2, 1, 3):
        got = [tuple(o.clone() for o in outs) for outs in pipe.infer_stream(iter(batches), lanes=lanes)]
        assert len(got) == len(want)
        for g, w in zip(got, want):
            for a, b in zip(g, w):
                assert torch.equal(a, b)
    # a consumer that stops early: the generator's cleanup joins the lane streams, a direct call afterwards is correct
    for k, outs in enumerate(pipe.infer_stream(iter(batches))):
        if k == 1:
            break
    again = tuple(o.cpu() for o in pipe(*[t.to(dev()) for t in batches[3]]))
    for a, b in zip(again, want[3]):
        assert torch.equal(a, b)


def test_video_stream_matches_single_calls():
    """VideoDepthStream (sliding reference tuples, padded last batch) returns per frame what a direct call returns."""
    T, H, W, B = 30, 32, 64, 2
    rng = np.random.default_rng(5)
    frames = torch.from_numpy(synth.smooth_images(rng, T, H, W))
    cams = torch.from_numpy(synth.cameras(rng, T, H, W)[0])
    dnet = cnmnet_b200.depthNet(idepth_scale=3.0, precision="f16f8").to(dev())
    rnet = cnmnet_b200.DepthRefineNet(idepth_scale=3.0, precision="f16f8").to(dev())
    pipe = cnmnet_b200.CNMPipeline(dnet, rnet, use_graph=True)
    got = list(cnmnet_b200.VideoDepthStream(pipe, batch=B, views=3).run(frames, cams))
    assert [g[0] for g in got] == [12, 15, 18]                                   # 3 frames: one full batch + a padded one
    for i, depth, normal, prob in got:
        d, n, p = pipe(frames[i:i + 1].expand(B, -1, -1, -1).contiguous().to(dev()),
                       torch.stack((frames[i + 10], frames[i - 10])).unsqueeze(0).expand(B, -1, -1, -1, -1).contiguous().to(dev()),
                       cams[i:i + 1].expand(B, -1, -1, -1).contiguous().to(dev()),
                       torch.stack((cams[i + 10], cams[i - 10])).unsqueeze(0).expand(B, -1, -1, -1, -1).contiguous().to(dev()))
        assert torch.equal(depth, d[0].cpu()) and torch.equal(normal, n[0].cpu()) and torch.equal(prob, p[0].cpu())


def test_weight_updates_are_picked_up():
    B, H, W = 1, 32, 32
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(9, B, H, W, 2)
    net = cnmnet_b200.depthNet(idepth_scale=3.0).to(dev())
    a, _ = net(ref.to(dev()), srcs[:, 0].to(dev()), ref_cam.to(dev()), src_cams[:, 0].to(dev()))
    with torch.no_grad():
        net.disp1[0].bias.add_(1.0)
    b, _ = net(ref.to(dev()), srcs[:, 0].to(dev()), ref_cam.to(dev()), src_cams[:, 0].to(dev()))
    assert not torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_shape_validation():
    net = cnmnet_b200.depthNet(idepth_scale=3.0).to(dev())
    z = torch.zeros((1, 3, 40, 64), device=dev())
    cam = torch.zeros((1, 2, 4, 4), device=dev())
    with pytest.raises(ValueError):
        net(z, z, cam, cam)                                   # 40 is not divisible by 32
    with pytest.raises(ValueError):
        cnmnet_b200.depthNet(idepth_scale=5.0).to(dev())(z[:, :, :32], z[:, :, :32], cam, cam)


# ------------------------------------------------------------------ the benched configurations, end to end
def _fresh_pipeline(dsd, rsd, precision="f16f8", planes=64, graph=True):
    dnet = cnmnet_b200.depthNet(idepth_scale=3.0, precision=precision, planes=planes)
    rnet = cnmnet_b200.DepthRefineNet(idepth_scale=3.0, precision=precision)
    dnet.load_state_dict(dsd)
    rnet.load_state_dict(rsd)
    return cnmnet_b200.CNMPipeline(dnet.to(dev()), rnet.to(dev()), k_size=9, use_graph=graph)


@pytest.mark.parametrize("tune_mode", ["cache", "off"])
@pytest.mark.parametrize("name,B,V", [("C4", 5, 2), ("C1", 1, 2), ("seven_views", 2, 6)])
def test_headline_configs_end_to_end(monkeypatch, name, B, V, tune_mode):
    """BASELINE configs at full size, through the public pipeline call, against the fp32 oracle: C4 = the benched
    configuration (256x192, 64 planes, batch 5, 2 source views, eval.py:440-455), C1 = the same at batch 1, and the
    seven-view front end (6 source views stacked on the batch axis, thirds averaged, eval.py:885-929).  Responsive
    weights (heads x0.02) so that the final depth actually depends on every layer.  Run with the
    shipped kernel-variant table ("cache": what bench.py runs) and with the untuned defaults ("off"): the variants
    differ in K order, so both must stay inside the north-star 1e-3 on the maximum."""
    from cnmnet_b200 import engine
    monkeypatch.setattr(engine, "TUNE_MODE", tune_mode)
    H, W = 192, 256
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(2468 + B + V, B, H, W, V)
    dsd = synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 71, 0.02)
    rsd = synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), 72, 0.02)
    pipe = _fresh_pipeline(dsd, rsd)
    d, n, p = pipe(ref.to(dev()), srcs.to(dev()), ref_cam.to(dev()), src_cams.to(dev()))
    st = pipe.static_state(B, V, H, W, dev())
    kinds = {op.backend for op in list(st["dplan"].convs.values()) + list(st["rplan"].convs.values())}
    if tune_mode == "cache" and name == "C4":
        assert "halo2" in kinds and all(op.tuned is not None for op in st["dplan"].convs.values()), kinds
    want_d, want_n, want_p = _oracle_pipeline(dsd, rsd, ref, srcs, ref_cam, src_cams, 3.0, 9)
    rel = np.abs(d.cpu().numpy() - want_d) / np.abs(want_d)
    print("%s/%s: depth rel err max %.2e p99 %.2e mean %.2e" % (name, tune_mode, rel.max(), np.percentile(rel, 99), rel.mean()))
    assert rel.max() < 1e-3 and np.percentile(rel, 99) < 3e-4 and rel.mean() < 1e-4, \
        (rel.max(), np.percentile(rel, 99), rel.mean())
    assert np.abs(p.cpu().numpy() - want_p).max() < 2e-3
    cos = np.clip((n.cpu().numpy() * want_n).sum(1), -1, 1)
    assert np.median(np.degrees(np.arccos(cos))) < 0.5


def test_weight_reload_invalidates_captured_graph():
    """Head biases and the f16f8 accumulator scale are launch arguments, frozen when the step is captured: reloading
    weights after the first graph replay must re-capture.  Compared with a fresh pipeline without graphs."""
    B, H, W, V = 1, 64, 64, 2
    ref, srcs, ref_cam, src_cams = [t.to(dev()) for t in synth.frame_tuples(31, B, H, W, V)]
    sd_a = (synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 81, 0.02),
            synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), 82, 0.02))
    sd_b = (synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 83, 0.02),
            synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), 84, 0.02))
    for sd in sd_b:                                  # different magnitude -> different accumulator exponent; biased heads
        for k in sd:
            if k.endswith(".weight") and sd[k].dim() == 4 and sd[k].shape[0] > 1:
                sd[k] = sd[k] * 3.0
            if k.endswith(".0.bias"):
                sd[k] = sd[k] + 0.25
    pipe = _fresh_pipeline(*sd_a, graph=True)
    pipe(ref, srcs, ref_cam, src_cams)
    pipe(ref, srcs, ref_cam, src_cams)               # graph replay with the first weights
    pipe.depth_network.load_state_dict(sd_b[0])
    pipe.refine_network.load_state_dict(sd_b[1])
    got = pipe(ref, srcs, ref_cam, src_cams)
    want = _fresh_pipeline(*sd_b, graph=False)(ref, srcs, ref_cam, src_cams)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


_HASH_SCRIPT = r"""
import hashlib, sys, torch
sys.path.insert(0, %r)
import cnmnet_b200
from cnmnet_b200 import synth
from oracle import nets as onets
dev = torch.device("cuda", 0)
B, V, H, W = 5, 2, 192, 256
ref, srcs, ref_cam, src_cams = synth.frame_tuples(2475, B, H, W, V)
dnet = cnmnet_b200.depthNet(idepth_scale=3.0)
rnet = cnmnet_b200.DepthRefineNet(idepth_scale=3.0)
dnet.load_state_dict(synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 71, 0.02))
rnet.load_state_dict(synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), 72, 0.02))
pipe = cnmnet_b200.CNMPipeline(dnet.to(dev), rnet.to(dev), k_size=9)
h = hashlib.sha256()
for t in pipe(ref.to(dev), srcs.to(dev), ref_cam.to(dev), src_cams.to(dev)):
    h.update(t.cpu().numpy().tobytes())
print("HASH", h.hexdigest())
"""


def test_outputs_are_bit_identical_across_processes():
    """Kernel variants are chosen from the shipped table, never by timing: two fresh processes running the benched
    configuration produce the same bytes (round 1's autotuner could pick different K orders per process)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k not in ("CNM_AUTOTUNE", "CNM_TUNE_CACHE")}
    hashes = []
    for _ in range(2):
        out = subprocess.run([sys.executable, "-c", _HASH_SCRIPT % root], capture_output=True, text=True, env=env,
                             timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        hashes.append([ln for ln in out.stdout.splitlines() if ln.startswith("HASH")][-1])
    assert hashes[0] == hashes[1]


def test_programmatic_dependent_launch_does_not_change_results():
    """cnm_set_pdl: with the step kernels chained by programmatic dependent launch (every kernel waits before its first
    global-memory access) or in plain stream order, eagerly and as a captured graph, the outputs are the same bytes."""
    from cnmnet_b200 import _lib
    lib = _lib.load()
    B, H, W, V = 2, 96, 128, 2
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(21, B, H, W, V)
    dsd = synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 41, 0.02)
    rsd = synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), 42, 0.02)
    args = [t.to(dev()) for t in (ref, srcs, ref_cam, src_cams)]
    outs = []
    try:
        for pdl in (1, 0):
            for graph in (False, True):
                lib.cnm_set_pdl(pdl)
                pipe = _fresh_pipeline(dsd, rsd, graph=graph)
                for _ in range(3):   # the third call of a graph pipeline is a replay
                    d, n, p = pipe(*args)
                torch.cuda.synchronize()
                outs.append(torch.cat([d.flatten(), n.flatten(), p.flatten()]).clone())
    finally:
        lib.cnm_set_pdl(1)
    assert torch.isfinite(outs[0]).all()
    for o in outs[1:]:
        assert torch.equal(outs[0].view(torch.int32), o.view(torch.int32))


def test_range_check_counts_fp16_overflow():
    """CNM_F16F8 stores activations as fp16 (saturating at 65504); the debug range check counts convolution outputs
    beyond that range: zero on ordinary weights, non-zero when a layer is scaled out of range."""
    from cnmnet_b200 import _lib
    lib = _lib.load()
    B, H, W = 1, 32, 64
    ref, srcs, ref_cam, src_cams = [t.to(dev()) for t in synth.frame_tuples(12, B, H, W, 2)]
    sd = synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 91, 0.02)
    lib.cnm_set_range_check(1)
    try:
        assert lib.cnm_range_overflow_count(1) >= 0
        net = cnmnet_b200.depthNet(idepth_scale=3.0, precision="f16f8")
        net.load_state_dict(sd)
        net.to(dev())(ref, srcs[:, 0].contiguous(), ref_cam, src_cams[:, 0].contiguous())
        assert lib.cnm_range_overflow_count(1) == 0
        hot = dict(sd)
        hot["conv1.1.weight"] = sd["conv1.1.weight"] * 1.0e5      # BatchNorm gamma of conv1.0: activations ~1e5
        net2 = cnmnet_b200.depthNet(idepth_scale=3.0, precision="f16f8")
        net2.load_state_dict(hot)
        net2.to(dev())(ref, srcs[:, 0].contiguous(), ref_cam, src_cams[:, 0].contiguous())
        assert lib.cnm_range_overflow_count(1) > 0
    finally:
        lib.cnm_set_range_check(0)
    assert lib.cnm_range_overflow_count(0) == -1


@pytest.mark.parametrize("switch,value", [("SIDE_DISP", True), ("FOLD_UPSAMPLE", False), ("DENSE_ROWTAPS", False),
                                           ("FOLD_MIN_BLOCKS", 2)])
def test_pipeline_matches_oracle_with_layer_program_switches(monkeypatch, switch, value):
    """The alternative layer programs behind engine switches -- disparity input in the epilogue (CNM_SIDE_DISP=1), upsample2x
    + conv3x3 instead of the folded upconv, one kernel row of image taps per K-step, folded upconv on small maps (every
    tile a border tile) -- give the same depth as the oracle, within the tolerance of the default program."""
    from cnmnet_b200 import engine
    monkeypatch.setattr(engine, switch, value)
    B, H, W, V = 2, 96, 160, 2
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(77, B, H, W, V)
    dsd = synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 51, 0.02)
    rsd = synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), 52, 0.02)
    pipe = _fresh_pipeline(dsd, rsd, graph=False)
    d, n, p = pipe(ref.to(dev()), srcs.to(dev()), ref_cam.to(dev()), src_cams.to(dev()))
    st = pipe.static_state(B, V, H, W, dev())
    kinds = {op.backend for op in list(st["dplan"].convs.values()) + list(st["rplan"].convs.values())}
    if switch == "FOLD_MIN_BLOCKS":
        assert "halo2f" in kinds or "halof" in kinds
    if switch == "SIDE_DISP":
        assert st["dplan"].convs["iconv1.0"].desc.side_disp
    want_d, want_n, want_p = _oracle_pipeline(dsd, rsd, ref, srcs, ref_cam, src_cams, 3.0, 9)
    rel = np.abs(d.cpu().numpy() - want_d) / np.abs(want_d)
    assert rel.max() < 1e-3 and np.percentile(rel, 99) < 3e-4 and rel.mean() < 1e-4, (rel.max(), np.percentile(rel, 99), rel.mean())
    assert np.abs(p.cpu().numpy() - want_p).max() < 2e-3
