"""Pins the CPU oracle (oracle/) against outputs of the unmodified reference, committed under
tests/golden/ by oracle/make_golden.py.  CPU only."""
import os

import numpy as np
import pytest
import torch

from cnmnet_b200 import synth
from oracle import cost_volume as ocv
from oracle import nets, normals as onrm


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("tag", ["s3", "s2"])
def test_cost_volume_matches_reference(golden_dir, tag):
    g = _load(golden_dir, "cost_volume_%s.npz" % tag)
    idmin, idmax = ocv.idepth_range(float(g["idepth_scale"]))
    vol64 = ocv.cost_volume(g["left"], g["right"], g["left_cam"], g["right_cam"], 64, idmin, idmax, np.float64)
    ref = g["volume"]
    # the reference is fp32: its own noise floor vs exact arithmetic is rel-L2 ~2e-6 on smooth images
    assert rel_l2(vol64, ref) < 1e-5
    err = np.abs(vol64 - ref)
    assert np.mean(err) < 2e-5
    # camera pre-processing: a = Hm.[x,y,1] sampled every 97th pixel of the x-major grid, and KT
    Hm, kt = ocv.homography(g["left_cam"], g["right_cam"])
    H, W = g["left"].shape[2:]
    flat = np.arange(0, H * W, 97)
    xs, ys = flat // H, flat % H  # x-major flattening (depth_util.py:15-18)
    a = np.einsum("bij,jn->bin", Hm, np.stack([xs, ys, np.ones_like(xs)]).astype(np.float64))
    assert np.allclose(a, g["KRKiUV"], rtol=2e-4, atol=2e-3)
    assert np.allclose(kt[..., None], g["KT"], rtol=2e-4, atol=2e-4)


def test_cost_volume_fp32_mode_is_reference_noise(golden_dir):
    g = _load(golden_dir, "cost_volume_s3.npz")
    vol32 = ocv.cost_volume(g["left"], g["right"], g["left_cam"], g["right_cam"], 64, 0.1, 3.0, np.float32)
    assert rel_l2(vol32, g["volume"]) < 2e-5


@pytest.mark.parametrize("k", [9, 5])
def test_normals_match_reference(golden_dir, k):
    g = _load(golden_dir, "normals_k%d.npz" % k)
    n64, p64, det = onrm.depth2normal(g["depth"], g["Kinv"], k, np.float64, return_det=True)
    ref_n = g["normals"]
    det = np.broadcast_to(det[:, None], ref_n.shape)
    # pixels whose NaN pattern agrees; reference is fp32 with cond(AtA) ~ 1e4 -> ~3e-4 mean error
    finite = np.isfinite(ref_n) & np.isfinite(n64)
    assert finite.mean() > 0.99
    stable = finite & (np.abs(det - 1e-5) > 1e-6)
    err = np.abs(n64 - ref_n)[stable]
    assert err.mean() < 2e-3, err.mean()
    assert np.percentile(err, 99) < 2e-2
    # all-invalid windows give exactly zero normals in both
    zero_ref = (ref_n == 0).all(axis=1)
    zero_orc = (n64 == 0).all(axis=1)
    assert (zero_ref == zero_orc).mean() > 0.999
    # points: pure fp32 products
    pf = np.isfinite(g["points"])
    assert np.allclose(p64[pf], g["points"][pf], rtol=1e-5, atol=1e-5)


def test_cnm_and_planes_match_reference(golden_dir):
    g = _load(golden_dir, "normals_k9.npz")
    out = onrm.get_normal_by_planes(g["gt_like"], g["segs"], g["planes_num"])
    assert np.allclose(out, g["normals_by_planes"], rtol=1e-4, atol=1e-5)
    # CNM on the reference's own (fp32) plane-fit normals isolates the replace-by-mean restatement
    n_in = np.moveaxis(g["normals"], 1, -1)
    rep, loss = onrm._replace_by_plane_mean(n_in, g["segs"], g["planes_num"], True, np.float64)
    ok = np.isfinite(g["normals_cnm"])
    assert np.allclose(np.moveaxis(rep, -1, 1)[ok], g["normals_cnm"][ok], rtol=1e-4, atol=2e-5)
    if np.isfinite(g["loss"]):
        assert abs(float(loss) - float(g["loss"])) < 1e-3 * max(1.0, abs(float(g["loss"])))


def _state(layers, g):
    sd = synth.kaiming_state(nets.state_shapes(layers), int(g["weight_seed"]), float(g["head_scale"]))
    tot = sum(float(v.double().sum()) for v in sd.values() if v.dtype.is_floating_point)
    asum = sum(float(v.double().abs().sum()) for v in sd.values() if v.dtype.is_floating_point)
    assert np.allclose([tot, asum], g["weight_checksum"], rtol=1e-9), "synthetic weights are not reproducible here"
    return sd


def test_depthnet_forward_matches_reference(golden_dir):
    g = _load(golden_dir, "depthnet_fwd.npz")
    sd = _state(nets.DEPTHNET_LAYERS, g)
    t = lambda k: torch.from_numpy(g[k])
    with torch.no_grad():
        disps, iconv1 = nets.depthnet_forward(sd, t("left"), t("right"), t("left_cam"), t("right_cam"), 3.0)
    assert rel_l2(iconv1.numpy(), g["iconv1"]) < 2e-4
    for i, k in enumerate(["disp1", "disp2", "disp3", "disp4"]):
        assert disps[i].shape == g[k].shape
        assert np.max(np.abs(disps[i].numpy() - g[k]) / np.abs(g[k])) < 1e-3, k
    # the responsive weights keep the sigmoid unsaturated, so the comparison is not vacuous
    assert 0.2 < float(g["disp1"].min()) and float(g["disp1"].max()) < 2.8


def test_refinenet_forward_matches_reference(golden_dir):
    g = _load(golden_dir, "refinenet_fwd.npz")
    sd = _state(nets.REFINENET_LAYERS, g)
    t = lambda k: torch.from_numpy(g[k])
    with torch.no_grad():
        dr, pm, vol = nets.refinenet_forward(sd, t("idepth01"), t("idepth02"), t("iconv01"), t("iconv02"), 3.0,
                                             return_volume=True)
    assert rel_l2(vol.numpy(), g["iconv1_depth"]) < 1e-4
    assert np.max(np.abs(dr.numpy() - g["disp_refined"]) / np.abs(g["disp_refined"])) < 1e-4
    assert np.max(np.abs(pm.numpy() - g["prob_map"])) < 1e-4


def test_reference_init_stream(golden_dir):
    g = _load(golden_dir, "init_stats.npz")
    for tag, layers in (("depthnet", nets.DEPTHNET_LAYERS), ("refinenet", nets.REFINENET_LAYERS)):
        sd = nets.reference_init_state(layers, 0)
        n = sum(v.numel() for k, v in sd.items() if "running" not in k and "num_batches" not in k)
        assert n == int(g[tag + "_nparams"])
        tot = sum(float(v.double().sum()) for v in sd.values() if v.dtype.is_floating_point)
        asum = sum(float(v.double().abs().sum()) for v in sd.values() if v.dtype.is_floating_point)
        # torch's CPU normal_ may take a different vector path on another CPU: tolerance, not bit-equality
        assert np.allclose([tot, asum], g[tag + "_checksum"], rtol=1e-4, atol=1.0)
        assert np.allclose(sd["conv1.0.weight"].flatten()[:16].numpy(), g[tag + "_conv1_head"], rtol=1e-4, atol=1e-6)


def test_inverse_warp_matches_reference(golden_dir):
    g = _load(golden_dir, "inverse_warp.npz")
    t = lambda k: torch.from_numpy(g[k])
    out = nets.inverse_warp(t("feat"), t("depth"), t("pose"), t("K"), t("Kinv"))
    assert np.allclose(out.numpy(), g["warped"], rtol=1e-4, atol=1e-4)


def test_fast_torch_ports_match_reference(golden_dir):
    """The fp32 torch ports used for the timed CPU baseline reproduce the reference's outputs."""
    g = _load(golden_dir, "cost_volume_s3.npz")
    t = lambda k: torch.from_numpy(g[k])
    vol = nets.cost_volume_torch(t("left"), t("right"), t("left_cam"), t("right_cam"), 3.0)
    assert rel_l2(vol.numpy(), g["volume"]) < 1e-6
    g = _load(golden_dir, "normals_k9.npz")
    n, p = nets.depth2normal_torch(torch.from_numpy(g["depth"]), torch.from_numpy(g["Kinv"]), 9)
    fin = np.isfinite(g["normals"]) & np.isfinite(n.numpy())
    assert fin.mean() > 0.99 and np.abs(n.numpy() - g["normals"])[fin].mean() < 1e-4


def test_depth_metrics_oracle_matches_reference_functions(golden_dir):
    """oracle/metrics.py vs the reference's own utils/metric.py functions called as eval.py:1030-1047 does
    (fixture written by oracle/make_golden_metrics.py): clamping, invalid / non-finite ground truth, an empty image."""
    from oracle import metrics as omet
    g = np.load(os.path.join(golden_dir, "depth_metrics.npz"))
    per, _ = omet.sequence_metrics(g["pred"], g["gt"])
    for b, m in enumerate(per):
        want = g["metrics"][b]
        assert m["num_valid"] == int(want[0])
        got = np.array([m[k] for k in omet.NAMES])
        if want[0] == 0:
            assert np.isnan(got).all() and np.isnan(want[1:]).all()
        else:
            assert np.allclose(got, want[1:], rtol=1e-12, atol=0)


def test_losses_oracle_matches_reference_module(golden_dir):
    """oracle/losses.py vs the reference's own depthnet/losses.py (fixture written by oracle/make_golden_losses.py);
    the reference evaluates in fp32, the oracle in fp64."""
    from oracle import losses as olos
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    t = {k: torch.from_numpy(g[k]) for k in ("gt", "pred", "prob", "disp2", "disp3", "disp4", "gt_pos", "n_pred", "n_gt", "valid")}
    got = {"idepth": olos.idepth_loss(t["pred"], t["gt"]), "idepth_log": olos.idepth_loss(t["pred"], t["gt"], log=True),
           "idepth_prob": olos.idepth_prob_loss(t["pred"], t["gt"], t["prob"]),
           "idepth_prob_log": olos.idepth_prob_loss(t["pred"], t["gt"], t["prob"], log=True),
           "idepth_234": olos.idepth_loss_234([None, t["disp2"], t["disp3"], t["disp4"]], t["gt_pos"])}
    got["normal"], got["normal_angle"] = olos.surface_normal_loss(t["n_pred"], t["n_gt"], t["valid"])
    got["normal_prob"], got["normal_prob_angle"] = olos.surface_normal_loss(t["n_pred"], t["n_gt"], t["valid"], t["prob"])
    for k, v in got.items():
        assert abs(v - float(g[k])) < 2e-6 * max(1.0, abs(float(g[k]))), (k, v, float(g[k]))
