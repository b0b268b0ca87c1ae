"""Generate tests/golden/*.npz by running the UNMODIFIED reference (via oracle/ref_shim.py)
on seeded synthetic inputs.  Build-container only (needs /root/reference); the produced
fixtures are committed and are what pins the oracle (tests/test_oracle_golden.py).

    python -m oracle.make_golden            # rewrites tests/golden/

Inputs are stored in the fixtures next to the outputs so nothing depends on regenerating
them bit-identically elsewhere.  Network weights are NOT stored (33.9 M parameters): they are
regenerated from ``cnmnet_b200.synth.kaiming_state`` (numpy PCG64, platform-stable) and a
checksum of them is stored instead.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from cnmnet_b200 import synth  # noqa: E402
from oracle import nets  # noqa: E402
from oracle.ref_shim import load_reference, reset_pixel_cache  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def state_checksum(sd):
    tot, asum = 0.0, 0.0
    for k, v in sd.items():
        if v.dtype.is_floating_point:
            tot += float(v.double().sum())
            asum += float(v.double().abs().sum())
    return np.array([tot, asum], dtype=np.float64)


def main():
    warnings.filterwarnings("ignore")
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    torch.set_grad_enabled(False)

    # ---- cost volume (depthNet.getVolume through the reference camera pre-processing) ----
    H, W, B = 32, 64, 2
    rng = np.random.default_rng(101)
    left = synth.smooth_images(rng, B, H, W)
    right = synth.smooth_images(rng, B, H, W)
    lcam, rcam = synth.cameras(rng, B, H, W)
    for scale, tag in ((3.0, "s3"), (2.0, "s2")):
        net = ref.depthNet(idepth_scale=scale)
        pix = ref.get_pixel_coordinates(H, W)
        KRKiUV, KT = ref.process_camera_parameters(torch.from_numpy(lcam), torch.from_numpy(rcam), pix)
        vol = net.getVolume(torch.from_numpy(left), torch.from_numpy(right), KRKiUV, KT)
        np.savez_compressed(os.path.join(OUT, "cost_volume_%s.npz" % tag), left=left, right=right, left_cam=lcam,
                            right_cam=rcam, idepth_scale=np.float64(scale), volume=vol.numpy(),
                            KRKiUV=KRKiUV.numpy()[:, :, ::97].copy(), KT=KT.numpy())

    # ---- depth -> normal, with and without plane segments; get_normal_by_planes ----
    H, W, B = 48, 64, 2
    rng = np.random.default_rng(202)
    depth = synth.depth_maps(rng, B, H, W)
    depth[1, 20, 30] = np.nan
    K = synth.intrinsics(H, W)
    Kinv = np.broadcast_to(np.linalg.inv(K).astype(np.float32), (B, 3, 3)).copy()
    segs, nums = synth.plane_segments(rng, B, H, W)
    for k_size in (9, 5):
        reset_pixel_cache(ref)
        d2n = ref.Depth2normal(k_size)
        normals, points = d2n(torch.from_numpy(depth), torch.from_numpy(Kinv))
        reset_pixel_cache(ref)
        n_cnm, loss, _ = d2n(torch.from_numpy(depth), torch.from_numpy(Kinv),
                             torch.from_numpy(segs.astype(bool)), [int(n) for n in nums])
        gt_like = np.random.default_rng(5).standard_normal((B, 3, H, W)).astype(np.float32)
        n_planes = ref.get_normal_by_planes(torch.from_numpy(gt_like), torch.from_numpy(segs.astype(bool)),
                                            [int(n) for n in nums])
        np.savez_compressed(os.path.join(OUT, "normals_k%d.npz" % k_size), depth=depth, Kinv=Kinv, segs=segs,
                            planes_num=nums, normals=normals.numpy(), points=points.numpy(),
                            normals_cnm=n_cnm.numpy(), loss=np.float64(float(loss)), gt_like=gt_like,
                            normals_by_planes=n_planes.numpy())

    # ---- depthNet forward (weights from synth.kaiming_state, heads x0.02) ----
    H, W, B = 32, 64, 1
    rng = np.random.default_rng(303)
    left = synth.smooth_images(rng, B, H, W)
    right = synth.smooth_images(rng, B, H, W)
    lcam, rcam = synth.cameras(rng, B, H, W)
    net = ref.depthNet(idepth_scale=3.0).eval()
    shapes = nets.state_shapes(nets.DEPTHNET_LAYERS)
    assert list(shapes.keys()) == list(net.state_dict().keys()), "state_dict key order differs from the reference"
    assert all(tuple(net.state_dict()[k].shape) == tuple(v) for k, v in shapes.items())
    sd = synth.kaiming_state(shapes, seed=11, head_scale=0.02)
    net.load_state_dict(sd)
    disps, iconv1 = net(torch.from_numpy(left), torch.from_numpy(right), torch.from_numpy(lcam), torch.from_numpy(rcam))
    np.savez_compressed(os.path.join(OUT, "depthnet_fwd.npz"), left=left, right=right, left_cam=lcam, right_cam=rcam,
                        weight_seed=np.int64(11), head_scale=np.float64(0.02), weight_checksum=state_checksum(sd),
                        disp1=disps[0].numpy(), disp2=disps[1].numpy(), disp3=disps[2].numpy(), disp4=disps[3].numpy(),
                        iconv1=iconv1.numpy())

    # ---- DepthRefineNet forward ----
    rnet = ref.DepthRefineNet(idepth_scale=3.0).eval()
    rshapes = nets.state_shapes(nets.REFINENET_LAYERS)
    assert list(rshapes.keys()) == list(rnet.state_dict().keys())
    assert all(tuple(rnet.state_dict()[k].shape) == tuple(v) for k, v in rshapes.items())
    rsd = synth.kaiming_state(rshapes, seed=12, head_scale=0.02)
    rnet.load_state_dict(rsd)
    rng = np.random.default_rng(404)
    id1 = (1.5 + 0.5 * synth.smooth_images(rng, B, H, W, channels=1)).astype(np.float32)
    id2 = (id1 + 0.1 * synth.smooth_images(rng, B, H, W, channels=1)).astype(np.float32)
    ic1 = np.abs(synth.smooth_images(rng, B, H, W, channels=64)).astype(np.float32)
    ic2 = np.abs(synth.smooth_images(rng, B, H, W, channels=64)).astype(np.float32)
    dr, pm, vol = rnet(torch.from_numpy(id1), torch.from_numpy(id2), torch.from_numpy(ic1), torch.from_numpy(ic2),
                       ReturnVolume=True)
    np.savez_compressed(os.path.join(OUT, "refinenet_fwd.npz"), idepth01=id1, idepth02=id2, iconv01=ic1, iconv02=ic2,
                        weight_seed=np.int64(12), head_scale=np.float64(0.02), weight_checksum=state_checksum(rsd),
                        disp_refined=dr.numpy(), prob_map=pm.numpy(), iconv1_depth=vol.numpy())

    # ---- reference random-init statistics (torch RNG stream; checked with tolerance) ----
    stats = {}
    for tag, ctor, layers in (("depthnet", lambda: ref.depthNet(idepth_scale=3.0), nets.DEPTHNET_LAYERS),
                              ("refinenet", lambda: ref.DepthRefineNet(idepth_scale=3.0), nets.REFINENET_LAYERS)):
        torch.manual_seed(0)
        rs = ctor().state_dict()
        stats[tag + "_checksum"] = state_checksum(rs)
        stats[tag + "_nparams"] = np.int64(sum(v.numel() for k, v in rs.items()
                                               if "running" not in k and "num_batches" not in k))
        first = "conv1.0.weight"
        stats[tag + "_conv1_head"] = rs[first].flatten()[:16].numpy().copy()
    np.savez_compressed(os.path.join(OUT, "init_stats.npz"), **stats)

    # ---- depth-driven inverse_warp ----
    H, W, B, C = 24, 32, 2, 4
    rng = np.random.default_rng(505)
    feat = synth.smooth_images(rng, B, H, W, channels=C)
    depth = synth.depth_maps(rng, B, H, W, with_invalid=False)
    K = synth.intrinsics(H, W).astype(np.float32)
    Kb = np.broadcast_to(K, (B, 3, 3)).copy()
    Kinvb = np.broadcast_to(np.linalg.inv(K.astype(np.float64)).astype(np.float32), (B, 3, 3)).copy()
    lcam, rcam = synth.cameras(rng, B, H, W)
    pose = np.stack([(rcam[b, 0].astype(np.float64) @ np.linalg.inv(lcam[b, 0].astype(np.float64)))[:3, :4]
                     for b in range(B)]).astype(np.float32)
    reset_pixel_cache(ref)
    warped = ref.inverse_warp(torch.from_numpy(feat), torch.from_numpy(depth), torch.from_numpy(pose),
                              torch.from_numpy(Kb), torch.from_numpy(Kinvb))
    np.savez_compressed(os.path.join(OUT, "inverse_warp.npz"), feat=feat, depth=depth, pose=pose, K=Kb, Kinv=Kinvb,
                        warped=warped.numpy())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
