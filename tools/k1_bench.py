"""Times the two cost-volume kernels at BASELINE config 2 (256x192, 64 planes, 2 views, batch 5) and at C5 size:
    python tools/k1_bench.py            -> gpurun_out/k1_bench.json
'sweep' = cnm_plane_sweep (TMA-staged), 'gather' = cnm_cost_volume (first generation, taps through L1; the source repack
is timed separately).  CUDA events, L2 flushed between repetitions."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cnmnet_b200 import _lib, engine, synth  # noqa: E402

dev = torch.device("cuda", 0)
lib, s = _lib.load(), _lib.stream_ptr()
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps=20):
    fn()
    ts = []
    for _ in range(reps):
        flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.mean(ts), min(ts)


rows = []
for name, B, V, H, W, D, seed in (("C2", 5, 2, 192, 256, 64, 1234), ("C2-seed77", 5, 2, 192, 256, 64, 77),
                                  ("C5-b8", 8, 2, 480, 640, 128, 1234)):
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(seed, B, H, W, V)
    src_vm = srcs.transpose(0, 1).contiguous().to(dev)
    cam_vm = src_cams.transpose(0, 1).contiguous().to(dev)
    refd, rcd = ref.to(dev), ref_cam.to(dev)
    hk = torch.empty((V * B, 12), dtype=torch.float32, device=dev)
    _lib.check(lib.cnm_camera_setup(rcd.data_ptr(), cam_vm.data_ptr(), hk.data_ptr(), B, V, s))
    tex = torch.empty((lib.cnm_packed_source_bytes(V * B, H, W),), dtype=torch.uint8, device=dev)
    _lib.check(lib.cnm_pack_source(src_vm.data_ptr(), tex.data_ptr(), V * B, H, W, s))
    out = torch.empty((V * B, D, H, W), dtype=torch.float32, device=dev)
    act = torch.empty((V * B, H, W, 2 * D), dtype=torch.float16, device=dev)
    alg = B * (V * 4 * D * H * W + (V + 1) * 12 * H * W)
    for layout, dt, buf, tag in ((_lib.VOLUME_NCHW_F32, 0, out, "nchw_f32"), (_lib.VOLUME_ACT, _lib.CNM_F16F8, act, "act_f16f8")):
        for kern in ("sweep", "gather"):
            if kern == "sweep":
                fn = lambda: _lib.check(lib.cnm_plane_sweep(refd.data_ptr(), src_vm.data_ptr(), hk.data_ptr(), buf.data_ptr(),
                                                            B, V, H, W, D, 0.1, 3.0, layout, dt, s))
            else:
                fn = lambda: _lib.check(lib.cnm_cost_volume(refd.data_ptr(), tex.data_ptr(), hk.data_ptr(), buf.data_ptr(),
                                                            B, V, H, W, D, 0.1, 3.0, layout, dt, s))
            mean, best = timeit(fn)
            rows.append({"config": name, "layout": tag, "kernel": kern, "us_mean": mean, "us_min": best,
                         "algorithmic_bytes": alg, "gbs": alg / mean / 1e3, "frac_of_hbm_peak": alg / mean / 1e3 / peak})
            print(rows[-1], flush=True)
    mean, best = timeit(lambda: _lib.check(lib.cnm_pack_source(src_vm.data_ptr(), tex.data_ptr(), V * B, H, W, s)))
    rows.append({"config": name, "kernel": "pack_source (gather path only)", "us_mean": mean, "us_min": best})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/k1_bench.json", "w"), indent=1)
