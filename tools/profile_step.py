"""One un-graphed pipeline step (after one warm-up) for ncu captures:
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc -c 3 \
        -o gpurun_out/prof python tools/profile_step.py
Also runs the C2 cost-volume micro-benchmark (NCHW fp32) and the C3 depth->normal call once."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import cnmnet_b200  # noqa: E402
from cnmnet_b200 import _lib, synth  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "f16f8"
dev = torch.device("cuda", 0)
B, V, H, W = bench.B_PER_GPU, bench.V, bench.H, bench.W
dnet, rnet = bench.make_weights(0)
dnet.precision = rnet.precision = cnmnet_b200.engine.parse_dtype(precision)
pipe = cnmnet_b200.CNMPipeline(dnet.to(dev), rnet.to(dev), k_size=9, use_graph=False)
st = pipe.static_state(B, V, H, W, dev)
pipe.load_inputs(st, *synth.frame_tuples(1234, B, H, W, V), non_blocking=False)
pipe.step(st)  # warm-up (lazy attribute setup)
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: the capture starts here
pipe.step(st)
torch.cuda.synchronize()
lib, s = _lib.load(), _lib.stream_ptr()
dplan = st["dplan"]
vol = torch.empty((V * B, 64, H, W), dtype=torch.float32, device=dev)
srcd = st["src"].view(V * B, 3, H, W)
# the C2 cost-volume micro-benchmark (NCHW fp32): second-generation kernel, then the first-generation pair
_lib.check(lib.cnm_plane_sweep(st["ref"].data_ptr(), srcd.data_ptr(), dplan.hk.data_ptr(), vol.data_ptr(), B, V, H, W, 64, 0.1,
                               3.0, _lib.VOLUME_NCHW_F32, 0, s))
tex = torch.empty((lib.cnm_packed_source_bytes(V * B, H, W),), dtype=torch.uint8, device=dev)
_lib.check(lib.cnm_pack_source(srcd.data_ptr(), tex.data_ptr(), V * B, H, W, s))
_lib.check(lib.cnm_cost_volume(st["ref"].data_ptr(), tex.data_ptr(), dplan.hk.data_ptr(), vol.data_ptr(), B, V, H, W, 64, 0.1,
                               3.0, _lib.VOLUME_NCHW_F32, 0, s))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", _lib.launch_count())
