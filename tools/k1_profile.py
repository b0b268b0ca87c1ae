"""Two cnm_plane_sweep launches at BASELINE config 2 (seeds 1234 and 77) and one first-generation launch, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:plane_sweep\\|cost_volume -o gpurun_out/prof_k1 python tools/k1_profile.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cnmnet_b200 import _lib, synth  # noqa: E402

dev = torch.device("cuda", 0)
lib, s = _lib.load(), _lib.stream_ptr()
B, V, H, W, D = 5, 2, 192, 256, 64
out = torch.empty((V * B, D, H, W), dtype=torch.float32, device=dev)
for seed in (1234, 77):
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(seed, B, H, W, V)
    src_vm = srcs.transpose(0, 1).contiguous().to(dev)
    cam_vm = src_cams.transpose(0, 1).contiguous().to(dev)
    refd, rcd = ref.to(dev), ref_cam.to(dev)
    hk = torch.empty((V * B, 12), dtype=torch.float32, device=dev)
    _lib.check(lib.cnm_camera_setup(rcd.data_ptr(), cam_vm.data_ptr(), hk.data_ptr(), B, V, s))
    _lib.check(lib.cnm_plane_sweep(refd.data_ptr(), src_vm.data_ptr(), hk.data_ptr(), out.data_ptr(), B, V, H, W, D, 0.1, 3.0,
                                   0, 0, s))
    torch.cuda.synchronize()
tex = torch.empty((lib.cnm_packed_source_bytes(V * B, H, W),), dtype=torch.uint8, device=dev)
_lib.check(lib.cnm_pack_source(src_vm.data_ptr(), tex.data_ptr(), V * B, H, W, s))
_lib.check(lib.cnm_cost_volume(refd.data_ptr(), tex.data_ptr(), hk.data_ptr(), out.data_ptr(), B, V, H, W, D, 0.1, 3.0, 0, 0, s))
torch.cuda.synchronize()
print("done")
