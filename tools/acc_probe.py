"""Accuracy probe: final-depth error of a precision mode vs the fp32 oracle for several sizes / weight seeds."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cnmnet_b200  # noqa: E402
from cnmnet_b200 import engine, synth  # noqa: E402
from oracle import nets as onets  # noqa: E402
from test_gpu_nets import _oracle_pipeline  # noqa: E402

dev = torch.device("cuda", 0)
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16x3"]
for (H, W, B, ws, halo) in [(64, 96, 2, 21, "auto"), (64, 96, 2, 51, "auto"), (96, 160, 3, 21, "auto"), (96, 160, 3, 51, "auto"),
                            (96, 160, 3, 51, "off"), (96, 128, 2, 51, "auto")]:
    engine.HALO_MODE = halo
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(500 + H, B, H, W, 2)
    dsd = synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), ws, 0.02)
    rsd = synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), ws + 1, 0.02)
    want_d, want_n, want_p = _oracle_pipeline(dsd, rsd, ref, srcs, ref_cam, src_cams, 3.0, 9)
    for mode in modes:
        dnet = cnmnet_b200.depthNet(idepth_scale=3.0, precision=mode)
        rnet = cnmnet_b200.DepthRefineNet(idepth_scale=3.0, precision=mode)
        dnet.load_state_dict(dsd)
        rnet.load_state_dict(rsd)
        pipe = cnmnet_b200.CNMPipeline(dnet.to(dev), rnet.to(dev), k_size=9, use_graph=False)
        d, n, p = pipe(ref.to(dev), srcs.to(dev), ref_cam.to(dev), src_cams.to(dev))
        rel = np.abs(d.cpu().numpy() - want_d) / np.abs(want_d)
        print("%dx%d B=%d wseed=%d halo=%s %-7s depth rel err mean %.2e p99 %.2e max %.2e" %
              (H, W, B, ws, halo, mode, rel.mean(), np.percentile(rel, 99), rel.max()), flush=True)
