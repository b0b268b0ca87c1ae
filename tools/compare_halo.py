"""Layer-by-layer comparison of the per-tap and row-halo tensor-core kernels on one DepthNet + RefineNet pass."""
import os
os.environ["CNM_PLAN_REUSE"] = "0"   # intermediate activations are compared: keep every buffer
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import cnmnet_b200  # noqa: E402
from cnmnet_b200 import engine, synth  # noqa: E402
from oracle import nets as onets  # noqa: E402

dev = torch.device("cuda", 0)
B, H, W, V = 1, 64, 64, 2
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
ref, srcs, ref_cam, src_cams = synth.frame_tuples(99, B, H, W, V)
dsd = synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 41, 0.02)
rsd = synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), 42, 0.02)
outs = {}
for mode in ("off", "force"):
    engine.HALO_MODE = mode
    dnet = cnmnet_b200.depthNet(idepth_scale=3.0, precision=prec)
    rnet = cnmnet_b200.DepthRefineNet(idepth_scale=3.0, precision=prec)
    dnet.load_state_dict(dsd)
    rnet.load_state_dict(rsd)
    pipe = cnmnet_b200.CNMPipeline(dnet.to(dev), rnet.to(dev), use_graph=False)
    st = pipe.static_state(B, V, H, W, dev)
    pipe(ref.to(dev), srcs.to(dev), ref_cam.to(dev), src_cams.to(dev))
    torch.cuda.synchronize()
    outs[mode] = {("d." + k): (op.backend, op.out.float().cpu()) for k, op in st["dplan"].convs.items()}
    outs[mode].update({("r." + k): (op.backend, op.out.float().cpu()) for k, op in st["rplan"].convs.items()})
for k in outs["off"]:
    a, b = outs["off"][k][1], outs["force"][k][1]
    print("%-22s %-5s rel_l2 %.3e  max %.3e  (|a| max %.2f)" % (k, outs["force"][k][0], float((a - b).norm() / a.norm()),
                                                               float((a - b).abs().max()), float(a.abs().max())))
