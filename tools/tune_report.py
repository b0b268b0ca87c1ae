"""Prints the autotuner's timing table (kernel x tile) for every conv layer of the C4 plans."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import cnmnet_b200  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "f16f8"
dev = torch.device("cuda", 0)
dnet, rnet = bench.make_weights(0)
dnet.precision = rnet.precision = cnmnet_b200.engine.parse_dtype(precision)
pipe = cnmnet_b200.CNMPipeline(dnet.to(dev), rnet.to(dev), k_size=9, use_graph=False)
st = pipe.static_state(bench.B_PER_GPU, bench.V, bench.H, bench.W, dev)
for tag, plan in (("depthnet", st["dplan"]), ("refine", st["rplan"])):
    for name, op in plan.convs.items():
        if op.tuned and op.tuned[3]:
            row = "  ".join("%s/%d%s %.1f" % (be, bn, "k%d" % ks if ks > 1 else "", 1e3 * t)
                            for t, be, bn, ks in sorted(op.tuned[3], key=lambda r: (r[1], -r[2], r[3])))
            print("%-9s %-18s best %s/%d/k%d | %s" % (tag, name, op.tuned[0], op.tuned[1], op.tuned[2], row))
if os.environ.get("CNM_SAVE_TUNE_CACHE"):
    cnmnet_b200.engine.save_tune_cache(os.environ["CNM_SAVE_TUNE_CACHE"])
