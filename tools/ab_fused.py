"""Interleaved A/B of the C4 step with and without the fused plane sweep (conv1.0 = halo2s vs halo2), one process,
alternating blocks of timed steps so that clock / power drift hits both arms alike.
    python tools/ab_fused.py [rounds] [steps per block]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import cnmnet_b200  # noqa: E402
from cnmnet_b200 import engine, synth  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
B, V, H, W = bench.B_PER_GPU, bench.V, bench.H, bench.W
host = synth.frame_tuples(1234, B, H, W, V)
arms = {}
for name, mode in (("unfused", "1"), ("fused", "force")):
    engine.FUSE_SWEEP = mode
    d, r = bench.make_weights(0)
    pipe = cnmnet_b200.CNMPipeline(d.to(dev), r.to(dev), k_size=9, use_graph=True)
    st = pipe.static_state(B, V, H, W, dev)
    pipe.load_inputs(st, *host, non_blocking=False)
    for _ in range(3):
        pipe.step(st)
    arms[name] = (pipe, st, st["dplan"].convs["conv1.0"].backend)
torch.cuda.synchronize()
flush_buf = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
res = {k: [] for k in arms}
for rd in range(rounds):
    for name in (("unfused", "fused") if rd % 2 == 0 else ("fused", "unfused")):
        pipe, st, _ = arms[name]
        ms, _ = bench.time_events(lambda: pipe.step(st), steps, lambda: flush_buf.zero_())
        res[name].append(ms)
for name, (_, _, be) in arms.items():
    v = res[name]
    print("%-8s conv1.0 = %-6s  ms/step per block: %s   median %.3f  min %.3f" %
          (name, be, " ".join("%.3f" % x for x in v), statistics.median(v), min(v)))
a, b = statistics.median(res["unfused"]), statistics.median(res["fused"])
print("fused / unfused = %.4f  (%+.1f %% step time)" % (b / a, 100.0 * (b / a - 1.0)))
