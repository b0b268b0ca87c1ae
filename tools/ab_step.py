"""Interleaved A/B of the C4 step between engine configurations, one process, alternating blocks of timed steps so that
clock / power drift hits every arm alike (a power-capped B200 drifts by several per cent over a run).
    python tools/ab_step.py [--rounds 6] [--steps 10] name:ATTR=value[,ATTR=value...][,tune=table.json] ...
ATTR is a module attribute of cnmnet_b200.engine (FOLD_UPSAMPLE, FOLD_MIN_BLOCKS, FUSE_SWEEP, BALANCE, ...), set while the
arm's pipeline is built; tune= merges a kernel-variant table (the format of cnmnet_b200/tune/b200.json) for that arm only.
    python tools/ab_step.py base: nofold:FOLD_UPSAMPLE=False bal:tune=gpurun_out/tune_bal.json"""
import ast
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import cnmnet_b200  # noqa: E402
from cnmnet_b200 import engine, synth  # noqa: E402

args = sys.argv[1:]
rounds, steps = 6, 10
while args and args[0].startswith("--"):
    if args[0] == "--rounds":
        rounds = int(args[1])
    elif args[0] == "--steps":
        steps = int(args[1])
    args = args[2:]
dev = torch.device("cuda", 0)
B, V, H, W = bench.B_PER_GPU, bench.V, bench.H, bench.W
host = synth.frame_tuples(1234, B, H, W, V)
arms = {}
for spec in args:
    name, _, rest = spec.partition(":")
    saved, table = {}, dict(engine._TUNE_CACHE)
    for item in filter(None, rest.split(",")):
        k, _, v = item.partition("=")
        if k == "tune":
            engine.load_tune_cache(v)
            continue
        saved[k] = getattr(engine, k)
        try:
            setattr(engine, k, ast.literal_eval(v))
        except (ValueError, SyntaxError):
            setattr(engine, k, v)
    d, r = bench.make_weights(0)
    pipe = cnmnet_b200.CNMPipeline(d.to(dev), r.to(dev), k_size=9, use_graph=True)
    st = pipe.static_state(B, V, H, W, dev)
    pipe.load_inputs(st, *host, non_blocking=False)
    for _ in range(3):
        pipe.step(st)
    kinds = {}
    for op in list(st["dplan"].convs.values()) + list(st["rplan"].convs.values()):
        kinds[op.backend] = kinds.get(op.backend, 0) + 1
    arms[name] = (pipe, st, kinds)
    for k, v in saved.items():
        setattr(engine, k, v)
    engine._TUNE_CACHE.clear()
    engine._TUNE_CACHE.update(table)
torch.cuda.synchronize()
flush_buf = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
res = {k: [] for k in arms}
names = list(arms)
for rd in range(rounds):
    for name in (names if rd % 2 == 0 else names[::-1]):
        pipe, st, _ = arms[name]
        ms, _ = bench.time_events(lambda: pipe.step(st), steps, lambda: flush_buf.zero_())
        res[name].append(ms)
base = statistics.median(res[names[0]])
for name, (_, _, kinds) in arms.items():
    v = res[name]
    print("%-10s median %.3f ms/step (%+.1f %% vs %s)  min %.3f  blocks: %s   kernels: %s" %
          (name, statistics.median(v), 100.0 * (statistics.median(v) / base - 1.0), names[0], min(v),
           " ".join("%.3f" % x for x in v), " ".join("%s x%d" % kv for kv in sorted(kinds.items()))))
