"""Per-K-step cost and fixed overhead of the convolution kernels on one layer shape: times a 3x3 layer at several input
widths (the K loop grows linearly with Cin) and fits  t = t0 + steps * t_step  per (kernel, tile).
    python tools/conv_scaling.py [N H W Cout]        (default 5 48 64 256: the refine network's 48x64 level)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cnmnet_b200 import _lib, engine  # noqa: E402

N, H, W, cout = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (5, 48, 64, 256)
dev = torch.device("cuda", 0)
dtype = _lib.CNM_F16F8
variants = [("tc", 256), ("tcm2", 256), ("halo2", 128), ("halo2", 256), ("halo2h", 256), ("halo2h", 128), ("halo", 128)]
cins = [128, 256, 512, 1024]
rows = {}
flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
for cin in cins:
    x = torch.zeros((N, H, W, cin * 2), dtype=engine.TORCH_DTYPE[dtype], device=dev)
    op = engine.ConvOp("t", [engine.act_src(x, 0, phys=2)], 3, 1, cout, N, H, W, dtype, dev, True, backend="tc")
    op.tunable = False
    op.load(torch.randn(cout, cin, 3, 3) * 0.01, torch.zeros(cout))
    for be, bn in variants:
        if be not in op.variants or cout % bn:
            continue
        op.select(be, bn)
        s = _lib.stream_ptr()
        for _ in range(3):
            op(s)
        ts = []
        for _ in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            op(s)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        rows.setdefault((be, bn), []).append((op.num_steps, float(np.median(ts))))
print("layer %dx%dx%d -> %d, 3x3, f16f8; K-steps of 64 channels (half of them fp8 rows); us, L2 flushed" % (N, H, W, cout))
for (be, bn), pts in rows.items():
    st = np.array([p[0] for p in pts], float)
    t = np.array([p[1] for p in pts], float)
    A = np.stack([np.ones_like(st), st], 1)
    (t0, ts_), *_ = np.linalg.lstsq(A, t, rcond=None)
    print("%-7s/%-3d " % (be, bn) + "  ".join("%4d:%6.1f" % (int(a), b) for a, b in pts) +
          "   fit t0 = %5.1f us, %5.3f us/step" % (t0, ts_))
