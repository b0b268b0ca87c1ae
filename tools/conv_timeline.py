"""Where the time of one per-tap convolution launch goes (CNM_CONV_TIMELINE=1): clock64 stamps of CTA 0.
    CNM_CONV_TIMELINE=1 python tools/conv_timeline.py [N H W Cin Cout]      (default 5 48 64 512 256)"""
import ctypes as C
import os
import sys

os.environ.setdefault("CNM_CONV_TIMELINE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cnmnet_b200 import _lib, engine  # noqa: E402

N, H, W, cin, cout = (int(v) for v in sys.argv[1:6]) if len(sys.argv) >= 6 else (5, 48, 64, 512, 256)
dev = torch.device("cuda", 0)
dtype = _lib.CNM_F16F8
lib = _lib.load()
x = torch.zeros((N, H, W, cin * 2), dtype=engine.TORCH_DTYPE[dtype], device=dev)
op = engine.ConvOp("t", [engine.act_src(x, 0, phys=2)], 3, 1, cout, N, H, W, dtype, dev, True, backend="tc")
op.tunable = False
op.load(torch.randn(cout, cin, 3, 3) * 0.01, torch.zeros(cout))
flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
names = ["entry", "prologue done", "dependency wait done", "first TMA issued", "first operands landed",
         "last MMA of tile 0 issued", "accumulator 0 complete", "stores drained", "exit", "pieces gathered (balanced)"]
for be, bn, cold in (("tc", 256, True), ("tc", 256, False), ("tc", 128, False), ("tcb", 256, False), ("tcb", 128, False)):
    op.select(be, bn)
    s = _lib.stream_ptr()
    for _ in range(3):
        op(s)
    torch.cuda.synchronize()
    if cold:
        flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    op(s)
    b.record()
    torch.cuda.synchronize()
    ts = (C.c_longlong * 12)()
    _lib.check(lib.cnm_debug_conv_timeline(ts))
    t0 = ts[0]
    print("%s/%d %s: %d K-steps, %.1f us by events; cycles since entry (CTA 0):" % (be, bn, "cold L2" if cold else "warm L2",
                                                                                  op.num_steps, a.elapsed_time(b) * 1e3))
    for n, v in zip(names, ts):
        print("    %-28s %8d" % (n, v - t0 if v else -1))
