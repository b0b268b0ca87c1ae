"""First-contact probe of the tcgen05 convolution on a real B200: smallest cases first, prints errors
instead of asserting, so one run tells how far the kernel gets.  Run under `timeout`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from cnmnet_b200 import _lib  # noqa: E402
from test_gpu_conv import CASES, oracle_conv, run_conv  # noqa: E402
from gpu_util import rel_l2  # noqa: E402

print(torch.cuda.get_device_name(0), _lib.load().cnm_version(), flush=True)
import numpy as np  # noqa: E402
for dtype in (_lib.CNM_BF16, _lib.CNM_TF32):
    for case in CASES:
        N, H, W, cins, cout, k, stride = case
        g = torch.Generator().manual_seed(1)
        xs = [torch.randn(N, c, H, W, generator=g) for c in cins]
        w = torch.randn(cout, sum(cins), k, k, generator=g) / np.sqrt(sum(cins) * k * k)
        b = torch.randn(cout, generator=g)
        wchs = [sum(cins[:i]) for i in range(len(cins))]
        want = oracle_conv(dtype, xs, w, b, k, stride)
        for backend in ("simt", "tc"):
            try:
                got = run_conv(backend, dtype, xs, wchs, w, b, k, stride)
                print("dtype", dtype, case, backend, "rel_l2 %.3e max %.3e" % (rel_l2(got, want), (got - want).abs().max()),
                      flush=True)
            except Exception as e:  # noqa: BLE001
                print("dtype", dtype, case, backend, "EXC", repr(e)[:300], flush=True)
