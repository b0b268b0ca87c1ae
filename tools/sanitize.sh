#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (run on a B200 box, e.g. through gpurun).  memcheck changes kernel
# timing enough to expose ordering bugs that the plain runs hide: round 1 found a missing barrier between staging the
# bias in shared memory and the epilogue's first read of it this way (test_fused_head failed only under memcheck).
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 \
    python -m pytest tests/test_gpu_conv.py tests/test_gpu_kernels.py -m gpu -q > gpurun_out/memcheck_conv.log 2>&1
tail -3 gpurun_out/memcheck_conv.log
CNM_AUTOTUNE=0 timeout 2400 compute-sanitizer --tool memcheck --print-limit 10 \
    python -m pytest tests/test_gpu_nets.py -m gpu -q > gpurun_out/memcheck_nets.log 2>&1
tail -3 gpurun_out/memcheck_nets.log
timeout 600 compute-sanitizer --tool initcheck --print-limit 10 \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/initcheck.log 2>&1
tail -2 gpurun_out/initcheck.log
