#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (run on a B200 box through gpurun; ~12 GPU-minutes).  Logs go to
# gpurun_out/ and their summaries are committed as profiles/<tag>_sanitizer.md by tools/make_sanitizer_summary.py.
#   memcheck   all convolution / kernel tests and the network tests (minus the full-size headline cases, whose CPU
#              oracle dominates; their kernels and plans are the ones the smaller cases run)
#   initcheck  smoke().  initcheck does not track writes made through the async proxy (cp.async.bulk.tensor stores: every
#              convolution output), so it reports the consumers of those tensors (upsample, packers, the next
#              convolution's SIMT twin, ...) as reading uninitialised memory; the summary lists WHICH kernels it flagged.
#   poison     what initcheck is for, done in a way that sees TMA stores: CNM_POISON=1 fills every plan buffer with 0xFF
#              (NaN in all act formats) and stops buffer recycling; a read of a never-written element fails parity.
#              Runs the end-to-end pipeline tests (incl. the fused plane sweep) that way.
#   racecheck  the shared-memory kernels WITHOUT TMA / tcgen05 (normals, CNM, packers, upsample, heads, metrics,
#              losses): racecheck does not model the async proxy (TMA writes, tcgen05 reads of shared memory), so the
#              tensor-core convolutions, the TMA-staged head kernel and the plane sweep are outside what it can check
# memcheck changes kernel timing enough to expose ordering bugs that the plain runs hide: round 1 found a missing
# barrier between staging the bias in shared memory and the epilogue's first read of it this way.
set -u
mkdir -p gpurun_out
git rev-parse HEAD > gpurun_out/sanitize_head.txt 2>/dev/null || echo "snapshot (no .git on the box)" > gpurun_out/sanitize_head.txt
date -u +%Y-%m-%dT%H:%M:%SZ >> gpurun_out/sanitize_head.txt
sha256sum cnmnet_b200/libcnmnet_b200.so >> gpurun_out/sanitize_head.txt
# the library is not bit-reproducible across builds; the sources are what ties the logs to a commit
echo "$(cat cnmnet_b200/csrc/*.cu cnmnet_b200/csrc/*.cuh include/*.h | sha256sum | cut -d' ' -f1)  cnmnet_b200/csrc/*.cu *.cuh include/*.h (concatenated)" >> gpurun_out/sanitize_head.txt
timeout 1200 compute-sanitizer --tool memcheck --print-limit 10 \
    python -m pytest tests/test_gpu_conv.py tests/test_gpu_kernels.py -m gpu -q > gpurun_out/memcheck_conv.log 2>&1
tail -3 gpurun_out/memcheck_conv.log
timeout 1800 compute-sanitizer --tool memcheck --print-limit 10 \
    python -m pytest tests/test_gpu_nets.py -m gpu -q -k "not headline and not bit_identical" > gpurun_out/memcheck_nets.log 2>&1
tail -3 gpurun_out/memcheck_nets.log
rm -f gpurun_out/initcheck.log
if [ "${CNM_SANITIZE_INITCHECK:-1}" = "1" ]; then   # ~10 GPU-minutes for a report that only restates the tool's blind spot
timeout 600 compute-sanitizer --tool initcheck --print-limit 300 \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/initcheck_full.log 2>&1
# keep the log small: the summary lines, smoke()'s own line and a tally of the flagged kernels (Host Frame = the C entry)
{ grep -E "^smoke:|ERROR SUMMARY" gpurun_out/initcheck_full.log; echo "flagged reads by entry point (first 300 reports):";
  grep -oE "Host Frame: ?cnm_[a-z0-9_]+" gpurun_out/initcheck_full.log | sort | uniq -c; } > gpurun_out/initcheck.log
rm -f gpurun_out/initcheck_full.log
tail -12 gpurun_out/initcheck.log
fi
CNM_POISON=1 timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_fused_sweep.py -m gpu -q \
    -k "pipeline_matches_oracle or odd_pyramid or 128_planes or fused or C1 or layer_program_switches" > gpurun_out/poison.log 2>&1
CNM_POISON=1 CNM_FUSE_SWEEP=force timeout 600 python -m pytest tests/test_gpu_nets.py -m gpu -q \
    -k "pipeline_matches_oracle and f16f8" >> gpurun_out/poison.log 2>&1
tail -4 gpurun_out/poison.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 \
    python -m pytest tests/test_gpu_kernels.py -m gpu -q \
    -k "normals or cnm or refine_aggregate or upsample or pack or metrics or losses or reciprocal or inverse_warp" \
    > gpurun_out/racecheck.log 2>&1
tail -3 gpurun_out/racecheck.log
