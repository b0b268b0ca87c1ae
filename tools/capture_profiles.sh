#!/bin/bash
# All ncu captures behind profiles/<tag>_*.md in one go (run on a B200 box through gpurun; ~8 GPU-minutes), then
# `python tools/make_profiles.py <tag> --full` here turns gpurun_out/ into the tracked summaries.
set -u
mkdir -p gpurun_out
N="ncu --profile-from-start off --clock-control none"
if [ "${CNM_CAPTURE_MAIN:-1}" = "1" ]; then
timeout 300 $N --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv \
    --log-file gpurun_out/launches_step.csv python tools/profile_step.py > gpurun_out/ls.log 2>&1
fi
F="--set full --import-source on --kernel-name-base demangled"
if [ "${CNM_CAPTURE_K1:-1}" = "1" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:plane_sweep -f -o gpurun_out/prof_k1 \
    python tools/k1_profile.py > gpurun_out/prof_k1.log 2>&1
fi
if [ "${CNM_CAPTURE_MAIN:-1}" = "1" ]; then
timeout 300 $N $F -k "regex:conv_halo2?_kernel<\(int\)64" -c 4 -f -o gpurun_out/prof_c64 python tools/profile_step.py > gpurun_out/prof_c64.log 2>&1
timeout 300 $N $F -k "regex:conv_halo2_kernel|conv_halo_kernel" -c 6 -f -o gpurun_out/prof_halo python tools/profile_step.py > gpurun_out/ph.log 2>&1
timeout 300 $N $F -k regex:conv_tc_kernel -c 4 -f -o gpurun_out/prof_tc python tools/profile_step.py > gpurun_out/pt.log 2>&1
# folded upsampling (conv_halo2f: DepthNet upconv1.1, RefineNet upconv1_depth.1) and the balanced CTA-pair kernel (conv1.3, conv2.3)
timeout 300 $N $F -k "regex:conv_halo2?f_kernel" -c 2 -f -o gpurun_out/prof_fold python tools/profile_step.py > gpurun_out/pfold.log 2>&1
timeout 300 $N $F -k "regex:conv_halo2?b_kernel|conv_tcb_kernel" -c 2 -f -o gpurun_out/prof_bal python tools/profile_step.py > gpurun_out/pbal.log 2>&1
timeout 300 $N $F -k "regex:upsample2x|head_partial|head_finish|pack_|refine_|plane_sweep" -c 12 -f -o gpurun_out/prof_small \
    python tools/profile_step.py > gpurun_out/psm.log 2>&1
fi
# the fused conv1.0 (hardware-counter sections only: ncu's SASS-patching metrics hang on a kernel that uses setmaxnreg)
[ "${CNM_CAPTURE_FUSED:-1}" = "1" ] && CNM_FUSE_SWEEP=force timeout 150 $N --section SpeedOfLight --section SchedulerStats --section WarpStateStats \
    --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --kernel-name-base demangled \
    -k regex:sweep_kernel -c 1 python tools/profile_step.py > gpurun_out/prof_fused.log 2>&1
[ "${CNM_CAPTURE_MAIN:-1}" = "1" ] && timeout 120 python tools/conv_timeline.py > gpurun_out/conv_timeline.log 2>&1
[ "${CNM_CAPTURE_SCALING:-1}" = "1" ] && timeout 200 python tools/conv_scaling.py > gpurun_out/conv_scaling.log 2>&1
# gpurun brings back at most 64 MiB: keep the raw metric pages (what tools/make_profiles.py reads), drop the reports
# unless CNM_KEEP_NCU_REP names the ones to keep (e.g. "prof_fold")
for r in gpurun_out/prof_*.ncu-rep; do
  [ -f "$r" ] || continue
  ncu -i "$r" --page raw --csv > "${r%.ncu-rep}.raw.csv" 2>/dev/null
  case " ${CNM_KEEP_NCU_REP:-} " in *" $(basename "${r%.ncu-rep}") "*) ;; *) rm -f "$r" ;; esac
done
ls -la gpurun_out/
