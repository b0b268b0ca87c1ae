"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/ (round tag as argv[1])."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

# ---- launch list ----
rows = list(csv.reader(open(os.path.join(G, "launches_step.csv"))))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, mi, idc, gi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "ID", "Grid Size"))
L = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    d = L.setdefault(r[idc], {"name": r[ki].split("(")[0].replace("void ", ""), "grid": r[gi]})
    d[r[mi]] = float(r[vi].replace(",", ""))
bench = {}
bp = os.path.join(G, "bench.json")
if os.path.exists(bp):
    bench = json.load(open(bp))
out = ["# %s launch list of one pipeline step" % tag, "",
       "Command: `ncu --profile-from-start off --metrics "
       "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv python tools/profile_step.py`",
       "(one un-graphed C4 step, %s mode, B=5 x 2 source views, 256x192x64, after one warm-up step; conv kernel / tile choices" % bench.get("dtype", "bf16x3"),
       "come from the shipped table cnmnet_b200/tune/b200.json; the last `plane_sweep_kernel<0>` and `cost_volume_kernel<16,0>` launches are",
       "the C2 NCHW fp32 micro-benchmark of K1, second and first generation).  Times are cold-cache and serialised by ncu: compare SHARES, not absolutes (the same step",
       "replayed as a CUDA graph takes %.2f ms in bench.py)." % bench.get("ms_per_step", float("nan")), "",
       "| # | kernel | grid | us | DRAM read+write MB |", "|---|---|---|---|---|"]
tot = collections.OrderedDict()
n = 0
for d in L.values():
    n += 1
    t, b = d["gpu__time_duration.sum"] / 1e3, (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / 1e6
    out.append("| %d | `%s` | %s | %.1f | %.1f |" % (n, d["name"], d["grid"], t, b))
    x = tot.setdefault(d["name"], [0, 0.0, 0.0])
    x[0] += 1
    x[1] += t
    x[2] += b
s = sum(v[1] for v in tot.values())
out += ["", "## Shares", "", "| kernel | launches | us | share | DRAM MB |", "|---|---|---|---|---|"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    out.append("| `%s` | %d | %.1f | %.1f %% | %.1f |" % (k, v[0], v[1], 100 * v[1] / s, v[2]))
out.append("| total | %d | %.1f | 100 %% | %.1f |" % (n, s, sum(v[2] for v in tot.values())))
cu = sum(v[1] for k, v in tot.items() if k.startswith("conv_"))
cm = sum(v[2] for k, v in tot.items() if k.startswith("conv_"))
out += ["", "Convolution kernels (every `conv_*` kernel: per-tap, row-halo, CTA-pair, balanced `...b`, folded upsampling `...f`): %.1f us = %.1f %% of the step, "
        "%.1f MB of DRAM traffic per step." % (cu, 100 * cu / s, cm)]
open(os.path.join(P, "%s_launches_step.md" % tag), "w").write("\n".join(out) + "\n")
cv = [d for d in L.values() if d["name"].startswith("plane_sweep_kernel<0>")] or [d for d in L.values() if d["name"].startswith("cost_volume_kernel<16, 0>")]
json.dump({"source": "profiles/%s_launches_step.md (ncu dram__bytes_read.sum + dram__bytes_write.sum, one C4 step, %s)" % (tag, bench.get("dtype", "bf16x3")),
           "conv_bytes_per_step": cm * 1e6, "conv_launches": sum(v[0] for k, v in tot.items() if k.startswith("conv_")),
           "cost_volume_c2_bytes": (cv[0]["dram__bytes_read.sum"] + cv[0]["dram__bytes_write.sum"]) if cv else None},
          open(os.path.join(P, "%s_traffic.json" % tag), "w"), indent=1)

# ---- full-set captures ----
PICKS = ["gpu__time_duration.sum", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
         "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
         "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
         "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
         "launch__registers_per_thread", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second"]


def dump(rep, outname, title, note):
    rp = os.path.join(G, rep)
    rc = rp.replace(".ncu-rep", ".raw.csv")   # written on the GPU box by tools/capture_profiles.sh (the reports stay there)
    if os.path.exists(rc):
        raw = open(rc).read()
    elif os.path.exists(rp):
        raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return
    rws = list(csv.reader(raw.splitlines()))
    h, u = rws[0], rws[1]
    lines = ["# " + title, "", note, ""]
    for r in rws[2:]:
        lines += ["## `%s`  grid %s" % (r[h.index("Kernel Name")].split("(")[0], r[h.index("launch__grid_size")]), "",
                  "| metric | value | unit |", "|---|---|---|"]
        lines += ["| %s | %s | %s |" % (p, r[h.index(p)], u[h.index(p)]) for p in PICKS if p in h]
        lines.append("")
    open(os.path.join(P, outname), "w").write("\n".join(lines))


if "--full" in sys.argv:
  dump("prof_k1.ncu-rep", "%s_plane_sweep_ncu.md" % tag,
     "%s ncu --set full: plane_sweep_kernel (K1, TMA-staged plane sweep)" % tag,
     "Capture: `ncu --set full --clock-control none --import-source on -k regex:plane_sweep -o gpurun_out/prof_k1 python "
     "tools/k1_profile.py`: C2 (256x192x64, 2 views, batch 5, fp32 NCHW, 134 676 480 algorithmic bytes), first launch with the "
     "bench cameras (seed 1234), second with seed 77.  Reading: issue slots 65-73 % active and l1tex LSU wavefronts (shared-memory "
     "taps, 1.5x the conflict-free count when the local zoom of the homography exceeds 1) 66-71 % of peak are the co-limiters; "
     "DRAM 12 %, L2 20 %.")
  dump("prof_c64.ncu-rep", "%s_conv_cout64_ncu.md" % tag,
     "%s ncu --set full: the Cout = 64 full-resolution layers (conv_halo2_kernel<64> / conv_halo_kernel<64>)" % tag,
     "Capture: `ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "
     "\"regex:conv_halo2?_kernel<\\(int\\)64\" -c 4 python tools/profile_step.py` (DepthNet upconv1.1, iconv1.0 with the fused "
     "disp1 head, RefineNet upconv1_depth.1, iconv1_depth.0).  Reading: tensor pipe 33-49 % active -- with N = 64 an MMA step "
     "reads 32 KB of A for 256 cycles of math (M = 256 pair), so the 128 B/clk shared-memory port (A reads + slab fills + "
     "staging) bounds these layers, not the tensor pipe; see DESIGN.md section 3.")
if "--full" in sys.argv:
  dump("prof_halo.ncu-rep", "%s_conv_halo_ncu.md" % tag,
     "%s ncu --set full: conv_halo_kernel / conv_halo2_kernel (row-halo tcgen05 conv, 1-SM and CTA-pair)" % tag,
     "Capture: `ncu --profile-from-start off --set full --clock-control none --import-source on -k "
     "regex:\"conv_halo2_kernel|conv_halo_kernel\" -c 6 python tools/profile_step.py` (%s mode, C4 step, first six halo launches of " % bench.get("dtype", "f16f8") +
     "the DepthNet pass: conv1.0, conv1.3, conv2.0, conv3.0, ... in launch order; see the launch list for grids).")
if "--full" in sys.argv:
  dump("prof_tc.ncu-rep", "%s_conv_tc_final_ncu.md" % tag,
     "%s ncu --set full: conv_tc_kernel (per-tap tcgen05 conv) after the warp-uniform issue loop" % tag,
     "Capture: `ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 4 "
     "python tools/profile_step.py` (%s mode, C4 step; the layers the autotuner leaves on the per-tap kernel: small and <= 48x64 layers)." % bench.get("dtype", "f16f8"))
if "--full" in sys.argv:
  dump("prof_fold.ncu-rep", "%s_conv_fold_ncu.md" % tag,
     "%s ncu --set full: conv_halo2f_kernel (bilinear x2 upsampling folded into the 3x3 convolution, cnm_conv_desc.fold)" % tag,
     "Capture: `ncu --profile-from-start off --set full --clock-control none --import-source on -k \"regex:conv_halo2?f_kernel\" -c 2 "
     "python tools/profile_step.py` (DepthNet upconv1.1: 10 x 96x128x128 -> 10 x 192x256x64 as a 256-column convolution on the "
     "low-res tensor, and RefineNet upconv1_depth.1 at batch 5).  Compare with %s_conv_cout64_ncu.md / r02 history: the same layers as "
     "64-column convolutions on the upsampled tensor ran the tensor pipe at 47-50 %% and moved 348 / 160 MB of DRAM traffic "
     "each, after upsample2x launches of 253 / 96 MB." % tag)
  dump("prof_bal.ncu-rep", "%s_conv_balanced_ncu.md" % tag,
     "%s ncu --set full: conv_halo2b_kernel (balanced / stream-K schedule of the CTA-pair row-halo kernel)" % tag,
     "Capture: `ncu --profile-from-start off --set full --clock-control none --import-source on -k "
     "\"regex:conv_halo2?b_kernel|conv_tcb_kernel\" -c 2 python tools/profile_step.py` (DepthNet conv1.3 and conv2.3, the two "
     "layers the table runs on the balanced schedule).")
if "--full" in sys.argv:
  dump("prof_small.ncu-rep", "%s_small_kernels_ncu.md" % tag,
     "%s ncu --set full: the kernels around the convolutions (cost volume, tap packers, upsample, heads, refine aggregation)" % tag,
     "Capture: `ncu --profile-from-start off --set full --clock-control none --import-source on -k "
     "regex:\"upsample2x|head_partial|head_finish|pack_|refine_|cost_volume\" -c 12 python tools/profile_step.py` "
     "(%s mode, C4 step, first twelve matching launches in launch order)." % bench.get("dtype", "f16f8"))

def text_profile(logs, outname, title, note, keep=None):
    """A committed summary made of the text a tool printed (ncu sections without a report file, timeline tools)."""
    body = []
    for lg in logs:
        pth = os.path.join(G, lg)
        if not os.path.exists(pth):
            continue
        lines = [l.rstrip() for l in open(pth, errors="replace") if not l.startswith("==") and l.strip()]
        if keep:
            lines = [l for l in lines if keep(l)]
        body += ["```"] + lines + ["```", ""]
    if body:
        open(os.path.join(P, outname), "w").write("\n".join(["# " + title, "", note, ""] + body))


text_profile(["prof_fused.log"], "%s_fused_sweep_ncu.md" % tag,
             "%s ncu (hardware-counter sections): conv_halo2_sweep_kernel<128> -- conv1.0 with the plane sweep fused in" % tag,
             "Capture: `CNM_FUSE_SWEEP=force ncu --profile-from-start off --section SpeedOfLight --section SchedulerStats --section "
             "WarpStateStats --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --clock-control none -k "
             "regex:sweep_kernel -c 1 python tools/profile_step.py` (C4 step).  `--set full` is not usable on this kernel: ncu's "
             "SASS-patching metrics hang on a kernel that re-balances registers with setmaxnreg.  Reading: DESIGN.md section 3, (f)-1.",
             keep=lambda l: not l.lstrip().startswith(("OPT", "INF", "WRN", "-----")) and "Est." not in l and "done " not in l)
text_profile(["conv_timeline.log", "conv_scaling.log"], "%s_conv_timeline.md" % tag,
             "%s where the time of a mid-size convolution launch goes" % tag,
             "`tools/conv_timeline.py` (clock64 stamps of CTA 0 of the per-tap kernel, CNM_CONV_TIMELINE=1) and "
             "`tools/conv_scaling.py` (t = t0 + steps x t_step per kernel variant; single launches timed with events, so t0 "
             "includes ~15 us of host-side launch work that a CUDA-graph replay does not pay) on the refine network's "
             "5 x 48x64 -> 256 layer.  Reading: DESIGN.md section 3, \"What bounds the 48x64 / 24x32 layers\".")
print("ok", cu, cm, s)
