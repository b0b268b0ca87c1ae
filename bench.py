#!/usr/bin/env python
"""Benchmark of the CNMNet inference hot path on B200 (one JSON line on stdout, rank 0).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on host cores

Workload (BASELINE.json configs[3], the one the metric "depth frames/s @256x192x64 planes" is quoted
on): full refine model, 1 ref + 2 source views, 256x192, 64 planes, k_size 9, batch 5 frame tuples per
GPU, random-init weights (reference init under torch.manual_seed(0)), synthetic frames.  A step is one
pass of the whole pipeline over one batch: 2 cost volumes -> 2 x DepthNet -> aggregation -> DepthRefineNet
-> 1/idepth -> depth->normal, replayed as one CUDA graph.  Multi-GPU = weak scaling: every rank runs its
own batch, then ONE NCCL all-gather of the packed outputs on a side stream (overlapping the next step).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H, W, D, V, B_PER_GPU, K_SIZE = 192, 256, 64, 2, 5, 9
METRIC = "depth frames/s @256x192x64 planes"
WORKLOAD = "C4: full refine model, 1 ref + 2 src views, 256x192, 64 planes, k_size 9, batch 5 per GPU"
# algorithmic work (SURVEY.md section 8d / BASELINE.md section 3)
DEPTHNET_GFLOP, REFINE_GFLOP = 161.43, 112.79
C2_COST_VOLUME_BYTES = 5 * (2 * 4 * D * H * W + 3 * 4 * 3 * H * W)  # 134 676 480


# final-depth error of each tensor-core mode vs the fp32 reference restatement, as measured by tests/test_gpu_nets.py and
# __graft_entry__.smoke() on B200 (round 1); the north-star tolerance is 1e-3 relative
PRECISION_NOTES = {
    "bf16": "bf16 operands, fp32 accumulate (speed mode): final depth mean rel err 4e-3 or more, p99 >= 1e-2",
    "tf32": "tf32 operands, fp32 accumulate: final depth mean rel err 2e-4..2e-3, p99 6e-4..7e-3 depending on weights",
    "bf16x3": "split bf16 (hi+lo operands, 3 bf16 tensor-core passes, fp32 accumulate): final depth max rel err <= 2.3e-4 "
              "over all weights tried -- meets the 1e-3 tolerance",
    "f16f8": "fp16 operands + fp8 (e4m3) first-order corrections: one kind::f16 pass hi*W16 and one kind::f8f6f4 pass "
             "[lo8|hi8]*[Whi8|Wlo8] into the same fp32 accumulator (2 bf16-pass-equivalents of tensor time instead of 3); "
             "operand error ~2^-17 like bf16x3; final depth max rel err <= 3.9e-4 over all weights tried -- meets the 1e-3 "
             "tolerance",
}


def load_traffic():
    """ncu-measured DRAM bytes (read + write) of the profiled kernels, committed under profiles/ (null if absent)."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"],
                "bf16_tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 50 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.lines, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def make_weights(seed=0):
    """Reference random init (kaiming_normal fan_out, BN 1/0) under torch.manual_seed(seed)."""
    import cnmnet_b200
    torch.manual_seed(seed)
    d = cnmnet_b200.depthNet(idepth_scale=3.0, planes=D)
    r = cnmnet_b200.DepthRefineNet(idepth_scale=3.0)
    return d, r


def reference_modules(device):
    """The UNMODIFIED reference modules through the import shim (oracle/ref_shim.py; baseline/_ref on the GPU box), with
    the reference's own random init under torch.manual_seed(0).  None when the reference files are not present."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        return None
    ns = ref_shim.load_reference()
    torch.manual_seed(0)
    d = ns.depthNet(idepth_scale=3.0).eval().to(device)
    r = ns.DepthRefineNet(idepth_scale=3.0).eval().to(device)
    return ns, d, r, ns.Depth2normal(K_SIZE)


def reference_step(mods, ref, srcs, ref_cam, src_cams):
    """One batch through the reference's own modules in the order of eval_refine (eval.py:440-455)."""
    ns, d, r, d2n = mods
    from oracle import ref_shim
    ref_shim.reset_pixel_cache(ns)
    with torch.no_grad():
        p1, ic1 = d(ref, srcs[:, 0].contiguous(), ref_cam, src_cams[:, 0].contiguous())
        p2, ic2 = d(ref, srcs[:, 1].contiguous(), ref_cam, src_cams[:, 1].contiguous())
        idr, prob = r(idepth01=p1[0], idepth02=p2[0], iconv01=ic1, iconv02=ic2)
        depth = torch.div(1.0, idr.squeeze(1))
        normal, _ = d2n(depth, torch.inverse(ref_cam[:, 1, 0:3, 0:3]))
    return depth, normal, prob


def cpu_reference_fps(steps, warmup, batch):
    """The reference on the host cores: its own modules when baseline/_ref (or /root/reference) is present
    (kind "reference"), else the op-for-op torch-CPU port in oracle/nets.py (kind "port").  Bounded sample: `steps`
    batches of `batch` frame tuples.  -> (frames/s, per-step seconds, kind)"""
    from cnmnet_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(1234, batch, H, W, V)
    mods = reference_modules(torch.device("cpu"))
    if mods is not None:
        kind = "reference"

        def run():
            reference_step(mods, ref, srcs, ref_cam, src_cams)
    else:
        from oracle import nets as onets
        kind = "port"
        dsd = onets.reference_init_state(onets.DEPTHNET_LAYERS, 0)
        rsd = onets.reference_init_state(onets.REFINENET_LAYERS, 0)

        def run():
            onets.reference_pipeline(dsd, rsd, ref, srcs, ref_cam, src_cams, 3.0, K_SIZE)
    from oracle import ref_shim
    with ref_shim.cpu_mode():   # the reference hard-codes torch.cuda.FloatTensor; this run is on the host cores
        for _ in range(warmup):
            run()
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            run()
            ts.append(time.perf_counter() - t0)
    return batch * len(ts) / sum(ts), ts, kind


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores, same workload
    (batch of B_PER_GPU tuples per step), a bounded number of steps."""
    if rank != 0:
        return
    steps = max(1, min(args.steps, 4))      # ~5-8 s of CPU work per batch-5 step
    warmup = max(1, min(args.warmup, 1))
    fps, ts, kind = cpu_reference_fps(steps, warmup, B_PER_GPU)
    sample = "%d timed steps (+%d warm-up) of one batch of %d frame tuples at 256x192x64, all host threads" % (
        steps, warmup, B_PER_GPU)
    emit({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * sum(ts) / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B_PER_GPU, "views": V + 1, "planes": D, "k_size": K_SIZE,
                   "weights": "reference random init, torch.manual_seed(0)", "sample": sample,
                   "host": "torch %s CPU, %d threads" % (torch.__version__, torch.get_num_threads()),
                   "what": "the reference's own depthNet / DepthRefineNet / Depth2normal modules (unmodified files in "
                           "baseline/_ref behind oracle/ref_shim.py), call order of eval.py:440-455" if kind == "reference"
                           else "op-for-op torch-CPU port of the reference pipeline (oracle/nets.py): baseline/_ref absent"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def reference_gpu_block(dev, host, reps=5):
    """Informational: the reference's own modules on THIS GPU through stock PyTorch (cuDNN / ATen library kernels) -- the
    bar SURVEY.md section 2.1 names ("the reference does have a Blackwell path -- the stock library kernels").  Call sites:
    depthNet_model.py:208-223 (~1000 ATen launches per getVolume), :235-261 and :335-365 (cuDNN convolutions, TF32 by
    default on this torch).  Same batch, same inputs, device-resident, CUDA events; plus the same modules in
    channels_last under bf16 autocast (what a user would try first to make the reference fast)."""
    mods = reference_modules(dev)
    if mods is None:
        return {"unavailable": "baseline/_ref not present (the reference files are vendored by __graft_entry__.build() in "
                               "the build container)"}
    ns, d, r, d2n = mods
    ref, srcs, ref_cam, src_cams = [t.to(dev) for t in host]
    B = ref.shape[0]

    def ev(fn, n=reps):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    out = {"what": "reference modules (baseline/_ref, unmodified) on this GPU, stock torch %s / cuDNN %s, batch %d, inputs "
                   "resident, median of %d" % (torch.__version__, torch.backends.cudnn.version(), B, reps),
           "allow_tf32_conv": bool(torch.backends.cudnn.allow_tf32), "unit": "frames/s"}
    try:
        with torch.no_grad():
            ms_all = ev(lambda: reference_step(mods, ref, srcs, ref_cam, src_cams))
            pix = ns.get_pixel_coordinates(H, W)
            s0, c0 = srcs[:, 0].contiguous(), src_cams[:, 0].contiguous()
            krk, kt = ns.process_camera_parameters(ref_cam, c0, pix)
            ms_vol = ev(lambda: d.getVolume(ref, s0, krk, kt))
            ms_net = ev(lambda: d(ref, s0, ref_cam, c0))
            p1, ic1 = d(ref, s0, ref_cam, c0)
            ms_ref = ev(lambda: r(idepth01=p1[0], idepth02=p1[0], iconv01=ic1, iconv02=ic1))
            depth = torch.div(1.0, p1[0].squeeze(1))
            Kinv = torch.inverse(ref_cam[:, 1, 0:3, 0:3])
            ms_nrm = ev(lambda: d2n(depth, Kinv))
        out["fp32_tf32"] = {"value": B / (ms_all * 1e-3), "ms_per_step": ms_all,
                            "stages_ms": {"getVolume (one source view)": ms_vol,
                                          "depthNet.forward (one source view, incl. getVolume)": ms_net,
                                          "DepthRefineNet.forward": ms_ref, "Depth2normal": ms_nrm}}
        # channels_last + bf16 autocast on the two networks
        d2 = d.to(memory_format=torch.channels_last)
        r2 = r.to(memory_format=torch.channels_last)
        mods2 = (ns, d2, r2, d2n)

        def fast():
            # autocast around the two networks only: Depth2normal's batched LU has no bf16 kernel in torch
            from oracle import ref_shim
            ref_shim.reset_pixel_cache(ns)
            with torch.no_grad():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    p1, ic1 = d2(ref, srcs[:, 0].contiguous(), ref_cam, src_cams[:, 0].contiguous())
                    p2, ic2 = d2(ref, srcs[:, 1].contiguous(), ref_cam, src_cams[:, 1].contiguous())
                    idr, prob = r2(idepth01=p1[0], idepth02=p2[0], iconv01=ic1, iconv02=ic2)
                depth = torch.div(1.0, idr.float().squeeze(1))
                normal, _ = d2n(depth, torch.inverse(ref_cam[:, 1, 0:3, 0:3]))
            return depth

        try:
            ms_fast = ev(fast)
            out["channels_last_bf16_autocast"] = {"value": B / (ms_fast * 1e-3), "ms_per_step": ms_fast}
        except Exception as e:  # the reference mixes dtypes in places autocast does not cover
            out["channels_last_bf16_autocast"] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
    except Exception as e:
        out["unavailable"] = "%s: %s" % (type(e).__name__, str(e)[:300])
    del mods
    torch.cuda.empty_cache()
    return out


def _kernel_census(st):
    """Which convolution kernel variant runs how many layers of the step (engine.ConvOp names: tc per-tap, halo / halo2 row-halo
    on one CTA / a CTA pair, ...b balanced (stream-K) schedule, halo2f / halof bilinear x2 upsampling folded in)."""
    kinds = {}
    for op in list(st["dplan"].convs.values()) + list(st["rplan"].convs.values()):
        kinds[op.backend] = kinds.get(op.backend, 0) + 1
    return kinds


def time_events(fn, reps, flush):
    """Average device time of fn() over reps, CUDA events on the current stream, L2 flushed between reps."""
    ts = []
    for _ in range(reps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.mean(ts), ts


_REAL_STDOUT = None


def emit(line):
    """Prints the one JSON line on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def quiet_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1: route fd 1 to stderr for the run, keep the real one."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cnmnet_b200", choices=["cnmnet_b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CNM_PRECISION", "f16f8"),
                    choices=["bf16", "tf32", "bf16x3", "f16f8"],
                    help="tensor-core mode; f16f8 (default) and bf16x3 meet the 1e-3 final-depth tolerance")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5", "c5-64"],
                    help="c4 = BASELINE configs[3] (default, the metric's config); c5 = configs[4]: 640x480, 128 planes "
                         "(depthNet(planes=128): conv1 widened to 131 inputs, an extension -- the reference hard-codes "
                         "64), batch 32 per GPU; c5-64 = the same at the reference's 64 planes")
    ap.add_argument("--lanes", type=int, default=0,
                    help="concurrent pipeline lanes of the streaming API in the e2e measurement (0 = 2 at C4, 1 at C5 "
                         "where a second buffer set does not fit)")
    ap.add_argument("--cpu-reps", type=int, default=2)
    ap.add_argument("--no-reference-gpu", action="store_true",
                    help="skip the informational run of the reference's own modules on this GPU (stock PyTorch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-modes", action="store_true", help="skip the bf16 / tf32 / bf16x3 lines (A/B experiments)")
    ap.add_argument("--layers", action="store_true", help="also print a per-layer timing table on stderr")
    args = ap.parse_args()
    global H, W, D, B_PER_GPU, WORKLOAD, DEPTHNET_GFLOP, REFINE_GFLOP, METRIC
    if args.workload in ("c5", "c5-64"):
        H, W, B_PER_GPU = 480, 640, 32
        D = 128 if args.workload == "c5" else 64
        WORKLOAD = "C5%s: full refine model, 1 ref + 2 src views, 640x480, %d planes, k_size 9, batch 32 per GPU" % (
            "" if D == 128 else " at 64 planes", D)
        METRIC = "depth frames/s @640x480x%d planes" % D
        # 1008.94 GFLOP at 64 planes; conv1.0 grows from 67 to 131 input channels at 128
        DEPTHNET_GFLOP, REFINE_GFLOP = (1255.5 if D == 128 else 1008.94), 704.96
        args.no_cpu_baseline = True
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import cnmnet_b200
    from cnmnet_b200 import _lib, synth
    from cnmnet_b200.dist import OutputGatherer
    lib = _lib.load()
    peaks = load_peaks()
    traffic = load_traffic()
    B = B_PER_GPU
    dnet, rnet = make_weights(0)
    dnet.precision = rnet.precision = cnmnet_b200.engine.parse_dtype(args.precision)
    dnet, rnet = dnet.to(dev), rnet.to(dev)
    pipe = cnmnet_b200.CNMPipeline(dnet, rnet, k_size=K_SIZE, use_graph=True)
    st = pipe.static_state(B, V, H, W, dev)
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(1234 + rank, B, H, W, V)
    host = [t.pin_memory() for t in (ref, srcs, ref_cam, src_cams)]
    pipe.load_inputs(st, *host, non_blocking=False)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def flush():
        flush_buf.zero_()

    # the only collective: ONE all-gather per step of the packed output block (depth, normal, prob), on a side stream so
    # that it overlaps the next step
    gatherer = OutputGatherer(st["packed"].numel(), dev) if world > 1 else None

    def step():
        pipe.step(st)
        if gatherer is not None:
            gatherer.submit(st["packed"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches_per_step = st.get("launches_per_step") or 0

    # ---- timed region: exactly K steps, device time per step by CUDA events, L2 flushed in between ----
    with ClockSampler(local_rank) as clocks:
        barrier()
        wall0 = time.perf_counter()
        total_ms = 0.0
        for i in range(args.steps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            if gatherer is not None and i == args.steps - 1:
                gatherer.join()   # every gather of the timed region completes inside it
            b.record()
            b.synchronize()
            total_ms += a.elapsed_time(b)
        barrier()
        wall = time.perf_counter() - wall0
        # diagnostic: the same K steps replayed back to back (no flush, no host sync in between) -- the sustained rate
        # the streaming API can approach; power capping makes it lower than the isolated per-step figure
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            step()
        if gatherer is not None:
            gatherer.join()
        b.record()
        b.synchronize()
        b2b_ms = a.elapsed_time(b) / args.steps
        barrier()
        # diagnostic: resident inputs, K steps alternating over two lanes (two buffer sets / graphs / streams): what
        # the streaming API can reach when transfers are free
        two_lane_ms = None
        if (args.lanes or (2 if args.workload == "c4" else 1)) >= 2:
            st1 = pipe._prepare(B, V, H, W, dev, lane=1)
            pipe.load_inputs(st1, *host, non_blocking=False)
            pipe.step(st1)
            torch.cuda.synchronize()
            cur, side = torch.cuda.current_stream(), torch.cuda.Stream()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            side.wait_stream(cur)
            for i in range(args.steps):
                if i & 1:
                    with torch.cuda.stream(side):
                        pipe.step(st1)
                else:
                    pipe.step(st)
            cur.wait_stream(side)
            b.record()
            b.synchronize()
            two_lane_ms = a.elapsed_time(b) / args.steps
            barrier()
        # ---- end to end through the public streaming API (CNMPipeline.infer_stream): every step uploads its inputs
        # from pinned host memory and downloads depth / normal / prob to the host; uploads of step i+1 and
        # downloads of step i-1 overlap the compute of step i on a copy stream ----
        def gather(_st):
            if gatherer is not None:
                gatherer.submit(_st["packed"])

        def host_batches(n):
            for _ in range(n):
                yield tuple(host)

        lanes = args.lanes or (2 if args.workload == "c4" else 1)
        for _ in pipe.infer_stream(host_batches(4), after_step=gather, lanes=lanes):
            pass
        barrier()
        e0 = time.perf_counter()
        n_out = 0
        for outs in pipe.infer_stream(host_batches(args.steps), after_step=gather, lanes=lanes):
            n_out += 1
        if gatherer is not None:
            gatherer.join()
        torch.cuda.synchronize()
        assert n_out == args.steps
        barrier()
        e2e_s = time.perf_counter() - e0
        out_host = {k: st[k] for k in ("depth", "normal", "prob")}
    t = torch.tensor([total_ms, e2e_s * 1e3, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, wall_ms = [float(x) for x in t.tolist()]
    ms_per_step = total_ms / args.steps
    fps = B * world / (ms_per_step * 1e-3)
    e2e_fps = B * world / (e2e_ms * 1e-3 / args.steps)
    h2d = sum(x.numel() * 4 for x in host)
    d2h = sum(x.numel() * 4 for x in out_host.values())

    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision, "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": B * world, "views": V + 1, "planes": D, "k_size": K_SIZE,
                   "weights": "reference random init, torch.manual_seed(0)", "parallelism": "dp%d" % world,
                   "l2": "flushed between timed steps (256 MiB memset, outside the event pair)",
                   "timing": "value / ms_per_step: K isolated steps, one batch in flight, device events per step; e2e: the "
                             "streaming API with host buffers and %d batch(es) in flight (lanes), which is why it can exceed "
                             "value; two_lane_resident_ms_per_step is the same overlap with resident inputs" % lanes,
                   "cuda_graph": True, "precision_note": PRECISION_NOTES[args.precision],
                   "conv_kernels": _kernel_census(st)},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "lanes": lanes},
        "gpu_launches": int(launches_per_step * args.steps),
        "launches_per_step": int(launches_per_step),
        "wall_ms_per_step_incl_flush": wall_ms / args.steps,
        "back_to_back_ms_per_step": b2b_ms,
        "two_lane_resident_ms_per_step": two_lane_ms,
        "plan_activation_bytes_per_lane": int(st["dplan"].bytes_allocated + st["rplan"].bytes_allocated),
    }

    if rank == 0:
        # ---- per-kernel live timing for the rooflines (rank 0, outside the timed region) ----
        s = _lib.stream_ptr()
        dplan, rplan = st["dplan"], st["rplan"]
        conv_ms, conv_rows = 0.0, []
        for plan, tag in ((dplan, "depthnet"), (rplan, "refine")):
            for name, op in plan.convs.items():
                ms, _ = time_events(lambda: op(s), 5, flush)
                conv_ms += ms
                conv_rows.append(("%s.%s [%s/%d%s]" % (tag, name, op.backend, op.desc.reserved,
                                                      "/k%d" % op.desc.ksplit if op.desc.ksplit > 1 else ""), ms,
                                  op.flops_bf16_time))
        alg_flop = B * (V * DEPTHNET_GFLOP + REFINE_GFLOP) * 1e9  # real channels only (no padding), SURVEY 8d
        issued_bf16_time = sum(r[2] for r in conv_rows)             # bf16-pass-equivalents issued, padding included
        tf = alg_flop / (conv_ms * 1e-3) / 1e12
        # every convolution is timed alone (5 launches of 50-250 us with an L2 flush in between): the burst peak is
        # the denominator; the sustained one is given beside it
        peak = peaks["bf16_tflops"]
        passes = {"bf16": 1.0, "tf32": 2.0, "bf16x3": 3.0, "f16f8": 2.0}[args.precision]
        line["roofline"] = {
            "kernel": "conv_tc / conv_halo / conv_halo2 kernels (tcgen05 implicit GEMM, all %d conv launches of one step)"
                      % len(conv_rows),
            "bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
            "achieved_note": "ALGORITHMIC FLOPs (SURVEY 8d: 2*Cout*Cin*k*k*Ho*Wo over real channels, %.3f TFLOP per step) / "
                             "summed conv kernel time" % (alg_flop / 1e12),
            "peak_source": "%s bf16 burst (kernels timed in isolation); sustained %.1f -> frac %.3f" % (
                peaks["source"], peaks["bf16_tflops_sustained"], tf / peaks["bf16_tflops_sustained"]),
            "traffic": traffic.get("conv_bytes_per_step") if (args.precision == "f16f8" and args.workload == "c4") else None,
            "traffic_note": "DRAM read+write bytes of all conv launches of one step, ncu (profiles/)",
            "ms_per_step": conv_ms, "share_of_step": conv_ms / ms_per_step,
            # what the tensor pipe is actually asked to do: this precision mode spends `passes` bf16-pass-equivalents per
            # algorithmic MAC (f16f8: one fp16 pass + one fp8 pass of twice the K at twice the rate)
            "tensor_pipe_equiv": {"bf16_pass_equivalents_per_mac": passes, "achieved": tf * passes,
                                  "frac_of_burst": tf * passes / peak,
                                  "frac_of_sustained": tf * passes / peaks["bf16_tflops_sustained"],
                                  "issued_incl_padding": issued_bf16_time / (conv_ms * 1e-3) / 1e12},
        }
        # cost volume at BASELINE config 2: NCHW fp32, batch 5, 2 views
        refd, srcd = st["ref"], st["src"].view(V * B, 3, H, W)
        hk = dplan.hk
        vol = torch.empty((V * B, D, H, W), dtype=torch.float32, device=dev)
        tex = torch.empty((lib.cnm_packed_source_bytes(V * B, H, W),), dtype=torch.uint8, device=dev)

        def cv_sweep():
            _lib.check(lib.cnm_plane_sweep(refd.data_ptr(), srcd.data_ptr(), hk.data_ptr(), vol.data_ptr(), B, V, H, W, D,
                                           0.1, 3.0, _lib.VOLUME_NCHW_F32, 0, s))

        def cv_gather():
            _lib.check(lib.cnm_cost_volume(refd.data_ptr(), tex.data_ptr(), hk.data_ptr(), vol.data_ptr(), B, V, H, W, D,
                                           0.1, 3.0, _lib.VOLUME_NCHW_F32, 0, s))

        _lib.check(lib.cnm_pack_source(srcd.data_ptr(), tex.data_ptr(), V * B, H, W, s))
        cv_sweep()
        ms_cvk, _ = time_events(cv_sweep, 20, flush)
        ms_cvg, _ = time_events(cv_gather, 20, flush)
        cv_bytes = B * (V * 4 * D * H * W + (V + 1) * 4 * 3 * H * W)   # = C2_COST_VOLUME_BYTES on the default workload
        gbs = cv_bytes / (ms_cvk * 1e-3) / 1e9
        line["roofline_cost_volume"] = {
            "kernel": "plane_sweep_kernel (%s: %dx%dx%d, %d views, batch %d, fp32 NCHW; TMA-staged source boxes)" % (
                "C2" if args.workload == "c4" else "C5 size", W, H, D, V, B), "bound": "hbm",
            "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "traffic": traffic.get("cost_volume_c2_bytes") if args.workload == "c4" else None,
            "us_per_launch": ms_cvk * 1e3, "algorithmic_bytes": cv_bytes, "peak_source": peaks["source"],
            "first_generation_kernel_us": ms_cvg * 1e3,
        }
        del vol, tex
        # depth -> normal at C3 and the aggregation at C4
        nrm = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        ms_n, _ = time_events(lambda: _lib.check(lib.cnm_depth2normal(st["depth"].data_ptr(), st["Kinv"].data_ptr(),
                                                                      nrm.data_ptr(), None, B, H, W, K_SIZE, s)), 20, flush)
        nb = B * 16 * H * W
        line["roofline_normals"] = {"kernel": "depth2normal_kernel (C3: 256x192, batch 5, k=9)", "bound": "hbm",
                                    "achieved": nb / (ms_n * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": nb / (ms_n * 1e-3) / 1e9 / peaks["hbm_gbs"], "us_per_launch": ms_n * 1e3,
                                    "algorithmic_bytes": nb}
        if args.workload == "c4":
            # the same kernel at C5 size (640x480, batch 32: 157 MB), where it leaves the launch-latency regime
            B5, H5, W5 = 32, 480, 640
            d5 = torch.rand((B5, H5, W5), dtype=torch.float32, device=dev) * 4.0 + 0.5
            k5 = st["Kinv"][:1].expand(B5, 3, 3).contiguous()
            n5 = torch.empty((B5, 3, H5, W5), dtype=torch.float32, device=dev)
            ms_n5, _ = time_events(lambda: _lib.check(lib.cnm_depth2normal(d5.data_ptr(), k5.data_ptr(), n5.data_ptr(), None,
                                                                           B5, H5, W5, K_SIZE, s)), 10, flush)
            nb5 = B5 * 16 * H5 * W5
            line["roofline_normals_c5"] = {"kernel": "depth2normal_kernel (C5 size: 640x480, batch 32, k=9)", "bound": "hbm",
                                           "achieved": nb5 / (ms_n5 * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                           "frac": nb5 / (ms_n5 * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                           "us_per_launch": ms_n5 * 1e3, "algorithmic_bytes": nb5}
            del d5, n5
        # Combined Normal Map (K3b) on the same normals: synthetic disjoint plane masks, reference semantics
        segs_np, pn_np = synth.plane_segments(np.random.default_rng(3), B, H, W)
        segs = torch.from_numpy(segs_np).to(dev)
        pn = torch.tensor([int(v) for v in pn_np], dtype=torch.int32, device=dev)
        P = segs.shape[1]
        cws = torch.empty((lib.cnm_plane_regularize_workspace_bytes(B, P, H, W),), dtype=torch.uint8, device=dev)
        ms_c, _ = time_events(lambda: _lib.check(lib.cnm_plane_regularize(nrm.data_ptr(), segs.data_ptr(), pn.data_ptr(), None,
                                                                          cws.data_ptr(), B, P, H, W, s)), 20, flush)
        cb = B * (2 * 12 * H * W + P * H * W)   # normals read + written, P mask bytes per pixel
        line["roofline_cnm"] = {"kernel": "cnm_accumulate / solve / apply kernels (C3: 256x192, batch 5, %d mask slots, %.1f planes per image)"
                                          % (P, float(pn.float().mean())), "bound": "hbm",
                                "achieved": cb / (ms_c * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": cb / (ms_c * 1e-3) / 1e9 / peaks["hbm_gbs"], "us_per_launch": ms_c * 1e3,
                                "algorithmic_bytes": cb}
        # refine-input aggregation (K4) as the pipeline runs it: act-format iconv1 of both views in, refine conv1.0 operands out
        rpl = st["rplan"]
        ms_a, _ = time_events(lambda: rpl.ops[0](s), 20, flush)
        ab_fp32 = B * (2 * 65 * 4 * H * W + 67 * 4 * H * W)             # SURVEY 8d figure (fp32 NCHW tensors)
        esz = dplan.iconv1.element_size() * dplan.iconv1.shape[3]      # stored bytes per pixel of a 64-channel act tensor
        ab_moved = B * H * W * (V * esz + V * 4 + esz + rpl.taps.element_size() * rpl.taps.shape[3] + 8)
        line["roofline_refine_aggregate"] = {
            "kernel": "refine_taps_kernel + refine_sum_act_kernel (C4: batch 5, 2 views)", "bound": "hbm",
            "achieved": ab_fp32 / (ms_a * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": ab_fp32 / (ms_a * 1e-3) / 1e9 / peaks["hbm_gbs"], "us_per_step": ms_a * 1e3,
            "algorithmic_bytes": ab_fp32,
            "note": "algorithmic bytes = SURVEY 8d (fp32 NCHW tensors of the reference: read 2*(1+64)*4*HW, write 67*4*HW "
                    "per sample); the kernels move %d bytes of the stored act formats" % ab_moved}
        # the other tensor-core modes, same workload, same timing method (kept out of the headline value)
        line["other_modes"] = []
        others = [m for m in ("bf16", "tf32", "bf16x3", "f16f8") if m != args.precision]
        if args.workload != "c4" or args.no_other_modes:
            others = []   # one mode at a time at 640x480 x 32: a plan holds ~60-120 GB of activations
        for other in others:
            d2, r2 = make_weights(0)
            d2.precision = r2.precision = cnmnet_b200.engine.parse_dtype(other)
            pipe2 = cnmnet_b200.CNMPipeline(d2.to(dev), r2.to(dev), k_size=K_SIZE, use_graph=True)
            st2 = pipe2.static_state(B, V, H, W, dev)
            pipe2.load_inputs(st2, *host, non_blocking=False)
            for _ in range(3):
                pipe2.step(st2)
            ms2, _ = time_events(lambda: pipe2.step(st2), max(3, min(args.steps, 10)), flush)
            line["other_modes"].append({"dtype": other, "value": B / (ms2 * 1e-3), "unit": "frames/s", "ms_per_step": ms2,
                                        "n_gpus": 1})
            del pipe2, st2, d2, r2
            torch.cuda.empty_cache()
        if args.workload == "c4" and args.precision == "f16f8" and not args.no_other_modes:
            # (f)-1 as an A/B at step level: the same step with conv1.0 computing the cost volume itself
            # (cnm_conv2d_halo2_sweep, engine kernel name "halo2s"); same timing method, kept out of the headline value
            eng = cnmnet_b200.engine
            was = eng.FUSE_SWEEP
            try:
                eng.FUSE_SWEEP = "force"
                d2, r2 = make_weights(0)
                pipe2 = cnmnet_b200.CNMPipeline(d2.to(dev), r2.to(dev), k_size=K_SIZE, use_graph=True)
                st2 = pipe2.static_state(B, V, H, W, dev)
                op1 = st2["dplan"].convs["conv1.0"]
                pipe2.load_inputs(st2, *host, non_blocking=False)
                for _ in range(3):
                    pipe2.step(st2)
                # interleaved blocks of both arms: under the power cap the step time drifts by several per cent over a
                # run, more than the difference that is being measured
                blocks = {"default": [], "fused": []}
                for rd in range(6):
                    for arm in (("default", "fused") if rd % 2 == 0 else ("fused", "default")):
                        fn = (lambda: pipe.step(st)) if arm == "default" else (lambda: pipe2.step(st2))
                        blocks[arm].append(time_events(fn, 5, flush)[0])
                ms2, ms_def = statistics.median(blocks["fused"]), statistics.median(blocks["default"])
                ms_c, _ = time_events(lambda: op1(s), 10, flush)
                line["fused_sweep"] = {
                    "what": "the same step with conv1.0 computing the plane-sweep cost volume itself (SURVEY 8f rank 1: "
                            "cnm_conv2d_halo2_sweep; the volume never exists as a tensor); bit-identical outputs "
                            "(tests/test_gpu_fused_sweep.py); opt-in (CNM_FUSE_SWEEP=force), not the default, because it is slower",
                    "kernel": op1.backend, "value": B / (ms2 * 1e-3), "unit": "frames/s", "ms_per_step": ms2,
                    "conv1_0_ms_single_launch": ms_c, "default_ms_per_step_interleaved": ms_def,
                    "method": "6 alternating blocks of 5 timed steps per arm in this process, medians"}
                del pipe2, st2, d2, r2
                torch.cuda.empty_cache()
            finally:
                eng.FUSE_SWEEP = was
        if args.layers:
            for name, ms, fl in sorted(conv_rows, key=lambda r: -r[1]):
                sys.stderr.write("%-40s %8.3f ms  %7.1f TFLOP/s issued\n" % (name, ms, fl / ms / 1e9))
        if world == 1 and args.workload == "c4" and not args.no_reference_gpu:
            line["reference_gpu"] = reference_gpu_block(dev, host)
        if not args.no_cpu_baseline and world == 1:
            fps_cpu, ts, kind = cpu_reference_fps(max(1, args.cpu_reps), 1, B_PER_GPU)
            line["cpu_baseline"] = {"value": fps_cpu, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": kind,
                                    "sample": "%d timed steps (+1 warm-up) of one batch of %d frame tuples at 256x192x64: %s"
                                              % (len(ts), B_PER_GPU, "the reference's own modules (baseline/_ref)"
                                                 if kind == "reference" else "fp32 torch-CPU port of the reference")}
        emit(line)
        if os.environ.get("CNM_SAVE_TUNE_CACHE"):   # with CNM_AUTOTUNE=1: how cnmnet_b200/tune/b200.json is (re)made
            cnmnet_b200.engine.save_tune_cache(os.environ["CNM_SAVE_TUNE_CACHE"])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
