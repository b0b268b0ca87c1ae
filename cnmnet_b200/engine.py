"""Host-side layer programs for the B200 kernels: BN folding, weight packing, K-step lists and
static execution plans for DepthNet / DepthRefineNet.

A *plan* owns every activation buffer for one (batch, H, W, dtype) shape and a flat list of
pre-bound C-ABI calls, so running it is a loop of ctypes calls with fixed device pointers:
no allocation, no host sync, CUDA-graph capturable.  The wiring follows the reference
(depthnet/depthNet_model.py:226-263 and :331-370); torch.cat never happens -- concatenated
inputs are extra operand sources of the implicit GEMM (see csrc/conv_tc.cu).
"""
import ctypes as C
import os
from collections import namedtuple

import torch

from . import _lib
from ._lib import CNM_BF16, CNM_BF16X3, CNM_F16F8, CNM_TF32

BN_EPS = 1e-5
# which stride-1 layers use the row-halo kernel: "auto" (enough tiles for a wave), "force" (every eligible layer,
# used by tests), "off" (per-tap kernel everywhere)
HALO_MODE = os.environ.get("CNM_CONV_HALO", "auto")
# CNM_POISON=1 (a test mode, tools/sanitize.sh): every plan buffer starts as 0xFF bytes (NaN in every act format and in
# fp32) and buffers are not recycled, so a kernel that reads an element nobody wrote turns the outputs into NaN and
# fails the parity tests.  compute-sanitizer's initcheck cannot do this job here: it does not see TMA bulk stores (the
# async proxy) as writes, so every consumer of a convolution output is flagged.
POISON = os.environ.get("CNM_POISON", "0") == "1"
PLAN_REUSE = os.environ.get("CNM_PLAN_REUSE", "1") != "0" and not POISON   # activation buffers of a plan are recycled after their last reader
FUSE_HEADS = os.environ.get("CNM_FUSE_HEADS", "1") != "0"   # 64-channel heads in the producing conv's epilogue
DENSE_ROWTAPS = os.environ.get("CNM_DENSE_ROWTAPS", "1") != "0"   # conv1.0's image taps: three kernel rows per K-step
# iconv3 / iconv2 / iconv1: the upsampled disparity enters in the epilogue, in fp32 on unrounded values (cnm_conv_desc.side_disp),
# instead of as a 64-channel side tensor + K-step.  Exact and tested, saves 165 MB of buffers per lane at C4 -- and is NOT
# faster (tools/ab_step.py: +1.6 % step time on all three layers, +-0 on iconv3 / iconv2 only: 576 extra FMAs per pixel and 64
# output channels in an epilogue that the Cout = 64 layer cannot hide), so it is opt-in: CNM_SIDE_DISP=1.
SIDE_DISP = os.environ.get("CNM_SIDE_DISP", "0") == "1"
SIDE_DISP_MIN_CIN = int(os.environ.get("CNM_SIDE_DISP_MIN_CIN", "64"))   # layers with at least this many input channels
PLANE_SWEEP = os.environ.get("CNM_PLANE_SWEEP", "1") != "0"   # K1 = csrc/plane_sweep.cu (TMA-staged); 0: cost_volume.cu
HALO2_ENABLED = os.environ.get("CNM_CONV_HALO2", "1") != "0"   # CTA-pair (cta_group::2) variant as a tuning candidate
TC_MULTICAST = os.environ.get("CNM_TC_MULTICAST", "1") != "0"   # per-tap kernel on clusters sharing the weight tiles (tcm2 / tcm4)
# conv1.0 with the plane sweep fused in (cnm_conv2d_halo2_sweep, kernel name "halo2s"): "1" = available to the variant
# table / autotuner, "force" = used wherever it applies, "0" = off
FUSE_SWEEP = os.environ.get("CNM_FUSE_SWEEP", "1")
TORCH_DTYPE = {CNM_BF16: torch.bfloat16, CNM_TF32: torch.float32, CNM_BF16X3: torch.bfloat16, CNM_F16F8: torch.float16}
F8_LO_SCALE, F8_HI_SCALE = 128.0, 0.0625   # csrc/common.cuh: lo8 = e4m3((v - hi) * 2^7), hi8 = e4m3(hi * 2^-4)


def chunk_channels(dtype):
    """Channels per K-chunk (one 128-byte operand row)."""
    return 32 if dtype == CNM_TF32 else 64


def phys_factor(dtype):
    """Physical channels per logical channel: CNM_BF16X3 stores every value as a (hi, lo) bf16 pair, laid out as a
    128-byte hi row followed by a 128-byte lo row per 64-channel logical chunk (see csrc/common.cuh: bf16s)."""
    return 2 if dtype in (CNM_BF16X3, CNM_F16F8) else 1


# (activation part, weight part) passes of one logical K-chunk: x*w ~ hi*Whi + lo*Whi + hi*Wlo for split bf16;
# fp16 row x fp16 weights, then fp8 row [lo8 | hi8] x fp8 weights [Whi8 | Wlo8] for CNM_F16F8 (both rows 64 physical
# channels apart, like the hi / lo rows of split bf16)
SPLIT_PASSES = (("hi", "hi"), ("lo", "hi"), ("hi", "lo"))
F16F8_PASSES = (("hi", "h16"), ("lo", "f8"))


def split_passes(dtype):
    """False for the one-row dtypes, else the pass list of the two-row (split) act layouts."""
    return SPLIT_PASSES if dtype == CNM_BF16X3 else (F16F8_PASSES if dtype == CNM_F16F8 else False)


def parse_dtype(d):
    if d in (CNM_BF16, "bf16", torch.bfloat16):
        return CNM_BF16
    if d in (CNM_TF32, "tf32", torch.float32):
        return CNM_TF32
    if d in (CNM_BF16X3, "bf16x3"):
        return CNM_BF16X3
    if d in (CNM_F16F8, "f16f8"):
        return CNM_F16F8
    raise ValueError("precision must be 'bf16', 'tf32', 'bf16x3' or 'f16f8', got %r" % (d,))


def round_tf32(t):
    """Round-to-nearest (ties away) onto the 10-bit tf32 mantissa, like cvt.rna.tf32.f32."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & -8192).view(torch.float32)


def fold_bn(sd, prefix, ci):
    """conv (no bias) + BatchNorm2d(eval) -> (w', b') with w' = w*g/sqrt(v+eps), b' = beta - mu*g/sqrt(v+eps)."""
    w = sd["%s.%d.weight" % (prefix, ci)].double()
    p = "%s.%d." % (prefix, ci + 1)
    scale = sd[p + "weight"].double() / torch.sqrt(sd[p + "running_var"].double() + BN_EPS)
    wf = w * scale.view(-1, 1, 1, 1)
    bf = sd[p + "bias"].double() - sd[p + "running_mean"].double() * scale
    return wf.float(), bf.float()


# An operand source of a convolution.
#   kind 'act'      full act tensor, channels [coff, coff+C) <-> weight input channels [wch, wch+C)
#   kind 'rowtaps'  side tensor whose channel s*nreal+c holds tap s of input channel wch+c (one K-step per kernel row)
#   kind 'rowtaps3' side tensor [N, H+2, W, chunk] (one apron row above and below the image: row j is centred on image row
#                   j-1) whose channel (q*k + s)*nreal+c holds tap (row q-1, column s) of input channel wch+c: three kernel
#                   rows per K-step (2-byte dtypes: 3*k*nreal <= 64), taps 3 rows apart (cnm_pack_row_taps3).  The apron
#                   keeps the rows that are centred outside the image but reach into it.
#   kind 'taps3x3'  side tensor whose channel 9c+3r+s holds tap (r,s) of input channel wch+c (a single K-step)
Src = namedtuple("Src", "buf kind wch C coff nreal")


def act_src(buf, wch, C=None, coff=0, phys=1):
    """Full act tensor as an operand source; C / coff are LOGICAL channels (`phys` physical channels each)."""
    return Src(buf, "act", wch, buf.shape[3] // phys if C is None else C, coff, 0)


def _chunk_passes(split):
    """`split` is False or a pass list (split_passes); True means split bf16."""
    return (SPLIT_PASSES if split is True else split) if split else ((None, None),)


def _pc0(c_logical, apart, split):
    """Physical first channel of the logical chunk starting at c_logical (hi row, or lo row 64 further)."""
    return (2 * c_logical + (64 if apart == "lo" else 0)) if split else c_logical


def make_steps(srcs, k, stride, ch, split=False):
    """-> (steps [(src, sub, dx, dy, c0 physical)], cols [packing recipe per step]).  See include/cnmnet_b200.h."""
    pad = (k - 1) // 2
    steps, cols = [], []
    for si, s in enumerate(srcs):
        if s.kind == "act":
            assert s.C % ch == 0 and s.coff % ch == 0
            for r in range(k):
                for t in range(k):
                    ty, tx = r - pad, t - pad
                    if stride == 1:
                        sub, dy, dx = 0, ty, tx
                    else:  # input row 2*oy + ty = 2*(oy + dy) + py
                        py, px = ty % 2, tx % 2
                        dy, dx = (ty - py) // 2, (tx - px) // 2
                        sub = py * 2 + px
                    for c0 in range(0, s.C, ch):
                        for ap, wp_ in _chunk_passes(split):
                            steps.append((si, sub, dx, dy, _pc0(s.coff + c0, ap, split)))
                            cols.append(("act", s.wch + c0, r, t, wp_, ap))
        elif s.kind == "rowtaps":
            assert stride == 1 and k * s.nreal <= ch
            for r in range(k):
                for ap, wp_ in _chunk_passes(split):
                    steps.append((si, 0, 0, r - pad, _pc0(0, ap, split)))
                    cols.append(("rowtaps", s.wch, s.nreal, r, wp_, ap))
        elif s.kind == "rowtaps3":
            assert stride == 1 and 3 * k * s.nreal <= ch
            for T in range((k + 2) // 3):   # kernel rows 3T .. 3T+2, read at their middle row (+1: the apron row)
                for ap, wp_ in _chunk_passes(split):
                    steps.append((si, 0, 0, 3 * T + 2 - pad, _pc0(0, ap, split)))
                    cols.append(("rowtaps3", s.wch, s.nreal, T, wp_, ap))
        elif s.kind == "taps3x3":
            assert stride == 1 and k == 3 and 9 * s.nreal <= ch
            for ap, wp_ in _chunk_passes(split):
                steps.append((si, 0, 0, 0, _pc0(0, ap, split)))
                cols.append(("taps3x3", s.wch, s.nreal, 0, wp_, ap))
        else:
            raise ValueError(s.kind)
    order = _corrections_first(cols)
    return [steps[i] for i in order], [cols[i] for i in order]


def _corrections_first(cols):
    """Order of the K loop for the two-row (split) modes: every correction step (bf16x3: lo*Whi, hi*Wlo; f16f8: the fp8
    row) before every leading-term step (hi*Whi / fp16 row).  The fp32 accumulation inside tcgen05.mma loses ~1 ulp of the
    running sum per MMA instruction (tools/acc_floor.py); the correction terms are 2^-8 ... 2^-12 of the leading ones, so
    accumulating them while the sum is still tiny costs nothing, and the leading terms then see one third (bf16x3) / one
    half (f16f8) of the updates at full magnitude.  The kernels only need to know that the fp8 steps of f16f8 are the first
    half of the list."""
    main = [i for i, c in enumerate(cols) if len(c) > 4 and c[4] in ("hi", "h16") and _apart(c) == "hi"]
    if not main or len(main) == len(cols):
        return list(range(len(cols)))
    ms = set(main)
    return [i for i in range(len(cols)) if i not in ms] + main


def _apart(col):
    return col[5] if len(col) > 5 else "hi"


GROUP_TAP_STRIDE_SHIFT = 28   # cnm_conv_group.ntaps bits 28-30: rows between consecutive taps, minus one


def group_taps(nt):
    """(taps, rows between taps) of a group's ntaps entry as make_groups returns it."""
    return nt & 15, ((nt >> GROUP_TAP_STRIDE_SHIFT) & 7) + 1


def make_groups(srcs, k, ch, stride=1, split=False):
    """Row-halo K loop (csrc/conv_halo.cu): -> (groups [(map, c0 physical, dx, ntaps)], cols, map_info [12][2]).
    One group = one slab load serving `ntaps` vertical taps (ntaps may carry a tap stride: group_taps); the weight steps
    (cols) follow group order.  map = src*4 + sub; map_info[map] = (rows of the slab above the tile, slab rows)."""
    pad = (k - 1) // 2
    groups, cols = [], []
    info = [[0, 0] for _ in range(12)]
    for si, s in enumerate(srcs):
        if s.kind == "act" and stride == 1:
            info[si * 4] = [pad, 16 + 2 * pad]
            for c0 in range(0, s.C, ch):
                for t in range(k):
                    for ap, wp_ in _chunk_passes(split):
                        groups.append((si * 4, _pc0(s.coff + c0, ap, split), t - pad, k))
                        cols += [("act", s.wch + c0, r, t, wp_, ap) for r in range(k)]
        elif s.kind == "act":  # stride 2: taps grouped by (column tap, row parity) on the parity sub-images
            for c0 in range(0, s.C, ch):
                for t in range(k):
                    tx = t - pad
                    px, qx = tx % 2, (tx - tx % 2) // 2
                    for py in (0, 1):
                        rs = [r for r in range(k) if (r - pad) % 2 == py]
                        if not rs:
                            continue
                        qs = [(r - pad - py) // 2 for r in rs]
                        assert qs == list(range(qs[0], qs[0] + len(qs)))
                        m = si * 4 + py * 2 + px
                        info[m] = [-qs[0], 16 + qs[-1] - qs[0]]
                        for ap, wp_ in _chunk_passes(split):
                            groups.append((m, _pc0(s.coff + c0, ap, split), qx, len(rs)))
                            cols += [("act", s.wch + c0, r, t, wp_, ap) for r in rs]
        elif s.kind == "rowtaps":
            assert stride == 1 and k * s.nreal <= ch
            info[si * 4] = [pad, 16 + 2 * pad]
            for ap, wp_ in _chunk_passes(split):
                groups.append((si * 4, _pc0(0, ap, split), 0, k))
                cols += [("rowtaps", s.wch, s.nreal, r, wp_, ap) for r in range(k)]
        elif s.kind == "rowtaps3":
            assert stride == 1 and 3 * k * s.nreal <= ch and k >= 5
            nT = (k + 2) // 3
            info[si * 4] = [pad - 2, 16 + 3 * (nT - 1)]   # first tap at tensor row 2 - pad (apron included), taps 3 rows apart
            for ap, wp_ in _chunk_passes(split):
                groups.append((si * 4, _pc0(0, ap, split), 0, nT | (2 << GROUP_TAP_STRIDE_SHIFT)))
                cols += [("rowtaps3", s.wch, s.nreal, T, wp_, ap) for T in range(nT)]
        elif s.kind == "taps3x3":
            assert stride == 1 and k == 3 and 9 * s.nreal <= ch
            info[si * 4] = [0, 16]
            for ap, wp_ in _chunk_passes(split):
                groups.append((si * 4, _pc0(0, ap, split), 0, 1))
                cols.append(("taps3x3", s.wch, s.nreal, 0, wp_, ap))
        else:
            raise ValueError(s.kind)
    # corrections first, group-wise (a group's weight steps stay together and in order)
    starts, pos = [], 0
    for g in groups:
        starts.append(pos)
        pos += g[3] & 15
    gorder = _corrections_first([cols[st] for st in starts])
    groups2 = [groups[i] for i in gorder]
    cols2 = [c for i in gorder for c in cols[starts[i]:starts[i] + (groups[i][3] & 15)]]
    return groups2, cols2, info


def _gather_cols(w, cols, ch, want=None):
    """[Cout, Cin, k, k] -> [Cout, len(cols)*ch] fp32 by the per-step recipes; with `want`, only the column blocks whose
    weight part is in `want` are filled."""
    cout, _, k, _ = w.shape
    wp = torch.zeros((cout, len(cols) * ch), dtype=torch.float32, device=w.device)
    for i, col in enumerate(cols):
        if want is not None and (col[4] if len(col) > 4 else None) not in want:
            continue
        base = i * ch
        if col[0] == "act":
            _, c, r, t = col[:4]
            wp[:, base:base + ch] = w[:, c:c + ch, r, t]
        elif col[0] == "rowtaps":
            _, wch, n, r = col[:4]
            wp[:, base:base + k * n] = w[:, wch:wch + n, r, :].permute(0, 2, 1).reshape(cout, k * n)
        elif col[0] == "rowtaps3":
            _, wch, n, T = col[:4]
            for q in range(3):
                if 3 * T + q < k:
                    wp[:, base + q * k * n:base + (q + 1) * k * n] = \
                        w[:, wch:wch + n, 3 * T + q, :].permute(0, 2, 1).reshape(cout, k * n)
        else:
            _, wch, n, _ = col[:4]
            wp[:, base:base + 9 * n] = w[:, wch:wch + n].reshape(cout, 9 * n)
    return wp


def _e4m3(t):
    return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn)


def f16f8_weight_exp(wf):
    """acc_exp e of CNM_F16F8 weights: the largest power of two with max|w| * 2^(e-7) <= 448 (e4m3 range), which also
    keeps w * 2^e <= 57344 inside fp16."""
    m = float(wf.abs().max())
    if m == 0.0:
        return 0
    import math
    return int(math.floor(math.log2(448.0 / m))) + 7


def pack_weights(wf, cols, ch, dtype, acc_exp=None):
    """Folded weights [Cout, Cin, k, k] -> K-major operand [Cout, len(cols)*ch] in the act dtype.  For CNM_BF16X3 a
    column block holds either Whi = bf16(w) or Wlo = bf16(w - Whi), as its recipe says.  For CNM_F16F8 (include/
    cnmnet_b200.h) an 'h16' block is W16 = fp16(w * 2^e) and an 'f8' block is 128 bytes [e4m3(w * 2^(e-7)) x 64 |
    e4m3((w * 2^e - W16) * 2^4) x 64]; e = acc_exp (default f16f8_weight_exp(wf))."""
    if dtype == CNM_F16F8:
        e = f16f8_weight_exp(wf) if acc_exp is None else acc_exp
        ws = wf * (2.0 ** e)
        w16 = ws.clamp(-65504.0, 65504.0).to(torch.float16)
        cout = wf.shape[0]
        out = torch.zeros((cout, len(cols), 128), dtype=torch.uint8, device=wf.device)
        h16 = _gather_cols(w16.float(), cols, ch, ("h16",)).to(torch.float16).view(cout, len(cols), ch)
        hi8 = _e4m3(_gather_cols(ws / F8_LO_SCALE, cols, ch, ("f8",))).view(cout, len(cols), ch)
        lo8 = _e4m3(_gather_cols((ws - w16.float()) / F8_HI_SCALE, cols, ch, ("f8",))).view(cout, len(cols), ch)
        for i, col in enumerate(cols):
            if col[4] == "h16":
                out[:, i] = h16[:, i].contiguous().view(torch.uint8)
            else:
                out[:, i, :64] = hi8[:, i].contiguous().view(torch.uint8)
                out[:, i, 64:] = lo8[:, i].contiguous().view(torch.uint8)
        return out.view(cout, len(cols) * 128).view(torch.float16)
    parts = {None: wf, "hi": wf, "lo": wf}
    if dtype == CNM_BF16X3:
        whi = wf.to(torch.bfloat16).float()
        parts = {"hi": whi, "lo": wf - whi}
    wp = None
    for part, w in parts.items():
        blk = _gather_cols(w, cols, ch, (part,))
        wp = blk if wp is None else wp + blk
    if dtype == CNM_TF32:
        return round_tf32(wp)
    return wp.to(torch.bfloat16).contiguous()


_TUNE_CACHE = {}
SPLITK = os.environ.get("CNM_SPLITK", "1") != "0"   # split-K candidates for layers with fewer tiles than SMs
_SPLITK_WS = {}   # device -> shared fp32 workspace (layers run one after the other)


def splitk_workspace(device, nfloats, lane=0):
    """One workspace per (device, lane): plans of different lanes may run concurrently on different streams."""
    key = (str(device), lane)
    ws = _SPLITK_WS.get(key)
    if ws is None or ws.numel() < nfloats:
        ws = _SPLITK_WS[key] = torch.empty((nfloats,), dtype=torch.float32, device=device)
    return ws

_BALANCE_WS = {}   # (device, lane) -> zero-initialised workspace of the balanced (stream-K) schedules
BALANCE = os.environ.get("CNM_CONV_BALANCE", "1") != "0"   # "tcb" / "halob" / "halo2b" as variants (cnm_conv_desc.balance)


def balance_workspace(device, lane=0):
    """Flags + one accumulator slot per CTA (include/cnmnet_b200.h: cnm_conv_desc.balance).  Zeroed once; the kernels
    leave the flags zeroed.  One per (device, lane): launches sharing it must be stream-ordered."""
    key = (str(device), lane)
    ws = _BALANCE_WS.get(key)
    if ws is None:
        n = int(_lib.load().cnm_conv_balance_ws_bytes())
        ws = _BALANCE_WS[key] = torch.zeros((n // 4,), dtype=torch.float32, device=device)
    return ws


# Kernel-variant choice (which of the three convolution kernels, output-channel tile, split-K factor) per layer shape:
#   "cache"  (default)  the shipped table cnmnet_b200/tune/b200.json (+ CNM_TUNE_CACHE, a JSON file of the same format that
#                        overrides it); shapes it does not list run the deterministic default of ConvOp.__init__.  Nothing is
#                        timed, so two processes -- or two boxes -- run the same kernels in the same K order and produce
#                        bit-identical outputs.
#   "auto"   (CNM_AUTOTUNE=1)  shapes missing from the table time their candidates once on the device (how the table is
#                        made: tools/tune_report.py, bench.py with CNM_SAVE_TUNE_CACHE).  Variants differ in K order, so
#                        the result of an auto-tuned process is reproducible only through the table it saves.
#   "off"    (CNM_AUTOTUNE=off)  defaults only (tests cover the untuned variants this way).
TUNE_MODE = {"1": "auto", "auto": "auto", "off": "off", "none": "off"}.get(os.environ.get("CNM_AUTOTUNE", "0"), "cache")
SHIPPED_TUNE_CACHE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tune", "b200.json")
TUNE_CACHE_FILE = os.environ.get("CNM_TUNE_CACHE", "")


def load_tune_cache(path):
    """Merges a saved table ({repr(key): [kernel, tile, ksplit]}) into the process's table."""
    import json
    with open(path) as f:
        for k, v in json.load(f).items():
            _TUNE_CACHE[k] = (v[0], int(v[1]), int(v[2]) if len(v) > 2 else 1, None)


def save_tune_cache(path):
    import json
    with open(path, "w") as f:
        json.dump({k: [v[0], v[1], v[2]] for k, v in _TUNE_CACHE.items()}, f, indent=0, sort_keys=True)


for _p in (SHIPPED_TUNE_CACHE, TUNE_CACHE_FILE):
    if _p == SHIPPED_TUNE_CACHE and os.environ.get("CNM_NO_SHIPPED_TUNE", "0") == "1":
        continue   # re-making the table: every shape is timed afresh
    if _p and os.path.exists(_p):
        load_tune_cache(_p)


class ConvOp:
    """One pre-bound convolution call.  Holds the per-tap program (cnm_conv2d_tc / _simt) and, for stride-1 layers
    with at least one 16x16 tile, the row-halo program (cnm_conv2d_halo); `tune()` times the candidates
    (kernel x output-channel tile) on the device once per layer shape and keeps the fastest."""

    def __init__(self, name, srcs, k, stride, cout, N, Ho, Wo, dtype, device, relu=True, out=None, c_out_offset=0,
                 backend="tc", lane=0):
        self.name, self.k, self.stride, self.dtype, self.lane = name, k, stride, dtype, lane
        ch = chunk_channels(dtype)
        split = split_passes(dtype)
        pf = phys_factor(dtype)
        self.srcs = srcs
        self.variants = {}
        steps, cols = make_steps(srcs, k, stride, ch, split)
        st = torch.zeros((len(steps), 8), dtype=torch.int32)
        st[:, :5] = torch.tensor(steps, dtype=torch.int32)
        self.variants["tc"] = {"cols": cols, "steps_dev": st.to(device)}
        self.variants["simt"] = self.variants["tc"]
        if dtype != CNM_TF32 and cout % 128 == 0 and backend != "simt" and TC_MULTICAST:
            # same program on clusters of 2 / 4 CTAs sharing every weight tile by TMA multicast (cnm_conv_desc.cluster)
            self.variants["tcm2"] = self.variants["tcm4"] = self.variants["tc"]
        if dtype != CNM_TF32 and backend != "simt" and BALANCE:
            self.variants["tcb"] = self.variants["tc"]   # same program, balanced (stream-K) schedule
        halo_ok = Ho >= 16 and Wo >= 16 and backend in ("tc", "tcb", "halo", "halo2", "halo2h", "halob", "halo2b") and \
            (stride == 1 or all(s.kind == "act" for s in srcs))
        if halo_ok:
            groups, gcols, info = make_groups(srcs, k, ch, stride, split)
            assert len(gcols) == len(cols)
            flat = [v for pr in info for v in pr]
            gt = torch.tensor(groups, dtype=torch.int32)
            taps = gt[:, 3] & 15
            gt[:, 3] += (torch.cumsum(taps, 0) - taps) << 8   # cnm_conv_group.ntaps: taps | first weight step << 8 (| tap stride)
            self.variants["halo"] = {"cols": gcols, "groups_dev": gt.to(device),
                                     "ngroups": len(groups), "pads": (C.c_int32 * len(flat))(*flat)}
            if (N * ((Ho + 15) // 16) * ((Wo + 15) // 16)) % 2 == 0 and HALO2_ENABLED:
                self.variants["halo2"] = self.variants["halo"]   # same program, CTA-pair kernel
            if (N * ((Ho + 7) // 8) * ((Wo + 15) // 16)) % 2 == 0 and HALO2_ENABLED and dtype != CNM_TF32:
                self.variants["halo2h"] = self.variants["halo"]  # same program, CTA pairs on 16x8 blocks (tile_h = 8)
            if dtype != CNM_TF32 and BALANCE:   # same programs, balanced (stream-K) schedule; column tile <= 128
                self.variants["halob"] = self.variants["halo"]
                if "halo2" in self.variants:
                    self.variants["halo2b"] = self.variants["halo"]
        elif backend in ("halo", "halo2", "halo2h", "halob", "halo2b"):
            raise ValueError("%s: the row-halo kernel needs an output of at least 16x16" % name)
        self.num_steps = len(cols)
        self.out = out if out is not None else torch.empty((N, Ho, Wo, cout * pf), dtype=TORCH_DTYPE[dtype],
                                                           device=device)
        self.wp = torch.zeros((cout, self.num_steps * ch), dtype=TORCH_DTYPE[dtype], device=device)
        self.bias = torch.zeros((cout,), dtype=torch.float32, device=device)
        self.desc = _lib.ConvDesc(N, Ho, Wo, cout, self.out.shape[3], c_out_offset, stride, 1 if relu else 0,
                                  len(srcs), self.num_steps, dtype, 0)
        arr = (_lib.ConvSrc * len(srcs))()
        for i, s in enumerate(srcs):
            arr[i] = _lib.ConvSrc(_lib.ptr(s.buf), s.buf.shape[1], s.buf.shape[2], s.buf.shape[3], 1 if s.kind == "rowtaps3" else 0)
        self.srcs_arr = arr
        # issued MACs x 2, padding channels included.  CNM_F16F8: every second K-step is an fp8 row (K = 128);
        # `flops_bf16_time` weighs those at half (kind::f8f6f4 runs at twice the kind::f16 rate), i.e. it is the
        # tensor-pipe time of the layer expressed in bf16 FLOPs
        self.flops = 2.0 * N * Ho * Wo * cout * self.num_steps * ch * (1.5 if dtype == CNM_F16F8 else 1.0)
        self.flops_bf16_time = 2.0 * N * Ho * Wo * cout * self.num_steps * ch
        self._weights = None
        self.tuned = None
        self.sweep = None
        self.key = repr((dtype, N, Ho, Wo, cout, k, stride, tuple((s.kind, s.C, s.buf.shape[3]) for s in srcs)))
        # default choice without tuning
        halo_tiles = N * ((Ho + 15) // 16) * ((Wo + 15) // 16) * max(1, cout // 256)
        if backend == "tc" and halo_ok and HALO_MODE != "off" and (halo_tiles >= 120 or HALO_MODE == "force"):
            backend = "halo"
        self.tunable = backend in ("tc", "halo", "halo2", "halo2h", "tcb", "halob", "halo2b") and HALO_MODE == "auto"
        self.select(backend, 0)

    @property
    def cols(self):
        return self.variants[self.backend]["cols"]

    def select(self, backend, bn, ksplit=1):
        repack = getattr(self, "backend", backend) != backend and self._weights is not None
        self.backend = backend
        if backend == "halo2s" and not self.sweep.scratch:
            nbytes = int(_lib.load().cnm_sweep_scratch_bytes(self.sweep.D, self.k))
            self._sweep_scratch = torch.empty((nbytes,), dtype=torch.uint8, device=self.out.device)
            if POISON:
                self._sweep_scratch.fill_(0xFF)
            self.sweep.scratch, self.sweep.scratch_bytes = self._sweep_scratch.data_ptr(), nbytes
        self.desc.reserved = int(bn)
        self.desc.cluster = {"tcm2": 2, "tcm4": 4}.get(backend, 0)
        self.desc.tile_h = 8 if backend == "halo2h" else 0
        self.desc.ksplit = int(ksplit) if backend in ("tc", "tcm2", "tcm4") else 1
        self.desc.balance = 1 if backend in ("tcb", "halob", "halo2b") else 0
        if self.desc.balance:
            self._ws = balance_workspace(self.out.device, self.lane)
            self.desc.splitk_ws = self._ws.data_ptr()
        elif self.desc.ksplit > 1:
            d = self.desc
            ws = splitk_workspace(self.out.device, d.ksplit * d.N * d.Ho * d.Wo * d.Cout, self.lane)
            self._ws = ws
            d.splitk_ws = ws.data_ptr()
        else:
            self.desc.splitk_ws = None
        if repack:
            self.load(*self._weights)

    def candidates(self):
        """(kernel, output-channel tile, split-K factor) combinations worth timing for this layer."""
        d = self.desc
        out = []
        for be in ("tc", "halo", "halo2", "halo2h", "tcb", "halob", "halo2b"):
            if be in self.variants:
                out += [(be, bn, 1) for bn in (256, 128, 64) if d.Cout % bn == 0 and (bn <= 128 or be not in ("halob", "halo2b"))]
        # "halo2s" (fused plane sweep) is not timed here: it replaces the cost-volume launch as well, so it is compared
        # at step level (bench.py, CNM_FUSE_SWEEP=force) and entered into the table by hand if it wins
        mc = [be for be in ("tcm2", "tcm4") if be in self.variants]
        out += [(be, bn, 1) for be in mc for bn in (256, 128) if d.Cout % bn == 0]
        if SPLITK and "tc" in self.variants and not getattr(self, "_head", None) and not getattr(self, "_side", None):
            tw = 16 if d.Wo >= 16 else 8
            th = min(128 // tw, 1 << max(d.Ho - 1, 0).bit_length())
            nb = 128 // (tw * th)
            m_tiles = -(-d.Wo // tw) * -(-d.Ho // th) * -(-d.N // nb)
            for bn in (256, 128, 64):
                if d.Cout % bn:
                    continue
                tiles = m_tiles * (d.Cout // bn)
                for ks in (2, 3, 4, 6, 8):
                    if tiles < 120 and tiles * ks <= 304 and self.num_steps >= 8 * ks:
                        out.append(("tc", bn, ks))
                        out += [(be, bn, ks) for be in mc if bn >= 128]
        return out

    def tune(self, stream, reps=3):
        """Applies the table's choice for this layer shape; in "auto" mode a missing shape times every (kernel, tile,
        split-K) candidate with CUDA events once and enters the fastest into the table."""
        if not self.tunable or TUNE_MODE == "off" or (FUSE_SWEEP == "force" and "halo2s" in self.variants):
            return
        best = _TUNE_CACHE.get(self.key)
        if best is not None and best[0] not in self.variants:
            best = None
        if best is None:
            if TUNE_MODE != "auto":
                return   # not in the table: the deterministic default stays
            results = []
            for be, bn, ks in self.candidates():
                self.select(be, bn, ks)
                for _ in range(2):
                    self(stream)
                ts = []
                for _ in range(reps):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    self(stream)
                    b.record()
                    b.synchronize()
                    ts.append(a.elapsed_time(b))
                results.append((min(ts), be, bn, ks))
            results.sort()
            best = _TUNE_CACHE[self.key] = (results[0][1], results[0][2], results[0][3], results)
        self.tuned = best
        self.select(best[0], best[1], best[2])

    def attach_sweep(self, vol_src, Bref, planes, idmin, idmax, tex, hk):
        """Makes the fused plane-sweep kernel ("halo2s": csrc/conv_halo.cu, SWEEP instantiation) a variant of this layer:
        source `vol_src` is then computed inside the convolution from the packed source views `tex`, the homographies
        `hk` and the reference frames (pointer bound per call through `self.sweep.ref`) instead of being read."""
        d = self.desc
        if FUSE_SWEEP == "0" or "halo2" not in self.variants or self.dtype != CNM_F16F8 or self.stride != 1 or \
                d.Cout > 128 or getattr(self, "_head", None):
            return False
        self._sweep_keep = (tex, hk)
        # the scratch (37 MB at 64 planes) is allocated when the variant is first selected
        self.sweep = _lib.SweepFuse(None, tex.data_ptr(), hk.data_ptr(), None, 0, Bref, planes, vol_src, 0, idmin, idmax)
        self.variants["halo2s"] = self.variants["halo2"]   # same program; the volume's slabs come from the scratch
        if FUSE_SWEEP == "force":
            self.select("halo2s", d.Cout)
        return True

    def attach_side_disp(self, disp, wch):
        """Input channel `wch` of the layer is the nearest-upsampled 1-channel disparity `disp` [N, 1, Ho/2, Wo/2] fp32
        (depthNet_model.py:247-258): it is not among the operand sources; the kernels add its nine taps to the fp32
        accumulator in the epilogue (cnm_conv_desc.side_disp / side_w).  2-byte dtypes, tensor-core kernels."""
        d = self.desc
        assert self.dtype != CNM_TF32 and tuple(disp.shape) == (d.N, 1, d.Ho // 2, d.Wo // 2) and disp.dtype == torch.float32
        self._side = (disp, wch)
        self.side_w = torch.zeros((9, d.Cout), dtype=torch.float32, device=self.out.device)
        d.side_disp, d.side_w = disp.data_ptr(), self.side_w.data_ptr()
        # variants that cannot take it: balanced schedules, 16x8 blocks, multicast clusters
        for be in ("tcb", "halob", "halo2b", "halo2h", "tcm2", "tcm4"):
            self.variants.pop(be, None)
        self.key = self.key[:-1] + ", 'side')"   # its own rows in the variant table

    def fuse_head(self, w_tap_major, partial):
        """Fused 1-channel 3x3 head on this layer's output (cnm_conv_desc.head_w [9][Cout] / head_partial [Cout/64 slices])."""
        assert self.dtype != CNM_TF32
        self._head = (w_tap_major, partial)
        self.desc.head_w, self.desc.head_partial = w_tap_major.data_ptr(), partial.data_ptr()

    def load(self, wf, bf):
        self._weights = (wf, bf)
        if self.dtype == CNM_F16F8:
            self.desc.acc_exp = f16f8_weight_exp(wf)
        self.wp.copy_(pack_weights(wf.to(self.wp.device), self.cols, chunk_channels(self.dtype), self.dtype))
        self.bias.copy_(bf.to(self.bias.device))
        if getattr(self, "_side", None):
            cout = wf.shape[0]
            self.side_w.copy_(wf[:, self._side[1]].reshape(cout, 9).t().to(self.side_w.device))

    def __call__(self, stream):
        lib = _lib.load()
        v = self.variants[self.backend]
        if self.backend == "halo2s":
            if not self.sweep.ref:
                raise RuntimeError("%s: the fused plane sweep has no reference frames bound" % self.name)
            _lib.check(lib.cnm_conv2d_halo2_sweep(C.byref(self.desc), self.srcs_arr, v["pads"], v["groups_dev"].data_ptr(),
                                                  v["ngroups"], self.wp.data_ptr(), self.bias.data_ptr(),
                                                  self.out.data_ptr(), C.byref(self.sweep), stream), self.name)
            return
        if self.backend in ("halo", "halo2", "halo2h", "halob", "halo2b"):
            fn = lib.cnm_conv2d_halo if self.backend in ("halo", "halob") else lib.cnm_conv2d_halo2
            _lib.check(fn(C.byref(self.desc), self.srcs_arr, v["pads"], v["groups_dev"].data_ptr(),
                                           v["ngroups"], self.wp.data_ptr(), self.bias.data_ptr(),
                                           self.out.data_ptr(), stream), self.name)
            return
        fn = lib.cnm_conv2d_simt if self.backend == "simt" else lib.cnm_conv2d_tc
        _lib.check(fn(C.byref(self.desc), self.srcs_arr, v["steps_dev"].data_ptr(), self.wp.data_ptr(),
                      self.bias.data_ptr(), self.out.data_ptr(), stream), self.name)


# ---- bilinear x2 upsampling folded into the consuming 3x3 convolution ------------------------------------------------
# nn.Upsample(scale_factor=2, mode='bilinear') + conv3x3 (depthnet/depthNet_model.py:92-98, `upconv`) is a linear map of
# the LOW-RES tensor x: with u[2i] = .25 x[i-1] + .75 x[i], u[2i+1] = .75 x[i] + .25 x[i+1] (align_corners=False), output
# row 2i+a of the convolution reads low-res rows i-1, i, i+1 with the kernel rows combined by _FOLD_A[a] (low-res tap j x
# kernel row r), and the same along the columns.  So the layer is ONE 3x3 convolution on x with 4 x Cout output channels --
# a block of Cout per output phase (a, b) -- whose epilogue scatters block (a, b) of pixel (i, j) to out[2i+a][2j+b]
# (cnm_conv_desc.fold): the upsampled tensor is never written or read, and a 64-channel layer runs at a column tile of 128.
# At the border the interpolation clamps (x[-1] := x[0]) while the convolution pads u with zeros, and TMA hands the kernel
# zeros for x[-1]: the difference to the interior operator is, per dimension, a multiple of x[0] (first row / column:
# _FOLD_T[a]) or x[last] (_FOLD_B[a]) only.  Those corrections enter the same accumulator as extra K groups whose operand is
# a one-row / one-column / one-pixel VIEW of x (cnm_conv_src.x_off ...; everything outside the view reads as zero) and which
# run only on the tiles of that border class (bits 4-7 of cnm_conv_group.ntaps): rows need 3 column taps, columns 3 row
# taps, corners one -- 2-D operator = (A + dR) x (A + dC) = AxA + dRxA + AxdC + dRxdC.
FOLD_UPSAMPLE = os.environ.get("CNM_FOLD_UPSAMPLE", "1") != "0"
FOLD_MIN_BLOCKS = int(os.environ.get("CNM_FOLD_MIN_BLOCKS", "6"))   # low-res blocks (16 px) per side from which folding pays
FOLD_PAIRS = os.environ.get("CNM_FOLD_PAIRS", "1") != "0"   # CTA pairs (halo2f) where the block count is even, else single CTAs (halof)
_FOLD_A = (((.75, .25, 0.), (.25, .75, .75), (0., 0., .25)), ((.25, 0., 0.), (.75, .75, .25), (0., .25, .75)))
_FOLD_T = ((-.25, .25, 0.), (.25, 0., 0.))
_FOLD_B = ((0., 0., .25), (0., .25, -.25))
_FOLD_TOP, _FOLD_BOTTOM, _FOLD_LEFT, _FOLD_RIGHT = 1, 2, 4, 8


def fold_upconv_weights(wf):
    """[Cout, Cin, 3, 3] (BN folded) -> [4 Cout, 9 Cin, 3, 3] fp32: input-channel block s holds the weights of operand s
    (0 x, 1 top row, 2 bottom row, 3 left column, 4 right column, 5-8 corners TL TR BL BR), output block 2a+b is phase
    (a, b); kernel position [j, l] = low-res tap (j-1, l-1).  Computed in fp64."""
    w = wf.double()
    cout, cin = w.shape[:2]
    big = torch.zeros((4, cout, 9, cin, 3, 3), dtype=torch.float64, device=w.device)
    t64 = lambda v: torch.tensor(v, dtype=torch.float64, device=w.device)
    for a in (0, 1):
        for b in (0, 1):
            A_a, A_b, T_a, T_b, B_a, B_b = t64(_FOLD_A[a]), t64(_FOLD_A[b]), t64(_FOLD_T[a]), t64(_FOLD_T[b]), t64(_FOLD_B[a]), t64(_FOLD_B[b])
            o = big[2 * a + b]
            o[:, 0] = torch.einsum("jr,lt,ocrt->ocjl", A_a, A_b, w)
            o[:, 1, :, 1, :] = torch.einsum("r,lt,ocrt->ocl", T_a, A_b, w)
            o[:, 2, :, 1, :] = torch.einsum("r,lt,ocrt->ocl", B_a, A_b, w)
            o[:, 3, :, :, 1] = torch.einsum("jr,t,ocrt->ocj", A_a, T_b, w)
            o[:, 4, :, :, 1] = torch.einsum("jr,t,ocrt->ocj", A_a, B_b, w)
            for s_, (ra, cb) in enumerate(((T_a, T_b), (T_a, B_b), (B_a, T_b), (B_a, B_b))):
                o[:, 5 + s_, :, 1, 1] = torch.einsum("r,t,ocrt->oc", ra, cb, w)
    return big.reshape(4 * cout, 9 * cin, 3, 3).float()


class FoldedUpconvOp(ConvOp):
    """`upconv` of the reference (Upsample x2 bilinear + conv3x3 + BN + ReLU) as one convolution on the low-res act tensor
    `x` [N, h, w, Cin]; writes the [N, 2h, 2w, Cout] output.  See the section comment above; kernels conv_halo2f / conv_halof
    (csrc/conv_halo.cu, FOLD instantiations)."""

    def __init__(self, name, x, cout, N, h, w, dtype, device, out, relu=True, lane=0):
        if dtype == CNM_TF32 or cout % 64 or h < 16 or w < 16:
            raise ValueError("%s: folded upsampling needs a 2-byte dtype, Cout %% 64 == 0 and a low-res map of at least 16x16" % name)
        self.name, self.k, self.stride, self.dtype, self.lane = name, 3, 1, dtype, lane
        ch, split, pf = chunk_channels(dtype), split_passes(dtype), phys_factor(dtype)
        cin = x.shape[3] // pf
        assert cin % ch == 0 and tuple(x.shape[:3]) == (N, h, w) and tuple(out.shape) == (N, 2 * h, 2 * w, cout * pf)
        self.x, self.out, self.cin, self.cout_phase = x, out, cin, cout
        groups, cols = [], []   # group = (map, c0 physical, dx, ntaps, class flags)
        fl = {1: _FOLD_TOP, 2: _FOLD_BOTTOM, 3: _FOLD_LEFT, 4: _FOLD_RIGHT, 5: _FOLD_TOP | _FOLD_LEFT,
              6: _FOLD_TOP | _FOLD_RIGHT, 7: _FOLD_BOTTOM | _FOLD_LEFT, 8: _FOLD_BOTTOM | _FOLD_RIGHT}
        for c0 in range(0, cin, ch):
            for ap, wp_ in _chunk_passes(split):
                pc = _pc0(c0, ap, split)
                for t in range(3):                                    # x: a slab per column tap, three row taps
                    groups.append((0, pc, t - 1, 3, 0))
                    cols += [("act", c0, r, t, wp_, ap) for r in range(3)]
                for s_ in (1, 2):                                     # first / last row: three column taps, row tap 0
                    for t in range(3):
                        groups.append((s_, pc, t - 1, 1, fl[s_]))
                        cols.append(("act", s_ * cin + c0, 1, t, wp_, ap))
                for s_ in (3, 4):                                     # first / last column: column tap 0, three row taps
                    groups.append((s_, pc, 0, 3, fl[s_]))
                    cols += [("act", s_ * cin + c0, r, 1, wp_, ap) for r in range(3)]
                for s_ in (5, 6, 7, 8):                               # corners
                    groups.append((s_, pc, 0, 1, fl[s_]))
                    cols.append(("act", s_ * cin + c0, 1, 1, wp_, ap))
        starts, pos = [], 0
        for g in groups:
            starts.append(pos)
            pos += g[3]
        gorder = _corrections_first([cols[st] for st in starts])      # fp8 / lo rows first, group-wise (see make_groups)
        gcols = [c for i in gorder for c in cols[starts[i]:starts[i] + groups[i][3]]]
        groups = [groups[i] for i in gorder]
        first, pos = [], 0
        for g in groups:
            first.append(pos)
            pos += g[3]
        gt = torch.tensor([[g[0], g[1], g[2], g[3] | (g[4] << 4) | (f << 8)] for g, f in zip(groups, first)], dtype=torch.int32)
        info = [[0, 0] for _ in range(12)]
        for m in range(9):
            info[m] = [1, 18] if m in (0, 3, 4) else [0, 16]
        flat = [v for pr in info for v in pr]
        pairs_ok = (N * ((h + 15) // 16) * ((w + 15) // 16)) % 2 == 0 and HALO2_ENABLED and FOLD_PAIRS
        self.backend = "halo2f" if pairs_ok else "halof"
        self.variants = {self.backend: {"cols": gcols, "groups_dev": gt.to(device), "ngroups": len(groups),
                                        "pads": (C.c_int32 * len(flat))(*flat)}}
        self.num_steps = len(gcols)
        self.wp = torch.zeros((4 * cout, self.num_steps * ch), dtype=TORCH_DTYPE[dtype], device=device)
        self.bias = torch.zeros((4 * cout,), dtype=torch.float32, device=device)
        self.desc = _lib.ConvDesc(N, h, w, 4 * cout, out.shape[3], 0, 1, 1 if relu else 0, 9, self.num_steps, dtype, 128)
        self.desc.fold = 1
        es = x.element_size()
        pix = x.shape[3] * es
        arr = (_lib.ConvSrc * 9)()
        base = x.data_ptr()
        views = {1: (1, w, 0, 0), 2: (1, w, 0, h - 1), 3: (h, 1, 0, 0), 4: (h, 1, w - 1, 0), 5: (1, 1, 0, 0), 6: (1, 1, w - 1, 0),
                 7: (1, 1, 0, h - 1), 8: (1, 1, w - 1, h - 1)}   # source -> (H, W, x_off, y_off)
        arr[0] = _lib.ConvSrc(base, h, w, x.shape[3], 0)
        for s_, (vh, vw, xo, yo) in views.items():
            arr[s_] = _lib.ConvSrc(base + (yo * w + xo) * pix, vh, vw, x.shape[3], 0, xo, yo, w, h * w)
        self.srcs_arr, self.srcs = arr, []
        # issued MACs x 2 of the interior operator (the border groups add ~3 / 9 on edge tiles, 7 / 9 on corner tiles)
        main_steps = 9 * (cin // ch) * len(_chunk_passes(split))
        self.flops = 2.0 * N * h * w * 4 * cout * main_steps * ch * (1.5 if dtype == CNM_F16F8 else 1.0)
        self.flops_bf16_time = 2.0 * N * h * w * 4 * cout * main_steps * ch
        self._weights, self.sweep, self.tunable = None, None, False
        self.tuned = (self.backend, 128, 1, None)
        self.key = repr(("fold", dtype, N, h, w, cout, cin))

    @property
    def cols(self):
        return self.variants[self.backend]["cols"]

    def select(self, backend, bn, ksplit=1):
        if backend != self.backend:
            raise ValueError("%s: a folded upsampling layer runs %s only" % (self.name, self.backend))

    def tune(self, stream, reps=3):
        return

    def load(self, wf, bf):
        self._weights = (wf, bf)
        big = fold_upconv_weights(wf.to(self.wp.device))
        if self.dtype == CNM_F16F8:
            self.desc.acc_exp = f16f8_weight_exp(big)
        self.wp.copy_(pack_weights(big, self.cols, chunk_channels(self.dtype), self.dtype))
        self.bias.copy_(bf.to(self.bias.device).repeat(4))

    def __call__(self, stream):
        lib = _lib.load()
        v = self.variants[self.backend]
        fn = lib.cnm_conv2d_halo2 if self.backend == "halo2f" else lib.cnm_conv2d_halo
        _lib.check(fn(C.byref(self.desc), self.srcs_arr, v["pads"], v["groups_dev"].data_ptr(), v["ngroups"], self.wp.data_ptr(),
                      self.bias.data_ptr(), self.out.data_ptr(), stream), self.name)


class _ShapeOnly:
    """Stands in for an act tensor that is never materialised (the cost volume under the fused plane sweep): ConvOp
    reads its shape and passes a valid, aligned device pointer that no kernel dereferences."""

    def __init__(self, shape, anchor):
        self.shape, self._anchor, self.is_cuda = tuple(shape), anchor, True

    def is_contiguous(self):
        return True

    def data_ptr(self):
        return self._anchor.data_ptr()


class _Plan:
    def __init__(self, dtype, device, backend, lane=0):
        self.dtype, self.device, self.backend, self.lane = parse_dtype(dtype), torch.device(device), backend, lane
        if self.device.type != "cuda":
            raise RuntimeError("cnmnet_b200 runs on CUDA devices only (no CPU path); got %s" % (self.device,))
        self.lib = _lib.load()
        self.ops = []      # (callable(stream))
        self.convs = {}    # name -> ConvOp
        self.heads = {}    # name -> dict(w, bias, scale)
        self.ch = chunk_channels(self.dtype)
        self.pf = phys_factor(self.dtype)
        self._storage = {}   # data_ptr of a live buffer -> its uint8 backing tensor
        self._free = []      # backing tensors whose last reader has been enqueued: reusable by later ops
        self.bytes_allocated = 0

    def buf(self, *shape, dtype=None, fresh=False):
        """Plain buffer (dtype given) or act tensor [N,H,W,C logical] (C * pf physical channels).

        Buffers come from a small arena: ops run in creation order on one stream, so a buffer handed back with
        release() after its last reader was created can back any later tensor (PLAN_REUSE).  A plan then holds the peak
        live set instead of the sum of all activations (~1.7x less)."""
        if dtype is None:
            shape = tuple(shape[:-1]) + (shape[-1] * self.pf,)
        dt = TORCH_DTYPE[self.dtype] if dtype is None else dtype
        n = 1
        for v in shape:
            n *= int(v)
        need = (n * torch.empty((), dtype=dt).element_size() + 511) // 512 * 512
        best = None
        if PLAN_REUSE and not fresh:
            for i, t in enumerate(self._free):
                if t.numel() >= need and (best is None or t.numel() < self._free[best].numel()):
                    best = i
        if best is not None:
            store = self._free.pop(best)
        else:
            store = torch.empty((need,), dtype=torch.uint8, device=self.device)
            if POISON:
                store.fill_(0xFF)
            self.bytes_allocated += need
        out = store[:n * torch.empty((), dtype=dt).element_size()].view(dt).view(shape)
        self._storage[out.data_ptr()] = store
        return out

    def release(self, *tensors):
        """The last op reading these buffers has been created: their memory may back later tensors."""
        for t in tensors:
            store = self._storage.pop(t.data_ptr(), None)
            if store is not None:
                self._free.append(store)

    def src(self, buf, wch):
        return act_src(buf, wch, phys=self.pf)

    def conv(self, name, srcs, k, stride, cout, N, Ho, Wo, relu=True):
        op = ConvOp(name, srcs, k, stride, cout, N, Ho, Wo, self.dtype, self.device, relu, out=self.buf(N, Ho, Wo, cout),
                    backend=self.backend, lane=self.lane)
        self.convs[name] = op
        self.ops.append(op)
        return op.out

    def upsample(self, x):
        N, h, w, Cc = x.shape
        Cc //= self.pf
        out = self.buf(N, 2 * h, 2 * w, Cc)
        lib, dt = self.lib, self.dtype
        self.ops.append(lambda s: _lib.check(lib.cnm_upsample2x(x.data_ptr(), out.data_ptr(), N, h, w, Cc, dt, s),
                                             "upsample2x"))
        return out

    def upconv(self, name, x, cout, N, H, W):
        """The reference's `upconv` block (Upsample x2 bilinear + conv3x3 + BN + ReLU, depthNet_model.py:92-98) on the
        low-res act tensor x -> [N, H, W, cout]; x is released.  Maps of at least FOLD_MIN_BLOCKS 16-pixel blocks per side
        run it as ONE convolution on x (FoldedUpconvOp: the upsampled tensor is never written); smaller ones -- where the
        border-correction groups would cost more than the upsampling pass -- keep the two kernels."""
        h, w = H // 2, W // 2
        if FOLD_UPSAMPLE and self.dtype != CNM_TF32 and self.backend == "tc" and HALO_MODE != "off" and \
                min(h, w) // 16 >= FOLD_MIN_BLOCKS and cout % 64 == 0:
            out = self.buf(N, H, W, cout)
            op = FoldedUpconvOp(name, x, cout, N, h, w, self.dtype, self.device, out, lane=self.lane)
            self.convs[name] = op
            self.ops.append(op)
            self.release(x)
            return out
        u = self.upsample(x)
        self.release(x)
        y = self.conv(name, [self.src(u, 0)], 3, 1, cout, N, H, W)
        self.release(u)
        return y

    def head(self, name, x, scale, want_logits=False):
        N, H, W, Cc = x.shape
        Cc //= self.pf
        out = self.buf(N, 1, H, W, dtype=torch.float32)
        logits = self.buf(N, 1, H, W, dtype=torch.float32) if want_logits else None
        hw = {"w": torch.zeros((9 * Cc,), dtype=torch.float32, device=self.device), "bias": 0.0, "scale": float(scale),
              "out": out}
        self.heads[name] = hw
        lib, dt = self.lib, self.dtype
        lp = logits.data_ptr() if logits is not None else None
        producer = [op for op in self.convs.values() if op.out is x]
        if FUSE_HEADS and dt != CNM_TF32 and self.backend != "simt" and producer:
            # the head rides in the producing convolution's epilogue: every 64-channel chunk of the fp32 accumulators
            # leaves nine per-tap dot products per pixel in its own slice; what is left is a 3x3 gather of scalars
            # (summing the Cc/64 slices in order) + bias + sigmoid
            op = producer[0]
            S = Cc // 64
            # fresh: the producing conv (created earlier) writes it while reading inputs that may already be released
            partial = self.buf(S * N, H, W, 12, dtype=torch.float32, fresh=True)
            hw["w_tap_major"] = torch.zeros((9, Cc), dtype=torch.float32, device=self.device)
            hw["partial"] = partial
            op.fuse_head(hw["w_tap_major"], partial)
            self.ops.append(lambda s: _lib.check(
                lib.cnm_head_finish(partial.data_ptr(), S, hw["bias"], hw["scale"], hw["out"].data_ptr(), lp, N, H, W, s), name))
            self.release(partial)
            return out, logits
        if FUSE_HEADS:
            # two passes: per-pixel tap dot products (each input element read and decoded once), then the 3x3 gather
            partial = self.buf(N, H, W, 12, dtype=torch.float32)
            hw["w_tap_major"] = torch.zeros((9, Cc), dtype=torch.float32, device=self.device)
            hw["partial"] = partial

            def two_pass(s):
                _lib.check(lib.cnm_head_partial(x.data_ptr(), hw["w_tap_major"].data_ptr(), partial.data_ptr(), N, H, W, Cc,
                                                dt, s), name)
                _lib.check(lib.cnm_head_finish(partial.data_ptr(), 1, hw["bias"], hw["scale"], hw["out"].data_ptr(), lp, N, H, W, s),
                           name)

            self.ops.append(two_pass)
            self.release(partial)
            return out, logits
        self.ops.append(lambda s: _lib.check(
            lib.cnm_head3x3(x.data_ptr(), hw["w"].data_ptr(), hw["bias"], hw["scale"], hw["out"].data_ptr(), lp, N, H, W, Cc,
                            dt, s), name))
        return out, logits

    def bind_head_output(self, name, out):
        """Redirects a head's output [N,1,H,W] fp32 to a caller-owned buffer (CNMPipeline's packed output block).  The
        owner of the plan must do this before capturing a graph."""
        hw = self.heads[name]
        if tuple(out.shape) != tuple(hw["out"].shape) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("head output must be a contiguous fp32 tensor of shape %s" % (tuple(hw["out"].shape),))
        hw["out"] = out

    def iconv(self, name, srcs, disp, wch, cout, N, Ho, Wo):
        """iconv3 / iconv2 / iconv1 of DepthNet (depthNet_model.py:247-260): conv3x3 over cat(srcs, nearest-upsampled
        disparity `disp`); the disparity is input channel `wch`.  Tensor-core plans in the 2-byte modes add its nine taps in
        the epilogue, in fp32 (ConvOp.attach_side_disp); the others list them as a 64-channel side tensor."""
        cin = sum(s.C for s in srcs)
        if SIDE_DISP and self.dtype != CNM_TF32 and self.backend != "simt" and cin >= SIDE_DISP_MIN_CIN:
            y = self.conv(name, srcs, 3, 1, cout, N, Ho, Wo)
            self.convs[name].attach_side_disp(disp, wch)
            return y
        t = self.disp_taps(disp)
        y = self.conv(name, srcs + [Src(t, "taps3x3", wch, self.ch, 0, 1)], 3, 1, cout, N, Ho, Wo)
        self.release(t)
        return y

    def disp_taps(self, disp):
        N, _, h, w = disp.shape
        out = self.buf(N, 2 * h, 2 * w, self.ch)
        lib, dt = self.lib, self.dtype
        self.ops.append(lambda s: _lib.check(lib.cnm_pack_disp_taps(disp.data_ptr(), out.data_ptr(), N, h, w, dt, s),
                                             "pack_disp_taps"))
        return out

    def load_head(self, name, sd):
        hw = self.heads[name]
        w = sd[name + ".0.weight"].to(self.device, torch.float32)  # [1, C, 3, 3] -> (r, s, c)
        hw["w"].copy_(w[0].permute(1, 2, 0).reshape(-1))
        if "w_tap_major" in hw:
            hw["w_tap_major"].copy_(w[0].permute(1, 2, 0).reshape(9, -1))
        hw["bias"] = float(sd[name + ".0.bias"].reshape(-1)[0])

    def tune(self):
        """Kernel-variant choice for every convolution of the plan (see TUNE_MODE); weights are still zero."""
        s = _lib.stream_ptr()
        for op in self.convs.values():
            op.tune(s)
        torch.cuda.synchronize(self.device)

    def run(self, stream=None):
        s = _lib.stream_ptr() if stream is None else stream
        for op in self.ops:
            op(s)

    @property
    def conv_flops(self):
        return sum(op.flops for op in self.convs.values())


class DepthNetPlan(_Plan):
    """camera setup -> cost volume -> DepthNet trunk for N (view, batch) pairs sharing Bref reference frames.

    Inputs are bound at run time: ref [Bref,3,H,W], src [N,3,H,W] (view-major: pair p = v*Bref + b),
    ref_cam [Bref,2,4,4], src_cam [N,2,4,4].  Outputs live in plan buffers: disp1..4 fp32 [N,1,h,w],
    iconv1 act [N,H,W,64]."""

    def __init__(self, N, Bref, H, W, idepth_scale=3.0, dtype="bf16", device="cuda", backend="tc", planes=64,
                 want_logits=False, lane=0):
        super().__init__(dtype, device, backend, lane)
        if H % 32 or W % 32:
            raise ValueError("DepthNet needs H and W divisible by 32 (five stride-2 stages); got %dx%d" % (H, W))
        if N % Bref:
            raise ValueError("N=%d pairs must be a multiple of Bref=%d reference frames" % (N, Bref))
        if float(idepth_scale) == 2.0:
            self.idmin, self.idmax = 0.02, 2.0
        elif float(idepth_scale) == 3.0:
            self.idmin, self.idmax = 0.1, 3.0
        else:  # the reference raises UnboundLocalError here (depthNet_model.py:186-191)
            raise ValueError("idepth_scale must be 2.0 or 3.0 (the reference defines no other plane range)")
        if planes < 64 or planes % 64:
            raise ValueError("planes must be a positive multiple of 64 (the reference network: 64)")
        self.N, self.Bref, self.H, self.W, self.D, self.scale = N, Bref, H, W, planes, float(idepth_scale)
        lib, dt, ch = self.lib, self.dtype, self.ch
        self.hk = self.buf(N, 12, dtype=torch.float32)
        # K1: the TMA-staged plane sweep reads the source views in place; the first-generation pair (repack + L1 gather,
        # CNM_PLANE_SWEEP=0) stays as its twin
        self.use_sweep = PLANE_SWEEP and W % 4 == 0 and planes <= 128
        self.may_fuse = FUSE_SWEEP != "0" and self.dtype == CNM_F16F8 and planes <= 128
        self.tex = None if (self.use_sweep and not self.may_fuse) else \
            torch.empty((lib.cnm_packed_source_bytes(N, H, W),), dtype=torch.uint8, device=self.device)
        # with the fused plane sweep forced, the volume never exists: conv1.0 only needs its shape
        self.cost = _ShapeOnly((N, H, W, planes * self.pf), self.hk) if (self.may_fuse and FUSE_SWEEP == "force") \
            else self.buf(N, H, W, planes)
        # the 3 image channels of conv1.0 as a side tensor: 7 column taps x 3 kernel rows per K-step where a chunk holds them
        # (63 of 64 channels: 3 K-steps instead of 7), else 7 column taps per kernel row
        self.dense_taps = DENSE_ROWTAPS and 3 * 7 * 3 <= ch
        self.img_taps = self.buf(N, H + 2, W, ch) if self.dense_taps else self.buf(N, H, W, ch)
        self.inputs = {}

        def front(s):
            i = self.inputs
            _lib.check(lib.cnm_camera_setup(i["ref_cam"], i["src_cam"], self.hk.data_ptr(), Bref, N // Bref, s), "camera")
            if self.convs["conv1.0"].backend == "halo2s":
                # (f)-1: conv1.0 computes the volume itself; it only needs the source views as zero-bordered texels
                _lib.check(lib.cnm_pack_source(i["src"], self.tex.data_ptr(), N, H, W, s), "pack_source")
            elif self.use_sweep:
                _lib.check(lib.cnm_plane_sweep(i["ref"], i["src"], self.hk.data_ptr(), self.cost.data_ptr(), Bref,
                                               N // Bref, H, W, planes, self.idmin, self.idmax, _lib.VOLUME_ACT, dt, s),
                           "plane_sweep")
            else:
                _lib.check(lib.cnm_pack_source(i["src"], self.tex.data_ptr(), N, H, W, s), "pack_source")
                _lib.check(lib.cnm_cost_volume(i["ref"], self.tex.data_ptr(), self.hk.data_ptr(), self.cost.data_ptr(),
                                               Bref, N // Bref, H, W, planes, self.idmin, self.idmax, _lib.VOLUME_ACT,
                                               dt, s), "cost_volume")
            _lib.check((lib.cnm_pack_row_taps3 if self.dense_taps else lib.cnm_pack_row_taps)(
                i["ref"], self.img_taps.data_ptr(), N, Bref, 3, H, W, 7, dt, s), "pack_row_taps")

        self.ops.append(front)
        self._trunk(want_logits)

    def _trunk(self, want_logits):
        N, H, W = self.N, self.H, self.W
        cv, rel = self.conv, self.release
        x = cv("conv1.0", [self.src(self.cost, 3), Src(self.img_taps, "rowtaps3" if self.dense_taps else "rowtaps", 0, self.ch, 0, 3)],
               7, 1, 128, N, H, W)
        if self.may_fuse:
            self.convs["conv1.0"].attach_sweep(0, self.Bref, self.D, self.idmin, self.idmax, self.tex, self.hk)
        rel(self.cost, self.img_taps)
        conv1 = cv("conv1.3", [self.src(x, 0)], 7, 2, 128, N, H // 2, W // 2)
        rel(x)
        x = cv("conv2.0", [self.src(conv1, 0)], 5, 1, 256, N, H // 2, W // 2)
        conv2 = cv("conv2.3", [self.src(x, 0)], 5, 2, 256, N, H // 4, W // 4)
        rel(x)
        x = cv("conv3.0", [self.src(conv2, 0)], 3, 1, 512, N, H // 4, W // 4)
        conv3 = cv("conv3.3", [self.src(x, 0)], 3, 2, 512, N, H // 8, W // 8)
        rel(x)
        x = cv("conv4.0", [self.src(conv3, 0)], 3, 1, 512, N, H // 8, W // 8)
        conv4 = cv("conv4.3", [self.src(x, 0)], 3, 2, 512, N, H // 16, W // 16)
        rel(x)
        x = cv("conv5.0", [self.src(conv4, 0)], 3, 1, 512, N, H // 16, W // 16)
        conv5 = cv("conv5.3", [self.src(x, 0)], 3, 2, 512, N, H // 32, W // 32)
        rel(x)

        up5 = self.upconv("upconv5.1", conv5, 512, N, H // 16, W // 16)
        iconv5 = cv("iconv5.0", [self.src(up5, 0), self.src(conv4, 512)], 3, 1, 512, N, H // 16, W // 16)
        rel(up5, conv4)
        up4 = self.upconv("upconv4.1", iconv5, 512, N, H // 8, W // 8)
        iconv4 = cv("iconv4.0", [self.src(up4, 0), self.src(conv3, 512)], 3, 1, 512, N, H // 8, W // 8)
        rel(up4, conv3)
        self.disp4, l4 = self.head("disp4", iconv4, self.scale, want_logits)
        up3 = self.upconv("upconv3.1", iconv4, 256, N, H // 4, W // 4)
        iconv3 = self.iconv("iconv3.0", [self.src(up3, 0), self.src(conv2, 256)], self.disp4, 512, 256, N, H // 4, W // 4)
        rel(up3, conv2)
        self.disp3, l3 = self.head("disp3", iconv3, self.scale, want_logits)
        up2 = self.upconv("upconv2.1", iconv3, 128, N, H // 2, W // 2)
        iconv2 = self.iconv("iconv2.0", [self.src(up2, 0), self.src(conv1, 128)], self.disp3, 256, 128, N, H // 2, W // 2)
        rel(up2, conv1)
        self.disp2, l2 = self.head("disp2", iconv2, self.scale, want_logits)
        up1 = self.upconv("upconv1.1", iconv2, 64, N, H, W)
        self.iconv1 = self.iconv("iconv1.0", [self.src(up1, 0)], self.disp2, 64, 64, N, H, W)
        rel(up1)
        self.disp1, l1 = self.head("disp1", self.iconv1, self.scale, want_logits)
        self.logits = [l1, l2, l3, l4]
        self.tune()

    def load_state(self, sd):
        for name, op in self.convs.items():
            prefix, ci = name.rsplit(".", 1)
            op.load(*fold_bn(sd, prefix, int(ci)))
        for name in self.heads:
            self.load_head(name, sd)

    def bind(self, ref, src, ref_cam, src_cam):
        for t, shape in ((ref, (self.Bref, 3, self.H, self.W)), (src, (self.N, 3, self.H, self.W)),
                         (ref_cam, (self.Bref, 2, 4, 4)), (src_cam, (self.N, 2, 4, 4))):
            if tuple(t.shape) != shape or t.dtype != torch.float32:
                raise ValueError("expected fp32 tensor of shape %s, got %s %s" % (shape, tuple(t.shape), t.dtype))
        self.inputs = {"ref": _lib.ptr(ref), "src": _lib.ptr(src), "ref_cam": _lib.ptr(ref_cam),
                       "src_cam": _lib.ptr(src_cam)}
        sweep = self.convs["conv1.0"].sweep
        if sweep is not None:
            sweep.ref = self.inputs["ref"]
        self._keep = (ref, src, ref_cam, src_cam)

    def iconv1_nchw(self, stream=None):
        N, H, W, Cc = self.iconv1.shape
        Cc //= self.pf
        out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=self.device)
        s = _lib.stream_ptr() if stream is None else stream
        _lib.check(self.lib.cnm_act_to_nchw(self.iconv1.data_ptr(), out.data_ptr(), N, Cc, H, W, Cc, 0, self.dtype, s),
                   "act_to_nchw")
        return out


class RefinePlan(_Plan):
    """refine-input aggregation -> DepthRefineNet for B frames.

    Inputs: idepth [V,B,H,W] fp32 and iconv1 either as the DepthNet plan's act tensor [V*B,H,W,64] or, at
    the public boundary, NCHW fp32 [2,B,64,H,W].  V = 2*group source views (eval.py:659-663, 920-929)."""

    def __init__(self, B, H, W, idepth_scale=2.0, dtype="bf16", device="cuda", backend="tc", group=1,
                 nchw_input=False, want_logits=False, lane=0):
        super().__init__(dtype, device, backend, lane)
        if H % 8 or W % 8:
            raise ValueError("DepthRefineNet needs H and W divisible by 8; got %dx%d" % (H, W))
        self.B, self.H, self.W, self.group, self.scale = B, H, W, group, float(idepth_scale)
        lib, dt, ch = self.lib, self.dtype, self.ch
        self.sum_act = self.buf(B, H, W, 64)
        self.taps = self.buf(B, H, W, ch)
        self.id_mean = self.buf(2, B, H, W, dtype=torch.float32)
        self.inputs = {}

        def front(s):
            i = self.inputs
            if nchw_input:
                _lib.check(lib.cnm_refine_aggregate_nchw(i["idepth"], i["iconv"], self.sum_act.data_ptr(),
                                                         self.taps.data_ptr(), B, H, W, dt, s), "refine_aggregate")
            else:
                _lib.check(lib.cnm_refine_aggregate(i["idepth"], i["iconv"], self.sum_act.data_ptr(),
                                                    self.taps.data_ptr(), self.id_mean.data_ptr(), B, H, W, group, dt,
                                                    s), "refine_aggregate")

        self.ops.append(front)
        cv, rel = self.conv, self.release
        x = cv("conv1.0", [self.src(self.sum_act, 3), Src(self.taps, "taps3x3", 0, ch, 0, 3)], 3, 1, 128, B, H, W)
        rel(self.sum_act, self.taps)
        conv1 = cv("conv1.3", [self.src(x, 0)], 3, 2, 128, B, H // 2, W // 2)
        rel(x)
        x = cv("conv2.0", [self.src(conv1, 0)], 3, 1, 256, B, H // 2, W // 2)
        conv2 = cv("conv2.3", [self.src(x, 0)], 3, 2, 256, B, H // 4, W // 4)
        rel(x)
        x = cv("conv3.0", [self.src(conv2, 0)], 3, 1, 512, B, H // 4, W // 4)
        conv3 = cv("conv3.3", [self.src(x, 0)], 3, 2, 512, B, H // 8, W // 8)
        rel(x)
        up_conv3 = self.upsample(conv3)  # shared by both decoders
        rel(conv3)
        self.out, self.logits, self.iconv1 = {}, {}, {}
        for br, head, sc in (("depth", "disp_refine", self.scale), ("prob", "prob", 1.0)):
            u3 = cv("upconv3_%s.1" % br, [self.src(up_conv3, 0)], 3, 1, 256, B, H // 4, W // 4)
            i3 = cv("iconv3_%s.0" % br, [self.src(u3, 0), self.src(conv2, 256)], 3, 1, 256, B, H // 4, W // 4)
            rel(u3)
            u2 = self.upconv("upconv2_%s.1" % br, i3, 128, B, H // 2, W // 2)
            i2 = cv("iconv2_%s.0" % br, [self.src(u2, 0), self.src(conv1, 128)], 3, 1, 128, B, H // 2, W // 2)
            rel(u2)
            u1 = self.upconv("upconv1_%s.1" % br, i2, 64, B, H, W)
            i1 = cv("iconv1_%s.0" % br, [self.src(u1, 0)], 3, 1, 64, B, H, W)
            rel(u1)
            self.iconv1[br] = i1          # kept: ReturnVolume hands iconv1_depth out
            self.out[br], self.logits[br] = self.head(head, i1, sc, want_logits)
        rel(up_conv3, conv2, conv1)
        self.tune()

    load_state = DepthNetPlan.load_state

    def bind(self, idepth, iconv):
        self.inputs = {"idepth": _lib.ptr(idepth), "iconv": _lib.ptr(iconv)}
        self._keep = (idepth, iconv)

    def iconv1_depth_nchw(self, stream=None):
        x = self.iconv1["depth"]
        N, H, W, Cc = x.shape
        Cc //= self.pf
        out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=self.device)
        s = _lib.stream_ptr() if stream is None else stream
        _lib.check(self.lib.cnm_act_to_nchw(x.data_ptr(), out.data_ptr(), N, Cc, H, W, Cc, 0, self.dtype, s),
                   "act_to_nchw")
        return out
