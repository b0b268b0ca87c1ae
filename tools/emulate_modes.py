"""CPU emulation of the conv precision modes (operand rounding only, fp32 accumulation) on the oracle's torch port:
predicts the final-depth error of a mode before a kernel exists for it.  Development tool, not shipped.

    python tools/emulate_modes.py [modes] [HxW] [weight seeds]
    modes: fp32, bf16, tf32, bf16x3, bf16f8  (comma separated)

bf16f8 = one bf16 pass hi_a*hi_w plus ONE fp8 pass over the K-concatenated corrections [lo_a | hi_a] x [hi_w | lo_w]
(e4m3 operands, power-of-two scales, separate accumulator)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from cnmnet_b200 import synth  # noqa: E402
from oracle import nets as onets  # noqa: E402

MODE = "fp32"
F8 = torch.float8_e4m3fn
LO_SCALE, HI_SCALE = 2.0 ** 8, 2.0 ** -2


def bf16(x):
    if MODE == "fp16f8":
        return x.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32)
    return x.to(torch.bfloat16).to(torch.float32)


def lo_scale():
    return 2.0 ** 7 if MODE == "fp16f8" else LO_SCALE


def tf32(x):
    b = x.contiguous().view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


def e4m3(x):
    return x.clamp(-448.0, 448.0).to(F8).to(torch.float32)


def pow2_scale(w, top=256.0):
    m = float(w.abs().max())
    return 2.0 ** np.floor(np.log2(top / m)) if m > 0 else 1.0


def stored(a):
    """what a non-conv consumer reads back from the act tensor of the mode"""
    if MODE == "bf16":
        return bf16(a)
    if MODE == "tf32":
        return tf32(a)
    if MODE == "bf16x3":
        hi = bf16(a)
        return hi + bf16(a - hi)
    if MODE in ("bf16f8", "fp16f8"):
        hi = bf16(a)
        return hi + e4m3((a - hi) * lo_scale()) / lo_scale()
    return a


def qconv(a, w, stride, pad):
    if MODE == "fp32":
        return F.conv2d(a, w, None, stride=stride, padding=pad)
    if MODE == "bf16":
        return F.conv2d(bf16(a), bf16(w), None, stride=stride, padding=pad)
    if MODE == "tf32":
        return F.conv2d(tf32(a), tf32(w), None, stride=stride, padding=pad)
    ah, wh = bf16(a), bf16(w)
    al, wl = a - ah, w - wh
    main = F.conv2d(ah, wh, None, stride=stride, padding=pad)
    if MODE == "bf16x3":
        return main + F.conv2d(bf16(al), wh, None, stride=stride, padding=pad) + \
            F.conv2d(ah, bf16(wl), None, stride=stride, padding=pad)
    sw = pow2_scale(w, 448.0)
    # both products must carry the same total scale S: LO_SCALE*sw == HI_SCALE*swl
    ls = lo_scale()
    hs = 2.0 ** -4 if MODE == "fp16f8" else HI_SCALE
    swl = ls * sw / hs
    S = ls * sw
    corr = F.conv2d(e4m3(al * ls), e4m3(w * sw), None, stride=stride, padding=pad) + \
        F.conv2d(e4m3(ah * hs), e4m3(wl * swl), None, stride=stride, padding=pad)
    return main + corr / S


def cbr(sd, prefix, ci, x, k, stride):
    w = sd["%s.%d.weight" % (prefix, ci)].float()
    p = "%s.%d." % (prefix, ci + 1)
    scale = sd[p + "weight"].float() / torch.sqrt(sd[p + "running_var"].float() + onets.BN_EPS)
    bias = sd[p + "bias"].float() - sd[p + "running_mean"].float() * scale
    y = qconv(stored(x), w * scale.view(-1, 1, 1, 1), stride, (k - 1) // 2) + bias.view(1, -1, 1, 1)
    return F.relu(y)


def head_logit(sd, name, x):
    return F.conv2d(stored(x), sd[name + ".0.weight"].float(), sd[name + ".0.bias"].float(), padding=1)


def up(sd, name, x):
    x = F.interpolate(stored(x), scale_factor=2, mode="bilinear", align_corners=False)
    return cbr(sd, name, 1, x, 3, 1)


def main():
    global MODE
    modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16", "bf16x3", "bf16f8"]
    H, W = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "64x96").split("x")]
    seeds = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "21,51").split(",")]
    orig = (onets._cbr, onets._head_logit, onets._up)
    for ws in seeds:
        ref, srcs, ref_cam, src_cams = synth.frame_tuples(500 + H, 2, H, W, 2)
        dsd = {k: torch.from_numpy(np.asarray(v)) for k, v in
               synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), ws, 0.02).items()}
        rsd = {k: torch.from_numpy(np.asarray(v)) for k, v in
               synth.kaiming_state(onets.state_shapes(onets.REFINENET_LAYERS), ws + 1, 0.02).items()}
        onets._cbr, onets._head_logit, onets._up = orig
        want = onets.reference_pipeline(dsd, rsd, ref, srcs, ref_cam, src_cams, 3.0, 9)[0].numpy()
        onets._cbr, onets._head_logit, onets._up = cbr, head_logit, up
        for MODE in modes:
            d = onets.reference_pipeline(dsd, rsd, ref, srcs, ref_cam, src_cams, 3.0, 9)[0].numpy()
            rel = np.abs(d - want) / np.abs(want)
            print("%dx%d wseed=%d %-7s depth rel err mean %.2e p99 %.2e max %.2e" %
                  (H, W, ws, MODE, rel.mean(), np.percentile(rel, 99), rel.max()), flush=True)
    onets._cbr, onets._head_logit, onets._up = orig


if __name__ == "__main__":
    main()
