"""(f)-1: conv1.0 with the plane sweep fused in (cnm_conv2d_halo2_sweep) against the unfused path it replaces
(getVolume + torch.cat + conv1[0]: depthnet/depthNet_model.py:185-224, 231-235).

The fused kernel's volume producers run the arithmetic of cnm_cost_volume (csrc/sweep_l1.cuh is the one body both
compile) and its MMAs run the K order of cnm_conv2d_halo2, so the bar is BIT equality of conv1.0's output with
cnm_pack_source + cnm_cost_volume + cnm_conv2d_halo2; the distance of both to the oracle is covered by the network
tests, which run with whatever the variant table picks."""
import numpy as np
import pytest
import torch

from cnmnet_b200 import _lib, engine, synth
from gpu_util import dev
from oracle import nets as onets

pytestmark = pytest.mark.gpu


def _plan(N, Bref, H, W, planes, fused, monkeypatch):
    monkeypatch.setattr(engine, "TUNE_MODE", "off")
    monkeypatch.setattr(engine, "PLANE_SWEEP", False)   # the unfused twin computes the volume with cnm_cost_volume
    monkeypatch.setattr(engine, "FUSE_SWEEP", "1")
    plan = engine.DepthNetPlan(N, Bref, H, W, 3.0, "f16f8", dev(), planes=planes)
    op = plan.convs["conv1.0"]
    assert "halo2s" in op.variants
    op.select("halo2s" if fused else "halo2", 128)
    return plan


def _run_front_and_conv1(plan, ref, src, ref_cam, src_cam, sd):
    plan.load_state(sd)
    plan.bind(ref, src, ref_cam, src_cam)
    s = _lib.stream_ptr()
    plan.ops[0](s)                      # camera setup, volume (or source repack only), image taps
    assert plan.ops[1] is plan.convs["conv1.0"]
    plan.ops[1](s)
    torch.cuda.synchronize()
    return plan.convs["conv1.0"].out.clone()


# (B, V, H, W, planes): one tile per CTA pair; several tiles per CTA (the scratch slots are re-used, it >= 2); tiles
# cut by the image border on all four sides are in every case (the region reaches 3 pixels outside the image)
CASES = [(2, 2, 64, 96, 64), (2, 2, 256, 192, 64), (1, 2, 96, 160, 128), (5, 2, 256, 192, 64)]


@pytest.mark.parametrize("B,V,H,W,planes", CASES)
def test_fused_conv1_is_bit_identical_to_the_unfused_path(monkeypatch, B, V, H, W, planes):
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(7 + H, B, H, W, V)
    N = B * V
    src = srcs.transpose(0, 1).reshape(N, 3, H, W).contiguous().to(dev())        # view-major pairs p = v*B + b
    src_cam = src_cams.transpose(0, 1).reshape(N, 2, 4, 4).contiguous().to(dev())
    ref, ref_cam = ref.to(dev()), ref_cam.to(dev())
    shapes = onets.state_shapes(onets.DEPTHNET_LAYERS)
    if planes != 64:
        shapes = dict(shapes)
        shapes["conv1.0.weight"] = (128, 3 + planes, 7, 7)
    sd = synth.kaiming_state(shapes, 5, 0.02)
    outs = []
    for fused in (False, True):
        plan = _plan(N, B, H, W, planes, fused, monkeypatch)
        assert plan.convs["conv1.0"].backend == ("halo2s" if fused else "halo2")
        outs.append(_run_front_and_conv1(plan, ref, src, ref_cam, src_cam, sd))
    a, b = outs[0].view(torch.uint8), outs[1].view(torch.uint8)
    assert a.shape == b.shape
    assert float(outs[0].float().abs().max()) > 0.0
    diff = int((a != b).sum())
    assert diff == 0, "%d of %d output bytes differ" % (diff, a.numel())


def test_fused_sweep_twice_gives_the_same_bits(monkeypatch):
    """Two launches on the same plan (slots dirty from the first one) agree: no stale slot is ever read."""
    B, V, H, W = 2, 2, 128, 192
    ref, srcs, ref_cam, src_cams = synth.frame_tuples(3, B, H, W, V)
    N = B * V
    src = srcs.transpose(0, 1).reshape(N, 3, H, W).contiguous().to(dev())
    src_cam = src_cams.transpose(0, 1).reshape(N, 2, 4, 4).contiguous().to(dev())
    sd = synth.kaiming_state(onets.state_shapes(onets.DEPTHNET_LAYERS), 5, 0.02)
    plan = _plan(N, B, H, W, 64, True, monkeypatch)
    first = _run_front_and_conv1(plan, ref.to(dev()), src, ref_cam.to(dev()), src_cam, sd)
    ref2, srcs2, _, _ = synth.frame_tuples(4, B, H, W, V)
    src2 = srcs2.transpose(0, 1).reshape(N, 3, H, W).contiguous().to(dev())
    other = _run_front_and_conv1(plan, ref2.to(dev()), src2, ref_cam.to(dev()), src_cam, sd)
    again = _run_front_and_conv1(plan, ref.to(dev()), src, ref_cam.to(dev()), src_cam, sd)
    assert not torch.equal(first.view(torch.uint8), other.view(torch.uint8))
    assert torch.equal(first.view(torch.uint8), again.view(torch.uint8))


def test_fused_sweep_rejects_what_it_cannot_run():
    lib = _lib.load()
    assert lib.cnm_sweep_scratch_bytes(64, 7) == 2 * 148 * 22 * 22 * 256
    assert lib.cnm_sweep_scratch_bytes(64, 4) == 0
    plan = engine.DepthNetPlan(2, 1, 64, 64, 3.0, "bf16", dev())
    assert "halo2s" not in plan.convs["conv1.0"].variants     # CNM_F16F8 rows only
