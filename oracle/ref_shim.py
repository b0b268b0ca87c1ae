"""Import shim for the UNMODIFIED reference at /root/reference (build container only).

Recipe: SURVEY.md Appendix D.  The reference imports tensorflow (unused on the hot
path), hard-codes torch.cuda.FloatTensor and uses Tensor.get_device() as a device;
on a CPU-only box those three are patched and nothing else.  Never used at run time
on the GPU box (``/root/reference`` does not exist there) and never by the product.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CNM_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "depthnet"))


def load_reference():
    """Returns a namespace with depthNet, DepthRefineNet, Depth2normal, ... from the reference."""
    import torch

    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("tensorflow", "tensorflow.python", "tensorflow.python.lib",
                 "tensorflow.python.lib.io", "tensorflow.python.lib.io.file_io"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["tensorflow.python.lib.io"].file_io = sys.modules["tensorflow.python.lib.io.file_io"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not torch.cuda.is_available():
        torch.cuda.FloatTensor = torch.FloatTensor
    torch.Tensor.get_device = lambda self: self.device
    import depthnet.depthNet_model as m
    import depthnet.depth_util as du
    import depthnet.inverse_warp as iw

    ns = types.SimpleNamespace(
        depthNet=m.depthNet, DepthRefineNet=m.DepthRefineNet,
        Depth2normal=du.Depth2normal, get_normal_by_planes=du.get_normal_by_planes,
        process_camera_parameters=du.process_camera_parameters,
        get_pixel_coordinates=du.get_pixel_coordinates,
        pixel2cam=iw.pixel2cam, inverse_warp=iw.inverse_warp, _iw=iw)
    return ns


def reset_pixel_cache(ns):
    """The reference caches a pixel grid in a module global keyed on height only."""
    ns._iw.pixel_coords = None
