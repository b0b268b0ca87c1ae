"""Import shim for the UNMODIFIED reference (test / benchmark infrastructure, never used by the product).

Recipe: SURVEY.md Appendix D.  The reference imports tensorflow (unused on the hot path), hard-codes
torch.cuda.FloatTensor and uses Tensor.get_device() as a device; on a CPU-only box those three are patched and nothing
else.  Where it is found, in this order: $CNM_REFERENCE_ROOT, /root/reference (the build container), baseline/_ref
(git-ignored copy of the six files the import chain needs, made by __graft_entry__.build() in the build container so
that bench.py can time the reference itself on the GPU box's host cores and on its GPU -- BASELINE.md section 4.2).
"""
import os
import shutil
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VENDORED_ROOT = os.path.join(_REPO, "baseline", "_ref")
VENDORED_FILES = ("depthnet/__init__.py", "depthnet/depthNet_model.py", "depthnet/depth_util.py",
                  "depthnet/inverse_warp.py", "scannet/__init__.py", "scannet/preprocess.py")


def _find_root():
    for root in (os.environ.get("CNM_REFERENCE_ROOT", ""), "/root/reference", VENDORED_ROOT):
        if root and os.path.isfile(os.path.join(root, "depthnet", "depthNet_model.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def vendor_reference(src_root="/root/reference"):
    """Copies the reference files of the import chain, unmodified, to the git-ignored baseline/_ref (build container
    only; a no-op where the reference tree is absent).  Returns True if baseline/_ref is populated afterwards."""
    global REFERENCE_ROOT
    if os.path.isfile(os.path.join(src_root, "depthnet", "depthNet_model.py")):
        for rel in VENDORED_FILES:
            dst = os.path.join(VENDORED_ROOT, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(src_root, rel), dst)
    ok = all(os.path.isfile(os.path.join(VENDORED_ROOT, rel)) for rel in VENDORED_FILES)
    REFERENCE_ROOT = _find_root()
    return ok


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "depthnet", "depthNet_model.py"))


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


class cpu_mode:
    """Context manager for running the reference on the CPU of a box that HAS a GPU: the reference allocates
    torch.cuda.FloatTensor in get_pixel_coordinates / getVolume (depth_util.py:20,53-54; depthNet_model.py:199,204);
    while active that name is aliased to torch.FloatTensor, as load_reference() does permanently on CPU-only boxes."""

    def __enter__(self):
        import torch
        self._saved = torch.cuda.FloatTensor
        torch.cuda.FloatTensor = torch.FloatTensor
        return self

    def __exit__(self, *exc):
        import torch
        torch.cuda.FloatTensor = self._saved
        return False


def reset_pixel_cache(ns):
    """The reference caches a pixel grid in a module global keyed on height only."""
    ns._iw.pixel_coords = None
