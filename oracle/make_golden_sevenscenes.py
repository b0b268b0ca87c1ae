"""Generates tests/golden/sevenscenes_loader.npz: outputs of the reference's own LoadSevenScenes (eval.py:26-159) on a
small synthetic sequence written in the 7-Scenes file format.  Run in the build container (needs /root/reference):

    python oracle/make_golden_sevenscenes.py

eval.py cannot be imported (sacred / easydict are not installed): the class definition is cut out of the file with
`ast` and executed unmodified, with the names its methods use (cv2, np, torch, os, copy).  One adaptation, recorded in
the fixture: read_pose_file is overridden to split on whitespace, because NumPy >= 2 rejects the reference's
two-character delimiter '\\t ' (eval.py:124); the files written here use that delimiter."""
import ast
import os
import sys
import tempfile
from copy import copy

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("CNM_REFERENCE_ROOT", "/root/reference")


def write_sequence(root, scene="chess", seq="seq-03", frames=4, h=48, w=64, seed=0):
    """frame-XXXXXX.color.png (8-bit BGR on disk), .depth.png (uint16 millimetres), .pose.txt (4x4, tab+space separated)."""
    rng = np.random.default_rng(seed)
    d = os.path.join(root, scene, seq)
    os.makedirs(d, exist_ok=True)
    for i in range(frames):
        stem = os.path.join(d, "frame-%06d" % i)
        cv2.imwrite(stem + ".color.png", rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        cv2.imwrite(stem + ".depth.png", rng.integers(300, 6000, (h, w)).astype(np.uint16))
        th = 0.1 * i
        pose = np.eye(4)
        pose[:3, :3] = [[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]]
        pose[:3, 3] = rng.normal(0, 0.5, 3)
        with open(stem + ".pose.txt", "w") as f:
            for row in pose:
                f.write("\t ".join("%.7e" % v for v in row) + "\t \n")
    return d


def reference_class():
    src = open(os.path.join(REF, "eval.py")).read()
    node = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "LoadSevenScenes"][0]
    ns = {"cv2": cv2, "np": np, "torch": torch, "os": os, "copy": copy}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "eval.py", "exec"), ns)
    cls = ns["LoadSevenScenes"]
    cls.read_pose_file = lambda self, filepath: np.loadtxt(filepath, dtype="f")
    return cls


def main():
    with tempfile.TemporaryDirectory() as root:
        write_sequence(root)
        loader = reference_class()(root)
        paths = loader.get_filepaths("chess", "seq-03")
        out = {"num_frames": len(paths), "names": np.array([p["pred_depth_name"] for p in paths]),
               "test_seqs": np.array(["%s/%s" % t for t in loader.test_seqs_list]), "intrinsics": loader.intrinsics}
        for i, sp in enumerate(paths):
            rgb, depth, cam, name = loader.load_sample(sp, 24, 32)
            out["rgb%d" % i], out["depth%d" % i], out["cam%d" % i] = rgb.numpy(), depth.numpy(), cam.numpy()
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sevenscenes_loader.npz"), **out)
        print("wrote", sorted(out))


if __name__ == "__main__":
    main()
