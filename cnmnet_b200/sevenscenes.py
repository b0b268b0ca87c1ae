"""7-Scenes front end: the reference's file loader and evaluation driver around the GPU pipeline.

Reference behaviour mirrored here:
  * ``LoadSevenScenes`` (eval.py:26-159): the 18 test sequences, ``frame-XXXXXX.{color.png, depth.png, pose.txt}``
    triplets, colour image resized to the network size with cv2 INTER_LINEAR and ImageNet-normalised, depth in metres
    at its ORIGINAL size (millimetre PNG / 1000), pose (camera->world) inverted to the world->camera extrinsic,
    ``cam[2,4,4]`` = (extrinsic, K with fx, fy, cx, cy scaled by the resize factors).  Same method names, arguments and
    return values.
  * the evaluation loop of ``eval_refine`` (eval.py:408-455): every 3rd frame with sources at +10 / -10
    (cnmnet_b200.video.reference_tuples), and the stored prediction ``1 / (idepth + 1e-4)`` with values above 100 m
    zeroed (eval.py:493-494).
  * ``cal_metrics`` (eval.py:995-1090): prediction resized to the ground-truth size (cv2 INTER_LINEAR), clamped to
    [0.3, 8] m, scored on 0.3 < gt < 8 -- the per-image reductions run on the GPU (cnmnet_b200.depth_metrics).

One deliberate difference: pose files are parsed on any whitespace.  The reference passes ``delimiter='\\t '`` to
np.loadtxt (eval.py:124), which current NumPy rejects (multi-character delimiter); the files are tab + space separated,
so the parsed matrix is the same.

Host-side I/O only (cv2 + numpy); all arithmetic of the depth path stays in the CUDA pipeline.
"""
import os
from copy import copy

import numpy as np
import torch

from .video import VideoDepthStream

TEST_SEQUENCES = [("chess", "seq-03"), ("chess", "seq-05"), ("fire", "seq-03"), ("fire", "seq-04"), ("heads", "seq-01"),
                  ("office", "seq-02"), ("office", "seq-06"), ("office", "seq-07"), ("office", "seq-09"),
                  ("pumpkin", "seq-01"), ("pumpkin", "seq-07"), ("redkitchen", "seq-03"), ("redkitchen", "seq-04"),
                  ("redkitchen", "seq-06"), ("redkitchen", "seq-12"), ("redkitchen", "seq-14"), ("stairs", "seq-01"),
                  ("stairs", "seq-04")]
IMAGENET_MEAN, IMAGENET_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def _cv2():
    import cv2
    return cv2


class LoadSevenScenes(object):
    """eval.py:26-159."""

    def __init__(self, root_dir):
        self.root_dir = root_dir
        self.test_seqs_list = list(TEST_SEQUENCES)
        self.intrinsics = np.asarray([[585, 0, 320], [0, 585, 240], [0, 0, 1]])

    def get_filepaths(self, scene, seq):
        """-> list of {'rgb', 'depth', 'pose', 'pred_depth_name'} for every colour frame of the sequence, sorted."""
        seq_dir = os.path.join(self.root_dir, scene, seq)
        out = []
        for filename in sorted(os.listdir(seq_dir)):
            if "color" in filename:
                out.append({"rgb": os.path.join(seq_dir, filename),
                            "depth": os.path.join(seq_dir, filename.replace("color", "depth")),
                            "pose": os.path.join(seq_dir, filename.replace("color.png", "pose.txt")),
                            "pred_depth_name": filename.replace("color", "pred_depth")})
        return out

    def load_sample(self, sample_path, image_height_expected, image_width_expected):
        """-> rgb [3,H,W] fp32 (normalised), depth [h0,w0] fp32 metres (original size), cam [2,4,4] fp32, pred_depth_name."""
        cv2 = _cv2()
        rgb = cv2.imread(sample_path["rgb"], -1)
        if rgb is None:
            raise FileNotFoundError(sample_path["rgb"])
        rgb = cv2.cvtColor(rgb, cv2.COLOR_BGR2RGB)
        depth = cv2.imread(sample_path["depth"], -1)
        if depth is None:
            raise FileNotFoundError(sample_path["depth"])
        depth = depth / 1000.0
        pose = self.read_pose_file(sample_path["pose"])
        cam = self.get_cam(self.intrinsics, self.pose2extrinsics(pose))
        scale_x = float(image_width_expected) / rgb.shape[1]
        scale_y = float(image_height_expected) / rgb.shape[0]
        cam = self.scale_cam(cam, scale_x, scale_y)
        rgb = self.scale_img(rgb, image_height_expected, image_width_expected, "linear")
        rgb = self.normalize_image(rgb)
        rgb, depth, cam = self.toTensor(rgb, depth, cam)
        return rgb, depth, cam, sample_path["pred_depth_name"]

    def toTensor(self, rgb, depth, cam):
        return (torch.from_numpy(np.float32(rgb)).permute(2, 0, 1), torch.from_numpy(np.float32(depth)),
                torch.from_numpy(np.float32(cam)))

    def scale_img(self, image, expected_height, expected_width, interpolation):
        cv2 = _cv2()
        if interpolation == "linear":
            return cv2.resize(image, (expected_width, expected_height), interpolation=cv2.INTER_LINEAR)
        if interpolation == "nearest":
            return cv2.resize(image, (expected_width, expected_height), interpolation=cv2.INTER_NEAREST)
        raise Exception("interpolation cannot be None")

    def normalize_image(self, img):
        img = img / 255.
        return ((img - IMAGENET_MEAN) / IMAGENET_STD).astype(np.float32)

    def read_pose_file(self, filepath):
        return np.loadtxt(filepath, dtype="f")

    def pose2extrinsics(self, pose):
        return np.linalg.inv(pose)

    def get_cam(self, intrinsics, extrinsics):
        cam = np.zeros((2, 4, 4))
        cam[0] = extrinsics
        cam[1][0:3, 0:3] = self.intrinsics
        return cam

    def scale_cam(self, cam, scale_x, scale_y):
        new_cam = copy(cam)
        new_cam[1][0][0] *= scale_x
        new_cam[1][1][1] *= scale_y
        new_cam[1][0][2] *= scale_x
        new_cam[1][1][2] *= scale_y
        return new_cam

    # ---- whole sequences ----
    def load_sequence(self, scene, seq, image_height, image_width):
        """-> frames [T,3,H,W], gt depths (list of [h0,w0] tensors), cams [T,2,4,4], pred_depth names."""
        paths = self.get_filepaths(scene, seq)
        rgbs, depths, cams, names = [], [], [], []
        for sp in paths:
            rgb, depth, cam, name = self.load_sample(sp, image_height, image_width)
            rgbs.append(rgb)
            depths.append(depth)
            cams.append(cam)
            names.append(name)
        return torch.stack(rgbs), depths, torch.stack(cams), names


def stored_depth(depth):
    """The prediction as the reference stores it (eval.py:493-494): 1 / (idepth + 1e-4), zero beyond 100 m."""
    pred = 1.0 / (1.0 / depth + 1e-4)
    pred = torch.where(pred > 100.0, torch.zeros_like(pred), pred)
    return pred


def evaluate(pipeline, root_dir, sequences=None, image_height=192, image_width=256, batch=5, views=3, device=None):
    """Runs `pipeline` (a CNMPipeline) over 7-Scenes sequences the way eval_refine does and scores like cal_metrics.

    -> (means {l1, abs_relative, ..., a3} over all evaluated frames, per_frame [(scene, seq, index, 9 metrics)])."""
    from .metrics import NAMES, depth_metrics
    cv2 = _cv2()
    loader = LoadSevenScenes(root_dir)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    stream = VideoDepthStream(pipeline, batch=batch, views=views)
    rows = []
    for scene, seq in (loader.test_seqs_list if sequences is None else sequences):
        frames, gts, cams, _ = loader.load_sequence(scene, seq, image_height, image_width)
        for index, depth, _normal, _prob in stream.run(frames, cams):
            gt = gts[index]
            h0, w0 = gt.shape
            pred = stored_depth(depth).numpy()
            pred = cv2.resize(pred, (w0, h0), interpolation=cv2.INTER_LINEAR)      # eval.py:1030
            per_image, _ = depth_metrics(torch.from_numpy(pred)[None].to(dev), gt[None].to(dev))
            rows.append((scene, seq, index) + tuple(per_image[0, 1:].tolist()))
    if not rows:
        return {}, rows
    vals = np.array([r[3:] for r in rows], dtype=np.float64)
    return dict(zip(NAMES, vals.mean(0).tolist())), rows
