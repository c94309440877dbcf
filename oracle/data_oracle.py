"""TEST INFRASTRUCTURE — not product code.  numpy restatement of the data-side operators next to the MG-GAN hot path
(SURVEY.md 8f #2, #3), pinned against the unmodified reference by tests/golden/scene_crop.npz and
tests/golden/evaluation.npz (oracle/make_golden_eval.py, tests/test_eval_crop_cpu.py).  Only tests/, smoke() and the
CPU legs of bench.py may import this file.
"""
import numpy as np

CROP = 33
MARGIN = 16


def image_features_small(image_u8, last_xy, scaling_small):
    """`BaseDataset.ImageFeatures_small` (reference mggan/data_utils/BaseTrajectories.py:254-288), format "meter",
    margin_in = margin_out = 16.

    image_u8 (H, W, 3) uint8 = the scene's `small_image`; last_xy (2,) float32 = last observed position in metres.
    centre pixel = int(last_xy * (1 / scaling_small)) computed in float32 and truncated toward zero (:265-268), crop box
    [c - 16, c + 17) with zeros outside the image (PIL `Image.crop` semantics, :270-277), channels 0-2 =
    -1 + u8 * 2 / 256 (:284), channel 3 = one-hot at [16, 16] (:279-282).  -> (4, 33, 33) float32.
    """
    h, w, _ = image_u8.shape
    scale = np.float32(1.0 / scaling_small)
    c = (np.asarray(last_xy, dtype=np.float32) * scale).astype(np.int64)        # astype(int): truncation toward zero
    out = np.empty((4, CROP, CROP), dtype=np.float32)
    ys = c[1] - MARGIN + np.arange(CROP)
    xs = c[0] - MARGIN + np.arange(CROP)
    inside = ((ys >= 0) & (ys < h))[:, None] & ((xs >= 0) & (xs < w))[None, :]
    patch = image_u8[np.clip(ys, 0, h - 1)[:, None], np.clip(xs, 0, w - 1)[None, :]].astype(np.float64)   # (33, 33, 3)
    patch = np.where(inside[..., None], patch, 0.0)
    out[:3] = (-1.0 + patch * 2.0 / 256.0).transpose(2, 0, 1).astype(np.float32)
    out[3] = 0.0
    out[3, MARGIN, MARGIN] = 1.0
    return out


def tube_inside(manifold, tests, radius):
    """`Manifold.compute_inside` (reference mggan/manifold.py:9-18, :70-77): manifold (m, T, 2), tests (n, T, 2) float32,
    tube radius linspace(radius / T, radius, T) in float64; a test trajectory is inside when at EVERY step it is closer
    than the step's radius to SOME manifold sample.  Distances are float32 like numpy's norm of float32 inputs."""
    T = manifold.shape[1]
    rad = np.linspace(radius / T, radius, T, endpoint=True)
    res = np.zeros(len(tests), dtype=bool)
    for i in range(len(tests)):
        diff = manifold - tests[i][None]
        d = np.sqrt((diff * diff).sum(-1))
        res[i] = (d < rad[None]).any(0).all(0)
    return res
