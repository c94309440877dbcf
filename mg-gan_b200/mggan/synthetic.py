"""Synthetic pedestrian-scene batches in the exact `seq_collate_scene` layout.

The reference's datasets (`data.zip`) are not part of the checkout, so every
benchmark, test and golden vector is built from seeded synthetic scenes that
have the same dict layout, dtypes and value ranges as the reference's collate
function (reference: mggan/data_utils/trajectories_scene.py:40-78 for the dict,
:211-214,366 for `in_dxdy = in_xy[1:] - in_xy[:-1]`, and
mggan/data_utils/BaseTrajectories.py:278-286 for the 4x33x33 crop: channels
0-2 in [-1, 1) as `-1 + u8 * 2 / 256`, channel 3 one-hot at [16, 16]).

Shapes (time-major, fp32): in_xy (8,N,2), in_dxdy (7,N,2), gt_xy (12,N,2),
gt_dxdy (12,N,2), features (N,4,33,33); seq_start_end is a python list of
[start, end) pairs, one per scene.
"""
import numpy as np

OBS_LEN = 8
PRED_LEN = 12
CROP = 33

SCENE_SHAPES = {
    # name: (min agents, max agents) per scene, BASELINE.json configs
    "tiny": (4, 4),          # cfg 1: one 4-agent scene
    "eth": (1, 6),           # cfg 2: sparse ETH/BIWI-like
    "sdd": (2, 16),          # cfg 3: Stanford-drone-like
    "univ": (32, 32),        # cfg 4: dense, 32 agents per scene
    "gofp": (2, 8),          # cfg 5: multi-future (replicated observations, NaN-masked agents)
}


def scene_sizes(shape, num_scenes, rng):
    lo, hi = SCENE_SHAPES[shape]
    if lo == hi:
        return [lo] * num_scenes
    return rng.integers(lo, hi + 1, size=num_scenes).tolist()


def _lowpass_u8(rng, n):
    """Seeded smooth u8 texture, (n, 3, 33, 33); content is irrelevant to timing."""
    coarse = rng.integers(0, 256, size=(n, 3, 5, 5)).astype(np.float32)
    # bilinear upsample 5x5 -> 33x33 (separable, align_corners)
    pos = np.linspace(0.0, 4.0, CROP, dtype=np.float32)
    i0 = np.minimum(np.floor(pos).astype(np.int64), 3)
    f = pos - i0
    rows = coarse[:, :, i0, :] * (1 - f)[None, None, :, None] + coarse[:, :, i0 + 1, :] * f[None, None, :, None]
    img = rows[:, :, :, i0] * (1 - f)[None, None, None, :] + rows[:, :, :, i0 + 1] * f[None, None, None, :]
    return np.clip(np.rint(img), 0, 255)


def make_batch(sizes, seed=0, with_img=True, nan_frac=0.0, multi_future=1):
    """Build one collated batch.

    sizes: agents per scene.  multi_future=R replicates every scene R times with
    identical observations and R different futures (GoFP-shape); nan_frac masks
    that fraction of agents' futures with NaN (reference:
    mggan/data_utils/trajectories_scene.py:169-174).
    Returns a dict of numpy arrays + the python `seq_start_end` list.
    """
    rng = np.random.default_rng(seed)
    branch = np.deg2rad(np.array([-45.0, 0.0, 45.0, 90.0], dtype=np.float64))
    obs_all, fut_all, sse = [], [], []
    cursor = 0
    for n in sizes:
        start = rng.uniform(0.0, 15.0, size=(n, 2))
        heading = rng.uniform(0.0, 2 * np.pi, size=n)
        speed = np.abs(rng.normal(1.3, 0.3, size=n)) * 0.4
        vel = np.stack([np.cos(heading), np.sin(heading)], -1) * speed[:, None]
        steps = np.arange(OBS_LEN)[:, None, None]
        obs = start[None] + steps * vel[None] + rng.normal(0.0, 0.05, size=(OBS_LEN, n, 2))
        turn0 = rng.normal(0.0, 0.03, size=n)
        for r in range(multi_future):
            turn = turn0 + (branch[r % 4] / PRED_LEN if multi_future > 1 else 0.0)
            pos = obs[-1].copy()
            ang = np.arctan2(vel[:, 1], vel[:, 0])
            fut = np.empty((PRED_LEN, n, 2))
            for t in range(PRED_LEN):
                ang = ang + turn
                pos = pos + np.stack([np.cos(ang), np.sin(ang)], -1) * speed[:, None]
                fut[t] = pos + rng.normal(0.0, 0.05, size=(n, 2))
            obs_all.append(obs)
            fut_all.append(fut)
            sse.append([cursor, cursor + n])
            cursor += n
    in_xy = np.concatenate(obs_all, 1).astype(np.float32)
    gt_xy = np.concatenate(fut_all, 1).astype(np.float32)
    N = in_xy.shape[1]
    in_dxdy = in_xy[1:] - in_xy[:-1]
    gt_dxdy = np.concatenate([gt_xy[:1] - in_xy[-1:], gt_xy[1:] - gt_xy[:-1]], 0)
    if nan_frac > 0:
        masked = rng.random(N) < nan_frac
        if masked.all():
            masked[0] = False
        gt_xy[:, masked] = np.nan
        gt_dxdy[:, masked] = np.nan
    batch = {
        "in_xy": in_xy,
        "in_dxdy": in_dxdy.astype(np.float32),
        "gt_xy": gt_xy,
        "gt_dxdy": gt_dxdy.astype(np.float32),
        "seq_start_end": sse,
    }
    if with_img:
        feat = np.empty((N, 4, CROP, CROP), dtype=np.float32)
        feat[:, :3] = -1.0 + _lowpass_u8(rng, N) * (2.0 / 256.0)
        feat[:, 3] = 0.0
        feat[:, 3, CROP // 2, CROP // 2] = 1.0
        batch["features"] = feat
    return batch


def make_config_batch(shape, num_scenes, seed=0, with_img=True):
    """Batch for one of the BASELINE.json scene shapes."""
    rng = np.random.default_rng(seed + 7919)
    sizes = scene_sizes(shape, num_scenes, rng)
    if shape == "gofp":
        return make_batch(sizes, seed=seed, with_img=with_img, nan_frac=0.25, multi_future=4)
    return make_batch(sizes, seed=seed, with_img=with_img)


# ---- scene images + crops: the reference cuts every agent's 33 x 33 crop from its scene's
# `small_image` on the host (BaseTrajectories.py:254-288); the B200 path can keep the images in HBM and cut the crops in
# `mggan_scene_crop` instead (mggan/data_utils/scene_images.py).  Both need scene images, which the per-agent textures
# of `make_batch` are not: these helpers make one seeded image per scene.
SCALING_SMALL = 0.5          # metres per small-image pixel (BaseTrajectories.py:40)
IMAGE_HW = (64, 64)


def make_scene_image(seed, hw=IMAGE_HW):
    """Seeded smooth u8 RGB texture (H, W, 3)."""
    rng = np.random.default_rng(seed)
    h, w = hw
    coarse = rng.integers(0, 256, size=(3, 9, 9)).astype(np.float32)
    py, px = np.linspace(0.0, 8.0, h, dtype=np.float32), np.linspace(0.0, 8.0, w, dtype=np.float32)
    y0, x0 = np.minimum(np.floor(py).astype(np.int64), 7), np.minimum(np.floor(px).astype(np.int64), 7)
    fy, fx = py - y0, px - x0
    rows = coarse[:, y0, :] * (1 - fy)[None, :, None] + coarse[:, y0 + 1, :] * fy[None, :, None]
    img = rows[:, :, x0] * (1 - fx)[None, None, :] + rows[:, :, x0 + 1] * fx[None, None, :]
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8).transpose(1, 2, 0))


def host_crop_features(image_u8, last_xy, scaling_small=SCALING_SMALL):
    """The reference's HOST pipeline for a scene's agents: (n, 2) last observed positions -> (n, 4, 33, 33) features
    (`ImageFeatures_small`, BaseTrajectories.py:254-288: centre = int(xy / scaling_small) in float32, box [c-16, c+17)
    zero-padded like PIL's crop, -1 + u8 * 2 / 256, one-hot position channel).  Dataset-side code, like the reference's."""
    h, w, _ = image_u8.shape
    n = len(last_xy)
    c = (np.asarray(last_xy, np.float32) * np.float32(1.0 / scaling_small)).astype(np.int64)
    ys = c[:, 1, None] - CROP // 2 + np.arange(CROP)[None]                    # (n, 33)
    xs = c[:, 0, None] - CROP // 2 + np.arange(CROP)[None]
    ok = ((ys >= 0) & (ys < h))[:, :, None] & ((xs >= 0) & (xs < w))[:, None, :]
    patch = image_u8[np.clip(ys, 0, h - 1)[:, :, None], np.clip(xs, 0, w - 1)[:, None, :]]      # (n, 33, 33, 3)
    patch = np.where(ok[..., None], patch, 0).astype(np.float32)
    feat = np.zeros((n, 4, CROP, CROP), np.float32)
    feat[:, :3] = (-1.0 + patch * (2.0 / 256.0)).transpose(0, 3, 1, 2)
    feat[:, 3, CROP // 2, CROP // 2] = 1.0
    return feat


def make_image_batch(sizes, seed=0, resident=False, first_image_id=0, **kw):
    """`make_batch` with one scene image per entry of `sizes` (the replicas of a multi-future scene share theirs).
    resident=False: `features` are cut on the host from the scene images (the reference's pipeline); resident=True: the
    batch carries `image_ids` (N,) int32 instead and the crops are cut on the device from a `SceneImageStore`.
    Returns (batch, [scene images])."""
    batch = make_batch(sizes, seed=seed, with_img=False, **kw)
    sse = batch["seq_start_end"]
    rep = len(sse) // len(sizes)                     # multi_future replicas per scene
    images = [make_scene_image(seed * 1000 + i) for i in range(len(sizes))]
    if resident:
        batch["image_ids"] = np.concatenate([np.full(e - s, first_image_id + j // rep, np.int32)
                                             for j, (s, e) in enumerate(sse)])
    else:
        last = batch["in_xy"][-1]
        batch["features"] = np.concatenate([host_crop_features(images[j // rep], last[s:e])
                                            for j, (s, e) in enumerate(sse)])
    return batch, images
