"""`get_dataloader` (reference: mggan/data_utils/data_loaders.py:10-100).

The reference's datasets live in a `data.zip` that is not part of its repository, and its
loaders depend on removed numpy / Pillow APIs, so they are outside this hot
path.  What IS in scope is the batch layout its collate function emits
(mggan/data_utils/trajectories_scene.py:40-78): `synthetic_*` datasets produce exactly that
dict from seeded synthetic scenes (mggan/synthetic.py), one item per scene, collated by
concatenating agents and rebuilding `seq_start_end`.
"""
import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from mggan.synthetic import SCALING_SMALL, SCENE_SHAPES, make_batch, make_image_batch, make_scene_image, scene_sizes

_PHASE_SEED = {"train": 0, "val": 1, "test": 2}


class SyntheticScenes(Dataset):
    """One item = one scene in the `seq_collate_scene` field layout."""

    def __init__(self, shape, num_scenes, phase="train", with_img=True, seed=42, images="agent"):
        """images: "agent" = an independent texture per agent (default); "host_crop" = `features` cut on the host from one
        image per scene (the reference's pipeline); "resident" = items carry `image_ids` and the crops are cut on the
        device from `scene_image_store()` (mggan/data_utils/scene_images.py)."""
        assert images in ("agent", "host_crop", "resident"), images
        rng = np.random.default_rng(seed * 31 + _PHASE_SEED.get(phase, 3))
        self.shape, self.with_img, self.images = shape, with_img, images
        self.sizes = scene_sizes(shape, num_scenes, rng)
        self.seeds = rng.integers(0, 2 ** 31 - 1, size=num_scenes).tolist()
        self.multi_future = 4 if shape == "gofp" else 1
        self.nan_frac = 0.25 if shape == "gofp" else 0.0
        self.ratio = 1.0          # metres; pixel datasets of the reference scale errors by 1/ratio

        self.dataset_name = "synthetic_" + shape
        self._eval = None

    def __len__(self):
        return len(self.sizes)

    def _eval_arrays(self):
        # agent-major views used by mggan.evaluation (reference dataset attributes obs_traj / pred_traj)
        if self._eval is None:
            items = [self[i] for i in range(len(self))]
            col = seq_collate_scene(items)
            self._eval = (col["in_xy"].permute(1, 0, 2), col["gt_xy"].permute(1, 0, 2), col["seq_start_end"])
        return self._eval

    obs_traj = property(lambda self: self._eval_arrays()[0])
    pred_traj = property(lambda self: self._eval_arrays()[1])
    seq_start_end = property(lambda self: self._eval_arrays()[2])
    scene_list = property(lambda self: [self.dataset_name] * len(self._eval_arrays()[2]))

    def __getitem__(self, i):
        if self.with_img and self.images != "agent":
            # multi-future scenes are replicas of one scene: they share its image
            b, _ = make_image_batch([self.sizes[i]], seed=self.seeds[i], resident=self.images == "resident",
                                    nan_frac=self.nan_frac, multi_future=self.multi_future)
            if "image_ids" in b:
                b["image_ids"][:] = self.image_id_offset + i
            return b
        return make_batch([self.sizes[i]], seed=self.seeds[i], with_img=self.with_img, nan_frac=self.nan_frac,
                          multi_future=self.multi_future)

    image_id_offset = 0          # added to the item index when several datasets share one SceneImageStore
    scaling_small = SCALING_SMALL

    def scene_image_list(self):
        return [make_scene_image(self.seeds[i] * 1000) for i in range(len(self))]

    def scene_image_store(self, device="cuda"):
        """The dataset's scene images uploaded once (image id = image_id_offset + item index)."""
        from mggan.data_utils.scene_images import SceneImageStore
        assert self.image_id_offset == 0
        return SceneImageStore(self.scene_image_list(), SCALING_SMALL, device)


def seq_collate_scene(items):
    out, sse, cur = {}, [], 0
    for it in items:
        for s, e in it["seq_start_end"]:
            sse.append([cur + s, cur + e])
        cur += it["in_xy"].shape[1]
    for key in ("in_xy", "in_dxdy", "gt_xy", "gt_dxdy"):
        out[key] = torch.from_numpy(np.concatenate([it[key] for it in items], 1))
    if "features" in items[0]:
        out["features"] = torch.from_numpy(np.concatenate([it["features"] for it in items], 0))
    if "image_ids" in items[0]:
        out["image_ids"] = torch.from_numpy(np.concatenate([it["image_ids"] for it in items], 0))
    out["seq_start_end"] = sse
    return out


def get_dataloader(dataset, phase="train", augment=False, batch_size=8, workers=0, shuffle=False, split=None,
                   num_scenes=64, with_img=True, seed=42, images="agent"):
    if not dataset.startswith("synthetic_"):
        raise NotImplementedError(
            f"dataset '{dataset}': the reference datasets (data.zip) are not shipped; use synthetic_"
            f"{{{','.join(SCENE_SHAPES)}}}")
    shape = dataset[len("synthetic_"):]
    ds = SyntheticScenes(shape, num_scenes, phase, with_img, seed, images)
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=workers, collate_fn=seq_collate_scene)
