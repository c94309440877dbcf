"""Scene images resident in HBM and the per-agent crop features cut from them on the device.

The reference builds `features (N, 4, 33, 33)` on the host, one PIL crop per agent and batch
(`BaseDataset.ImageFeatures_small`, mggan/data_utils/BaseTrajectories.py:254-288, called from
trajectories_scene.py:343-351) and ships 17,424 bytes per agent to the GPU every step.  Here the scenes' `small_image`s
(u8 RGB, a few hundred KB each) are uploaded once; a batch then carries only `image_ids` (one int32 per agent) and the
crop is cut by `mggan_scene_crop` from the agents' last observed positions, bit-identical to the host crop.
"""
import numpy as np
import torch

from mggan import kernels as K


class SceneImageStore:
    def __init__(self, images, scaling_small=0.5, device="cuda", fmt="meter"):
        """images: list of (H, W, 3) uint8 arrays (the scenes' `small_image`); scaling_small: metres per pixel of the small
        image (one number or one per image; the loaders use 0.5 / 0.7 / 1.2, data_loaders.py:36,74,85); fmt "pixel" means
        the trajectories already are small-image pixels (scale 1, BaseTrajectories.py:257-260)."""
        if len(images) == 0:
            raise ValueError("SceneImageStore needs at least one image")
        scal = np.broadcast_to(np.asarray(scaling_small, dtype=np.float64), (len(images),))
        offs, wh, cur = [], [], 0
        for im in images:
            im = np.asarray(im)
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
                raise ValueError(f"scene images must be (H, W, 3) uint8, got {im.dtype} {im.shape}")
            offs.append(cur)
            wh.append((im.shape[1], im.shape[0]))
            cur += im.size
        atlas = np.concatenate([np.ascontiguousarray(im).reshape(-1) for im in images])
        self.device = torch.device(device)
        self.n_images = len(images)
        self.atlas = torch.from_numpy(atlas).to(self.device)
        self.img_off = torch.tensor(offs, dtype=torch.int64, device=self.device)
        self.img_wh = torch.tensor(wh, dtype=torch.int32, device=self.device)
        # float32(1 / scaling_small): the reference multiplies a float32 position by this Python float (numpy keeps float32)
        scale = np.ones(len(images), np.float32) if fmt == "pixel" else (1.0 / scal).astype(np.float32)
        self.img_scale = torch.from_numpy(scale).to(self.device)

    def nbytes(self):
        return int(self.atlas.numel())

    def crop(self, image_ids, last_xy, out=None):
        """image_ids (N) int (host or device), last_xy (N, 2) device fp32 = in_xy[-1] -> features (N, 4, 33, 33)
        (written into `out` when given)."""
        ids = torch.as_tensor(image_ids)
        if ids.dtype != torch.int32:
            ids = ids.to(torch.int32)
        ids = ids.to(self.device, non_blocking=True)
        return K.scene_crop(self.atlas, self.img_off, self.img_wh, self.img_scale, ids, last_xy, out=out)


class DeferredCrop:
    """Crop features of a staged batch that have not been cut yet.  The training loop stages batch i + 1 on a copy stream
    while batch i computes; cutting its crops there would write 17,424 B per agent into a staging buffer that the main
    stream then copies once more into the captured iteration's input.  The loop stages only the image ids and the
    trajectories, and the iteration cuts the crops itself on the main stream, straight into the buffer it reads
    (`materialize(out=...)`): one write of the crops per step instead of a write, a read and a write."""

    def __init__(self, store, ids, in_xy):
        self.store, self.ids, self.in_xy = store, ids, in_xy          # ids (N) int32 device, in_xy (obs_len, N, 2) device
        self.shape = (int(ids.numel()), 4, 33, 33)

    def materialize(self, out=None):
        return self.store.crop(self.ids, self.in_xy[-1], out=out)
