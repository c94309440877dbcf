"""Trainer base (reference: mggan/abstract_train.py:25-322): device, two AdamW optimisers, two
cosine schedules, the epoch / iteration loop, validation with best-checkpoint tracking and the
checkpoint layout `{generator, discriminator, gen_opt, disc_opt}` under
`<log_dir>/<experiment>/<name>/version_<v>/checkpoints/`.

Differences from the reference, all behind the same method names:
  * the optimisers are `FusedAdamW` (clip + AdamW in two launches, same state_dict layout);
  * per-iteration loss scalars stay on the device and are averaged once per epoch (the
    reference calls .item() seven times per iteration);
  * no global `torch.set_default_tensor_type`: tensors are created on `self.device` explicitly;
  * gan_obj NS / MM / LS are built; W raises NotImplementedError (it cannot run in the reference either);
  * optional data-parallel mode (`mggan.distributed`): scenes are sharded over ranks and the
    step functions all-reduce gradients, loss normalisers and BatchNorm statistics.
"""
import abc
import math
from argparse import Namespace
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

from mggan.data_utils.data_loaders import get_dataloader
from mggan.logging import Experiment
from mggan.model.config import get_parser
from mggan.optim import FusedAdamW
from mggan.utils import get_argparse_defaults, load_hparams_from_tags_csv


def _mean(values):
    if len(values) == 0:
        return float("nan")
    if torch.is_tensor(values[0]):
        return float(torch.stack([v.detach().float().reshape(()) for v in values]).mean().item())
    return float(np.mean(values))


class MultiGeneratorGAN(abc.ABC):
    def __init__(self, generator, discriminator, config, writer, dist_ctx=None):
        self.writer = writer
        self.config = config
        self.device = torch.device("cuda" if config.gpus else "cpu")
        if self.device.type != "cuda":
            raise RuntimeError("the B200 path has no CPU fallback: run with a CUDA device (--gpus 0)")
        self.D = discriminator.to(self.device)
        self.G = generator.to(self.device)
        self.l2_weight = self.config.l2_loss_weight
        self.gan_type = self.config.gan_type
        self.dist = dist_ctx
        self.log_dir = Path(self.writer.get_data_path(self.writer.name, self.writer.version))
        self.model_save_dir = self.log_dir / "checkpoints"
        self.model_save_dir.mkdir(exist_ok=True, parents=True)

        self.optimizerD = FusedAdamW(self.D.parameters(), lr=self.config.d_lr, betas=(config.beta1, 0.999))
        self.optimizerG = FusedAdamW(self.G.parameters(), lr=self.config.g_lr, betas=(config.beta1, 0.999))
        self.lr_schedulerD = torch.optim.lr_scheduler.CosineAnnealingLR(self.optimizerD, config.epochs, eta_min=0)
        self.lr_schedulerG = torch.optim.lr_scheduler.CosineAnnealingLR(self.optimizerG, config.epochs, eta_min=0)
        self.epoch = 0
        self._graph = None                 # mggan.graph.GraphedIteration while capturing
        self._graphs = []                  # captured iterations, one per batch structure, most recently used first (LRU)
        self._graph_seen = {}              # structure key -> eager iterations seen (a structure is captured on its 2nd one)
        self.graph_hits = self.graph_misses = 0
        self._graph_failed = set()         # structures whose capture raised: they stay eager
        self._stage_ring = {}              # slot -> {name: persistent device staging buffer} (train_iterations prefetch)
        self.scene_images = None           # SceneImageStore: crops are cut on the device for batches carrying `image_ids`
        # GAN objective (reference abstract_train.py:61-85): phi_1 (D on real), phi_2 (D on fake), phi_3 (G on fake), each a
        # (loss kernel, which label, sign) triple applied to the discriminator output with a scalar smoothed label
        from mggan import kernels as K
        if self.config.gan_obj == "NS":
            self.phi_1, self.phi_2, self.phi_3 = (K.bce_scalar_label, "real", 1.0), (K.bce_scalar_label, "fake", 1.0), (K.bce_scalar_label, "real", 1.0)
        elif self.config.gan_obj == "MM":
            self.phi_1, self.phi_2, self.phi_3 = (K.bce_scalar_label, "real", 1.0), (K.bce_scalar_label, "fake", 1.0), (K.bce_scalar_label, "fake", -1.0)
        elif self.config.gan_obj == "LS":
            self.phi_1, self.phi_2, self.phi_3 = (K.mse_scalar_label, "real", 1.0), (K.mse_scalar_label, "fake", 1.0), (K.mse_scalar_label, "real", 1.0)
        else:       # "W" needs calc_gradient_penalty, which cannot run in the reference either
            raise NotImplementedError("gan_obj='%s' is outside the B200 hot path" % config.gan_obj)
        if dist_ctx is not None:
            dist_ctx.attach(self.G, self.D)

    # ------------------------------------------------------------------ loop
    def _to_device(self, batch, slot=None):
        """Host (or device) batch -> device tensors.  `slot` (0 / 1): copy into the persistent staging buffers of that ring
        slot instead of allocating (the loop's prefetch: a fresh 285 MB allocation per step on the copy stream keeps the
        caching allocator growing for many iterations and costs milliseconds per step until it settles)."""
        ring = None
        if slot is not None:
            ring = self._stage_ring.setdefault(slot, {})

        def put(name, src, dtype=None):
            if ring is None:
                return src.to(self.device, non_blocking=True)
            buf = ring.get(name)
            if buf is None or buf.shape != src.shape or buf.dtype != (dtype or src.dtype):
                buf = ring[name] = torch.empty(src.shape, device=self.device, dtype=dtype or src.dtype)
            buf.copy_(src, non_blocking=True)
            return buf

        in_xy = put("in_xy", batch["in_xy"])
        in_dxdy = put("in_dxdy", batch["in_dxdy"])
        b = in_xy.size(1)
        sub_batches = batch["seq_start_end"] if "seq_start_end" in batch else list(zip(range(b), range(1, b + 1)))
        gt_xy = put("gt_xy", batch["gt_xy"])
        gt_dxdy = put("gt_dxdy", batch["gt_dxdy"])
        img = put("features", batch["features"]) if "features" in batch else None
        if img is None and "image_ids" in batch:
            # crops cut on the device from the resident scene images (mggan_scene_crop) instead of a host-built
            # `features` tensor: 4 bytes per agent cross PCIe instead of 17,424
            if self.scene_images is None:
                raise RuntimeError("the batch carries image_ids but no SceneImageStore is attached "
                                   "(trainer.attach_scene_images(dataset.scene_image_store()))")
            if ring is not None:
                # the loop's prefetch stages the ids only; the iteration cuts the crops on the main stream, straight into
                # the buffer it reads (scene_images.DeferredCrop)
                from mggan.data_utils.scene_images import DeferredCrop
                ids = torch.as_tensor(batch["image_ids"])
                img = DeferredCrop(self.scene_images, put("image_ids", ids if ids.dtype == torch.int32 else ids.to(torch.int32)),
                                   in_xy)
            else:
                img = self.scene_images.crop(batch["image_ids"], in_xy[-1])
        return in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, img

    def attach_scene_images(self, store):
        """Keep a dataset's scene images resident in HBM (mggan/data_utils/scene_images.py)."""
        self.scene_images = store

    def _prepare(self, batch, slot=None):
        """Collated batch (host or device tensors) -> device tensors + NaN loss mask.  The reference always
        builds the mask (abstract_train.py:130-132); when no future is masked it is dropped so the step runs
        without boolean-index gathers.  For host batches the NaN test runs on the host copy (no device sync)."""
        gt_host = batch["gt_xy"]
        has_nan = bool(torch.isnan(gt_host).any()) if not gt_host.is_cuda else None
        in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, img = self._to_device(batch, slot)
        if has_nan is None:
            has_nan = bool(torch.isnan(gt_xy).any())
        loss_mask = None
        if has_nan:
            loss_mask = ~gt_xy.isnan().any(2).any(0)
            gt_dxdy, gt_xy = gt_dxdy[:, loss_mask], gt_xy[:, loss_mask]
        return in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, img, loss_mask

    def _run_prepared(self, prepared, metrics, total_iterations=0, sums=None):
        in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, img, loss_mask = prepared
        if self.dist is not None and self._graph is None:
            self.dist.begin_iteration()
            if sums is not None:
                self.dist.set_sums(sums)           # already exchanged by _run_iteration
            else:
                self.dist.prefetch_sums(self._local_counts(prepared))
        backup = None
        if (total_iterations % self.config.num_gen_steps == 0) or (self.epoch >= self.config.keep_gen_steps):
            # reference abstract_train.py:139-153: num_unrolling_steps + 1 discriminator steps
            for u in range(self.config.num_unrolling_steps + 1):
                self.discriminator_step(in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, metrics, loss_mask, img)
                if u == 0 and self.config.num_unrolling_steps > 0:
                    backup = self.D.state_dict()
        self.generator_step(in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, metrics, loss_mask, img)
        self.net_chooser_step(in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, metrics, loss_mask, img)
        if backup is not None:
            # reference :161-162.  As executed there this restores nothing: state_dict() returns the live tensors, so the
            # "backup" follows the unrolled steps and is copied onto itself (frozen from the unmodified reference in
            # tests/golden/var_mse_unroll.npz: D after the iteration = D after num_unrolling_steps + 1 steps).  Kept as written.
            self.D.load_state_dict(backup)

    @staticmethod
    def _local_counts(prepared):
        n_agents = prepared[0].size(1)
        return {"agents": n_agents, "active": n_agents if prepared[6] is None else int(prepared[6].sum())}

    def train_iteration(self, batch, metrics, total_iterations=0):
        """One D step + G step + PM step on a collated batch (reference loop body :114-168)."""
        self._run_iteration(self._prepare(batch), metrics, total_iterations)

    # ------------------------------------------------------------------ CUDA-graph replay
    def _graph_eligible(self, prepared):
        cfg = self.config
        # data-parallel runs capture the NCCL collectives too (every rank captures and replays in lockstep; the host-side
        # normaliser sums are replayed from the eager iteration, DistContext.freeze)
        return (getattr(cfg, "cuda_graph", True) and prepared[6] is None
                and cfg.gan_obj in ("NS", "MM") and cfg.weighting_target in ("ml", "none")
                and cfg.num_gen_steps == 1 and cfg.num_unrolling_steps == 0)

    def _run_iteration(self, prepared, metrics, total_iterations=0):
        """Eager iteration, or the replay of a captured one when this batch has a structure (scene sizes, no masked
        futures) that was captured before: a structure is captured the second time it is seen (mggan/graph.py) and the
        captures live in an LRU cache of `--graph_cache` entries.
        Data-parallel: the choice is collective.  One host-side exchange per iteration tells every rank whether ALL
        ranks hold a matching graph (replay), whether all could capture now (eager + capture), or not (eager); it also
        carries the global agent counts, and a graph is replayed only under the counts it was captured with (its loss
        normalisers are baked in)."""
        eligible = self._graph_eligible(prepared)
        g = next((x for x in self._graphs if x.matches(prepared)), None) if eligible else None
        key = (tuple(prepared[0].shape), prepared[5] is None, tuple(tuple(s) for s in prepared[4]))
        ready = eligible and (g is not None or self._graph_seen.get(key, 0) >= 1)
        sums = None
        if self.dist is not None:
            from mggan.distributed import host_sum
            local = self._local_counts(prepared)
            v = host_sum([1.0 if g is not None else 0.0, 1.0 if ready else 0.0, local["agents"], local["active"]],
                         self.dist.group)
            sums = {"agents": (float(local["agents"]), v[2]), "active": (float(local["active"]), v[3])}
            world = self.dist.world_size
            if g is not None and not (v[0] == world and g.global_counts == (v[2], v[3])):
                g = None
            ready = v[1] == world
        if g is not None:
            self.graph_hits += 1
            if self._graphs[0] is not g:                     # LRU order
                self._graphs.remove(g)
                self._graphs.insert(0, g)
            for k, v_ in g.run(prepared).items():
                metrics[k].extend(t.clone() for t in v_)           # the graph's own tensors are overwritten by the next replay
            return
        self.graph_misses += 1
        if prepared[5] is not None and not torch.is_tensor(prepared[5]):
            # deferred crops (train_iterations): an eager iteration cuts them now; a replay cuts them into its own input
            prepared = prepared[:5] + (prepared[5].materialize(),) + prepared[6:]
        if eligible:
            if len(self._graph_seen) > 4096:
                self._graph_seen.clear()
            self._graph_seen[key] = self._graph_seen.get(key, 0) + 1
        self._run_prepared(prepared, metrics, total_iterations, sums)       # this batch runs eagerly
        if ready and key not in self._graph_failed:
            from mggan.graph import GraphedIteration
            try:
                new = GraphedIteration(self, prepared, total_iterations)
            except Exception as exc:        # the eager iteration above already ran: keep training, never capture this structure again
                # (data-parallel: a rank without a graph answers "no graph" in the per-iteration exchange, so no rank replays)
                self._graph_failed.add(key)
                torch.cuda.synchronize(self.device)
                print(f"[mggan] CUDA-graph capture failed ({exc!r}); this batch structure keeps running eagerly")
                return
            new.global_counts = (sums["agents"][1], sums["active"][1]) if sums is not None else None
            self._graphs.insert(0, new)
            # Ragged datasets: with the reference-default batch of 2 scenes a shuffled epoch has only a few dozen distinct
            # structures (pairs of scene sizes), so a modest LRU cache turns almost every iteration into a replay.
            del self._graphs[max(1, int(getattr(self.config, "graph_cache", 64))):]

    def train_iterations(self, batches, metrics, total_iterations=0, on_step=None):
        """The reference loop `for batch in loader: <D, G, PM step>` (abstract_train.py:114-168) with the
        host->device copy of batch i+1 issued on a side stream while batch i computes (pinned host buffers
        make the copies asynchronous).  `on_step(i, metrics)` runs after each iteration is enqueued.
        Returns the number of iterations run."""
        copy_stream = getattr(self, "_copy_stream", None)
        if copy_stream is None:
            copy_stream = self._copy_stream = torch.cuda.Stream(self.device)
        main = torch.cuda.current_stream(self.device)

        done = [None, None]                 # per ring slot: main-stream event after the iteration that consumed it

        def stage(batch, slot):
            with torch.cuda.stream(copy_stream):
                if done[slot] is not None:
                    copy_stream.wait_event(done[slot])        # the iteration that read this slot's buffers has finished
                prepared = self._prepare(batch, slot)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            return prepared, ev

        it = iter(batches)
        try:
            nxt = stage(next(it), 0)
        except StopIteration:
            return 0
        n = 0
        while nxt is not None:
            prepared, ev = nxt
            main.wait_event(ev)
            for t in prepared:               # tensors made on the copy stream (NaN-mask gathers) are read on the main one
                if torch.is_tensor(t):
                    t.record_stream(main)
            try:
                nxt = stage(next(it), (n + 1) & 1)
            except StopIteration:
                nxt = None
            self._run_iteration(prepared, metrics, total_iterations + n)
            done[n & 1] = torch.cuda.Event()
            done[n & 1].record(main)
            if on_step is not None:
                on_step(n, metrics)
            n += 1
        return n

    def _loaders(self):
        kw = dict(dataset=self.config.dataset, batch_size=self.config.batch_size, workers=self.config.workers,
                  num_scenes=getattr(self.config, "synthetic_scenes", 64),
                  with_img=getattr(self.config, "scene_dim", 64) > 0, seed=getattr(self.config, "seed", 42),
                  images="resident" if getattr(self.config, "resident_images", False) else "agent")
        return (get_dataloader(phase="train", augment=self.config.augment, shuffle=True, **kw),
                get_dataloader(phase="val", augment=False, shuffle=False, **kw))

    def train(self):
        train_loader, val_loader = self._loaders()
        if getattr(self.config, "resident_images", False) and getattr(self.config, "scene_dim", 64) > 0:
            # one store for both phases: validation image ids follow the training ones
            from mggan.data_utils.scene_images import SceneImageStore
            tr_ds, va_ds = train_loader.dataset, val_loader.dataset
            va_ds.image_id_offset = len(tr_ds)
            self.attach_scene_images(SceneImageStore(tr_ds.scene_image_list() + va_ds.scene_image_list(),
                                                     tr_ds.scaling_small, self.device))
        total_iterations = 0
        track_metric = "val/ADE k=20"
        min_track_metric = math.inf
        for epoch in range(self.config.epochs):
            self.epoch += 1
            self.D.train()
            self.G.train()
            metrics = defaultdict(list)
            total_iterations += self.train_iterations(train_loader, metrics, total_iterations)
            if self.epoch % self.config.val_every == 0:
                self.D.eval()
                self.G.eval()
                with torch.no_grad():
                    m = self.check_accuracy(val_loader, vis=True, prefix="val/", num_k=self.config.top_k_test)
                    for k, v in m.items():
                        metrics[f"val/{k}"].append(v)
                if track_metric in metrics:
                    cur = _mean(metrics[track_metric])
                    if cur < min_track_metric:
                        print(f'Saving best model... "{track_metric}: Before: {min_track_metric}, After: {cur}')
                        min_track_metric = cur
                        self.save(checkpoint_name="checkpoint_best.pth")
            metrics = {k: _mean(v) for k, v in metrics.items()}
            self.writer.log(metrics, epoch)
            if self.epoch % self.config.save_every == 0:
                self.save()
            self.l2_weight *= self.config.l2_decay_rate        # kept: has no effect in the reference either
            self.lr_schedulerD.step()
            self.lr_schedulerG.step()
            self.writer.save()

    # ------------------------------------------------------------------ abstract steps
    @abc.abstractmethod
    def generator_step(self, in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, train_metrics, loss_mask, img=None):
        pass

    @abc.abstractmethod
    def discriminator_step(self, in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, train_metrics, loss_mask, img=None):
        pass

    @abc.abstractmethod
    def net_chooser_step(self, in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, metrics, loss_mask, img):
        pass

    @abc.abstractmethod
    def check_accuracy(self, loader, vis=False, prefix="", num_k=20):
        pass

    @abc.abstractmethod
    def predict(self, in_dxdy, in_xy, sub_batches, img=None, num=20, noise=None):
        pass

    @staticmethod
    @abc.abstractmethod
    def construct_model(config):
        pass

    # ------------------------------------------------------------------ checkpoints
    def save(self, checkpoint_name=None):
        if self.dist is not None and self.dist.rank != 0:
            return
        save_obj = {
            "generator": self.G.state_dict(),
            "discriminator": self.D.state_dict(),
            "gen_opt": self.optimizerG.state_dict(),
            "disc_opt": self.optimizerD.state_dict(),
        }
        if not checkpoint_name:
            checkpoint_name = "checkpoint_{}.pth".format(self.epoch)
        torch.save(save_obj, self.model_save_dir / checkpoint_name)

    @classmethod
    def load(cls, log_path: Path, exp_name: str, version: int, checkpoint):
        version_dir = Path(log_path) / exp_name / "version_{}".format(version)
        checkpoint_dir = version_dir / "checkpoints"
        if checkpoint == "latest":
            epochs = [int(n.stem.split("_")[1]) for n in checkpoint_dir.iterdir() if n.stem.split("_")[1] != "best"]
            checkpoint = max(epochs)
        state_dicts = torch.load(checkpoint_dir / "checkpoint_{}.pth".format(checkpoint), map_location="cpu")
        config = load_hparams_from_tags_csv(version_dir / "meta_tags.csv")
        defaults = get_argparse_defaults(get_parser())
        defaults.update(config)
        # `--gpus` is a STRING flag used for truthiness (config.py:24: the default "0" selects the GPU); meta_tags.csv
        # round-trips it through `convert`, which turns "0" into the falsy int 0 -- read it back as the string it was
        if defaults.get("gpus") is not None and not isinstance(defaults["gpus"], bool):
            defaults["gpus"] = str(defaults["gpus"])
        config = Namespace(**defaults)
        g, d = cls.construct_model(config)
        writer = Experiment(log_path, name=exp_name, version=version)
        m = cls(g, d, config, writer)
        m.G.load_state_dict(state_dicts["generator"], strict=False)
        m.D.load_state_dict(state_dicts["discriminator"], strict=False)
        try:
            m.optimizerD.load_state_dict(state_dicts["disc_opt"])
            m.optimizerG.load_state_dict(state_dicts["gen_opt"])
        except Exception as e:          # best effort, like the reference
            print("Could not restore optimizers.", str(e))
        return m, config

    @classmethod
    def load_from_path(cls, version_path: Path, checkpoint="best"):
        version_path = Path(version_path)
        assert "version" in version_path.stem, "Input path should point to model version directory."
        exp_folder = version_path.parent.parent
        model_name = version_path.parent.name
        version = int(version_path.stem.split("_")[1])
        return cls.load(exp_folder, model_name, version, checkpoint)

    def test(self, num_k=20, batch_size=8, **kwargs):
        loader = get_dataloader(dataset=self.config.dataset, phase="test", augment=False, batch_size=batch_size,
                                workers=self.config.workers, shuffle=False,
                                with_img=getattr(self.config, "scene_dim", 64) > 0)
        return self.check_accuracy(loader, vis=False, num_k=num_k, **kwargs)
