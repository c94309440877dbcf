"""CUDA-graph replay of one whole training iteration (D step + G step + PM-Network step).

The iteration is ~650 kernel launches (ours + autograd glue); at the reference-default batch (64 agents) it is
bound by their host-side enqueue cost, and at 16k agents the launch gaps still leave the GPU idle ~7 % of the
step.  When consecutive batches have the same structure (same scene sizes, no NaN-masked futures) the trainer
captures the iteration once (`torch.cuda.graph`: forward, backward, clip + AdamW, all on the capture stream) and
replays it on static input buffers.

Everything that changes from one iteration to the next without changing the launch sequence lives in device
memory and is refreshed by the host before each replay (`ScalarFeed`):
  * the smoothed GAN labels (reference utils.py:18-25: numpy draws, made in the same order as the eager path);
  * AdamW's learning rate and per-tensor bias corrections (`MgganTensorTable.dyn`, csrc/optim.cu) -- the host
    keeps the per-tensor step counters exactly as the eager optimiser does;
  * the Philox offset of the PM-Network sampler (`mggan_gumbel_sample(..., dyn_offset, ...)`).
Scene noise comes from torch's CUDA generator, which is graph-safe (its offset advances per replay).
"""
from collections import defaultdict
import math

import numpy as np
import torch

RING = 8


class ScalarFeed:
    """Device-resident per-iteration scalars with a ring of pinned host mirrors (the host may run several
    replays ahead of the device, so one mirror is not enough)."""

    def __init__(self, device, n_float=4096, n_int=8):
        self.device = device
        self.df = torch.zeros(n_float, device=device, dtype=torch.float32)
        self.di = torch.zeros(n_int, device=device, dtype=torch.int64)
        self.hf = [torch.zeros(n_float, dtype=torch.float32).pin_memory() for _ in range(RING)]
        self.hi = [torch.zeros(n_int, dtype=torch.int64).pin_memory() for _ in range(RING)]
        self.events = [None] * RING
        self.slot = 0
        self.nf = 0
        self.f = np.zeros(n_float, dtype=np.float32)       # staging copy the trainer writes into
        self.i = np.zeros(n_int, dtype=np.int64)

    def alloc(self, n):
        off = self.nf
        self.nf += n
        assert self.nf <= self.f.size, "ScalarFeed too small"
        return off

    def upload(self):
        """Copy the staged scalars to the device on the current stream."""
        s = self.slot
        self.slot = (s + 1) % RING
        if self.events[s] is not None:
            self.events[s].synchronize()
        self.hf[s].numpy()[:] = self.f
        self.hi[s].numpy()[:] = self.i
        # A kernel reads the pinned mirror over UVA instead of a cudaMemcpyAsync: the copy engine is busy with the next
        # batch's host->device transfer (hundreds of MB), and a 16 KB memcpy queued behind it would stall the replay.
        from .cuda_ext import TensorTable, call
        tb = TensorTable()
        tb.p[0], tb.g[0], tb.n[0] = self.df.data_ptr(), self.hf[s].data_ptr(), max(self.nf, 1)
        tb.p[1], tb.g[1], tb.n[1] = self.di.data_ptr(), self.hi[s].data_ptr(), 2 * self.di.numel()
        call("mggan_multi_copy", tb, 2)
        ev = torch.cuda.Event()
        ev.record()
        self.events[s] = ev


class GraphedIteration:
    """One captured training iteration of `trainer` for batches shaped like `prepared`."""

    _uid = 0

    def __init__(self, trainer, prepared, total_iterations=0):
        in_xy, in_dxdy, gt_xy, gt_dxdy, sub_batches, img, loss_mask = prepared
        assert loss_mask is None, "batches with NaN-masked futures run eagerly"
        self.tr = trainer
        dev = in_xy.device
        self.sub_batches = sub_batches
        # the captured kernels read the scene CSR arrays of this structure: keep them alive for as long as the graph
        # (kernels.SceneIndex caches them, but that cache is bounded and cleared)
        from . import kernels as K
        self.scene_index = K.SceneIndex.get(sub_batches, dev)
        from .utils import _scene_ids
        self.scene_ids = _scene_ids(sub_batches, dev)       # same for the scene-id gather of the scene-shared noise
        self.static = [t.clone() if torch.is_tensor(t) else t for t in (in_xy, in_dxdy, gt_xy, gt_dxdy, img)]
        self.feed = ScalarFeed(dev)
        self.label_off = self.feed.alloc(8)
        self.label_calls = 0           # get_gan_labels calls made by one iteration (counted during capture)
        self.adam_calls = []           # (float offset, optimizer, param group, [params]) per clip_adamw table
        self.sampler_calls = 0
        self.replays = 0
        self.metrics = defaultdict(list)
        self.total_iterations = total_iterations
        self.global_counts = None      # data-parallel: (sum of agents, sum of unmasked agents) the normalisers were baked with
        GraphedIteration._uid += 1
        self.uid = GraphedIteration._uid

        trainer._graph = self
        for m in (trainer.G, trainer.D):
            m._graph = self
        for opt in (trainer.optimizerG, trainer.optimizerD):
            opt.graph = self
        if trainer.dist is not None:
            trainer.dist.freeze(True)      # host-side normaliser sums: replay the eager iteration that just ran
        try:
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            # every capture keeps its own memory pool: tensors a capture leaves behind (parameter .grad, metrics) are
            # freed by later eager iterations, and a shared pool would hand those blocks to the next capture while this
            # graph still writes them on replay
            with torch.cuda.graph(self.graph, capture_error_mode="relaxed"):
                s = self.static
                trainer._run_prepared((s[0], s[1], s[2], s[3], sub_batches, s[4], None), self.metrics, total_iterations)
        finally:
            if trainer.dist is not None:
                trainer.dist.freeze(False)
            trainer._graph = None
            for m in (trainer.G, trainer.D):
                m._graph = None
            for opt in (trainer.optimizerG, trainer.optimizerD):
                opt.graph = None
        trainer.G.drop_shared()

    # ---- hooks called by the trainer / optimiser / generator while capturing ----------------------------
    def next_labels(self):
        """(real, fake) device scalars of the next get_gan_labels call of the iteration."""
        i = self.label_calls
        self.label_calls += 1
        assert 2 * i + 1 < 8
        v = self.feed.df[self.label_off + 2 * i: self.label_off + 2 * i + 2]
        return v[0], v[1]

    def adam_table(self, optimizer, group, params):
        """Device pointer of the [lr, bc1[64], bc2_sqrt[64]] block of one clip_adamw table (<= 64 tensors)."""
        off = self.feed.alloc(129)
        self.adam_calls.append((off, optimizer, group, list(params)))
        return self.feed.df[off: off + 129]

    def sampler_offset(self):
        self.sampler_calls += 1
        assert self.sampler_calls <= 4, "the replay offset layout reserves 2 bits for the sampler call index"
        return self.feed.di[0:1]

    # ---- replay ----------------------------------------------------------------------------------------
    def matches(self, prepared):
        in_xy, _, _, _, sub_batches, img, loss_mask = prepared
        return (loss_mask is None and in_xy.shape == self.static[0].shape and (img is None) == (self.static[4] is None)
                and len(sub_batches) == len(self.sub_batches)
                and all(tuple(a) == tuple(b) for a, b in zip(sub_batches, self.sub_batches)))

    def _stage_scalars(self):
        from mggan.model import train as T
        f = self.feed.f
        for i in range(self.label_calls):                      # same numpy draws, same order as the eager steps
            real, fake = T._label_scalars(None)
            f[self.label_off + 2 * i], f[self.label_off + 2 * i + 1] = real, fake
        for off, opt, group, params in self.adam_calls:        # AdamW bookkeeping the eager step() does on the host
            b1, b2 = group["betas"]
            f[off] = group["lr"]
            for j, p in enumerate(params):
                st = opt._init_state(p)
                st["step"] += 1
                t = int(st["step"].item())
                f[off + 1 + j] = 1.0 - b1 ** t
                f[off + 65 + j] = math.sqrt(1.0 - b2 ** t)
        self.replays += 1
        # a fresh Philox offset range per replay, disjoint from the eager draws (bit 62) and from other captured
        # structures (uid); the 2 bits under it hold the sampler call's index inside the iteration
        self.feed.i[0] = (1 << 62) | ((self.uid & 0xFFFF) << 46) | ((self.replays & 0xFFFFFF) << 22)

    def run(self, prepared):
        """Copy the batch into the static buffers, refresh the per-iteration scalars, replay.  Returns the metrics
        dict of the captured iteration (its tensors are overwritten by the next replay)."""
        in_xy, in_dxdy, gt_xy, gt_dxdy, _, img, _ = prepared
        for dst, src in zip(self.static, (in_xy, in_dxdy, gt_xy, gt_dxdy, img)):
            if dst is None:
                continue
            if not torch.is_tensor(src):
                src.materialize(out=dst)           # deferred crops: cut from the resident images straight into the input
            elif dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self._stage_scalars()
        self.feed.upload()
        self.graph.replay()
        return self.metrics
