"""Data-parallel execution of the training step: scenes are the unit of sharding.

All cross-agent coupling in MG-GAN is inside a scene (social attention, scene-shared noise, the
scene-level min-L2), so a batch splits into contiguous scene ranges, one per rank, with the
weights replicated.  What must be exchanged for single-GPU-exact results:

  * gradients: ONE all-reduce(sum) of a flat buffer per optimiser step (D, G, PM) over NCCL
    (NVLink 5 / NVSwitch inside a box), issued from `FusedAdamW.step(reduce_fn=...)`; the clip
    norm is then taken on the reduced gradient;
  * loss normalisers: the global agent / sample counts and the per-generator draw counts
    (`sum_scalar`, `sum_tensor`);
  * train-mode BatchNorm statistics of the two scene CNNs: per-channel sums are all-reduced
    between the statistics pass and the normalisation pass (`AttentionGlobal.stat_group`).

The reference has no multi-GPU path at all (`--gpus` is a truthiness string).
"""
import torch
import torch.distributed as dist


def shard_scenes(sub_batches, world_size, rank, balance="agents"):
    """Contiguous scene range of `rank` balanced by agents (or by n^2 pairs).  Returns
    (scene_lo, scene_hi, agent_lo, agent_hi, local seq_start_end re-based to 0)."""
    sizes = [int(e) - int(s) for s, e in sub_batches]
    cost = [n * n for n in sizes] if balance == "pairs" else sizes
    total = sum(cost)
    bounds, acc, nxt = [0], 0, 1
    for i, c in enumerate(cost):
        acc += c
        while nxt < world_size and acc >= total * nxt / world_size:
            bounds.append(i + 1)
            nxt += 1
    while len(bounds) < world_size:
        bounds.append(len(sizes))
    bounds.append(len(sizes))
    lo, hi = bounds[rank], bounds[rank + 1]
    a_lo = int(sub_batches[lo][0]) if lo < len(sizes) else int(sub_batches[-1][1])
    a_hi = int(sub_batches[hi - 1][1]) if hi > lo else a_lo
    local = [[int(s) - a_lo, int(e) - a_lo] for s, e in sub_batches[lo:hi]]
    return lo, hi, a_lo, a_hi, local


def shard_batch(batch, world_size, rank, balance="agents"):
    """Slice a collated batch dict (agents along dim 1, images along dim 0) to this rank's scenes."""
    _, _, a_lo, a_hi, local = shard_scenes(batch["seq_start_end"], world_size, rank, balance)
    out = {"seq_start_end": local}
    for k, v in batch.items():
        if k == "seq_start_end":
            continue
        out[k] = v[a_lo:a_hi] if k == "features" else v[:, a_lo:a_hi]
    return out


def shard_items(n_items, world_size, rank):
    """Contiguous [lo, hi) range of dataset items (scenes) of `rank`: sizes differ by at most one, order preserved."""
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_predictions(preds, group, rank, world_size):
    """Per-rank prediction arrays (pred_len, k, n_rank, 2) (None for an empty shard) -> the dataset-order concatenation
    along the agent axis on rank 0, None elsewhere.  Object gather over the host group: evaluation sets are small (a few
    MB) and the ranks hold different agent counts."""
    import numpy as np
    cpu_group = _host_group(group)
    bucket = [None] * world_size if rank == 0 else None
    dist.gather_object(preds, bucket, dst=dist.get_global_rank(cpu_group, 0) if cpu_group is not None else 0, group=cpu_group)
    if rank != 0:
        return None
    parts = [p for p in bucket if p is not None]
    return np.concatenate(parts, 2) if parts else None


def _host_group(group):
    """The gloo companion of an NCCL group (the group itself when it is not NCCL)."""
    if dist.get_backend(group) != "nccl":
        return group
    key = id(group) if group is not None else None
    cpu = _CPU_GROUPS.get(key)
    if cpu is None:
        ranks = dist.get_process_group_ranks(group if group is not None else dist.group.WORLD)
        cpu = _CPU_GROUPS[key] = dist.new_group(ranks=ranks, backend="gloo")
    return cpu


_CPU_GROUPS = {}


def host_sum(values, group=None):
    """All-reduce(sum) of a few HOST numbers.  Under NCCL this goes through a companion gloo group on CPU tensors: the
    numbers (agent counts, crop counts) are known on the host, and routing them through the GPU costs a stream
    synchronisation per call (and a host<->device copy that queues behind the batch's H2D transfer on the copy engine).
    Collective: every rank of `group` must call it at the same point."""
    t = torch.tensor([float(v) for v in values], dtype=torch.float64)
    dist.all_reduce(t, group=_host_group(group))
    return t.tolist()


class PeerReducer:
    """The exchanges of the data-parallel step as ONE kernel each over NVLink peer memory (csrc/peer_reduce.cu,
    `mggan_peer_allreduce`): every rank maps every peer's symmetric arena (torch.distributed._symmetric_memory), writes
    its operand into its own region, meets the peers at an in-kernel flag barrier and sums all regions in rank order.
    No NCCL call, no staging `torch.cat`, and for gradients the squared norm of the clip comes out of the same launch.

    Regions are bump-allocated per call and the cursor is reset every iteration, so a region is reused one iteration
    later at the earliest (the kernel has no exit barrier; see peer_reduce.cu).  All ranks issue the same call sequence."""
    FLAG_BYTES = 64 * 16 * 4
    _CODES = {torch.float32: 0, torch.float64: 1, torch.int32: 2}

    def __init__(self, group, device, arena_bytes=16 << 20):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.arena = symm.empty(arena_bytes // 4, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.arena, self.group)
        self.arena.zero_()                                   # flags start at 0 on every rank ...
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                             # ... before any rank launches
        self.base = [int(p) for p in self.handle.buffer_ptrs]
        assert len(self.base) == self.world and self.base[self.rank] == self.arena.data_ptr()
        self.bytes, self.off = arena_bytes, self.FLAG_BYTES
        self.calls = 0

    def begin_iteration(self):
        self.off = self.FLAG_BYTES

    def _region(self, nbytes):
        from .cuda_ext import PeerTable
        off = self.off
        self.off = (off + nbytes + 255) & ~255
        if self.off > self.bytes:
            raise RuntimeError(f"peer arena exhausted ({self.off} > {self.bytes} bytes in one iteration)")
        tb = PeerTable()
        for r in range(self.world):
            tb.region[r] = self.base[r] + off
            tb.flags[r] = self.base[r]
        tb.rank, tb.world = self.rank, self.world
        self.calls += 1
        return tb, off

    def allreduce_(self, t):
        """In-place sum of a contiguous float32 / float64 / int32 tensor over the ranks."""
        from .cuda_ext import call, ptr
        assert t.is_contiguous() and t.dtype in self._CODES, (t.dtype, t.is_contiguous())
        if t.numel() == 0:
            return t
        tb, _ = self._region(t.numel() * t.element_size())
        call("mggan_peer_allreduce", tb, self._CODES[t.dtype], ptr(t), t.numel(), ptr(t), None)
        return t

    def allreduce_grads(self, grads, want_sqnorm=True):
        """Gradient tensors -> (views of the reduced flat buffer in the same order, squared L2 norm of the reduced
        gradient as a device double or None).  The gradients are packed straight into this rank's region by one
        pointer-table copy kernel per 64 tensors; the reduce kernel reads the regions of all ranks."""
        from . import kernels as K
        from .cuda_ext import call, ptr
        sizes = [g.numel() for g in grads]
        total = sum(sizes)
        tb, off = self._region(total * 4)
        mine = self.arena[off // 4: off // 4 + total]
        dsts, o = [], 0
        for n in sizes:
            dsts.append(mine[o:o + n])
            o += n
        K.multi_copy(dsts, [g.reshape(-1) for g in grads])
        flat = torch.empty(total, device=mine.device, dtype=torch.float32)
        sq = torch.zeros(1, device=mine.device, dtype=torch.float64) if want_sqnorm else None
        call("mggan_peer_allreduce", tb, 0, None, total, ptr(flat), ptr(sq))
        out, o = [], 0
        for g, n in zip(grads, sizes):
            out.append(flat[o:o + n].view_as(g))
            o += n
        return out, sq


_REDUCERS = {}          # id(process group) -> PeerReducer


def allreduce_tensor(t, group):
    """In-place sum over `group`: the peer-memory kernel when the group has a PeerReducer, else the backend's all_reduce."""
    red = _REDUCERS.get(id(group if group is not None else dist.group.WORLD))
    if red is not None and t.is_cuda and t.dtype in PeerReducer._CODES and t.is_contiguous():
        return red.allreduce_(t)
    dist.all_reduce(t, group=group)
    return t


class DistContext:
    def __init__(self, group=None, peer=None):
        """peer: None = use the peer-memory reducer when the backend is NCCL and symmetric memory can be set up on every
        rank (MGGAN_PEER_REDUCE=0 disables it); False = NCCL collectives only."""
        import os
        assert dist.is_initialized()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        self._log = []          # (local value, global sum) of every sum_scalar call of the current iteration
        self._replay = None     # while an iteration is captured as a CUDA graph: the previous iteration's log
        self.peer, self.peer_error = None, None
        want = peer is not False and os.environ.get("MGGAN_PEER_REDUCE", "1") != "0"
        if want and self.world_size > 1 and dist.get_backend(group) == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            try:
                red = PeerReducer(group, dev)
            except Exception as exc:                    # no fabric / symmetric-memory support: NCCL carries the exchanges
                red, self.peer_error = None, repr(exc)
            ok = torch.tensor([1.0 if red is not None else 0.0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)          # the choice is collective
            if float(ok.item()) == 1.0:
                self.peer = red
                _REDUCERS[id(group if group is not None else dist.group.WORLD)] = red

    def begin_iteration(self):
        self._log = []
        if self.peer is not None:
            self.peer.begin_iteration()

    def freeze(self, on):
        """Capture mode: `sum_scalar` has a host read-back, which cannot run inside a CUDA-graph capture; the captured
        iteration has the structure of the eager one that just ran on every rank, so its normalisers are replayed."""
        self._replay = list(self._log) if on else None

    def attach(self, G, D):
        # Called after the (identically seeded) weight initialisation: from here on every rank draws its own scene noise
        # (torch's CUDA generator) and generator indices (the Gumbel sampler is keyed on torch.initial_seed()), so the
        # shards of one global batch are statistically independent like the scenes of a single-GPU batch.
        if self.world_size > 1 and not getattr(self, "_reseeded", False):
            self._reseeded = True
            torch.manual_seed((torch.initial_seed() + 1000003 * (self.rank + 1)) % (1 << 63))
        for m in (G, D):
            enc = getattr(m, "scene_encoder", None)
            if enc is not None:
                enc.stat_group = self.group if self.group is not None else dist.group.WORLD

    def sum_scalar(self, value, device=None):
        if self._replay is not None:
            local, total = self._replay.pop(0)
            assert local == float(value), "captured iteration diverged from the eager one it was modelled on"
            return total
        total = host_sum([float(value)], self.group)[0]
        self._log.append((float(value), total))
        return total

    def prefetch_sums(self, named):
        """Sum a few host numbers ({name: local value}) over ranks once per iteration; `fetched(name, value)` answers
        from them.  The loss normalisers of an iteration (agents, unmasked agents) are known before its first kernel."""
        if self._replay is not None:
            return
        names = sorted(named)
        vals = [float(named[n]) for n in names]
        self._pref = {n: (v, g) for n, v, g in zip(names, vals, host_sum(vals, self.group))}

    def set_sums(self, table):
        """{name: (local value, global sum)} exchanged by the caller."""
        if self._replay is None:
            self._pref = dict(table)

    def fetched(self, name, value):
        """Global sum of the prefetched number `name`, or None when it was not prefetched with this local value."""
        pref = getattr(self, "_pref", None)
        if pref is None or name not in pref or pref[name][0] != float(value):
            return None
        return pref[name][1]

    def sum_tensor(self, t):
        return allreduce_tensor(t, self.group)

    def _device(self):
        return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else "cpu"

    def allreduce_grads(self, grads, want_sqnorm=True):
        """Sum a list of gradient tensors over ranks with one exchange; returns views of the reduced flat buffer in the
        same order -- or (views, squared norm) from the peer-memory kernel, which produces the clip norm in the same pass."""
        if self.peer is not None:
            return self.peer.allreduce_grads(grads, want_sqnorm)
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, group=self.group)
        out, off = [], 0
        for g in grads:
            n = g.numel()
            out.append(flat[off:off + n].view_as(g))
            off += n
        return out
