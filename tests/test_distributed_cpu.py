"""CPU (gloo, world_size 2): host-side logic of the data-parallel path — scene sharding, the flat
gradient all-reduce, global loss normalisers.  Also checks, with the oracle, the property the
sharding relies on: the sum over scene shards of (shard gradient of the shard's loss terms with
GLOBAL normalisers) equals the full-batch gradient."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_scenes_partitions_everything():
    from mggan.distributed import shard_batch, shard_scenes
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        sizes = rng.integers(1, 33, size=37).tolist()
        sse, c = [], 0
        for n in sizes:
            sse.append([c, c + n])
            c += n
        seen_scenes, seen_agents = [], []
        for r in range(world):
            lo, hi, a_lo, a_hi, local = shard_scenes(sse, world, r)
            seen_scenes += list(range(lo, hi))
            seen_agents += list(range(a_lo, a_hi))
            assert [e - s for s, e in local] == sizes[lo:hi]
            assert (not local) or local[0][0] == 0
        assert seen_scenes == list(range(len(sizes)))
        assert seen_agents == list(range(c))
    batch = {"in_xy": torch.arange(8 * c * 2.0).reshape(8, c, 2), "features": torch.arange(c * 1.0).reshape(c, 1),
             "seq_start_end": sse}
    parts = [shard_batch(batch, 4, r) for r in range(4)]
    assert torch.equal(torch.cat([p["in_xy"] for p in parts], 1), batch["in_xy"])
    assert torch.equal(torch.cat([p["features"] for p in parts], 0), batch["features"])


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mggan.distributed import DistContext
    ctx = DistContext()
    grads = [torch.full((3, 2), float(rank + 1)), torch.arange(5.0) * (rank + 1)]
    red = ctx.allreduce_grads(grads)
    ok = torch.equal(red[0], torch.full((3, 2), 3.0)) and torch.equal(red[1], torch.arange(5.0) * 3)
    ok = ok and ctx.sum_scalar(10 + rank, device="cpu") == 21.0
    t = ctx.sum_tensor(torch.tensor([rank, 1], dtype=torch.int32))
    ok = ok and t.tolist() == [1, 2]
    # loss normalisers prefetched once per iteration; a value that was not prefetched (or changed) is not answered
    ctx.prefetch_sums({"agents": 100 + rank, "active": 90 - rank})
    ok = ok and ctx.fetched("agents", 100 + rank) == 201.0 and ctx.fetched("active", 90 - rank) == 179.0
    ok = ok and ctx.fetched("agents", 7) is None and ctx.fetched("other", 1) is None
    # capture mode replays the normalisers of the eager iteration that just ran, in order
    ctx.begin_iteration()
    a = ctx.sum_scalar(5 + rank, device="cpu")
    ctx.freeze(True)
    ok = ok and ctx.sum_scalar(5 + rank, device="cpu") == a == 11.0
    ctx.freeze(False)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def _gather_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mggan.distributed import gather_predictions, shard_items
    n_items, agents_per_item = 5, [3, 1, 4, 2, 6]
    full = np.arange(12 * 2 * sum(agents_per_item) * 2, dtype=np.float32).reshape(12, 2, sum(agents_per_item), 2)
    lo, hi = shard_items(n_items, world, rank)
    a_lo, a_hi = sum(agents_per_item[:lo]), sum(agents_per_item[:hi])
    got = gather_predictions(full[:, :, a_lo:a_hi] if hi > lo else None, None, rank, world)
    ret[rank] = bool(np.array_equal(got, full)) if rank == 0 else got is None
    dist.destroy_process_group()


def test_sharded_evaluation_gathers_dataset_order_world2():
    """scripts/evaluate.py under torchrun (cfg-5): item ranges partition the dataset, rank 0 receives the predictions in
    dataset order."""
    from mggan.distributed import shard_items
    for n, w in ((5, 2), (8, 8), (3, 8), (64, 8)):
        cover = []
        for r in range(w):
            lo, hi = shard_items(n, w, r)
            cover += list(range(lo, hi))
        assert cover == list(range(n))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_gloo_allreduce_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_scene_sharded_gradients_sum_to_full_batch_oracle():
    """No-image configuration (BatchNorm is the only cross-scene coupling besides normalisers)."""
    import mggan_oracle as O
    from conftest import load_golden
    from mggan.distributed import shard_batch
    g = load_golden("cfg2_g4_eth_noimg")
    ng, k = g["meta"]["num_gens"], 6
    b = dict(g["batch"])
    b["seq_start_end"] = g["meta"]["seq_start_end"]
    N = b["in_xy"].shape[1]
    gen = torch.Generator().manual_seed(0)
    sse = b["seq_start_end"]
    noise = torch.stack([O.global_noise(8, sse, gen) for _ in range(k)])
    idx = torch.randint(0, ng, (N, k), generator=gen)

    def gstep_grads(batch, noise, idx, n_global, counts):
        tr = O.OracleTrainer(g["G0"], g["D0"], ng, num_samples=k)
        mask, gt_xy, gt_dxdy = tr.loss_mask(batch)
        (rel, ab), _, _ = tr._G(batch, noise, False, k, mask, idx)
        l2 = (ab - gt_xy[:, None]).norm(dim=-1).sum(0)
        min_l2 = sum(l2[:, a:e].sum(1).min() for a, e in batch["seq_start_end"]) / n_global
        out, branch = tr._D(batch, ab, rel, mask)
        w = 1.0 / counts[idx]
        adv = (O._bce(out, torch.full_like(out, 0.95)) * w).sum() / (n_global * k)
        clf = (torch.nn.functional.cross_entropy(branch.flatten(0, 1), idx.reshape(-1), reduction="none")
               .reshape(idx.shape) * w).sum() / (n_global * k)
        return tr._grads(tr.G, min_l2 + adv + clf)

    counts = torch.bincount(idx.flatten(), minlength=ng).float()
    full = gstep_grads(b, noise, idx, N, counts)
    acc = None
    for r in range(2):
        from mggan.distributed import shard_scenes
        _, _, a_lo, a_hi, _ = shard_scenes(sse, 2, r)
        part = shard_batch(b, 2, r)
        gr = gstep_grads(part, noise[:, a_lo:a_hi], idx[a_lo:a_hi], N, counts)
        acc = gr if acc is None else {n: (acc[n] + v if v is not None else acc[n]) for n, v in gr.items()}
    for n, v in full.items():
        if v is None:
            continue
        assert torch.allclose(acc[n], v, rtol=2e-4, atol=1e-7), n
