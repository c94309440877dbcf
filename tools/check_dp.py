"""On-hardware data-parallel parity: the parameters after ITERS training iterations on W GPUs
(scenes sharded, NCCL all-reduce of gradients / BatchNorm sums / per-generator counts) equal those of the SAME global batch
on one GPU.  Every random draw (scene noise, generator indices, smoothed labels) is injected, so the only difference between
the two runs is the sharding.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_dp.py

Rank 0 also runs the single-GPU trainer and prints one JSON line {"ok": bool, "max_abs_diff": ..., "worst": name, ...};
exit code 1 on mismatch.  `--graph` additionally lets the trainers capture / replay CUDA graphs (the NCCL-captured path).
"""
import argparse
import json
import os
import sys
import tempfile
from collections import defaultdict

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mg-gan_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--num_gens", type=int, default=4)
    ap.add_argument("--k", type=int, default=6)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--nccl", action="store_true", help="NCCL collectives instead of the peer-memory kernel")
    ap.add_argument("--tol", type=float, default=5e-5,
                    help="parameters after ITERS iterations: 5 %% of one AdamW step (lr 1e-3).  AdamW's first steps are "
                         "lr * g / (|g| + 1e-8): for weight elements whose gradient is ~1e-8 the update depends on the "
                         "gradient's round-off (summation order differs between 1 and W shards), measured 1.3e-5 on one "
                         "decoder tensor; the gradients themselves are compared at 5e-5 relative to each tensor's largest entry (fp32 atomics: two single-GPU runs differ at the 1e-5 level already)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    import mggan.model.modules.standard as S
    import mggan.model.train as T
    from mggan.distributed import DistContext, shard_batch, shard_scenes
    from mggan.logging import Experiment
    from mggan.model.config import get_parser
    from mggan.model.model_factory import construct_model
    from mggan.synthetic import make_batch

    sizes = [3, 1, 5, 2, 4, 6, 2, 3] * max(1, world // 2)  # ragged, including singleton scenes; >= 4 scenes per rank
    b = make_batch(sizes, seed=21, with_img=True)
    sse = b["seq_start_end"]
    N = b["in_xy"].shape[1]
    full = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in b.items()}
    gen = torch.Generator().manual_seed(5)
    rng = np.random.default_rng(6)

    def scene_noise():
        return torch.cat([torch.randn(1, 8, generator=gen).repeat(e - s, 1) for s, e in sse])

    draws = []
    for _ in range(a.iters):
        draws.append(dict(d_noise=scene_noise(), d_idx=torch.randint(0, a.num_gens, (N, 1), generator=gen),
                          g_noise=torch.stack([scene_noise() for _ in range(a.k)]),
                          g_idx=torch.randint(0, a.num_gens, (N, a.k), generator=gen), pm_noise=scene_noise(),
                          labels=[(float(rng.uniform(0.9, 1.0)), float(rng.uniform(0.0, 0.1))) for _ in range(3)]))

    class Inj:
        """Cyclic queues: an iteration makes 3 noise, 3 index and 3 label draws (D step, G step, PM step); a CUDA-graph
        capture pass repeats the noise / index calls of the iteration it models (and bakes those tensors into the graph)."""
        noise, idx, labels = [], [], []
        cur = {"noise": 0, "idx": 0, "labels": 0}

        @classmethod
        def take(cls, what):
            lst = getattr(cls, what)
            v = lst[cls.cur[what] % len(lst)]
            cls.cur[what] += 1
            return v

    def global_noise(dim, sub_batches, noise_type, device=None, num_samples=None):
        z = Inj.take("noise").to(dev)
        return z[None] if (num_samples is not None and z.dim() == 2) else z

    def gan_labels(shape, smoothness=0.1):
        real, fake = Inj.take("labels")
        return torch.zeros(shape) + real, torch.zeros(shape) + fake

    def get_samples(self, enc_h, num_samples=5):
        return self.pm_logits(enc_h), Inj.take("idx").to(dev)

    T.get_global_noise = S.get_global_noise = global_noise
    T.get_gan_labels = gan_labels
    S.MultiGenerator.get_samples = get_samples

    def run(ctx, batch, lo, hi):
        torch.manual_seed(77)
        cfg = get_parser().parse_args(["--num_gens", str(a.num_gens), "--num_samples", str(a.k),
                                       "--cuda_graph", "1" if a.graph else "0"])
        cfg.gpus = True
        G, D = construct_model(cfg)
        tr = T.PiNetMultiGeneratorGAN(G, D, cfg, Experiment(tempfile.mkdtemp(prefix="mggan_dp_"), "dp", version=rank),
                                      dist_ctx=ctx)
        tr.epoch = 1
        tr.G.train(); tr.D.train()
        prepared = tr._prepare(batch)
        for it in range(a.iters):
            d = draws[0 if a.graph else it]         # a captured iteration bakes its injected noise / indices: keep them fixed
            Inj.cur = {"noise": 0, "idx": 0, "labels": 0}
            Inj.noise = [d["d_noise"][lo:hi], d["g_noise"][:, lo:hi], d["pm_noise"][lo:hi]]
            Inj.idx = [d["d_idx"][lo:hi], d["g_idx"][lo:hi], torch.zeros(hi - lo, 1, dtype=torch.long)]
            Inj.labels = list(d["labels"][:2]) + [d["labels"][2]]
            if a.graph:
                tr._run_iteration(prepared, defaultdict(list), it)
            else:
                tr._run_prepared(prepared, defaultdict(list), it)
        torch.cuda.synchronize()
        # gradients left by the last optimiser step of the run (PM-Network step: encoder, scene CNN, social attention,
        # PM-Network); data-parallel ranks hold their shard's part, summed here exactly as the step summed it
        grads = {}
        for n, p in tr.G.named_parameters():
            if p.grad is not None and not n.startswith("G_"):
                g = p.grad.detach().clone()
                if ctx is not None:
                    dist.all_reduce(g)
                grads["Ggrad." + n] = g
        run.grads = grads
        return {("G." + n): p.detach().clone() for n, p in tr.G.named_parameters() if not n.startswith("G_")} | \
               {("D." + n): p.detach().clone() for n, p in tr.D.named_parameters()} | \
               {("Gbuf." + n): p.detach().clone().float() for n, p in tr.G.named_buffers() if not n.startswith("G_")} | \
               {("Dbuf." + n): p.detach().clone().float() for n, p in tr.D.named_buffers()}

    _, _, a_lo, a_hi, _ = shard_scenes(sse, world, rank)
    ctx = DistContext(peer=False if a.nccl else None)
    if not a.nccl and ctx.peer is None:
        print(f"[rank {rank}] peer-memory reducer unavailable ({ctx.peer_error}); NCCL carries the exchanges", file=sys.stderr)
    sharded = run(ctx, shard_batch(full, world, rank), a_lo, a_hi)
    sharded_grads = run.grads
    ok, report = True, {}
    if rank == 0:
        single = run(None, full, 0, N)
        grad_err, grad_worst = 0.0, None
        for n, v in run.grads.items():
            if n.endswith("Conv_1.bias"):
                continue
            e = float((sharded_grads[n] - v).abs().max() / (v.abs().max() + 1e-12))
            if e > grad_err:
                grad_err, grad_worst = e, n
        worst, worst_name, worst_rm = 0.0, None, 0.0
        ranked = sorted(((float((sharded[n] - v).abs().max()) if v.numel() else 0.0, n) for n, v in single.items()), reverse=True)
        for n, v in single.items():
            if n.endswith("Conv_1.bias"):
                continue          # zero true gradient under train-mode BatchNorm: AdamW amplifies round-off to +-lr
            dlt = float((sharded[n] - v).abs().max()) if v.numel() else 0.0
            if n.endswith("running_mean"):
                worst_rm = max(worst_rm, dlt)       # carries that conv-bias drift x momentum: bounded by the lr envelope
                continue
            if dlt > worst:
                worst, worst_name = dlt, n
        ok = worst <= a.tol and worst_rm <= 5e-4 * a.iters and grad_err <= 5e-5 and len(run.grads) >= 20
        report = {"ok": ok, "exchange": "peer-memory kernel" if ctx.peer is not None else "nccl",
                  "peer_calls": ctx.peer.calls if ctx.peer is not None else 0, "world": world, "iters": a.iters, "graph": a.graph, "max_abs_diff": worst, "worst": worst_name, "running_mean_max_abs_diff": worst_rm,
                  "tol": a.tol, "grad_rel_err": grad_err, "grad_worst": grad_worst, "grads_compared": len(run.grads),
                  "top5": ranked[:5], "tensors": len(single), "agents": N, "scenes": len(sse)}
        print(json.dumps(report))
        sys.stdout.flush()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    import threading
    threading.Timer(20.0, lambda: os._exit(0 if flag.item() else 1)).start()
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
