"""Which host-side lines launch the library (non-mggan) kernels of one training iteration?

    python tools/glue_profile.py [--scenes 512]      # needs a GPU; prints a table to stdout

Uses torch.profiler with stacks on ONE eager iteration (after warm-up) and attributes every CUDA kernel that is
not ours to the innermost frame inside mg-gan_b200/ (autograd-engine launches are attributed to `backward`)."""
import argparse
import collections
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mg-gan_b200"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=512)
    ap.add_argument("--agents", type=int, default=32)
    a = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench
    from collections import defaultdict
    from mggan.logging import Experiment
    from mggan.model.config import get_parser
    from mggan.model.model_factory import construct_model
    from mggan.model.train import PiNetMultiGeneratorGAN
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    cfg = get_parser().parse_args(["--num_gens", "8", "--num_samples", "20", "--cuda_graph", "0"])
    cfg.gpus = True
    G, D = construct_model(cfg)
    tr = PiNetMultiGeneratorGAN(G, D, cfg, Experiment(tempfile.mkdtemp(prefix="mggan_glue_"), "glue", version=0))
    tr.epoch = 1
    tr.G.train(); tr.D.train()
    host = bench.make_inputs(a.scenes, a.agents, seed=4000, pin=False)
    sse = host["seq_start_end"]
    b = {k: v.to(dev) for k, v in host.items() if k != "seq_start_end"}
    prepared = (b["in_xy"], b["in_dxdy"], b["gt_xy"], b["gt_dxdy"], sse, b["features"], None)
    m = defaultdict(list)
    for _ in range(3):
        tr._run_prepared(prepared, m)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        tr._run_prepared(prepared, m)
        torch.cuda.synchronize()
    # correlate kernels with the CPU op that launched them
    by_line = collections.defaultdict(lambda: [0, 0.0, collections.Counter()])
    events = prof.events()
    for ev in events:
        if ev.device_type != torch.autograd.DeviceType.CPU or not ev.kernels:
            continue
        for k in ev.kernels:
            name = k.name
            if "mggan" in name or name.startswith("void (anonymous namespace)") or "<unnamed>" in name:
                continue
            frame = "backward (autograd engine)"
            for fr in (ev.stack or []):
                if "mg-gan_b200" in fr:
                    frame = fr.split("mg-gan_b200/")[-1]
                    break
            rec = by_line[(frame, ev.name)]
            rec[0] += 1
            rec[1] += k.duration
            rec[2][name[:60]] += 1
    rows = sorted(by_line.items(), key=lambda kv: -kv[1][0])
    total = sum(r[0] for _, r in rows)
    print(f"library kernels in one iteration: {total}")
    for (frame, op), (n, us, names) in rows[:60]:
        print(f"{n:4d} launches {us:9.1f} us  {op:28s} {frame}   [{', '.join(f'{k}x{v}' for k, v in names.most_common(2))}]")


if __name__ == "__main__":
    main()
