"""Library (torch) ops of one eager training iteration grouped by (op, input shapes): which host-side glue launches kernels.

    python tools/glue_shapes.py [--scenes 2]      # needs a GPU
"""
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
    ap.add_argument("--scenes", type=int, default=2)
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
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        tr._run_prepared(prepared, m)
        torch.cuda.synchronize()
    agg = collections.Counter()
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CPU or not ev.kernels:
            continue
        n = sum(1 for k in ev.kernels if "mggan" not in k.name and "<unnamed>" not in k.name and "anonymous namespace)::" not in k.name.split("<")[0])
        if n:
            agg[(ev.name, str(ev.input_shapes)[:110])] += n
    print("library kernels:", sum(agg.values()))
    for (name, shapes), n in agg.most_common(70):
        print(f"{n:4d}  {name:26s} {shapes}")


if __name__ == "__main__":
    main()
