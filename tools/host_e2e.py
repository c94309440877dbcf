import sys, os, time, tempfile
ROOT="/root/repo"
sys.path.insert(0, os.path.join(ROOT, "mg-gan_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from collections import defaultdict
import bench
from mggan.logging import Experiment
from mggan.model.config import get_parser
from mggan.model.model_factory import construct_model
from mggan.model.train import PiNetMultiGeneratorGAN
from mggan.data_utils.scene_images import SceneImageStore
from mggan.synthetic import SCALING_SMALL, make_scene_image
dev=torch.device("cuda",0); torch.manual_seed(1234)
cfg=get_parser().parse_args(["--num_gens","8","--num_samples","20"]); cfg.gpus=True
G,D=construct_model(cfg)
tr=PiNetMultiGeneratorGAN(G,D,cfg,Experiment(tempfile.mkdtemp(),"h",version=0)); tr.epoch=1; tr.G.train(); tr.D.train()
host=bench.make_inputs(512,32,seed=4000,pin=True)
sse=host["seq_start_end"]
images=[make_scene_image(9000+i) for i in range(512)]
tr.attach_scene_images(SceneImageStore(images, SCALING_SMALL, dev))
ids=np.concatenate([np.full(e-s,i,np.int32) for i,(s,e) in enumerate(sse)])
host_res={k:v for k,v in host.items() if k!="features"}; host_res["image_ids"]=torch.from_numpy(ids).pin_memory()
loss_ring=torch.zeros(64,8).pin_memory()
stamps=[]
def read_back(i,mm):
    keys=sorted(kk for kk in mm if kk.startswith("train/"))
    vals=torch.stack([mm[kk][-1].float().reshape(()) for kk in keys])
    loss_ring[i%64,:vals.numel()].copy_(vals,non_blocking=True); mm.clear()
    stamps.append(time.perf_counter())
m=defaultdict(list)
tr.train_iterations((host_res for _ in range(6)), m, on_step=read_back); torch.cuda.synchronize()
import cProfile, pstats, io
pr=cProfile.Profile(); pr.enable()
tr.train_iterations((host_res for _ in range(30)), m, on_step=read_back)
pr.disable(); torch.cuda.synchronize()
st=io.StringIO(); pstats.Stats(pr,stream=st).sort_stats("cumulative").print_stats(45); print(st.getvalue()[:9000])
for rep in range(1):
    stamps.clear(); torch.cuda.synchronize(); t0=time.perf_counter()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    tr.train_iterations((host_res for _ in range(30)), m, on_step=read_back)
    e1.record(); t_host=time.perf_counter()-t0; torch.cuda.synchronize()
    d=np.diff(np.array(stamps))*1e3
    print(f"rep {rep}: host enqueue {t_host/30*1e3:.2f} ms/iter (median step gap {np.median(d):.2f}, max {d.max():.2f}); device {e0.elapsed_time(e1)/30:.2f} ms/iter")
