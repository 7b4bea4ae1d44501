"""One training step of the bench workload between cudaProfilerStart/Stop (for ncu
--profile-from-start off).  Usage: python scripts/profile_step.py [batch]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pn2_b200.train_step import Trainer

b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pc, labels, smpw = bench.make_batch(b, 8192, 100)
dev = torch.device("cuda", 0)
d_pc, d_lab, d_w = (torch.as_tensor(x).to(dev) for x in (pc, labels, smpw))
tr = Trainer(bench.HP, bench.NUM_CLASS, device=dev, seed=0, world_size=1, wgrad_sms=0)  # every kernel at its full grid
for _ in range(3):
    tr.step(d_pc, d_lab, d_w)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(d_pc, d_lab, d_w)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
