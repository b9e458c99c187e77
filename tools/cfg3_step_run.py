#!/usr/bin/env python
"""A few cfg3 training steps (full-resolution KITTI-Raw frames) -- the command ncu wraps for the launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import bench_workloads as bw
from bench import peaks
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
r = bw.cfg3_train_step(dev, peaks(), steps=int(sys.argv[1]) if len(sys.argv) > 1 else 3, with_reference=False)
print({k: v for k, v in r.items() if k not in ("what", "roofline")})
