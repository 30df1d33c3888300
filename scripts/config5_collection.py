"""Config 5 (10M x 768 streaming upserts + batch-64 queries) through B200Collection alone, with a host-side profile of the
step: what the id table, validation and result hand-back cost next to the device's share.
    python scripts/config5_collection.py [rows] [dim]"""
import cProfile
import pstats
import sys
import time

import torch

import bench

rows0 = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 768
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
out = bench.leg_config5_collection(dev, rows0, dim, 20, 64, 8192, 1.0)
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in out.items()})
pr = cProfile.Profile()
pr.enable()
out = bench.leg_config5_collection(dev, rows0 // 10, dim, 20, 64, 8192, 1.0)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
