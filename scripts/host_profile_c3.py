"""cProfile of the host side of a C3 step (4 x 200k-voxel Mask3D self-training step): where the Python time goes."""
import cProfile
import io
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch

import bench_configs as bc

dev = torch.device("cuda")
net, crit, wd = bc.build_mask3d(dev)
batch = bc.scene_batch(4, 200_000, 100, dev)
for _ in range(4):
    bc.train_step(net, crit, wd, batch, 1)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    bc.train_step(net, crit, wd, batch, 1)
torch.cuda.synchronize()
pr.disable()
for key in ("tottime", "cumulative"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
    print("\n".join(l[:170] for l in s.getvalue().split("\n")[:48]))
