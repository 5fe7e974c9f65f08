"""Per-iteration diagnostics of the CUDA NCut path on the golden scene, each iteration started from the reference's
painting: bit graph vs oracle matrix, Lanczos vector vs scipy on the same matrix and vs the golden vector."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from scipy.linalg import eigh
import unscene3d_b200  # noqa
from unscene3d_b200 import pseudo_masks as pm
import oracle.ncut_cpu as nc
import test_ncut as T

g, case = T.load_gold()
rec = T._oracle_replay(g, case)
agg = (torch.from_numpy(g["agg_a"]).cuda(), torch.from_numpy(g["agg_b"]).cuda())
for k, r in enumerate(rec):
    painted = torch.from_numpy(r["painted"]).cuda()
    keep = (~painted).float()[:, None]
    graph = pm.get_affinity_matrix((keep * agg[0], keep * agg[1]), tau=0.65, painted=painted)
    info = {}
    vec = pm.second_smallest_eigenvector(graph, info=info).cpu().numpy()
    W = graph.dense().cpu().numpy(); d = graph.degree.cpu().numpy()
    ref_same = nc.fiedler(W, np.diag(d))
    ref_same = ref_same if np.dot(ref_same, vec) >= 0 else -ref_same
    gold = g["eigvecs"][k]; gold = gold if np.dot(gold, vec) >= 0 else -gold
    w = eigh(np.diag(d) - W, np.diag(d), eigvals_only=True, subset_by_index=[0, 3])
    print(f"it {k:2d} painted {int(r['painted'].sum()):3d} vec_ok {bool(r['vec_ok'])!s:5s} vs scipy(same W) {np.abs(vec - ref_same).max() / np.abs(ref_same).max():.2e} "
          f"vs golden {np.abs(vec - gold).max() / np.abs(gold).max():.2e} steps {info['steps']} ritz {['%.6f' % v for v in info['ritz_values']]} "
          f"1-eig {['%.6f' % (1 - v) for v in w]} min beta {min(info['beta']):.1e} last betas {['%.1e' % v for v in info['beta'][-3:]]}", flush=True)
