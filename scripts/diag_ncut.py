"""Per-iteration diagnostics of the CUDA NCut path on the golden scene: bit graph vs oracle matrix, matvec vs dense,
Lanczos vector vs scipy on the same matrix."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from scipy.linalg import eigh
import unscene3d_b200  # noqa
from unscene3d_b200 import pseudo_masks as pm
from unscene3d_b200._lib import lib, check
from unscene3d_b200.engine.coords import _stream
import oracle.ncut_cpu as nc
import test_ncut as T

g, case = T.load_gold()
mats = []
orig = nc.fiedler
nc.fiedler = lambda A, D: (mats.append((A.copy(), D.copy())), orig(A, D))[1]
nc.unscene3d(torch.from_numpy(g["agg_a"]), torch.from_numpy(g["agg_b"]), torch.from_numpy(g["unique_segments"]), case["seg_connectivity"],
             affinity_tau=0.65, sign_hook=T.follow(g["eigvecs"]))
nc.fiedler = orig

it_state = {"k": 0}
orig_ssev = pm.second_smallest_eigenvector


def spy(graph, **kw):
    k = it_state["k"]; it_state["k"] += 1
    W = graph.dense().cpu().numpy()
    A, D = mats[k]
    d = graph.degree.cpu().numpy()
    S = graph.n
    x = torch.randn(S, dtype=torch.float64, device="cuda")
    y = torch.empty(S, dtype=torch.float64, device="cuda")
    xs = x.sum().reshape(1)
    check(lib.us3d_ncut_matvec(graph.bits.data_ptr(), S, float(graph.eps), x.contiguous().data_ptr(), xs.data_ptr(), y.data_ptr(), _stream()))
    mv_err = float((y.cpu().numpy() - W @ x.cpu().numpy()).__abs__().max())
    v = orig_ssev(graph, **kw)
    vv = v.cpu().numpy()
    ref = nc.fiedler(W, np.diag(d))
    ref = ref if np.dot(ref, vv) >= 0 else -ref
    w = eigh(np.diag(d) - W, np.diag(d), eigvals_only=True, subset_by_index=[0, 3])
    print(f"it {k:2d}: bits differing from oracle {(W != A).sum():4d}  degree err {np.abs(d - np.diag(D)).max():.2e}  matvec err {mv_err:.2e}  "
          f"lanczos vs scipy(same W) {np.abs(vv - ref).max() / np.abs(ref).max():.2e}  eigvals {w}", flush=True)
    return v


pm.second_smallest_eigenvector = spy
agg = (torch.from_numpy(g["agg_a"]).cuda(), torch.from_numpy(g["agg_b"]).cuda())
uniq = torch.from_numpy(g["unique_segments"]).cuda()
pm.unscene3d(agg, uniq, case["seg_connectivity"].cuda(), affinity_tau=0.65, sign_rule=T.follow(g["eigvecs"]))
