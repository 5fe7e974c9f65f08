"""Per-key deviation of the CUDA Mask3D step from the golden vectors (diagnostics for tests/test_mask3d.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import unscene3d_b200  # noqa
from unscene3d_b200 import engine, models
from golden.make_golden import run_mask3d_case, mask3d_inputs
from oracle import ops_cpu
from scipy.optimize import linear_sum_assignment

gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "mask3d_step.npz")))
matcher = models.HungarianMatcher(cost_class=2.0, cost_mask=5.0, cost_dice=2.0, cost_noise_robust=0.0, num_points=-1)
for rep in range(2):
    res = run_mask3d_case(models, engine, matcher, device="cuda")
    for k, g in gold.items():
        if k.startswith("match"):
            print(rep, k, "ours", res[k].tolist(), "gold", g.tolist())
            continue
        scale = max(float(np.abs(g).max()), 1e-12)
        print(rep, f"{k:60s} err/max {float(np.abs(res[k] - g).max()) / scale:.3e}  max|gold| {scale:.3e}")
    tg = mask3d_inputs()[4]
    for b in range(2):
        cg = np.asarray(ops_cpu.matcher_cost(torch.from_numpy(gold["pred_logits"][b]).float(), torch.from_numpy(gold[f"pred_masks{b}"]).float(),
                                             tg[b]["segment_mask"], tg[b]["labels"], 2.0, 5.0, 2.0), dtype=np.float64)
        co = np.asarray(ops_cpu.matcher_cost(torch.from_numpy(res["pred_logits"][b]).float(), torch.from_numpy(res[f"pred_masks{b}"]).float(),
                                             tg[b]["segment_mask"], tg[b]["labels"], 2.0, 5.0, 2.0), dtype=np.float64)
        print(rep, "scene", b, "cost matrices: max |ours - gold|", np.abs(co - cg).max(), "range of gold", cg.min(), cg.max(),
              "spread across queries per target", (cg.max(0) - cg.min(0)).tolist())
        i, j = linear_sum_assignment(cg); i2, j2 = linear_sum_assignment(co)
        print(rep, "  optimum gold-cost", cg[i, j].sum(), "ours-under-gold", cg[i2, j2].sum(), "ours-under-ours", co[i2, j2].sum(), "gold-under-ours", co[i, j].sum())
        m = res[f"match{b}"]
        print(rep, "  matcher(CUDA) under ours", co[m[0], m[1]].sum(), "under gold", cg[m[0], m[1]].sum())
