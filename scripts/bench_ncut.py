"""C4 timing: NCut pseudo-mask extraction on a synthetic 300k-point scene with S ~ 2048 segments (BASELINE configs[3]).

Device path (unscene3d_b200.pseudo_masks: segment means, affinity bit graph, device-resident Lanczos, greedy loop) against
the CPU oracle restatement of the reference functions (oracle/ncut_cpu.py) on the same scene, on this host's cores.
    python scripts/bench_ncut.py [--segments 2048] [--iters 20] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=300_000)
    ap.add_argument("--segments", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--tau", type=float, default=0.6)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-iters", type=int, default=3, help="NCut iterations of the CPU leg (bounded sample)")
    args = ap.parse_args()
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200 import pseudo_masks as pm
    from unscene3d_b200.synthetic import make_ncut_scene as make_c4_scene

    seg, fa, fb, conn = make_c4_scene(args.points, args.segments)
    dev = torch.device("cuda")
    seg_d, fa_d, fb_d, conn_d = (torch.from_numpy(x).to(dev) for x in (seg, fa, fb, conn))

    def run():
        agg_a, uniq = pm.aggregate_features(fa_d, seg_d, conn_d)
        agg_b, _ = pm.aggregate_features(fb_d, seg_d, conn_d)
        return pm.unscene3d((agg_a, agg_b), uniq, conn_d, affinity_tau=args.tau, max_number_of_instances=args.iters, min_segment_size=4)

    masks = run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        masks = run()
    torch.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) / reps * 1e3
    out = {"workload": f"NCut pseudo masks, {args.points} points, S = {args.segments} segments, 384-d + 96-d features, tau = {args.tau}, "
                       f"{args.iters} iterations incl. aggregation", "gpu_ms_per_scene": gpu_ms, "masks": int(masks.shape[0]),
           "mask_sizes": [int(m.sum()) for m in masks][:20]}
    if not args.no_cpu:
        from oracle import ncut_cpu

        seg_t, conn_t = torch.from_numpy(seg), torch.from_numpy(conn)
        t0 = time.perf_counter()
        agg_a, uniq = ncut_cpu.aggregate_features(torch.from_numpy(fa), seg_t, conn_t)
        agg_b, _ = ncut_cpu.aggregate_features(torch.from_numpy(fb), seg_t, conn_t)
        t_agg = time.perf_counter() - t0
        t0 = time.perf_counter()
        res = ncut_cpu.unscene3d(agg_a, agg_b, uniq, conn_t, affinity_tau=args.tau, max_number_of_instances=args.cpu_iters, min_segment_size=4)
        cpu_s = time.perf_counter() - t0
        out["cpu_baseline"] = {"aggregation_seconds": t_agg, "seconds": cpu_s, "iterations": args.cpu_iters, "cores": torch.get_num_threads(), "kind": "port",
                               "seconds_per_scene_extrapolated": t_agg + cpu_s / args.cpu_iters * args.iters,
                               "sample": f"aggregation + {args.cpu_iters} of {args.iters} NCut iterations of oracle/ncut_cpu.py (numpy / scipy eigh restatement of "
                                         "pseudo_masks/unscene3d_pseudo_main.py:89-146, 350-502) on the same scene"}
        out["cpu_masks"] = int(res.shape[0]) if hasattr(res, "shape") else None
    print(json.dumps(out))


if __name__ == "__main__":
    main()
