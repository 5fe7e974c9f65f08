"""Golden vectors for the pseudo-mask NCut path, produced by the UNMODIFIED reference functions.

pseudo_masks/unscene3d_pseudo_main.py cannot be imported as a module here (hydra, omegaconf, pyviz3d, the dataset
package and a CUDA ray-caster are imported at its top), so the function definitions this path consists of are taken
out of the file's syntax tree — source untouched — and executed in a namespace that holds only numpy, torch and
scipy's eigh.  Run in the build container (needs /root/reference):

    python tests/golden/make_ncut_golden.py        # writes tests/golden/ncut_scene.npz
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_FILE = "/root/reference/pseudo_masks/unscene3d_pseudo_main.py"
WANTED = ("normalize_mat", "get_affinity_matrix", "get_masked_affinity_matrix", "second_smallest_eigenvector",
          "get_salient_areas", "separate_segments", "segment_ids_to_mask", "aggregate_features", "unscene3d")


def reference_functions(trace=None):
    """Namespace with the reference's own function objects; `trace` collects every eigenvector it computes."""
    import torch.nn.functional as F
    from scipy.linalg import eigh

    tree = ast.parse(open(REFERENCE_FILE).read(), REFERENCE_FILE)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    assert len(body) == len(WANTED), "reference file layout changed"
    ns = {"np": np, "torch": torch, "F": F, "eigh": eigh}
    # cosine_sim / l2_sim of the single-modality branch live in utils/freemask_utils.py (which imports open3d, hdbscan, ME)
    utils_file = os.path.join(os.path.dirname(os.path.dirname(REFERENCE_FILE)), "utils", "freemask_utils.py")
    utree = ast.parse(open(utils_file).read(), utils_file)
    ubody = [n for n in utree.body if isinstance(n, ast.FunctionDef) and n.name in ("cosine_sim", "l2_sim")]
    assert len(ubody) == 2, "utils/freemask_utils.py layout changed"
    exec(compile(ast.Module(body=ubody, type_ignores=[]), utils_file, "exec"), ns)
    exec(compile(ast.Module(body=body, type_ignores=[]), REFERENCE_FILE, "exec"), ns)
    if trace is not None:
        inner = ns["second_smallest_eigenvector"]

        def recording(A, D):
            out = inner(A, D)
            trace.append(np.array(out[1]))
            return out

        ns["second_smallest_eigenvector"] = recording
    return types.SimpleNamespace(**ns)


def make_case(n_segments=240, n_objects=8, noise=0.8, seed=7, points_per_segment=12):
    """Segments on a jittered grid; each belongs to one of n_objects blobs (feature = blob centre + noise, two
    modalities); a few segments have no valid point features (filled from neighbours), one point row in ten is zero."""
    g = torch.Generator().manual_seed(seed)
    side = int(np.ceil(np.sqrt(n_segments)))
    gx, gy = torch.meshgrid(torch.arange(side), torch.arange(side), indexing="ij")
    cell = torch.stack([gx.flatten(), gy.flatten()], 1)[:n_segments]
    centres = torch.rand(n_objects, 2, generator=g) * side
    label = torch.cdist(cell.float(), centres).argmin(1)
    seg_ids_unique = torch.sort(torch.randperm(5 * n_segments, generator=g)[:n_segments])[0]
    ca, cb = torch.randn(n_objects, 24, generator=g), torch.randn(n_objects, 40, generator=g)
    segment_ids = seg_ids_unique.repeat_interleave(points_per_segment)
    lab_pts = label.repeat_interleave(points_per_segment)
    fa = ca[lab_pts] + noise * torch.randn(len(lab_pts), 24, generator=g)
    fb = cb[lab_pts] + noise * torch.randn(len(lab_pts), 40, generator=g)
    dead_rows = torch.rand(len(lab_pts), generator=g) < 0.1
    dead_segments = torch.zeros(n_segments, dtype=torch.bool)
    dead_segments[torch.randperm(n_segments, generator=g)[:3]] = True
    dead_rows |= dead_segments.repeat_interleave(points_per_segment)
    fa[dead_rows] = 0
    fb[dead_rows] = 0
    perm = torch.randperm(len(lab_pts), generator=g)
    segment_ids, fa, fb = segment_ids[perm], fa[perm], fb[perm]
    # 4-neighbour connectivity on the grid, listed in both directions except for every 7th edge (the reference
    # treats the list as directed)
    index = {tuple(c.tolist()): i for i, c in enumerate(cell)}
    edges = []
    for (x, y), i in index.items():
        for dx, dy in ((1, 0), (0, 1)):
            j = index.get((x + dx, y + dy))
            if j is not None:
                edges.append((i, j))
                if len(edges) % 7:
                    edges.append((j, i))
    conn = seg_ids_unique[torch.tensor(edges)]
    coords = torch.cat([cell[torch.searchsorted(seg_ids_unique, segment_ids)].float() + torch.rand(len(perm), 2, generator=g),
                        torch.rand(len(perm), 1, generator=g)], 1)
    return dict(segment_ids=segment_ids, feats_a=fa, feats_b=fb, seg_connectivity=conn, coords=coords)


def run_reference(case, tau=0.65, aggregation_mode="mean", separation_mode="max", single=False):
    trace = []
    ref = reference_functions(trace)
    cfg = types.SimpleNamespace(freemask=types.SimpleNamespace(aggregation_mode=aggregation_mode))
    agg_a, uniq = ref.aggregate_features(case["feats_a"], case["segment_ids"], case["seg_connectivity"], cfg)
    agg_b, _ = ref.aggregate_features(case["feats_b"], case["segment_ids"], case["seg_connectivity"], cfg)
    feats = agg_a if single else (agg_a, agg_b)
    A, D = ref.get_affinity_matrix(feats, tau=tau, eps=1e-5)
    masks = ref.unscene3d(feats, uniq, case["seg_connectivity"], case["segment_ids"], case["coords"],
                          torch.zeros_like(case["coords"]), affinity_tau=tau, separation_mode=separation_mode)
    return dict(agg_a=agg_a.numpy(), agg_b=agg_b.numpy(), unique_segments=uniq.numpy(), affinity_on=(A == 1.0),
                degree=np.diag(D).copy(), eigvecs=np.stack(trace), masks=masks)


def main():
    case = make_case()
    out = run_reference(case)
    out.update({"in_" + k: v.numpy() for k, v in case.items()})
    path = os.path.join(HERE, "ncut_scene.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()}, "masks:", out["masks"].sum(1))


if __name__ == "__main__":
    sys.exit(main())
