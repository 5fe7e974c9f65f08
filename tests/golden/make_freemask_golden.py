"""Golden vectors for the FreeMask-style pseudo-mask variant, produced by the reference's OWN source.

pseudo_masks/freemask_main.py holds this path as the body of the scene loop of `main()` (hydra entry point, open3d
visualisation calls in between), so it can be neither imported nor called.  The statements of the segment branch are
therefore cut out of the file by their marker comments — source untouched —, wrapped in a one-pass loop (they `continue`
when a scene yields nothing) and executed in a namespace holding the prepared inputs, torch, numpy and the reference's own
`cosine_sim` / `matrix_nms` (utils/freemask_utils.py, utils/pc_utils.py, taken out of their files' syntax trees).
Run in the build container (needs /root/reference):

    python tests/golden/make_freemask_golden.py        # writes tests/golden/freemask_scene.npz
"""
import ast
import os
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
MAIN_FILE = f"{REF}/pseudo_masks/freemask_main.py"


def _functions(path, names):
    tree = ast.parse(open(path).read(), path)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(body) == len(names), f"{path}: layout changed"
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


def _between(lines, first_marker, last_marker, last_occurrence=False):
    a = next(i for i, l in enumerate(lines) if first_marker in l)
    hits = [i for i, l in enumerate(lines) if last_marker in l and i >= a]
    b = hits[-1] if last_occurrence else hits[0]
    return textwrap.dedent("".join(lines[a:b + 1]))


def reference_program():
    """The reference statements of the segment branch as one compiled code object."""
    lines = open(MAIN_FILE).readlines()
    part1 = _between(lines, "# Get queries by averaging over the segment point feats", "connectivity_dict[s_id.item()] = set(")
    part2 = _between(lines, "# Use FPS sampled queries only if requested", "soft_masks = cosine_sim(key_feats, queries)")
    part3 = _between(lines, "# Filter out the zero features (probably from the image projection)",
                     "candidates after NMS maskness threshold filter")
    tail = "soft_masks = soft_masks[keep]\nmaskness = maskness[keep]\nfinished = True\n"
    src = "for _once in (0,):\n" + textwrap.indent(part1 + part2 + part3 + tail, "    ")
    return compile(src, MAIN_FILE + " (scene-loop excerpt)", "exec")


def make_case(side=14, n_objects=9, n_prototypes=4, points_per_segment=14, noise=0.55, seed=3, dim=48):
    """Segments = cells of a side x side floor grid (one of n_objects blobs each, feature = blob prototype + noise; blobs share
    n_prototypes prototypes, so one query activates several non-connected blobs and the separation step has work); every tenth
    point row is zero, three segments have no valid row at all (dropped by the reference), connectivity = 4-neighbourhood,
    directed, with every 7th reverse edge missing."""
    g = torch.Generator().manual_seed(seed)
    gx, gy = torch.meshgrid(torch.arange(side), torch.arange(side), indexing="ij")
    cell = torch.stack([gx.flatten(), gy.flatten()], 1)
    S = cell.shape[0]
    centres = torch.rand(n_objects, 2, generator=g) * side
    label = torch.cdist(cell.float(), centres).argmin(1)
    seg_ids_unique = torch.sort(torch.randperm(5 * S, generator=g)[:S])[0]
    obj_feat = torch.randn(n_prototypes, dim, generator=g)[torch.arange(n_objects) % n_prototypes]  # far-apart blobs look alike
    seg_of_point = torch.arange(S).repeat_interleave(points_per_segment)
    feats = obj_feat[label[seg_of_point]] + noise * torch.randn(len(seg_of_point), dim, generator=g)
    dead_rows = torch.rand(len(seg_of_point), generator=g) < 0.1
    dead_segments = torch.zeros(S, dtype=torch.bool)
    dead_segments[torch.randperm(S, generator=g)[:3]] = True
    dead_rows |= dead_segments[seg_of_point]
    feats[dead_rows] = 0
    xyz = torch.cat([cell[seg_of_point].float() * 10 + torch.rand(len(seg_of_point), 2, generator=g) * 9,
                     torch.rand(len(seg_of_point), 1, generator=g) * 30], 1).floor()
    perm = torch.randperm(len(seg_of_point), generator=g)
    feats, xyz, seg_of_point = feats[perm], xyz[perm], seg_of_point[perm]
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
    coords = torch.cat([torch.zeros(len(xyz), 1), xyz], 1).int()
    return {"keys_F": feats.float(), "matching_segment_ids": seg_ids_unique[seg_of_point].long(), "seg_connectivity": conn.long(),
            "lr_coords": xyz.numpy().astype(np.int32), "coords": coords}


def run_reference(case, cfg):
    cosine_sim, l2_sim = _functions(f"{REF}/utils/freemask_utils.py", ["cosine_sim", "l2_sim"])
    (matrix_nms,) = _functions(f"{REF}/utils/pc_utils.py", ["matrix_nms"])
    config = types.SimpleNamespace(freemask=types.SimpleNamespace(
        use_fps_sampling=False, similarity_metric="cos", hard_mask_threshold=cfg.hard_mask_threshold,
        instance_to_scene_max_ratio=cfg.instance_to_scene_max_ratio, max_instance_num=cfg.max_instance_num,
        nms_maskness_threshold=cfg.nms_maskness_threshold))
    ns = {
        "torch": torch, "np": np, "config": config, "cosine_sim": cosine_sim, "l2_sim": l2_sim, "matrix_nms": matrix_nms,
        "print": lambda *a, **k: None, "scene_name": ["synthetic"], "finished": False,
        "keys": types.SimpleNamespace(F=case["keys_F"]), "matching_segment_ids": case["matching_segment_ids"],
        "unique_segments": case["matching_segment_ids"].unique(), "seg_connectivity": case["seg_connectivity"],
        "segment_ids": [0],  # the file only asks whether it is the empty list (`segment_ids != []`)
        "lr_coords": case["lr_coords"], "coords": case["coords"],
    }
    exec(reference_program(), ns)
    assert ns["finished"], "the reference skipped the scene"
    return ns["soft_masks"], ns["maskness"]


def main():
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import freemask_cpu

    case = make_case()
    soft, maskness = run_reference(case, freemask_cpu.DEFAULTS)
    print(f"reference: {soft.shape[0]} masks over {soft.shape[1]} points, maskness {maskness.numpy().round(3)}")
    np.savez_compressed(os.path.join(HERE, "freemask_scene.npz"), soft_masks=soft.numpy(), maskness=maskness.numpy(),
                        **{f"in_{k}": (v.numpy() if torch.is_tensor(v) else v) for k, v in case.items()})


if __name__ == "__main__":
    main()
