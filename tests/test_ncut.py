"""Pseudo-mask NCut path against golden vectors that the UNMODIFIED reference functions of
pseudo_masks/unscene3d_pseudo_main.py produced (tests/golden/make_ncut_golden.py).

CPU: the oracle restatement (oracle/ncut_cpu.py) reproduces the reference's segment features, thresholded affinity,
degrees, every eigenvector (up to LAPACK's arbitrary sign) and the final masks; where /root/reference exists the
oracle is also compared with the reference functions on further random scenes.
GPU: the CUDA path stage by stage — segment means, affinity bits + degrees (bit-exact away from entries that sit within
float round-off of the threshold), Lanczos eigenvector against a dense LAPACK solve of the same graph (1e-7, fp64),
and the complete greedy extraction (identical masks).
"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from golden import make_ncut_golden as gen  # noqa: E402

GOLD = os.path.join(HERE, "golden", "ncut_scene.npz")


def follow(trace):
    """sign_hook that orients eigenvector number k like trace[k] (the reference's LAPACK sign)."""
    state = {"k": 0}

    def hook(vec):
        ref = trace[state["k"]]
        state["k"] += 1
        return 1.0 if float(np.dot(vec, ref)) >= 0 else -1.0

    return hook


def load_gold():
    g = dict(np.load(GOLD))
    case = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("in_")}
    return g, case


def test_oracle_reproduces_reference_golden():
    from oracle import ncut_cpu

    g, case = load_gold()
    agg_a, uniq = ncut_cpu.aggregate_features(case["feats_a"], case["segment_ids"], case["seg_connectivity"])
    agg_b, _ = ncut_cpu.aggregate_features(case["feats_b"], case["segment_ids"], case["seg_connectivity"])
    assert np.array_equal(uniq.numpy(), g["unique_segments"])
    assert np.allclose(agg_a.numpy(), g["agg_a"], atol=1e-6) and np.allclose(agg_b.numpy(), g["agg_b"], atol=1e-6)
    A, D = ncut_cpu.affinity(torch.from_numpy(g["agg_a"]), torch.from_numpy(g["agg_b"]), 0.65)
    assert np.array_equal(A == 1.0, g["affinity_on"])
    assert np.allclose(np.diag(D), g["degree"], rtol=1e-12)
    trace = []
    masks = ncut_cpu.unscene3d(torch.from_numpy(g["agg_a"]), torch.from_numpy(g["agg_b"]), uniq, case["seg_connectivity"],
                               affinity_tau=0.65, sign_hook=follow(g["eigvecs"]), trace=trace)
    assert len(trace) == len(g["eigvecs"])
    for k, (v, r) in enumerate(zip(trace, g["eigvecs"])):
        assert np.abs(v - r).max() < 1e-8 * np.abs(r).max(), f"eigenvector {k}"
    assert np.array_equal(masks, g["masks"])


@pytest.mark.skipif(not os.path.exists(gen.REFERENCE_FILE), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("n,k,noise,seed", [(90, 4, 0.5, 1), (200, 8, 0.3, 5), (150, 5, 1.0, 9)])
def test_oracle_matches_reference_functions(n, k, noise, seed):
    from oracle import ncut_cpu

    case = gen.make_case(n, k, noise, seed)
    ref = gen.run_reference(case)
    agg_a, uniq = ncut_cpu.aggregate_features(case["feats_a"], case["segment_ids"], case["seg_connectivity"])
    agg_b, _ = ncut_cpu.aggregate_features(case["feats_b"], case["segment_ids"], case["seg_connectivity"])
    assert np.allclose(agg_a.numpy(), ref["agg_a"], atol=1e-6)
    masks = ncut_cpu.unscene3d(agg_a, agg_b, uniq, case["seg_connectivity"], affinity_tau=0.65, sign_hook=follow(ref["eigvecs"]))
    assert np.array_equal(masks, ref["masks"])


@pytest.mark.skipif(not os.path.exists(gen.REFERENCE_FILE), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("aggregation_mode,separation_mode,single,tau", [("max", "max", False, 0.65), ("mean", "avg", False, 0.65),
                                                                         ("mean", "largest", False, 0.65), ("mean", "all", False, 0.65),
                                                                         ("mean", "max", True, 0.55), ("max", "largest", True, 0.55)])
def test_oracle_matches_reference_functions_in_the_other_modes(aggregation_mode, separation_mode, single, tau):
    """aggregation_mode 'max' (:366), separation modes avg / largest / all (:234-250) and the single-modality affinity
    (:92-98, asymmetric cosine_sim, column-sum degrees, lower triangle in the eigen-solve) against the reference functions."""
    from oracle import ncut_cpu

    case = gen.make_case(160, 6, 0.4, 11)
    ref = gen.run_reference(case, tau=tau, aggregation_mode=aggregation_mode, separation_mode=separation_mode, single=single)
    agg_a, uniq = ncut_cpu.aggregate_features(case["feats_a"], case["segment_ids"], case["seg_connectivity"], mode=aggregation_mode)
    agg_b, _ = ncut_cpu.aggregate_features(case["feats_b"], case["segment_ids"], case["seg_connectivity"], mode=aggregation_mode)
    assert np.allclose(agg_a.numpy(), ref["agg_a"], atol=1e-6)
    if single:
        A, D = ncut_cpu.affinity_single(agg_a, tau)
        assert np.array_equal(A == 1.0, ref["affinity_on"]) and np.allclose(np.diag(D), ref["degree"], rtol=1e-12)
    masks = ncut_cpu.unscene3d(agg_a, None if single else agg_b, uniq, case["seg_connectivity"], affinity_tau=tau,
                               sign_hook=follow(ref["eigvecs"]), separation_mode=separation_mode)
    assert masks.shape[0] >= 1 and np.array_equal(masks, ref["masks"])


@pytest.mark.gpu
@pytest.mark.parametrize("aggregation_mode,separation_mode,single,tau", [("max", "max", False, 0.65), ("mean", "avg", False, 0.65),
                                                                         ("mean", "largest", False, 0.65), ("mean", "all", False, 0.65),
                                                                         ("mean", "max", True, 0.55), ("max", "largest", True, 0.55)])
def test_cuda_other_modes_match_the_oracle(aggregation_mode, separation_mode, single, tau):
    from oracle import ncut_cpu
    from unscene3d_b200 import pseudo_masks as pm

    case = gen.make_case(160, 6, 0.4, 11)
    agg_a, uniq = ncut_cpu.aggregate_features(case["feats_a"], case["segment_ids"], case["seg_connectivity"], mode=aggregation_mode)
    agg_b, _ = ncut_cpu.aggregate_features(case["feats_b"], case["segment_ids"], case["seg_connectivity"], mode=aggregation_mode)
    ga, gu = pm.aggregate_features(case["feats_a"].cuda(), case["segment_ids"].cuda(), case["seg_connectivity"].cuda(), aggregation_mode)
    gb, _ = pm.aggregate_features(case["feats_b"].cuda(), case["segment_ids"].cuda(), case["seg_connectivity"].cuda(), aggregation_mode)
    assert np.array_equal(gu.cpu().numpy(), uniq.numpy())
    assert np.abs(ga.cpu().numpy() - agg_a.numpy()).max() < 5e-7
    if single:
        A, D = ncut_cpu.affinity_single(agg_a, tau)
        graph = pm.get_affinity_matrix(agg_a.cuda(), tau=tau)
        low = np.tril(A == 1.0)
        assert np.array_equal(graph.dense().cpu().numpy() == 1.0, low | low.T)
        assert np.allclose(graph.degree.cpu().numpy(), np.diag(D), rtol=1e-12)
    trace = []
    # once segments are painted the single-modality pencil has repeated eigenvalues and LAPACK's / Lanczos' member of the
    # eigenspace is arbitrary (see _oracle_replay): the free-running comparison covers the iterations before that
    n_inst = 1 if single else 20
    want = ncut_cpu.unscene3d(agg_a, None if single else agg_b, uniq, case["seg_connectivity"], affinity_tau=tau, trace=trace,
                              separation_mode=separation_mode, max_number_of_instances=n_inst)
    got = pm.unscene3d(agg_a.cuda() if single else (agg_a.cuda(), agg_b.cuda()), uniq.cuda(), case["seg_connectivity"].cuda(),
                       affinity_tau=tau, separation_mode=separation_mode, sign_rule=follow(trace), max_number_of_instances=n_inst)
    assert want.shape[0] >= 1 and np.array_equal(got, want)


def test_host_blob_growing_equals_the_oracle_on_random_graphs():
    """unscene3d_b200.pseudo_masks.separate_segments (libus3d host function, no GPU involved) against the literal restatement,
    all four modes, directed adjacency with ids that are not segments of the scene."""
    import unscene3d_b200  # noqa: F401
    from oracle import ncut_cpu
    from unscene3d_b200 import pseudo_masks as pm

    rng = np.random.default_rng(0)
    for trial in range(120):
        S = int(rng.integers(5, 60))
        ids = np.sort(rng.choice(500, S, replace=False))
        E = int(rng.integers(S, 4 * S))
        conn = np.stack([rng.choice(ids, E), rng.choice(np.concatenate([ids, [777]]), E)], 1)
        bip = rng.random(S) < 0.6
        if not bip.any():
            continue
        vec = rng.normal(size=S)
        vec[~bip] -= 10
        for mode in ("max", "avg", "largest", "all"):
            got = pm.separate_segments(bip, vec, torch.from_numpy(ids), torch.from_numpy(conn), mode)
            want = ncut_cpu.separate_segments(bip, vec, torch.from_numpy(ids), torch.from_numpy(conn), mode)
            assert got == set(int(x) for x in want), (trial, mode)


def test_blob_growing_keeps_reference_quirks():
    """Directed neighbour lists and the skipped blob after a merge (reference :207-224) change the result; the literal
    restatement must keep both."""
    from oracle import ncut_cpu

    ids = torch.arange(6)
    # 0-1 and 2-3 are blobs; 4 bridges both; 5 lists only 3 (directed: 3 does not list 5)
    conn = torch.tensor([[1, 0], [3, 2], [4, 0], [4, 2], [5, 3]])
    vec = np.array([0.0, 0.1, 0.2, 0.3, 0.9, 0.4])
    blob = ncut_cpu.separate_segments_max(np.ones(6, dtype=bool), vec, ids, conn)
    assert blob == {0, 1, 2, 3, 4, 5}
    # without the bridge the seed's blob is only what its own list reaches
    blob = ncut_cpu.separate_segments_max(np.array([1, 1, 1, 1, 0, 1], dtype=bool), np.array([0, 0, 0, 0, 0, 1.0]), ids, conn)
    assert blob == {2, 3, 5}


# ------------------------------------------------------------------------------------------------ GPU


def _oracle_soft_affinity(fa, fb):
    """The averaged normalised affinity before thresholding (float32, as the reference computes it)."""
    import torch.nn.functional as F

    from oracle import ncut_cpu

    fa, fb = F.normalize(fa, p=2, dim=-1), F.normalize(fb, p=2, dim=-1)
    return (ncut_cpu.normalize_mat((fa @ fa.T).numpy()) + ncut_cpu.normalize_mat((fb @ fb.T).numpy())) / 2


@pytest.mark.gpu
def test_cuda_segment_features_match_golden():
    from unscene3d_b200 import pseudo_masks as pm

    g, case = load_gold()
    for key in ("a", "b"):
        agg, uniq = pm.aggregate_features(case["feats_" + key].cuda(), case["segment_ids"].cuda(), case["seg_connectivity"].cuda())
        assert np.array_equal(uniq.cpu().numpy(), g["unique_segments"])
        assert np.abs(agg.cpu().numpy() - g["agg_" + key]).max() < 5e-7  # fp64 sums rounded once vs torch's fp32 mean


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["golden", "random1500"])
def test_cuda_affinity_bits_and_degrees(source):
    from oracle import ncut_cpu
    from unscene3d_b200 import pseudo_masks as pm

    if source == "golden":
        g, _ = load_gold()
        fa, fb = torch.from_numpy(g["agg_a"]), torch.from_numpy(g["agg_b"])
    else:
        gt = torch.Generator().manual_seed(0)
        lab = torch.randint(0, 20, (1500,), generator=gt)
        fa = torch.randn(20, 32, generator=gt)[lab] + 0.8 * torch.randn(1500, 32, generator=gt)
        fb = torch.randn(20, 96, generator=gt)[lab] + 0.8 * torch.randn(1500, 96, generator=gt)
    tau, eps = 0.65, 1e-5
    A, D = ncut_cpu.affinity(fa, fb, tau, eps)
    graph = pm.get_affinity_matrix((fa.cuda(), fb.cuda()), tau=tau, eps=eps)
    W = graph.dense().cpu().numpy()
    soft = _oracle_soft_affinity(fa, fb)
    differs = W != A
    assert not differs[np.abs(soft - tau) > 2e-6].any(), "affinity bit differs away from the threshold"
    assert differs.sum() <= 4
    if not differs.any():
        assert np.allclose(graph.degree.cpu().numpy(), np.diag(D), rtol=1e-12)
    # painted rows / columns drop to eps, degrees stay those of the unpainted graph
    painted = torch.zeros(fa.shape[0], dtype=torch.bool)
    painted[::7] = True
    keep = (~painted).float()[:, None]
    g2 = pm.get_affinity_matrix(((keep * fa).cuda(), (keep * fb).cuda()), tau=tau, eps=eps, painted=painted.cuda())
    A2, D2 = ncut_cpu.affinity(keep * fa, keep * fb, tau, eps)
    deg_ok = np.allclose(g2.degree.cpu().numpy(), np.diag(D2), rtol=1e-12)
    A2[painted.numpy()] = eps
    A2[:, painted.numpy()] = eps
    d2 = g2.dense().cpu().numpy() != A2
    assert d2.sum() <= 4 and (deg_ok or d2.any())


@pytest.mark.gpu
@pytest.mark.parametrize("S,K,noise", [(240, 8, 0.8), (1500, 20, 0.8), (700, 3, 0.05), (2048, 24, 0.8), (3001, 40, 0.6)])
def test_cuda_lanczos_eigenvector_matches_dense_solver(S, K, noise):
    from oracle import ncut_cpu
    from unscene3d_b200 import pseudo_masks as pm

    gt = torch.Generator().manual_seed(S)
    lab = torch.randint(0, K, (S,), generator=gt)
    fa = torch.randn(K, 32, generator=gt)[lab] + noise * torch.randn(S, 32, generator=gt)
    fb = torch.randn(K, 96, generator=gt)[lab] + noise * torch.randn(S, 96, generator=gt)
    graph = pm.get_affinity_matrix((fa.cuda(), fb.cuda()), tau=0.65)
    v = pm.second_smallest_eigenvector(graph).cpu().numpy()
    ref = ncut_cpu.fiedler(graph.dense().cpu().numpy(), np.diag(graph.degree.cpu().numpy()))
    v = v if np.dot(v, ref) >= 0 else -v
    assert np.abs(v - ref).max() < 1e-7 * np.abs(ref).max()


@pytest.mark.gpu
@pytest.mark.parametrize("S,painted_frac", [(240, 0.0), (777, 0.0), (1500, 0.9), (2048, 0.0)])
def test_device_resident_lanczos_equals_the_host_driven_recurrence(S, painted_frac):
    """us3d_ncut_lanczos (all steps in one cooperative launch) against the same recurrence driven step by step from the host
    over us3d_ncut_matvec: recurrence coefficients to 1e-9 while the Krylov space lasts, eigenvector to 1e-8, the same
    breakdown step on a mostly painted graph (tiny Krylov space), and one launch for the first 512 steps."""
    from unscene3d_b200 import pseudo_masks as pm

    gt = torch.Generator().manual_seed(S + 1)
    K = 12
    lab = torch.randint(0, K, (S,), generator=gt)
    fa = torch.randn(K, 32, generator=gt)[lab] + 0.7 * torch.randn(S, 32, generator=gt)
    fb = torch.randn(K, 96, generator=gt)[lab] + 0.7 * torch.randn(S, 96, generator=gt)
    painted = None
    if painted_frac:
        painted = torch.rand(S, generator=gt) < painted_frac
        keep = (~painted).float()[:, None]
        fa, fb = keep * fa, keep * fb
    graph = pm.get_affinity_matrix((fa.cuda(), fb.cuda()), tau=0.65, painted=None if painted is None else painted.cuda())
    info_d, info_h = {}, {}
    v_d = pm.second_smallest_eigenvector(graph, info=info_d)
    pm.set_fused_lanczos(False)
    try:
        v_h = pm.second_smallest_eigenvector(graph, info=info_h)
    finally:
        pm.set_fused_lanczos(True)
    # a breakdown (beta < 1e-10) is a rounding-level event: the two summation orders may cross the threshold one step apart
    assert abs(info_d["steps"] - info_h["steps"]) <= 2 or min(info_d["steps"], info_h["steps"]) >= 512, (info_d["steps"], info_h["steps"])
    # the first coefficients agree to rounding; later ones amplify the rounding differences of the two summation orders
    # (the Lanczos recurrence is not backward stable in its coefficients, only in the Ritz pairs)
    n = min(info_d["steps"], info_h["steps"], 12) - 1
    bd, bh = np.asarray(info_d["beta"][:n]), np.asarray(info_h["beta"][:n])
    assert np.abs(bd - bh).max() < 1e-9 * max(np.abs(bh).max(), 1e-300), (bd, bh)
    assert abs(info_d["ritz_values"][-1] - info_h["ritz_values"][-1]) < 1e-11
    v_d, v_h = v_d.cpu().numpy(), v_h.cpu().numpy()
    v_d = v_d if np.dot(v_d, v_h) >= 0 else -v_d
    # a degenerate top eigenvalue (mostly painted graph) leaves the eigenvector arbitrary within its eigenspace
    if painted is None:
        assert np.abs(v_d - v_h).max() < 1e-8 * np.abs(v_h).max()
    assert info_d["launches"] <= 1 + max(0, (info_d["steps"] - 512 + 31) // 32)


def _oracle_replay(g, case, tau=0.65, margin=2e-6, gap=1e-9):
    """Replays the oracle on the reference's segment features and records, per NCut iteration, the painting that goes in,
    the oracle's eigenvector and extracted part, and which of these decisions are WELL DEFINED (the replay runs on the
    host the test runs on: where a blob seed is decided by LAPACK's rounding noise it may paint differently from the
    machine that wrote the golden file, which is why the device path is compared with this replay, iteration by
    iteration, and the oracle itself is pinned to the golden file in test_oracle_reproduces_reference_golden):
      * vec_ok   the Fiedler value is a simple eigenvalue and no soft affinity lies within `margin` of the threshold (once
                 most segments are painted the pencil (D - A, D) has repeated eigenvalues and LAPACK returns an arbitrary
                 member of the eigenspace);
      * part_ok  additionally the foreground test v > mean(v) has no entry within 1e-9 of the mean, and every segment whose
                 value ties with max(v) to 1e-9 lies in the extracted blob (segments of one cluster share their value to the
                 last bits, so argmax(v) — the blob seed, :236 — is otherwise decided by LAPACK's rounding noise)."""
    from oracle import ncut_cpu
    from scipy.linalg import eigh

    rec, state = [], {}
    orig_aff, orig_fied, orig_sep = ncut_cpu.affinity, ncut_cpu.fiedler, ncut_cpu.separate_segments_max
    ids = g["unique_segments"]

    def spy_aff(a, b, t, eps=1e-5):
        state["painted"] = (a.abs().sum(1) == 0).numpy() & (b.abs().sum(1) == 0).numpy()
        state["border"] = float(np.abs(_oracle_soft_affinity(a, b) - t).min()) < margin
        return orig_aff(a, b, t, eps)

    def spy_fied(A, D):
        w = eigh(D - A, D, eigvals_only=True, subset_by_index=[0, 3])
        state["vec_ok"] = (not state["border"]) and (w[1] - w[0]) > gap and (w[2] - w[1]) > gap * max(1.0, abs(w[1]) / 1e-4)
        state["vec"] = orig_fied(A, D)
        return state["vec"]

    def spy_sep(bip, vec, u, c):
        part = orig_sep(bip, vec, u, c)
        mean = vec.sum() / len(vec)
        tied = ids[vec >= vec.max() - 1e-9 * np.abs(vec).max()]
        part_ok = state["vec_ok"] and np.abs(vec - mean).min() > 1e-9 * np.abs(vec).max() and set(tied.tolist()) <= set(part)
        rec.append({"painted": state["painted"].copy(), "vec_ok": state["vec_ok"], "part_ok": bool(part_ok), "part": set(part),
                    "vec": state["vec"].copy()})
        return part

    ncut_cpu.affinity, ncut_cpu.fiedler, ncut_cpu.separate_segments_max = spy_aff, spy_fied, spy_sep
    try:
        ncut_cpu.unscene3d(torch.from_numpy(g["agg_a"]), torch.from_numpy(g["agg_b"]), torch.from_numpy(ids),
                           case["seg_connectivity"], affinity_tau=tau, sign_hook=follow(g["eigvecs"]))
    finally:
        ncut_cpu.affinity, ncut_cpu.fiedler, ncut_cpu.separate_segments_max = orig_aff, orig_fied, orig_sep
    return rec


def _check_against_golden(pm, agg, g, case, tau=0.65, max_extent_ratio=0.8):
    """Every NCut iteration of the golden scene, started from the painting the reference had at that point: affinity bits
    -> fp64 Lanczos Fiedler vector (1e-6 where the reference's vector is well defined) -> foreground -> blob (identical
    set where the reference's seed is well defined)."""
    rec = _oracle_replay(g, case, tau)
    assert len(rec) == len(g["eigvecs"]) and sum(r["vec_ok"] for r in rec) >= 8 and sum(r["part_ok"] for r in rec) >= 2
    uniq = torch.from_numpy(g["unique_segments"]).cuda()
    conn = case["seg_connectivity"].cuda()
    for k, r in enumerate(rec):
        painted = torch.from_numpy(r["painted"]).cuda()
        keep = (~painted).float()[:, None]
        graph = pm.get_affinity_matrix((keep * agg[0], keep * agg[1]), tau=tau, painted=painted)
        vec = pm.second_smallest_eigenvector(graph).cpu().numpy()
        ref = r["vec"]
        vec = vec if np.dot(vec, ref) >= 0 else -vec
        ref = ref * (1.0 if np.dot(ref, g["eigvecs"][k]) >= 0 else -1.0)  # orientation the replay used (sign_hook)
        vec = vec if np.dot(vec, ref) >= 0 else -vec
        if r["vec_ok"]:
            assert np.abs(vec - ref).max() < 1e-6 * np.abs(ref).max(), f"eigenvector {k}"
        if r["part_ok"]:
            bip = vec > vec.sum() / len(vec)
            if bip.sum() / len(bip) > max_extent_ratio:
                bip, vec = np.logical_not(bip), -vec
            assert pm.separate_segments(bip, vec, uniq, conn, mode="max") == r["part"], f"extracted blob {k}"
    return uniq


@pytest.mark.gpu
def test_cuda_ncut_from_golden_segment_features_is_exact():
    """Affinity bits, fp64 Lanczos Fiedler vectors and the blob extraction, fed the REFERENCE's per-segment features."""
    from unscene3d_b200 import pseudo_masks as pm

    g, case = load_gold()
    _check_against_golden(pm, (torch.from_numpy(g["agg_a"]).cuda(), torch.from_numpy(g["agg_b"]).cuda()), g, case)


@pytest.mark.gpu
def test_cuda_pseudo_masks_match_reference_golden():
    """End to end from points: CUDA segment means (fp64 sums, rounded once) -> affinity -> NCut -> masks."""
    from unscene3d_b200 import pseudo_masks as pm

    g, case = load_gold()
    agg = tuple(pm.aggregate_features(case["feats_" + k].cuda(), case["segment_ids"].cuda(), case["seg_connectivity"].cuda())[0]
                for k in ("a", "b"))
    uniq = _check_against_golden(pm, agg, g, case)
    # the whole greedy loop with the library's own orientation rule: masks are disjoint sets of segments, and the first
    # one (extracted before any noise-decided seed can steer the painting) is the reference's
    own = pm.unscene3d(agg, uniq, case["seg_connectivity"].cuda(), affinity_tau=0.65)
    followed = pm.unscene3d(agg, uniq, case["seg_connectivity"].cuda(), affinity_tau=0.65, sign_rule=follow(g["eigvecs"]))
    assert np.array_equal(followed[0], g["masks"][0])
    assert own.dtype == bool and own.shape[1] == len(uniq) and own.sum(0).max() <= 1


@pytest.mark.gpu
def test_cuda_scene_features_3d_equal_kdtree_nearest_voxel():
    """A17 (pseudo_masks/unscene3d_pseudo_main.py:332-348): every full-resolution voxel takes the features of its nearest
    res_2 voxel.  The reference asks a scipy KDTree; ours is the parent map.  Same voxel wherever the nearest one is unique,
    an equally near one otherwise (the KDTree's pick among ties is implementation-defined)."""
    from scipy.spatial import KDTree

    import unscene3d_b200  # noqa: F401
    from helpers import Cfg, deterministic_state, random_scene
    from unscene3d_b200 import engine, models
    from unscene3d_b200 import pseudo_masks as pm

    c = random_scene(4000, 90, batch=1, extent=30)
    torch.manual_seed(1)
    f = torch.randn(c.shape[0], 3)
    net = models.Res16UNet34CMultiRes(3, 20, Cfg(), D=3, out_fpn=True)
    net.load_state_dict(deterministic_state(net, 3))
    net = net.cuda().eval()
    with torch.no_grad():
        x = engine.SparseTensor(f.cuda(), torch.from_numpy(c).cuda())
        out = net(x)                                                    # one forward pass serves both sides
        got = pm.encode_scene_feats_3d(lambda _: out, x, resolution_scale=2)
    enc = out[1]["res_2"]
    lr = enc.C[:, 1:].cpu().numpy()
    dist, idx = KDTree(lr).query(c[:, 1:], k=2)
    ref = enc.F[torch.from_numpy(idx[:, 0]).cuda()]
    unique = dist[:, 0] < dist[:, 1] - 1e-9
    assert unique.mean() > 0.1  # dense random scene: most voxels have an equally near neighbour of their parent
    assert got.shape == ref.shape
    assert torch.equal(got[torch.from_numpy(unique).cuda()], ref[torch.from_numpy(unique).cuda()])
    # ties: the voxel we picked is exactly as near as the tree's
    parent_xyz = (c[:, 1:] // 2) * 2
    assert np.allclose(np.linalg.norm(c[:, 1:] - parent_xyz, axis=1), dist[:, 0])
