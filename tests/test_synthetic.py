"""The synthetic ScanNet-shaped scene generator must yield a scene of exactly the requested size for EVERY seed:
the ranks of a multi-GPU bench run use seed = rank (a 4-GPU run of round 1 died on seed 2)."""
import hashlib

import numpy as np
import pytest

from unscene3d_b200.synthetic import level_sizes, make_scene


@pytest.mark.parametrize("seed", range(8))
def test_every_rank_seed_yields_a_full_size_scene(seed):
    s = make_scene(200_000, seed=seed, with_masks=False)
    assert s.coords.shape == (200_000, 3) and s.coords.dtype == np.int32
    assert np.unique(s.coords, axis=0).shape[0] == 200_000          # voxels are unique
    assert s.colors.shape == (200_000, 3) and s.point2segment.shape == (200_000,)
    assert s.coords.min() < 0 < s.coords.max()                      # mean-centred: negative coordinates occur


def test_seed_zero_scene_is_the_one_the_round_1_numbers_were_measured_on():
    s = make_scene(200_000, seed=0, with_masks=False)
    assert hashlib.md5(s.coords.tobytes()).hexdigest() == "30138071eb8e33eb6e6db68465a8876f"
    c4 = np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)
    assert level_sizes(c4) == [200000, 50121, 11676, 2223, 513]


@pytest.mark.parametrize("n,seed", [(5_000, 3), (50_000, 1), (300_000, 7)])
def test_other_sizes(n, seed):
    s = make_scene(n, seed=seed, with_masks=True)
    assert s.n == n and s.masks.shape[1] == n and s.segment_mask.shape[0] == s.masks.shape[0] == s.labels.shape[0]
