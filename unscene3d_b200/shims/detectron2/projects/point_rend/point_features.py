"""Imported but never called by models/criterion.py:13-16."""


def get_uncertain_point_coords_with_randomness(*a, **k):
    raise NotImplementedError("unused by the reference's criterion")


def point_sample(*a, **k):
    raise NotImplementedError("unused by the reference's criterion")
