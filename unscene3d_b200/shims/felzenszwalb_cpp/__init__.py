"""`felzenszwalb_cpp` (the reference's pybind module utils/cpp_utils/segmentator.cpp:152-262) on libus3d's host function.

    comps, connectivity = felzenszwalb_cpp.segment_mesh(vertices, faces, colors, kthr=0.005, segMinVerts=20)

vertices / colors: [V, 3] float32, faces: [F, 3] int32 (the reference reads the buffers as float / int without converting; here
other dtypes are converted).  Returns comps [V] (segment ids 0..S-1) and connectivity [P, 2] (directed adjacent segment pairs,
lexicographically sorted), int32 like the reference's arrays."""
import numpy as np

from unscene3d_b200._lib import check, lib


def segment_mesh(np_vertices, np_faces, np_colors, kthr=0.005, segMinVerts=20):
    v = np.ascontiguousarray(np_vertices, dtype=np.float32).reshape(-1, 3)
    f = np.ascontiguousarray(np_faces, dtype=np.int32).reshape(-1, 3)
    c = np.ascontiguousarray(np_colors, dtype=np.float32).reshape(-1, 3)
    if c.shape[0] != v.shape[0]:
        raise ValueError("felzenszwalb_cpp.segment_mesh: one colour per vertex expected")
    comps = np.empty(v.shape[0], dtype=np.int32)
    cap = max(3 * f.shape[0], 1)
    pairs = np.empty((cap, 2), dtype=np.int32)
    n = lib.us3d_felzenszwalb_segment_h(v.ctypes.data, f.ctypes.data, c.ctypes.data, v.shape[0], f.shape[0], float(kthr), int(segMinVerts),
                                        comps.ctypes.data, pairs.ctypes.data, cap)
    if n < 0:
        check(n)
    return comps, pairs[:n].copy()
