// Felzenszwalb graph segmentation of a coloured triangle mesh into geometric segments — HOST function of the
// C ABI.  Produces the `segment_ids` / `seg_connectivity` both pseudo-mask paths and the self-training targets consume.
//
// Replaces the reference's pybind module `felzenszwalb_cpp.segment_mesh` (/root/reference/utils/cpp_utils/segmentator.cpp:17-247,
// disjoint-set forest and edge type in include/segmentator.h:20-96; callers datasets/freemask_semseg.py:212,
// pseudo_masks/datasets/scannet.py:182).  The algorithm is sequential by nature (edges in weight order over a union-find forest),
// the reference runs it on the host and so does this: a flat implementation over caller-provided arrays — no per-vertex
// vec3 vectors, no std::map for the adjacency (pairs are collected, sorted and de-duplicated once), forest as three int arrays.
//
// Results are bit-identical to the reference module (tests/test_felzenszwalb.py compares with the reference source compiled
// as it is, oracle/_ref).  That requires the same arithmetic (float32, same expression order: running-average vertex normals
// in face order, weight = (1 - n1.n2) * sum|dc|, squared for convex, similarly coloured edges) and — because equal weights are
// common (flat, uniformly coloured regions have weight 0) and the small-segment pass depends on edge order — the same
// permutation of equal-weight edges: both sort with std::sort and a weight-only comparison.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

#include "common.cuh"

namespace us3d {
namespace fz {

struct Edge {
    float w;
    int a, b;
};
inline bool operator<(const Edge &x, const Edge &y) { return x.w < y.w; }

struct Forest {
    std::vector<int> parent, rank, size;
    explicit Forest(int n) : parent(n), rank(n, 0), size(n, 1) {
        for (int i = 0; i < n; ++i) parent[i] = i;
    }
    int find(int x) {  // the reference compresses only the queried element's link (segmentator.h:63-70)
        int y = x;
        while (y != parent[y]) y = parent[y];
        parent[x] = y;
        return y;
    }
    void join(int x, int y) {
        if (rank[x] > rank[y]) {
            parent[y] = x;
            size[x] += size[y];
        } else {
            parent[x] = y;
            size[y] += size[x];
            if (rank[x] == rank[y]) ++rank[y];
        }
    }
};

struct V3 {
    float x, y, z;
};

static inline V3 unit_cross(const V3 &u, const V3 &v) {
    V3 c = {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
    const float n = sqrtf(c.x * c.x + c.y * c.y + c.z * c.z);
    c.x /= n;
    c.y /= n;
    c.z /= n;
    return c;
}
static inline V3 blend(const V3 &a, const V3 &b, float v) {
    const float u = 1.0f - v;
    return V3{v * b.x + u * a.x, v * b.y + u * a.y, v * b.z + u * a.z};
}

}  // namespace fz
}  // namespace us3d

using namespace us3d;

extern "C" {

/* vertices / colors float[n_verts][3], faces int32[n_faces][3] -> comps int32[n_verts] (segment ids 0..S-1 in the order of their
 * representative vertex), pairs int32[<= cap][2] (directed adjacent segment pairs (s1, s2), s1 != s2, lexicographically sorted).
 * Returns the number of pairs (only the first `cap` are written) or a negative error code. */
int us3d_felzenszwalb_segment_h(const float *vertices, const int32_t *faces, const float *colors, int n_verts, int n_faces, float kthr,
                                int seg_min_verts, int32_t *comps, int32_t *pairs, int cap) {
    US3D_CHECK_ARG(n_verts >= 0 && n_faces >= 0 && vertices != nullptr && colors != nullptr && comps != nullptr, "felzenszwalb: bad arguments");
    for (long long i = 0; i < 3LL * n_faces; ++i)
        US3D_CHECK_ARG(faces[i] >= 0 && faces[i] < n_verts, "felzenszwalb: face %lld refers to vertex %d of %d", i / 3, faces[i], n_verts);
    const fz::V3 *P = reinterpret_cast<const fz::V3 *>(vertices);
    const fz::V3 *C = reinterpret_cast<const fz::V3 *>(colors);
    const size_t n_edges = (size_t)n_faces * 3;
    std::vector<fz::Edge> edges(n_edges);
    std::vector<fz::V3> normal(n_verts, fz::V3{0.f, 0.f, 0.f});
    std::vector<int> count(n_verts, 0);
    // face normals blended into running-average vertex normals, in face order (segmentator.cpp:58-83)
    for (int f = 0; f < n_faces; ++f) {
        const int i1 = faces[3 * f], i2 = faces[3 * f + 1], i3 = faces[3 * f + 2];
        edges[3 * f] = {0.f, i1, i2};
        edges[3 * f + 1] = {0.f, i1, i3};
        edges[3 * f + 2] = {0.f, i3, i2};
        const fz::V3 p1 = P[i1], p2 = P[i2], p3 = P[i3];
        const fz::V3 n = fz::unit_cross(fz::V3{p2.x - p1.x, p2.y - p1.y, p2.z - p1.z}, fz::V3{p3.x - p1.x, p3.y - p1.y, p3.z - p1.z});
        normal[i1] = fz::blend(normal[i1], n, 1.0f / (count[i1] + 1.0f));
        normal[i2] = fz::blend(normal[i2], n, 1.0f / (count[i2] + 1.0f));
        normal[i3] = fz::blend(normal[i3], n, 1.0f / (count[i3] + 1.0f));
        ++count[i1];
        ++count[i2];
        ++count[i3];
    }
    // edge weights (:86-123)
    for (size_t e = 0; e < n_edges; ++e) {
        const int a = edges[e].a, b = edges[e].b;
        const fz::V3 &n1 = normal[a], &n2 = normal[b], &p1 = P[a], &p2 = P[b];
        float dx = p2.x - p1.x, dy = p2.y - p1.y, dz = p2.z - p1.z;
        const float dd = sqrtf(dx * dx + dy * dy + dz * dz);
        dx /= dd;
        dy /= dd;
        dz /= dd;
        const float dot = n1.x * n2.x + n1.y * n2.y + n1.z * n2.z;
        const float normal_dist = 1.0f - dot;
        const float color_dist = (fabsf(C[a].x - C[b].x) + fabsf(C[a].y - C[b].y) + fabsf(C[a].z - C[b].z));
        float dist = normal_dist * color_dist;
        const float dot2 = n2.x * dx + n2.y * dy + n2.z * dz;
        if (dot2 > 0 && color_dist < 0.05) dist = dist * dist;
        edges[e].w = dist;
    }
    // segment_graph (:17-46)
    std::sort(edges.begin(), edges.end());
    fz::Forest u(n_verts);
    {
        std::vector<float> threshold(n_verts, kthr);
        for (size_t e = 0; e < n_edges; ++e) {
            int a = u.find(edges[e].a);
            const int b = u.find(edges[e].b);
            if (a != b && edges[e].w <= threshold[a] && edges[e].w <= threshold[b]) {
                u.join(a, b);
                a = u.find(a);
                threshold[a] = edges[e].w + (kthr / u.size[a]);
            }
        }
    }
    // small segments join a neighbour, in edge order (:128-135)
    for (size_t e = 0; e < n_edges; ++e) {
        const int a = u.find(edges[e].a), b = u.find(edges[e].b);
        if (a != b && (u.size[a] < seg_min_verts || u.size[b] < seg_min_verts)) u.join(a, b);
    }
    // compact ids in the order of the representative vertices (:206-228)
    std::vector<int> root(n_verts), id(n_verts, -1);
    for (int q = 0; q < n_verts; ++q) root[q] = u.find(q);
    int n_seg = 0;
    {
        std::vector<char> is_root(n_verts, 0);
        for (int q = 0; q < n_verts; ++q) is_root[root[q]] = 1;
        for (int q = 0; q < n_verts; ++q)
            if (is_root[q]) id[q] = n_seg++;
    }
    for (int q = 0; q < n_verts; ++q) comps[q] = id[root[q]];
    // adjacency: distinct (s1, s2) over the edges, sorted (:141-149, 230-251)
    std::vector<std::pair<int, int>> adj;
    adj.reserve(n_edges / 8 + 16);
    for (size_t e = 0; e < n_edges; ++e) {
        const int s1 = comps[edges[e].a], s2 = comps[edges[e].b];
        if (s1 != s2) adj.emplace_back(s1, s2);
    }
    std::sort(adj.begin(), adj.end());
    adj.erase(std::unique(adj.begin(), adj.end()), adj.end());
    const int n_pairs = (int)adj.size();
    if (pairs != nullptr)
        for (int i = 0; i < n_pairs && i < cap; ++i) {
            pairs[2 * i] = adj[i].first;
            pairs[2 * i + 1] = adj[i].second;
        }
    return n_pairs;
}

}  // extern "C"
