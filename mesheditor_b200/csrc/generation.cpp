// Generation-job glue (first slice): what the reference's modal generation job does on either side of
// modal::mesh2modes (src/audio/AudioSystem.cpp:838-862) apart from simplifying and tetrahedralizing the surface:
//   * the sample surface: the mesh's triangulation collapsed onto the excitation vertices (SampleSurfaceTriangles, :701-746),
//     then relabelled onto the sample points the solve merged them into (RelabelSampleTriangles, :761-769), both through
//     UniqueSampleTriangles (:675-695);
//   * ModalModes::Vertices from the solve's SamplePointOfExcitation (CompactExcitationVertices, :750-757);
//   * the display TetMeshData a `.modal` file stores (BuildTetMeshData, src/mesh/Tets.cpp:268-293).
// Host code, like the reference's: integer passes over the surface, nowhere near the solve's cost.
#include "common.h"

#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace me {
namespace {

using Triple = std::array<uint32_t, 3>;

template<typename T>
T *Export(const std::vector<T> &v) {
    auto *out = static_cast<T *>(std::malloc(std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!out) Fail(ME_OUT_OF_MEMORY, "host allocation failed");
    if (!v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(T));
    return out;
}

// One triangle per distinct set of three different points, ordered by the sorted set. Which winding of a repeated set
// survives is what the reference leaves to its sort (UniqueSampleTriangles sorts by the set alone, then keeps the first of
// every run): the same std::sort over the same sequence compared the same way gives the same survivor as the reference
// built with the same standard library - the first seen whenever the sort happens to be stable (16 triangles or fewer).
std::vector<uint32_t> DistinctTriangles(const std::vector<Triple> &windings) {
    struct Entry {
        Triple Set, Winding;
    };
    std::vector<Entry> entries;
    entries.reserve(windings.size());
    for (const auto &w : windings) {
        if (w[0] == w[1] || w[1] == w[2] || w[0] == w[2]) continue;
        Triple set = w;
        std::sort(set.begin(), set.end());
        entries.push_back({set, w});
    }
    std::sort(entries.begin(), entries.end(), [](const Entry &a, const Entry &b) { return a.Set < b.Set; });
    std::vector<uint32_t> out;
    out.reserve(entries.size() * 3);
    for (size_t i = 0; i < entries.size(); ++i) {
        if (i && entries[i].Set == entries[i - 1].Set) continue;
        out.insert(out.end(), entries[i].Winding.begin(), entries[i].Winding.end());
    }
    return out;
}

} // namespace
} // namespace me

using namespace me;

extern "C" {

MeStatus me_sample_surface_triangles(const uint32_t *triangle_indices, uint32_t n_triangle_indices, uint32_t vertex_count, const uint32_t *excitation_vertices, uint32_t n_excitation_vertices,
                                     uint32_t **out, uint32_t *n_out) {
    return Guard([&] {
        if (!out || !n_out || (n_triangle_indices && !triangle_indices) || (n_excitation_vertices && !excitation_vertices)) Fail(ME_BAD_ARG, "null argument");
        *out = nullptr, *n_out = 0;
        std::vector<uint32_t> result;
        const uint32_t n_tri = n_triangle_indices / 3;
        if (n_excitation_vertices >= 3 && n_tri >= 1) {
            for (uint32_t i = 0; i < n_triangle_indices; ++i)
                if (triangle_indices[i] >= vertex_count) Fail(ME_BAD_ARG, "triangle index %u outside the %u vertices", triangle_indices[i], vertex_count);
            // Neighbour lists in compressed rows: every corner lists its triangle's next and next-but-one corner, triangles in order
            // (the order decides which of two equally near excitation vertices a vertex takes).
            std::vector<uint32_t> row(size_t(vertex_count) + 1, 0);
            for (uint32_t i = 0; i < n_triangle_indices; ++i) row[triangle_indices[i] + 1] += 2;
            for (uint32_t v = 0; v < vertex_count; ++v) row[v + 1] += row[v];
            std::vector<uint32_t> neighbour(row[vertex_count]), cursor(row.begin(), row.end() - 1);
            for (uint32_t t = 0; t < n_tri; ++t) {
                const uint32_t *c = triangle_indices + size_t(3) * t;
                for (uint32_t k = 0; k < 3; ++k) {
                    neighbour[cursor[c[k]]++] = c[(k + 1) % 3];
                    neighbour[cursor[c[k]]++] = c[(k + 2) % 3];
                }
            }
            // Every vertex takes the excitation vertex it reaches in the fewest edges: one breadth-first sweep from all of them.
            constexpr uint32_t kNone = ~0u;
            std::vector<uint32_t> nearest(vertex_count, kNone), frontier;
            frontier.reserve(vertex_count);
            for (uint32_t s = 0; s < n_excitation_vertices; ++s) {
                const uint32_t v = excitation_vertices[s];
                if (v < vertex_count && nearest[v] == kNone) nearest[v] = s, frontier.push_back(v);
            }
            for (size_t at = 0; at < frontier.size(); ++at) {
                const uint32_t v = frontier[at];
                for (uint32_t i = row[v]; i < row[v + 1]; ++i) {
                    const uint32_t u = neighbour[i];
                    if (nearest[u] == kNone) nearest[u] = nearest[v], frontier.push_back(u);
                }
            }
            // A mesh triangle whose corners took three different excitation vertices contributes a triangle over them; a corner in a
            // shell without an excitation vertex has nothing to collapse onto.
            std::vector<Triple> collapsed;
            collapsed.reserve(n_tri);
            for (uint32_t t = 0; t < n_tri; ++t) {
                const Triple w{nearest[triangle_indices[3 * size_t(t)]], nearest[triangle_indices[3 * size_t(t) + 1]], nearest[triangle_indices[3 * size_t(t) + 2]]};
                if (w[0] != kNone && w[1] != kNone && w[2] != kNone) collapsed.push_back(w);
            }
            result = DistinctTriangles(collapsed);
        }
        *out = Export(result), *n_out = uint32_t(result.size());
    });
}

MeStatus me_compact_excitation_vertices(const uint32_t *vertices, uint32_t n_vertices, const uint32_t *sample_point_of, uint32_t n_sample_point_of, uint32_t **out, uint32_t *n_out) {
    return Guard([&] {
        if (!out || !n_out || (n_vertices && !vertices) || (n_sample_point_of && !sample_point_of)) Fail(ME_BAD_ARG, "null argument");
        // Sample points are numbered in the order their first excitation position appears, so the first vertex of each is the
        // one whose sample point number equals the count gathered so far.
        std::vector<uint32_t> firsts;
        const uint32_t n = std::min(n_vertices, n_sample_point_of);
        for (uint32_t i = 0; i < n; ++i)
            if (sample_point_of[i] == firsts.size()) firsts.push_back(vertices[i]);
        *out = Export(firsts), *n_out = uint32_t(firsts.size());
    });
}

MeStatus me_relabel_sample_triangles(const uint32_t *triangles, uint32_t n_triangle_indices, const uint32_t *sample_point_of, uint32_t n_sample_point_of, uint32_t **out, uint32_t *n_out) {
    return Guard([&] {
        if (!out || !n_out || (n_triangle_indices && !triangles) || (n_sample_point_of && !sample_point_of)) Fail(ME_BAD_ARG, "null argument");
        std::vector<uint32_t> result;
        if (n_sample_point_of) {
            std::vector<Triple> moved(n_triangle_indices / 3);
            for (size_t t = 0; t < moved.size(); ++t)
                for (int c = 0; c < 3; ++c) {
                    const uint32_t corner = triangles[3 * t + c];
                    if (corner >= n_sample_point_of) Fail(ME_BAD_ARG, "triangle corner %u has no sample point (%u excitation positions)", corner, n_sample_point_of);
                    moved[t][c] = sample_point_of[corner];
                }
            result = DistinctTriangles(moved);
        }
        *out = Export(result), *n_out = uint32_t(result.size());
    });
}

MeStatus me_desired_solve_vertices(uint32_t requested, uint32_t num_vertices, uint32_t **out, uint32_t *n_out) {
    return Guard([&] {
        if (!out || !n_out) Fail(ME_BAD_ARG, "null argument");
        if (!num_vertices) Fail(ME_BAD_ARG, "a mesh without vertices has nothing to excite");
        // Evenly spaced over the mesh's vertex order, capped at the vertex count so that none is taken twice; the products stay in
        // 32 bits, as the reference's do.
        const uint32_t count = std::clamp(requested, 1u, num_vertices);
        std::vector<uint32_t> picked(count);
        for (uint32_t i = 0; i < count; ++i) picked[i] = i * num_vertices / count;
        *out = Export(picked), *n_out = count;
    });
}

MeStatus me_build_tet_mesh_data(const double *points_xyz, uint32_t n_points, const uint32_t *tets, uint32_t n_tets, const float scale[3], float **positions_xyz, uint32_t **edge_indices,
                                uint32_t *n_edge_indices) {
    return Guard([&] {
        if (!positions_xyz || !edge_indices || !n_edge_indices || !scale || (n_points && !points_xyz) || (n_tets && !tets)) Fail(ME_BAD_ARG, "null argument");
        // Positions go back to the node's local frame in double, then to float.
        const double inverse[3] = {1.0 / double(scale[0]), 1.0 / double(scale[1]), 1.0 / double(scale[2])};
        std::vector<float> local(size_t(3) * n_points);
        for (size_t i = 0; i < local.size(); ++i) local[i] = float(points_xyz[i] * inverse[i % 3]);
        // The distinct edges of the tets, ascending by (low corner, high corner).
        std::vector<uint64_t> keys;
        keys.reserve(size_t(6) * n_tets);
        for (uint32_t t = 0; t < n_tets; ++t) {
            const uint32_t *c = tets + size_t(4) * t;
            for (int a = 0; a < 4; ++a) {
                if (c[a] >= n_points) Fail(ME_BAD_ARG, "tet %u refers to point %u of %u", t, c[a], n_points);
                for (int b = a + 1; b < 4; ++b) keys.push_back(uint64_t(std::min(c[a], c[b])) << 32 | std::max(c[a], c[b]));
            }
        }
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        std::vector<uint32_t> pairs(keys.size() * 2);
        for (size_t e = 0; e < keys.size(); ++e) pairs[2 * e] = uint32_t(keys[e] >> 32), pairs[2 * e + 1] = uint32_t(keys[e]);
        float *p = Export(local);
        uint32_t *q = nullptr;
        try {
            q = Export(pairs);
        } catch (...) {
            std::free(p);
            throw;
        }
        *positions_xyz = p, *edge_indices = q, *n_edge_indices = uint32_t(pairs.size());
    });
}

} // extern "C"
