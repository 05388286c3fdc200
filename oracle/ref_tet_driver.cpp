// TEST INFRASTRUCTURE ONLY (oracle / fixture generation). Never linked into the product library.
//
// extern "C" wrapper around the UNMODIFIED reference tetrahedralizer (/root/reference/src/mesh/Tetrahedralize.cpp,
// src/numeric/Predicates.cpp), compiled where the sources lie by oracle/Makefile into oracle/_ref/libme_ref_tet.so.
// Used only to regenerate the golden fixtures under tests/golden/ (tests/golden/make_golden.py): the reference's golden
// modal models (glTF_PhysicalAudio/samples) were solved over meshes produced by exactly this code.
#include "mesh/Tetrahedralize.h"

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {
tetra::Result g_result;
std::string g_error;
} // namespace

extern "C" {
// Tetrahedralizes the closed surface; returns 0 on success. Sizes are read back with the getters below.
int ref_tetrahedralize(const double *points_xyz, uint32_t n_points, const uint32_t *triangles, uint32_t n_triangle_indices, int quality) {
    std::vector<dvec3> points(n_points);
    for (uint32_t i = 0; i < n_points; ++i) points[i] = {points_xyz[3 * i], points_xyz[3 * i + 1], points_xyz[3 * i + 2]};
    auto result = tetra::Tetrahedralize(points, {triangles, n_triangle_indices}, {.Quality = quality != 0});
    if (!result) {
        g_error = result.error();
        return 1;
    }
    g_result = std::move(*result);
    return 0;
}
const char *ref_tet_error() { return g_error.c_str(); }
uint32_t ref_tet_point_count() { return uint32_t(g_result.Mesh.Points.size()); }
uint32_t ref_tet_count() { return uint32_t(g_result.Mesh.Tets.size()); }
void ref_tet_copy(double *points_xyz, uint32_t *tets) {
    for (size_t i = 0; i < g_result.Mesh.Points.size(); ++i) {
        points_xyz[3 * i] = g_result.Mesh.Points[i].x, points_xyz[3 * i + 1] = g_result.Mesh.Points[i].y, points_xyz[3 * i + 2] = g_result.Mesh.Points[i].z;
    }
    std::memcpy(tets, g_result.Mesh.Tets.data(), g_result.Mesh.Tets.size() * 4 * sizeof(uint32_t));
}
}
