// TEST INFRASTRUCTURE ONLY (oracle / fixture generation). Never linked into the product library.
//
// extern "C" wrapper around the UNMODIFIED reference tetrahedralizer (/root/reference/src/mesh/Tetrahedralize.cpp,
// src/numeric/Predicates.cpp), compiled where the sources lie by oracle/Makefile into oracle/_ref/libme_ref_tet.so.
// Used only to regenerate the golden fixtures under tests/golden/ (tests/golden/make_golden.py): the reference's golden
// modal models (glTF_PhysicalAudio/samples) were solved over meshes produced by exactly this code.
// Also the reference's own BuildTetMeshData and SimplifySurface (src/mesh/Tets.cpp, with lib/meshoptimizer), as the checker of
// me_build_tet_mesh_data and a fixture generator.
#include "mesh/Tetrahedralize.h"
#include "mesh/Tets.h"

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

// BuildTetMeshData (Tets.cpp:268-293) over a caller's tet mesh. Returns the edge index count; the arrays are read with ref_tet_data_copy.
namespace {
TetMeshData g_data;
}
uint32_t ref_build_tet_mesh_data(const double *points_xyz, uint32_t n_points, const uint32_t *tets, uint32_t n_tets, const float *scale) {
    TetMesh mesh;
    mesh.Points.resize(n_points);
    for (uint32_t i = 0; i < n_points; ++i) mesh.Points[i] = {points_xyz[3 * i], points_xyz[3 * i + 1], points_xyz[3 * i + 2]};
    mesh.Tets.resize(n_tets);
    std::memcpy(mesh.Tets.data(), tets, size_t(n_tets) * 4 * sizeof(uint32_t));
    g_data = BuildTetMeshData(mesh, vec3{scale[0], scale[1], scale[2]});
    return uint32_t(g_data.EdgeIndices.size());
}
void ref_tet_data_copy(float *positions_xyz, uint32_t *edge_indices) {
    std::memcpy(positions_xyz, g_data.Positions.data(), g_data.Positions.size() * sizeof(vec3));
    std::memcpy(edge_indices, g_data.EdgeIndices.data(), g_data.EdgeIndices.size() * sizeof(uint32_t));
}
// SimplifySurface (Tets.cpp:249-262) in place; returns the new position count, *n_triangle_indices the new index count.
uint32_t ref_simplify_surface(float *positions_xyz, uint32_t n_positions, uint32_t *triangle_indices, uint32_t *n_triangle_indices, float ratio) {
    std::vector<vec3> positions(n_positions);
    std::memcpy(positions.data(), positions_xyz, size_t(n_positions) * sizeof(vec3));
    std::vector<uint32_t> tris(triangle_indices, triangle_indices + *n_triangle_indices);
    SimplifySurface(positions, tris, ratio);
    std::memcpy(positions_xyz, positions.data(), positions.size() * sizeof(vec3));
    std::memcpy(triangle_indices, tris.data(), tris.size() * sizeof(uint32_t));
    *n_triangle_indices = uint32_t(tris.size());
    return uint32_t(positions.size());
}
}
