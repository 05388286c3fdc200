// Model interchange: the formats either side of a modal solve. Host code, like the reference's.
//   * `.modal` files: ModalModelData{Modes, Mass, Tets, Summary} (src/audio/ModalModelFile.h:13-20) in the byte layout the
//     reference's zpp::bits archive produces (ModalModelFile.cpp:15-22, :52-58): members in declaration order, arithmetic
//     values little-endian as they lie in memory, every std::vector as a 4-byte element count followed by its elements,
//     glm vectors component by component and the quaternion as x, y, z, w (src/action/SerializeGlm.h).
//   * the JSON that MeshEditorModalSolve prints (tests/ModalSolveTool.cpp:84-123), which glTF_PhysicalAudio's generator
//     embeds as a KHR_audio_rigid_bodies modal model: decayRates = ln 1000 / T60, shapes mode-major, triangles relabelled
//     onto the sample points.
#include "common.h"
#include "result.h"

#include <algorithm>
#include <charconv>
#include <cstdlib>
#include <memory>
#include <numbers>
#include <string>

struct MeModalFile {
    std::vector<uint32_t> Vertices, Indices, TetEdges, SolvedVertices;
    std::vector<float> TetPositions;
    MeModalFileExtras Extras{};
};

namespace me {
namespace {

struct Writer {
    std::vector<uint8_t> Bytes;
    template<typename T>
    void Put(const T &v) {
        const auto *p = reinterpret_cast<const uint8_t *>(&v);
        Bytes.insert(Bytes.end(), p, p + sizeof(T));
    }
    template<typename T>
    void PutArray(const T *v, size_t count) {
        if (count >= (uint64_t(1) << 32)) Fail(ME_BAD_ARG, "array of %zu elements does not fit the file's 32-bit counts", count);
        Put(uint32_t(count));
        const auto *p = reinterpret_cast<const uint8_t *>(v);
        Bytes.insert(Bytes.end(), p, p + count * sizeof(T));
    }
};

struct Reader {
    const uint8_t *At, *End;
    template<typename T>
    T Get() {
        if (size_t(End - At) < sizeof(T)) Fail(ME_BAD_ARG, "truncated .modal data");
        T v;
        std::memcpy(&v, At, sizeof(T));
        At += sizeof(T);
        return v;
    }
    template<typename T>
    void GetArray(std::vector<T> &out, size_t elements_per_count = 1) {
        const size_t count = size_t(Get<uint32_t>()) * elements_per_count;
        if (size_t(End - At) < count * sizeof(T)) Fail(ME_BAD_ARG, "truncated .modal data");
        out.resize(count);
        std::memcpy(out.data(), At, count * sizeof(T));
        At += count * sizeof(T);
    }
};

template<typename T>
void AppendNumber(std::string &s, T v) { // the shortest text that round-trips, which is what std::format("{}") prints
    char buf[64];
    const auto r = std::to_chars(buf, buf + sizeof buf, v);
    s.append(buf, r.ptr);
}

} // namespace
} // namespace me

using namespace me;

extern "C" {

MeStatus me_modal_file_serialize(const MeModalResult *r, const MeModalFileExtras *x, uint8_t **bytes, uint64_t *size) {
    return Guard([&] {
        if (!r || !x || !bytes || !size) Fail(ME_BAD_ARG, "null argument");
        const uint32_t modes = uint32_t(r->Modes.Freqs.size()), points = r->PointCount, eigen = uint32_t(r->Eigenvalues.size());
        if (r->Modes.Shapes.size() != size_t(points) * modes * 3 || r->Modes.Positions.size() != size_t(points) * 3) Fail(ME_BAD_ARG, "inconsistent modal result");
        Writer w;
        const auto nested = [&](const float *flat, uint32_t columns) { // std::vector<std::vector<vec3>>
            w.Put(points);
            for (uint32_t p = 0; p < points; ++p) {
                w.Put(columns);
                const auto *b = reinterpret_cast<const uint8_t *>(flat + size_t(p) * columns * 3);
                w.Bytes.insert(w.Bytes.end(), b, b + size_t(columns) * 12);
            }
        };
        const auto vec3s = [&](const float *xyz, uint32_t count) { // std::vector<vec3>
            w.Put(count);
            const auto *b = reinterpret_cast<const uint8_t *>(xyz);
            w.Bytes.insert(w.Bytes.end(), b, b + size_t(count) * 12);
        };
        // ModalModes (ModalModes.h:7-20)
        w.PutArray(r->Modes.Freqs.data(), modes);
        w.PutArray(r->Modes.T60s.data(), modes);
        nested(r->Modes.Shapes.data(), modes);
        w.PutArray(x->vertices, x->n_vertices);
        vec3s(r->Modes.Positions.data(), points);
        w.PutArray(x->indices, x->n_indices);
        w.Put(r->Modes.OriginalFundamentalFreq);
        w.Put(x->baked_scale[0]), w.Put(x->baked_scale[1]), w.Put(x->baked_scale[2]);
        // MassProperties (ContactModel.h:16-23); the quaternion is archived x, y, z, w
        w.Put(r->MassProps.mass);
        for (float v : r->MassProps.center_of_mass) w.Put(v);
        for (float v : r->MassProps.inertia_diagonal) w.Put(v);
        w.Put(r->MassProps.inertia_orientation[1]), w.Put(r->MassProps.inertia_orientation[2]), w.Put(r->MassProps.inertia_orientation[3]), w.Put(r->MassProps.inertia_orientation[0]);
        // TetMeshData (TetMeshData.h:8-13)
        vec3s(x->tet_positions_xyz, x->n_tet_positions);
        w.PutArray(x->tet_edge_indices, x->n_tet_edge_indices);
        // ModalEigenSummary (ModalEigenSummary.h:12-23)
        w.PutArray(r->Eigenvalues.data(), eigen);
        if (r->SummaryShapes.size() != size_t(points) * eigen * 3) Fail(ME_BAD_ARG, "inconsistent eigen summary");
        nested(r->SummaryShapes.data(), eigen);
        w.Put(x->solved_material.density), w.Put(x->solved_material.young_modulus), w.Put(x->solved_material.poisson_ratio), w.Put(x->solved_material.alpha), w.Put(x->solved_material.beta);
        w.Put(x->solved_min_mode_freq), w.Put(x->solved_max_mode_freq);
        w.Put(x->solved_num_modes);
        w.Put(x->tet_inputs_hash);
        w.PutArray(x->solved_vertices, x->n_solved_vertices);
        auto *out = static_cast<uint8_t *>(std::malloc(std::max<size_t>(w.Bytes.size(), 1)));
        if (!out) Fail(ME_OUT_OF_MEMORY, "host allocation failed");
        std::memcpy(out, w.Bytes.data(), w.Bytes.size());
        *bytes = out, *size = w.Bytes.size();
    });
}

MeStatus me_modal_file_parse(const uint8_t *bytes, uint64_t size, MeModalResult **result, MeModalFile **file) {
    return Guard([&] {
        if (!bytes || !result || !file) Fail(ME_BAD_ARG, "null argument");
        auto r = std::make_unique<MeModalResult>();
        auto f = std::make_unique<MeModalFile>();
        Reader in{bytes, bytes + size};
        const auto nested = [&](std::vector<float> &flat, uint32_t &points, uint32_t &columns) {
            points = in.Get<uint32_t>();
            columns = 0;
            flat.clear();
            for (uint32_t p = 0; p < points; ++p) {
                std::vector<float> row;
                in.GetArray(row, 3);
                if (p && row.size() != size_t(columns) * 3) Fail(ME_BAD_ARG, "ragged shape table in .modal data");
                columns = uint32_t(row.size() / 3);
                flat.insert(flat.end(), row.begin(), row.end());
            }
        };
        in.GetArray(r->Modes.Freqs), in.GetArray(r->Modes.T60s);
        uint32_t points = 0, columns = 0;
        nested(r->Modes.Shapes, points, columns);
        if (r->Modes.T60s.size() != r->Modes.Freqs.size()) Fail(ME_BAD_ARG, "%zu T60s for %zu frequencies in .modal data", r->Modes.T60s.size(), r->Modes.Freqs.size());
        if (points && columns != r->Modes.Freqs.size()) Fail(ME_BAD_ARG, "mode shapes do not match the mode count in .modal data");
        r->PointCount = points;
        in.GetArray(f->Vertices);
        in.GetArray(r->Modes.Positions, 3);
        if (r->Modes.Positions.size() != size_t(points) * 3) Fail(ME_BAD_ARG, "positions do not match the sample points in .modal data");
        in.GetArray(f->Indices);
        r->Modes.OriginalFundamentalFreq = in.Get<float>();
        for (float &v : f->Extras.baked_scale) v = in.Get<float>();
        r->MassProps.mass = in.Get<double>();
        for (float &v : r->MassProps.center_of_mass) v = in.Get<float>();
        for (float &v : r->MassProps.inertia_diagonal) v = in.Get<float>();
        r->MassProps.inertia_orientation[1] = in.Get<float>(), r->MassProps.inertia_orientation[2] = in.Get<float>(), r->MassProps.inertia_orientation[3] = in.Get<float>();
        r->MassProps.inertia_orientation[0] = in.Get<float>();
        in.GetArray(f->TetPositions, 3), in.GetArray(f->TetEdges);
        in.GetArray(r->Eigenvalues);
        uint32_t summary_points = 0, eigen = 0;
        nested(r->SummaryShapes, summary_points, eigen);
        // The accessors and me_rescale_modes index SummaryShapes[p * eigen + k] for every sample point: an archive that carries
        // eigenvalues must carry a summary row per point (the reference's nested vectors keep their own sizes; a flat table cannot).
        const bool summary_expected = !r->Eigenvalues.empty() && points > 0;
        if ((summary_points || summary_expected) && (summary_points != points || eigen != r->Eigenvalues.size())) Fail(ME_BAD_ARG, "eigen summary does not match in .modal data");
        auto &m = f->Extras.solved_material;
        m.density = in.Get<double>(), m.young_modulus = in.Get<double>(), m.poisson_ratio = in.Get<double>(), m.alpha = in.Get<double>(), m.beta = in.Get<double>();
        f->Extras.solved_min_mode_freq = in.Get<float>(), f->Extras.solved_max_mode_freq = in.Get<float>();
        f->Extras.solved_num_modes = in.Get<uint32_t>();
        f->Extras.tet_inputs_hash = in.Get<uint64_t>();
        in.GetArray(f->SolvedVertices);
        if (in.At != in.End) Fail(ME_BAD_ARG, "%zu trailing bytes in .modal data", size_t(in.End - in.At));
        auto &x = f->Extras;
        x.vertices = f->Vertices.data(), x.n_vertices = uint32_t(f->Vertices.size());
        x.indices = f->Indices.data(), x.n_indices = uint32_t(f->Indices.size());
        x.tet_positions_xyz = f->TetPositions.data(), x.n_tet_positions = uint32_t(f->TetPositions.size() / 3);
        x.tet_edge_indices = f->TetEdges.data(), x.n_tet_edge_indices = uint32_t(f->TetEdges.size());
        x.solved_vertices = f->SolvedVertices.data(), x.n_solved_vertices = uint32_t(f->SolvedVertices.size());
        *result = r.release(), *file = f.release();
    });
}

MeStatus me_modal_file_extras(const MeModalFile *f, MeModalFileExtras *out) {
    return Guard([&] {
        if (!f || !out) Fail(ME_BAD_ARG, "null argument");
        *out = f->Extras;
    });
}
void me_modal_file_free(MeModalFile *f) { delete f; }
void me_bytes_free(void *p) { std::free(p); }

MeStatus me_modal_solve_json(const MeModalResult *r, const uint32_t *triangles, uint32_t n_triangle_indices, char **json) {
    return Guard([&] {
        if (!r || !json || (n_triangle_indices && !triangles)) Fail(ME_BAD_ARG, "null argument");
        const uint32_t modes = uint32_t(r->Modes.Freqs.size()), points = r->PointCount;
        // The mesh's triangles, relabelled onto the sample points its vertices became; merged corners drop the triangle.
        std::vector<uint32_t> indices;
        for (uint32_t t = 0; t + 2 < n_triangle_indices; t += 3) {
            for (int c = 0; c < 3; ++c)
                if (triangles[t + c] >= r->SamplePointOfExcitation.size()) Fail(ME_BAD_ARG, "triangle index %u has no excitation sample", triangles[t + c]);
            const uint32_t a = r->SamplePointOfExcitation[triangles[t]], b = r->SamplePointOfExcitation[triangles[t + 1]], c = r->SamplePointOfExcitation[triangles[t + 2]];
            if (a == b || b == c || a == c) continue;
            indices.insert(indices.end(), {a, b, c});
        }
        constexpr float ln1000 = 3 * std::numbers::ln10_v<float>;
        std::string s = "{\n";
        const auto scalars = [&](const char *key, auto &&value_at, size_t count) {
            s += "  \"", s += key, s += "\": [";
            for (size_t i = 0; i < count; ++i) {
                if (i) s += ',';
                AppendNumber(s, value_at(i));
            }
            s += "],\n";
        };
        const auto triple = [&](const float *v, bool comma) {
            if (comma) s += ',';
            s += '[', AppendNumber(s, v[0]), s += ',', AppendNumber(s, v[1]), s += ',', AppendNumber(s, v[2]), s += ']';
        };
        scalars("frequencies", [&](size_t i) { return r->Modes.Freqs[i]; }, modes);
        scalars("decayRates", [&](size_t i) { return r->Modes.T60s[i] > 0 ? ln1000 / r->Modes.T60s[i] : 0.f; }, modes);
        s += "  \"positions\": [";
        for (uint32_t i = 0; i < points; ++i) triple(&r->Modes.Positions[size_t(3) * i], i != 0);
        s += "],\n  \"shapes\": [";
        for (uint32_t k = 0; k < modes; ++k) // mode-major, matching the model schema
            for (uint32_t i = 0; i < points; ++i) triple(&r->Modes.Shapes[(size_t(i) * modes + k) * 3], k || i);
        s += "],\n";
        scalars("indices", [&](size_t i) { return indices[i]; }, indices.size());
        s += "  \"mass\": ", AppendNumber(s, r->MassProps.mass), s += ",\n";
        s += "  \"centerOfMass\": ", triple(r->MassProps.center_of_mass, false), s += ",\n";
        s += "  \"inertiaDiagonal\": ", triple(r->MassProps.inertia_diagonal, false), s += "\n}\n";
        auto *out = static_cast<char *>(std::malloc(s.size() + 1));
        if (!out) Fail(ME_OUT_OF_MEMORY, "host allocation failed");
        std::memcpy(out, s.c_str(), s.size() + 1);
        *json = out;
    });
}

} // extern "C"
