// mesh.cpp -- binary-STL import and mesh voxelisation driver (/root/reference/src/mesh.rs).  All vector / matrix
// arithmetic is f32 with the operand order of mesh.rs:345-429, so the triangle coordinates handed to the voxeliser
// (and therefore the flags it writes) equal the reference's bit for bit.
#include <cmath>
#include <cstring>
#include <fstream>

#include "lbm.hpp"

namespace ionhost {

static const float PI_F = 3.14159274101257324219f;
static inline F32_3 v3(float x, float y, float z) { F32_3 v; v.x = x; v.y = y; v.z = z; return v; }
static inline F32_3 add(F32_3 a, F32_3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline F32_3 sub(F32_3 a, F32_3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline F32_3 mul(float s, F32_3 b) { return v3(s * b.x, s * b.y, s * b.z); }
static inline F32_3 mul(const F32_3_3& m, F32_3 v) {  // mesh.rs:391-402
    return v3(m.xx * v.x + m.xy * v.y + m.xz * v.z, m.yx * v.x + m.yy * v.y + m.yz * v.z, m.zx * v.x + m.zy * v.y + m.zz * v.z);
}
static inline F32_3_3 mul(const F32_3_3& a, const F32_3_3& m) {  // mesh.rs:418-428
    F32_3_3 r;
    r.xx = a.xx * m.xx + a.xy * m.yx + a.xz * m.zx; r.xy = a.xx * m.xy + a.xy * m.yy + a.xz * m.zy; r.xz = a.xx * m.xz + a.xy * m.yz + a.xz * m.zz;
    r.yx = a.yx * m.xx + a.yy * m.yx + a.yz * m.zx; r.yy = a.yx * m.xy + a.yy * m.yy + a.yz * m.zy; r.yz = a.yx * m.xz + a.yy * m.yz + a.yz * m.zz;
    r.zx = a.zx * m.xx + a.zy * m.yx + a.zz * m.zx; r.zy = a.zx * m.xy + a.zy * m.yy + a.zz * m.zy; r.zz = a.zx * m.xz + a.zy * m.yz + a.zz * m.zz;
    return r;
}
static F32_3_3 rotm_around_v(F32_3 v, float r) {  // mesh.rs:58-66
    const float sr = sinf(r), cr = cosf(r);
    auto sq = [](float x) { return x * x; };
    F32_3_3 m;
    m.xx = sq(v.x) + (1.0f - sq(v.x)) * cr; m.xy = v.x * v.y * (1.0f - cr) - v.z * sr; m.xz = v.x * v.z * (1.0f - cr) + v.y * sr;
    m.yx = v.x * v.y * (1.0f - cr) + v.z * sr; m.yy = sq(v.y) + (1.0f - sq(v.y)) * cr; m.yz = v.y * v.z * (1.0f - cr) - v.x * sr;
    m.zx = v.x * v.z * (1.0f - cr) - v.y * sr; m.zy = v.y * v.z * (1.0f - cr) + v.x * sr; m.zz = sq(v.z) + (1.0f - sq(v.z)) * cr;
    return m;
}
F32_3_3 construct_rotation_matrix(float rx, float ry, float rz) {  // mesh.rs:52-56
    return mul(mul(rotm_around_v(v3(1, 0, 0), rx), rotm_around_v(v3(0, 1, 0), ry)), rotm_around_v(v3(0, 0, 1), rz));
}

void Mesh::update_bounds() {  // mesh.rs:93-108
    p_min = p0[0];
    p_max = p0[0];
    for (uint32_t i = 0; i < triangle_number; i++) {
        const F32_3 a = p0[i], b = p1[i], c = p2[i];
        p_min.x = fminf(p_min.x, fminf(c.x, fminf(b.x, a.x)));
        p_min.y = fminf(p_min.y, fminf(c.y, fminf(b.y, a.y)));
        p_min.z = fminf(p_min.z, fminf(c.z, fminf(b.z, a.z)));
        p_max.x = fmaxf(p_max.x, fmaxf(c.x, fmaxf(b.x, a.x)));
        p_max.y = fmaxf(p_max.y, fmaxf(c.y, fmaxf(b.y, a.y)));
        p_max.z = fmaxf(p_max.z, fmaxf(c.z, fmaxf(b.z, a.z)));
    }
}
void Mesh::scale(float s) {  // mesh.rs:110-118
    for (uint32_t i = 0; i < triangle_number; i++) {
        p0[i] = add(mul(s, sub(p0[i], center)), center);
        p1[i] = add(mul(s, sub(p1[i], center)), center);
        p2[i] = add(mul(s, sub(p2[i], center)), center);
    }
    p_min = add(mul(s, sub(p_min, center)), center);
    p_max = add(mul(s, sub(p_max, center)), center);
}
void Mesh::translate(F32_3 t) {  // mesh.rs:120-129
    for (uint32_t i = 0; i < triangle_number; i++) {
        p0[i] = add(p0[i], t); p1[i] = add(p1[i], t); p2[i] = add(p2[i], t);
    }
    center = add(center, t);
    p_min = add(p_min, t);
    p_max = add(p_max, t);
}
void Mesh::rotate(const F32_3_3& r) {  // mesh.rs:131-138
    for (uint32_t i = 0; i < triangle_number; i++) {
        p0[i] = add(mul(r, sub(p0[i], center)), center);
        p1[i] = add(mul(r, sub(p1[i], center)), center);
        p2[i] = add(mul(r, sub(p2[i], center)), center);
    }
    update_bounds();
}

Mesh Mesh::read_stl_raw(const std::vector<uint8_t>& f, bool reposition, F32_3 box_size, F32_3 center, const F32_3_3& rotation, float size) {
    if (f.size() < 84) throw IonException(ION_ERR_INVALID, "Mesh import failed: file shorter than an STL header");
    uint32_t tn;
    memcpy(&tn, &f[80], 4);
    // mesh.rs:179-184 compares in 32 bits (and panics on the out-of-bounds read that a wrapped product would cause); here the
    // comparison is done in 64 bits so that a crafted triangle count cannot pass the check and drive reads past the buffer
    if (!(tn > 0 && (uint64_t)f.size() == 84ull + 50ull * (uint64_t)tn))
        throw IonException(ION_ERR_INVALID, "Mesh import failed: corrupted or unsupported file (only binary .stl)");
    Mesh mesh;
    mesh.triangle_number = tn;
    mesh.center = center;
    mesh.p0.resize(tn); mesh.p1.resize(tn); mesh.p2.resize(tn);
    size_t pos = 84;
    auto next3 = [&]() { float v[3]; memcpy(v, &f[pos], 12); pos += 12; return v3(v[0], v[1], v[2]); };
    for (uint32_t i = 0; i < tn; i++) {
        if (pos + 50 > f.size()) throw IonException(ION_ERR_RANGE, "Mesh import failed: triangle data past the end of the file");
        pos += 12;  // normal
        mesh.p0[i] = mul(rotation, next3());
        mesh.p1[i] = mul(rotation, next3());
        mesh.p2[i] = mul(rotation, next3());
        pos += 2;  // attribute bits
    }
    mesh.update_bounds();
    float scale;
    if (size == 0.0f) {  // get_scale_for_box_fit, mesh.rs:165-167
        scale = fminf(box_size.x / (mesh.p_max.x - mesh.p_min.x), fminf(box_size.y / (mesh.p_max.y - mesh.p_min.y), box_size.z / (mesh.p_max.z - mesh.p_min.z)));
    } else if (size > 0.0f) {  // size / get_max_size
        scale = size / fmaxf(mesh.p_max.x - mesh.p_min.x, fmaxf(mesh.p_max.y - mesh.p_min.y, mesh.p_max.z - mesh.p_min.z));
    } else {
        scale = -size;
    }
    const F32_3 offset = reposition ? mul(-0.5f, add(mesh.p_min, mesh.p_max)) : v3(0, 0, 0);
    for (uint32_t i = 0; i < tn; i++) {
        mesh.p0[i] = add(center, mul(scale, add(offset, mesh.p0[i])));
        mesh.p1[i] = add(center, mul(scale, add(offset, mesh.p1[i])));
        mesh.p2[i] = add(center, mul(scale, add(offset, mesh.p2[i])));
    }
    mesh.update_bounds();
    return mesh;
}

void Lbm::import_mesh(const std::string& path, float scale, float ox, float oy, float oz, float rx, float ry, float rz) {  // mesh.rs:233-239
    const F32_3_3 rot = construct_rotation_matrix(rx * PI_F / 180.0f, ry * PI_F / 180.0f, rz * PI_F / 180.0f);
    const float scale_lu = config.units.len_si_lu(scale);
    meshes.push_back(Mesh::read_stl_raw(file::read_file(path), false, v3(1, 1, 1), v3(ox, oy, oz), rot, -fabsf(scale_lu)));
}
void Lbm::import_mesh_reposition(const std::string& path, float cx, float cy, float cz, float rx, float ry, float rz, float size) {  // mesh.rs:248-253
    const F32_3_3 rot = construct_rotation_matrix(rx * PI_F / 180.0f, ry * PI_F / 180.0f, rz * PI_F / 180.0f);
    meshes.push_back(Mesh::read_stl_raw(file::read_file(path), true, v3((float)config.n_x, (float)config.n_y, (float)config.n_z), v3(cx, cy, cz), rot, size));
}
void Lbm::voxelise_mesh(size_t index, const ModelType& ctype) {  // mesh.rs:256-278
    if (index >= meshes.size()) throw IonException(ION_ERR_INVALID, "mesh index out of range");
    for (auto& d : domains) d.voxelize_mesh_on_device(meshes[index], ctype);
    finish_queues();
}

}  // namespace ionhost
