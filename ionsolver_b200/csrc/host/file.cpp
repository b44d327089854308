// file.cpp -- the on-disk formats either side of the hot path (/root/reference/src/file.rs, FILE_LAYOUT.txt):
// the little-endian `.ion` state snapshot and the serde_json dump of LbmConfig.
//
// `.ion`: 16-byte magic, 62-byte config block, then sections flags[N] u8, rho[N] f32, u[3N] f32 (x,y,z planes),
// charge[N] f32 (MHD only), N_C + 12*N_C fixed charges, N_M + 20*N_M magnets.  DDFs are not stored: a loaded state is
// re-initialised from rho/u/Q (file.rs:42-188, main.rs:274-278).
//
// The reference's encoder and decoder disagree with each other and with FILE_LAYOUT.txt; both behaviours
// are kept behind `reference_compatible`:
//   true  = bit-compatible with what file.rs really does: the encoder writes the float_type discriminant
//           (FP16S=0, FP16C=1) and never writes N_C/N_M; the decoder maps 0->FP16C, 1->FP16S and, for MHD files,
//           expects N_C/N_M (so it rejects the reference's own MHD files, where the Rust code panics); per-domain
//           sections are the first N/d elements of each domain buffer, halos included (file.rs:109,227).
//   false = FILE_LAYOUT.txt as written: float_type is the types.rs discriminant on both sides, N_C = N_M = 0 are
//           always present for MHD, and sections are in GLOBAL lattice order without halos, independent of the
//           domain split.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "lbm.hpp"

namespace ionhost {
namespace file {

namespace {
struct ByteBuffer {  // file.rs:338-377
    std::vector<uint8_t> b;
    void push(uint8_t x) { b.push_back(x); }
    void push32(uint32_t x) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)((x >> (i * 8)) & 0xFF)); }
    void push64(uint64_t x) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)((x >> (i * 8)) & 0xFF)); }
    void pushf(float f) { uint32_t u; memcpy(&u, &f, 4); push32(u); }
    void pushname() { const char* n = "IonSolver setup\n"; for (int i = 0; i < 16; i++) b.push_back((uint8_t)n[i]); }
};
struct ByteStream {  // file.rs:379-433
    const std::vector<uint8_t>& buf;
    size_t pos = 0;
    explicit ByteStream(const std::vector<uint8_t>& v) : buf(v) {}
    void need(size_t n) const { if (pos + n > buf.size()) throw IonException(ION_ERR_RANGE, "Not all data could be read, file may be corrupted (truncated)."); }
    uint8_t next_u8() { need(1); return buf[pos++]; }
    uint32_t next_u32() { need(4); uint32_t v = 0; for (int i = 0; i < 4; i++) v += (uint32_t)buf[pos++] << (i * 8); return v; }
    uint64_t next_u64() { need(8); uint64_t v = 0; for (int i = 0; i < 8; i++) v += (uint64_t)buf[pos++] << (i * 8); return v; }
    float next_f32() { uint32_t u = next_u32(); float f; memcpy(&f, &u, 4); return f; }
    bool at_end() const { return pos == buf.size(); }
};

// local-domain index -> global cell index (interior cells only)
template <typename F> void for_interior(const Lbm& lbm, const LbmDomain& d, F f) {
    const LbmConfig& c = lbm.config;
    const uint32_t hx = c.d_x > 1, hy = c.d_y > 1, hz = c.d_z > 1;
    for (uint32_t z = hz; z < d.n_z - hz; z++)
        for (uint32_t y = hy; y < d.n_y - hy; y++)
            for (uint32_t x = hx; x < d.n_x - hx; x++) {
                const uint64_t local = x + (y + (uint64_t)z * d.n_y) * d.n_x;
                const uint64_t gx = (uint64_t)((int64_t)x + d.o_x), gy = (uint64_t)((int64_t)y + d.o_y), gz = (uint64_t)((int64_t)z + d.o_z);
                f(local, gx + (gy + gz * c.n_y) * c.n_x);
            }
}
}  // namespace

std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw IonException(ION_ERR_INVALID, "Could not find file \"" + path + "\"");
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}

std::vector<uint8_t> encode(Lbm& lbm, bool reference_compatible) {  // file.rs:191-306
    if (lbm.world != 1) throw IonException(ION_ERR_UNSUPPORTED, "encode needs all domains in one process");
    const LbmConfig& c = lbm.config;
    ByteBuffer out;
    out.pushname();
    out.push((uint8_t)c.velocity_set);
    out.push((uint8_t)c.relaxation_time);
    out.push((uint8_t)c.float_type);
    out.pushf(c.units.m); out.pushf(c.units.kg); out.pushf(c.units.s); out.pushf(c.units.a);
    out.push32(c.n_x); out.push32(c.n_y); out.push32(c.n_z);
    out.push(0);  // fixed domain sizes
    out.push32(c.d_x); out.push32(c.d_y); out.push32(c.d_z);
    out.pushf(c.nu);
    out.pushf(c.f_x); out.pushf(c.f_y); out.pushf(c.f_z);
    out.push((uint8_t)((c.ext_equilibrium_boudaries ? 1 : 0) + ((c.ext_volume_force ? 1 : 0) << 1) + ((c.ext_force_field ? 1 : 0) << 2) + ((c.ext_magneto_hydro ? 1 : 0) << 3)));
    out.push(c.mhd_lod_depth);

    const uint64_t N = (uint64_t)c.n_x * c.n_y * c.n_z;
    const uint32_t domain_count = c.d_x * c.d_y * c.d_z;
    lbm.finish_queues();
    if (reference_compatible) {
        const uint64_t d_n = N / domain_count;  // file.rs:227: first d_n elements of every domain buffer
        for (auto& d : lbm.domains) { std::vector<uint8_t> t(d_n); d.read(ION_FIELD_FLAGS, t.data(), d_n); out.b.insert(out.b.end(), t.begin(), t.end()); }
        for (auto& d : lbm.domains) { std::vector<float> t(d_n); d.read(ION_FIELD_RHO, t.data(), d_n * 4); for (float v : t) out.pushf(v); }
        for (auto& d : lbm.domains) { std::vector<float> t(d_n * 3); d.read(ION_FIELD_U, t.data(), d_n * 12); for (float v : t) out.pushf(v); }
        if (c.ext_magneto_hydro)
            for (auto& d : lbm.domains) { std::vector<float> t(d_n); d.read(ION_FIELD_Q, t.data(), d_n * 4); for (float v : t) out.pushf(v); }
        return out.b;  // N_C / N_M are not written (commented out at file.rs:277-303)
    }
    std::vector<uint8_t> flags(N);
    std::vector<float> rho(N), u(3 * N), q(c.ext_magneto_hydro ? N : 0);
    for (auto& d : lbm.domains) {
        std::vector<uint8_t> lf(d.n);
        std::vector<float> lr(d.n), lu(3 * d.n), lq(c.ext_magneto_hydro ? d.n : 0);
        d.read(ION_FIELD_FLAGS, lf.data(), d.n);
        d.read(ION_FIELD_RHO, lr.data(), d.n * 4);
        d.read(ION_FIELD_U, lu.data(), d.n * 12);
        if (c.ext_magneto_hydro) d.read(ION_FIELD_Q, lq.data(), d.n * 4);
        for_interior(lbm, d, [&](uint64_t l, uint64_t g) {
            flags[g] = lf[l]; rho[g] = lr[l];
            u[g] = lu[l]; u[N + g] = lu[d.n + l]; u[2 * N + g] = lu[2 * d.n + l];
            if (c.ext_magneto_hydro) q[g] = lq[l];
        });
    }
    out.b.insert(out.b.end(), flags.begin(), flags.end());
    for (float v : rho) out.pushf(v);
    for (float v : u) out.pushf(v);
    if (c.ext_magneto_hydro) {
        for (float v : q) out.pushf(v);
        out.push32(0);  // N_C
        out.push32(0);  // N_M
    }
    return out.b;
}

Lbm* decode(const std::vector<uint8_t>& buffer, LbmConfig& config, bool reference_compatible, const std::vector<int>& devices) {  // file.rs:42-188
    ByteStream st(buffer);
    std::string header;
    for (int i = 0; i < 16; i++) header.push_back((char)st.next_u8());
    if (header != "IonSolver setup\n") throw IonException(ION_ERR_INVALID, "Invalid Format!");
    switch (st.next_u8()) { case 1: config.velocity_set = VelocitySet::D3Q15; break; case 2: config.velocity_set = VelocitySet::D3Q19; break;
                            case 3: config.velocity_set = VelocitySet::D3Q27; break; default: config.velocity_set = VelocitySet::D2Q9; }
    config.relaxation_time = st.next_u8() == 1 ? RelaxationTime::Trt : RelaxationTime::Srt;
    const uint8_t ft = st.next_u8();
    if (reference_compatible) config.float_type = ft == 1 ? FloatType::FP16S : ft == 2 ? FloatType::FP32 : FloatType::FP16C;  // file.rs:68-73
    else config.float_type = ft == 1 ? FloatType::FP16C : ft == 2 ? FloatType::FP32 : FloatType::FP16S;                       // types.rs:78-82
    config.units.m = st.next_f32(); config.units.kg = st.next_f32(); config.units.s = st.next_f32(); config.units.a = st.next_f32();
    config.n_x = st.next_u32(); config.n_y = st.next_u32(); config.n_z = st.next_u32();
    st.next_u8();  // fixed domains are skipped
    config.d_x = st.next_u32(); config.d_y = st.next_u32(); config.d_z = st.next_u32();
    config.nu = st.next_f32();
    config.f_x = st.next_f32(); config.f_y = st.next_f32(); config.f_z = st.next_f32();
    const uint8_t ext = st.next_u8();
    config.ext_equilibrium_boudaries = (ext & 0x1) != 0;
    config.ext_volume_force = (ext & 0x2) != 0;
    config.ext_force_field = (ext & 0x4) != 0;
    config.ext_magneto_hydro = (ext & 0x8) != 0;
    config.mhd_lod_depth = st.next_u8();
    if (!config.d_x || !config.d_y || !config.d_z) throw IonException(ION_ERR_INVALID, "zero domain count in file");
    // validate the payload size BEFORE allocating device memory
    const uint64_t N = (uint64_t)config.n_x * config.n_y * config.n_z;
    const uint64_t need = N * (1 + 4 + 12 + (config.ext_magneto_hydro ? 4 : 0));
    if (buffer.size() - st.pos < need) throw IonException(ION_ERR_RANGE, "Not all data could be read, file may be corrupted (truncated).");

    Lbm* lbm = Lbm::create(config, devices);
    try {
        const LbmConfig& c = lbm->config;
        const uint32_t d_total = c.d_x * c.d_y * c.d_z;
        const bool mhd = c.ext_magneto_hydro;
        if (reference_compatible) {
            const uint64_t d_n = (uint64_t)(c.n_x / c.d_x) * (c.n_y / c.d_y) * (c.n_z / c.d_z);  // file.rs:109
            for (uint32_t d = 0; d < d_total; d++) { lbm->domains[d].write(ION_FIELD_FLAGS, &buffer[st.pos], d_n); st.pos += d_n; }
            for (uint32_t d = 0; d < d_total; d++) { lbm->domains[d].write(ION_FIELD_RHO, &buffer[st.pos], d_n * 4); st.pos += d_n * 4; }
            for (uint32_t d = 0; d < d_total; d++) { lbm->domains[d].write(ION_FIELD_U, &buffer[st.pos], d_n * 12); st.pos += d_n * 12; }
            if (mhd) for (uint32_t d = 0; d < d_total; d++) { lbm->domains[d].write(ION_FIELD_Q, &buffer[st.pos], d_n * 4); st.pos += d_n * 4; }
        } else {
            const uint8_t* flags = &buffer[st.pos];
            const uint8_t* rho = flags + N;
            const uint8_t* u = rho + 4 * N;
            const uint8_t* q = u + 12 * N;
            st.pos += need;
            for (auto& d : lbm->domains) {
                std::vector<uint8_t> lf(d.n, 0);
                std::vector<float> lr(d.n, 1.0f), lu(3 * d.n, 0.0f), lq(mhd ? d.n : 0, 0.0f);
                for_interior(*lbm, d, [&](uint64_t l, uint64_t g) {
                    lf[l] = flags[g];
                    memcpy(&lr[l], rho + 4 * g, 4);
                    memcpy(&lu[l], u + 4 * g, 4); memcpy(&lu[d.n + l], u + 4 * (N + g), 4); memcpy(&lu[2 * d.n + l], u + 4 * (2 * N + g), 4);
                    if (mhd) memcpy(&lq[l], q + 4 * g, 4);
                });
                d.write(ION_FIELD_FLAGS, lf.data(), d.n);
                d.write(ION_FIELD_RHO, lr.data(), d.n * 4);
                d.write(ION_FIELD_U, lu.data(), d.n * 12);
                if (mhd) d.write(ION_FIELD_Q, lq.data(), d.n * 4);
            }
        }
        if (mhd && !reference_compatible && st.at_end()) {
            // a single-domain MHD file written by the reference's encoder ends here: file.rs:277-303 never writes N_C / N_M
        } else if (mhd) {
            const uint32_t n_charges = st.next_u32();  // read and discarded (lbm.charges is commented out, file.rs:166)
            for (uint32_t i = 0; i < n_charges; i++) { st.next_u64(); st.next_f32(); }
            const uint32_t n_magnets = st.next_u32();
            for (uint32_t i = 0; i < n_magnets; i++) { st.next_u64(); st.next_f32(); st.next_f32(); st.next_f32(); }
            if (!st.at_end()) throw IonException(ION_ERR_RANGE, "Not all data could be read, file may be corrupted.");
        }
        // non-MHD: the reference returns without the trailing-bytes check (file.rs:138-140)
    } catch (...) {
        delete lbm;
        throw;
    }
    return lbm;
}

// Default on-disk mode = FILE_LAYOUT.txt as written (reference_compatible = false): a file written here reloads to the same
// state for every storage codec, with MHD and with any domain split, and the reference's decoder accepts it too (it gets the
// N_C = N_M = 0 trailer its own encoder forgets).  Single-domain files written by the reference load as well; only its
// FP16S / FP16C discriminant swap (file.rs:68-73 vs :199) is NOT reproduced -- bug compatibility is the explicit opt-in.
void write(Lbm& lbm, const std::string& path, bool reference_compatible) {
    const std::vector<uint8_t> b = encode(lbm, reference_compatible);
    std::ofstream out(path, std::ios::binary);
    if (!out.write((const char*)b.data(), (std::streamsize)b.size())) throw IonException(ION_ERR_INVALID, "writing went wrong");
}
Lbm* read(const std::string& path, LbmConfig& config, bool reference_compatible) { return decode(read_file(path), config, reference_compatible, {}); }

// ---------------------------------------------------------------------------------------------------------------
// JSON (serde_json of LbmConfig, file.rs:310-334)
// ---------------------------------------------------------------------------------------------------------------
namespace {
std::string f32_json(float v) {  // serde_json prints the shortest round-trip f32 (ryu); any round-tripping text parses equal
    if (std::isnan(v) || std::isinf(v)) return "null";
    char buf[64];
    for (int prec = 1; prec <= 9; prec++) {
        snprintf(buf, sizeof(buf), "%.*g", prec, (double)v);
        if (strtof(buf, nullptr) == v) break;
    }
    std::string s(buf);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    return s;
}
const char* VS_NAMES[4] = {"D2Q9", "D3Q15", "D3Q19", "D3Q27"};
const char* RT_NAMES[2] = {"Srt", "Trt"};
const char* FT_NAMES[3] = {"FP16S", "FP16C", "FP32"};
const char* PROP_NAMES[6] = {"H", "He", "Ne", "Ar", "Kr", "Xe"};
const char* GRAPHICS_DEFAULT_REST =  // GraphicsConfig::new, graphics.rs:158-193 (everything after graphics_active)
    "\"background_color\":0,\"camera_width\":1920,\"camera_height\":1080,\"v_max\":0.25,\"q_min\":0.0001,\"f_max\":0.002,\"rho_delta\":0.5,"
    "\"streamline_every\":4,\"stream_line_lenght\":128,\"vec_vis_mode\":\"U\",\"slice_mode\":\"Off\",\"slice_x\":0,\"slice_y\":0,\"slice_z\":0,"
    "\"max_slice_x\":0,\"max_slice_y\":0,\"max_slice_z\":0,\"field_vis\":\"Vector\",\"streamline_mode\":false,\"field_mode\":false,"
    "\"field_slice_mode\":false,\"q_mode\":false,\"q_field_mode\":false,\"flags_mode\":false,\"flags_surface_mode\":false,\"axes_mode\":false,"
    "\"ecrc_mode\":false,\"render_intervals\":false,\"keyframes\":[]";

// minimal JSON reader: objects, strings, numbers, booleans, null, arrays (arrays/objects can be captured as raw text)
struct Json {
    const std::string& s;
    size_t p = 0;
    explicit Json(const std::string& t) : s(t) {}
    void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\t' || s[p] == '\r')) p++; }
    [[noreturn]] void bad(const char* what) { throw IonException(ION_ERR_INVALID, std::string("Could not parse file: ") + what + " at byte " + std::to_string(p)); }
    void expect(char c) { ws(); if (p >= s.size() || s[p] != c) bad("unexpected character"); p++; }
    bool peek(char c) { ws(); return p < s.size() && s[p] == c; }
    std::string str() {
        expect('"');
        std::string o;
        while (p < s.size() && s[p] != '"') { if (s[p] == '\\' && p + 1 < s.size()) p++; o.push_back(s[p++]); }
        if (p >= s.size()) bad("unterminated string");
        p++;
        return o;
    }
    double num() {
        ws();
        char* e = nullptr;
        const double v = strtod(s.c_str() + p, &e);
        if (e == s.c_str() + p) bad("number expected");
        p = (size_t)(e - s.c_str());
        return v;
    }
    bool boolean() {
        ws();
        if (s.compare(p, 4, "true") == 0) { p += 4; return true; }
        if (s.compare(p, 5, "false") == 0) { p += 5; return false; }
        bad("boolean expected");
    }
    std::string raw_value() {  // skip one value of any type, return its text
        ws();
        const size_t b = p;
        if (peek('"')) { str(); return s.substr(b, p - b); }
        if (peek('{') || peek('[')) {
            int depth = 0;
            bool in_str = false;
            for (; p < s.size(); p++) {
                const char c = s[p];
                if (in_str) { if (c == '\\') p++; else if (c == '"') in_str = false; continue; }
                if (c == '"') in_str = true;
                else if (c == '{' || c == '[') depth++;
                else if (c == '}' || c == ']') { depth--; if (depth == 0) { p++; break; } }
            }
            return s.substr(b, p - b);
        }
        while (p < s.size() && s[p] != ',' && s[p] != '}' && s[p] != ']') p++;
        return s.substr(b, p - b);
    }
    template <typename F> void object(F f) {
        expect('{');
        if (peek('}')) { p++; return; }
        for (;;) {
            const std::string key = str();
            expect(':');
            f(key);
            ws();
            if (peek(',')) { p++; continue; }
            expect('}');
            break;
        }
    }
};
template <size_t K> int enum_index(const char* (&names)[K], const std::string& v, Json& j) {
    for (size_t i = 0; i < K; i++) if (v == names[i]) return (int)i;
    j.bad("unknown enum variant");
}
}  // namespace

std::string config_to_json(const LbmConfig& c) {
    std::ostringstream o;
    o << "{\"velocity_set\":\"" << VS_NAMES[(int)c.velocity_set] << "\",\"relaxation_time\":\"" << RT_NAMES[(int)c.relaxation_time]
      << "\",\"float_type\":\"" << FT_NAMES[(int)c.float_type] << "\",\"units\":{\"m\":" << f32_json(c.units.m) << ",\"kg\":" << f32_json(c.units.kg)
      << ",\"s\":" << f32_json(c.units.s) << ",\"a\":" << f32_json(c.units.a) << ",\"k\":" << f32_json(c.units.k) << ",\"prop\":\""
      << PROP_NAMES[(int)c.units.prop] << "\"},\"n_x\":" << c.n_x << ",\"n_y\":" << c.n_y << ",\"n_z\":" << c.n_z << ",\"d_x\":" << c.d_x
      << ",\"d_y\":" << c.d_y << ",\"d_z\":" << c.d_z << ",\"nu\":" << f32_json(c.nu) << ",\"f_x\":" << f32_json(c.f_x) << ",\"f_y\":" << f32_json(c.f_y)
      << ",\"f_z\":" << f32_json(c.f_z) << ",\"ext_equilibrium_boudaries\":" << (c.ext_equilibrium_boudaries ? "true" : "false")
      << ",\"ext_volume_force\":" << (c.ext_volume_force ? "true" : "false") << ",\"ext_force_field\":" << (c.ext_force_field ? "true" : "false")
      << ",\"ext_magneto_hydro\":" << (c.ext_magneto_hydro ? "true" : "false") << ",\"ext_subgrid_ecr\":" << (c.ext_subgrid_ecr ? "true" : "false")
      << ",\"mhd_lod_depth\":" << (unsigned)c.mhd_lod_depth << ",\"ecr_freq\":" << f32_json(c.ecr_freq) << ",\"ecr_field_strength\":"
      << f32_json(c.ecr_field_strength) << ",\"graphics_config\":{\"graphics_active\":" << (c.graphics_config.graphics_active ? "true" : "false") << ","
      << (c.graphics_config.passthrough_json.empty() ? std::string(GRAPHICS_DEFAULT_REST) : c.graphics_config.passthrough_json) << "},\"run_steps\":"
      << c.run_steps << "}";
    return o.str();
}

LbmConfig config_from_json(const std::string& text) {
    LbmConfig c;
    Json j(text);
    j.object([&](const std::string& key) {
        if (key == "velocity_set") c.velocity_set = (VelocitySet)enum_index(VS_NAMES, j.str(), j);
        else if (key == "relaxation_time") c.relaxation_time = (RelaxationTime)enum_index(RT_NAMES, j.str(), j);
        else if (key == "float_type") c.float_type = (FloatType)enum_index(FT_NAMES, j.str(), j);
        else if (key == "units") j.object([&](const std::string& k) {
            if (k == "m") c.units.m = (float)j.num(); else if (k == "kg") c.units.kg = (float)j.num(); else if (k == "s") c.units.s = (float)j.num();
            else if (k == "a") c.units.a = (float)j.num(); else if (k == "k") c.units.k = (float)j.num();
            else if (k == "prop") c.units.prop = (Propellant)enum_index(PROP_NAMES, j.str(), j); else j.raw_value();
        });
        else if (key == "n_x") c.n_x = (uint32_t)j.num(); else if (key == "n_y") c.n_y = (uint32_t)j.num(); else if (key == "n_z") c.n_z = (uint32_t)j.num();
        else if (key == "d_x") c.d_x = (uint32_t)j.num(); else if (key == "d_y") c.d_y = (uint32_t)j.num(); else if (key == "d_z") c.d_z = (uint32_t)j.num();
        else if (key == "nu") c.nu = (float)j.num();
        else if (key == "f_x") c.f_x = (float)j.num(); else if (key == "f_y") c.f_y = (float)j.num(); else if (key == "f_z") c.f_z = (float)j.num();
        else if (key == "ext_equilibrium_boudaries") c.ext_equilibrium_boudaries = j.boolean();
        else if (key == "ext_volume_force") c.ext_volume_force = j.boolean();
        else if (key == "ext_force_field") c.ext_force_field = j.boolean();
        else if (key == "ext_magneto_hydro") c.ext_magneto_hydro = j.boolean();
        else if (key == "ext_subgrid_ecr") c.ext_subgrid_ecr = j.boolean();
        else if (key == "mhd_lod_depth") c.mhd_lod_depth = (uint8_t)j.num();
        else if (key == "ecr_freq") c.ecr_freq = (float)j.num();
        else if (key == "ecr_field_strength") c.ecr_field_strength = (float)j.num();
        else if (key == "run_steps") c.run_steps = (uint64_t)j.num();
        else if (key == "graphics_config") {
            std::string rest;
            j.object([&](const std::string& k) {
                if (k == "graphics_active") c.graphics_config.graphics_active = j.boolean();
                else { if (!rest.empty()) rest += ","; rest += "\"" + k + "\":" + j.raw_value(); }
            });
            c.graphics_config.passthrough_json = rest;
        } else j.raw_value();
    });
    return c;
}

void write_config(const std::string& path, const LbmConfig& cfg) {
    std::ofstream out(path, std::ios::binary);
    const std::string s = config_to_json(cfg);
    if (!out.write(s.data(), (std::streamsize)s.size())) throw IonException(ION_ERR_INVALID, "Could not write file");
}
LbmConfig read_config(const std::string& path) {
    const std::vector<uint8_t> b = read_file(path);
    return config_from_json(std::string(b.begin(), b.end()));
}

}  // namespace file
}  // namespace ionhost
