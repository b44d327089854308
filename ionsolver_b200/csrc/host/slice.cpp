// slice.cpp -- field read-back for visualisation: one plane of a field assembled over all local domains, and a
// minimal RGB PNG writer for it.
//
// Reference behaviour: slices are a display mode of the rasteriser -- GraphicsConfig::slice_mode (Off, X, Y, Z, ...),
// slice_x/y/z (src/lbm/graphics.rs:124-130), kernel graphics_field_slice (graphics_kernels.cl:669-706: velocity slices are
// coloured with iron_colormap(v_max * |u|), :689) -- and frames are saved as RGB PNG by draw_frame (graphics.rs:328-373).
// The rasteriser is out of scope; this file keeps the user-visible result (a coloured slice on disk) without a camera:
// one pixel per lattice cell, halo layers removed, the second in-plane axis pointing up.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "lbm.hpp"

namespace ionhost {
namespace slice {

// iron_colormap + color_from_floats, graphics_kernels.cl:412-428 and :93-95 (0xRRGGBB)
uint32_t iron_colormap(float x) {
    x = fminf(fmaxf(4.0f * (1.0f - x), 0.0f), 4.0f);
    float r = 1.0f, g = 0.0f, b = 0.0f;
    if (x < 0.66666667f) {  // white - yellow
        g = 1.0f;
        b = 1.0f - x * 1.5f;
    } else if (x < 2.0f) {  // yellow - red
        g = 1.5f - x * 0.75f;
    } else if (x < 3.0f) {  // red - violet
        r = 2.0f - x * 0.5f;
        b = x - 2.0f;
    } else {  // violet - black
        r = 2.0f - x * 0.5f;
        b = 4.0f - x;
    }
    auto ch = [](float v) { const int i = (int)fmaf(255.0f, v, 0.5f); return (uint32_t)(i < 0 ? 0 : i > 255 ? 255 : i); };
    return ch(r) << 16 | ch(g) << 8 | ch(b);
}

// ---- PNG: 8-bit RGB, zlib stream of stored (uncompressed) deflate blocks -- no compression library needed ----
namespace {
uint32_t crc_table[256];
bool crc_ready = false;
uint32_t crc32(const uint8_t* p, size_t n, uint32_t c = 0xFFFFFFFFu) {
    if (!crc_ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t v = i;
            for (int k = 0; k < 8; k++) v = (v & 1u) ? 0xEDB88320u ^ (v >> 1) : v >> 1;
            crc_table[i] = v;
        }
        crc_ready = true;
    }
    for (size_t i = 0; i < n; i++) c = crc_table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return c;
}
void be32(std::vector<uint8_t>& o, uint32_t v) { o.push_back(v >> 24); o.push_back(v >> 16); o.push_back(v >> 8); o.push_back(v); }
void chunk(std::vector<uint8_t>& o, const char type[4], const std::vector<uint8_t>& data) {
    be32(o, (uint32_t)data.size());
    const size_t start = o.size();
    o.insert(o.end(), type, type + 4);
    o.insert(o.end(), data.begin(), data.end());
    be32(o, crc32(o.data() + start, o.size() - start) ^ 0xFFFFFFFFu);
}
}  // namespace

std::vector<uint8_t> encode_png_rgb(const uint8_t* rgb, uint32_t w, uint32_t h) {
    if (!w || !h) throw IonException(ION_ERR_INVALID, "empty image");
    std::vector<uint8_t> raw;  // scanlines: filter byte 0 + w*3 bytes
    raw.reserve((size_t)h * (1 + 3 * (size_t)w));
    for (uint32_t y = 0; y < h; y++) {
        raw.push_back(0);
        raw.insert(raw.end(), rgb + (size_t)y * w * 3, rgb + (size_t)(y + 1) * w * 3);
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0;  // adler32
    for (uint8_t v : raw) { a = (a + v) % 65521u; b = (b + a) % 65521u; }
    for (size_t off = 0; off < raw.size(); off += 65535) {
        const size_t n = raw.size() - off < 65535 ? raw.size() - off : 65535;
        z.push_back(off + n == raw.size() ? 1 : 0);
        z.push_back(n & 0xFF); z.push_back(n >> 8); z.push_back(~n & 0xFF); z.push_back((~n >> 8) & 0xFF);
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
    }
    be32(z, b << 16 | a);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    be32(ihdr, w); be32(ihdr, h);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);  // 8 bit, RGB
    chunk(out, "IHDR", ihdr);
    chunk(out, "IDAT", z);
    chunk(out, "IEND", {});
    return out;
}

// One plane of `field` over the whole lattice (cells of domains that live in other processes stay 0).
// slice_mode: SliceMode of graphics.rs (1 = X, 2 = Y, 3 = Z); index: global cell coordinate along that axis.
// Layout: X -> out[gy + gz*Ny] (w = Ny, h = Nz); Y -> out[gx + gz*Nx]; Z -> out[gx + gy*Nx].
void read(Lbm& lbm, int field, int component, uint32_t slice_mode, uint32_t index, std::vector<float>& out, uint32_t& w, uint32_t& h) {
    const LbmConfig& c = lbm.config;
    if (slice_mode < 1 || slice_mode > 3) throw IonException(ION_ERR_INVALID, "slice_mode must be 1 (X), 2 (Y) or 3 (Z)");
    const uint32_t dir = slice_mode - 1;
    const uint32_t extent = dir == 0 ? c.n_x : dir == 1 ? c.n_y : c.n_z;
    if (index >= extent) throw IonException(ION_ERR_RANGE, "slice index outside the lattice");
    w = dir == 0 ? c.n_y : c.n_x;
    h = dir == 2 ? c.n_y : c.n_z;
    out.assign((size_t)w * h, 0.0f);
    const uint32_t hx = c.d_x > 1, hy = c.d_y > 1, hz = c.d_z > 1;
    std::vector<float> local;
    for (auto& d : lbm.domains) {
        const int64_t li = (int64_t)index - (dir == 0 ? d.o_x : dir == 1 ? d.o_y : d.o_z);  // local coordinate of the plane
        const uint32_t nl = dir == 0 ? d.n_x : dir == 1 ? d.n_y : d.n_z, hl = dir == 0 ? hx : dir == 1 ? hy : hz;
        if (li < (int64_t)hl || li >= (int64_t)(nl - hl)) continue;  // plane not in this domain's interior
        local.resize(d.get_area(dir));
        check(ion_read_slice(d.dev, field, component, dir, (uint32_t)li, local.data()));
        for (uint32_t z = (dir == 2 ? (uint32_t)li : hz); z < (dir == 2 ? (uint32_t)li + 1 : d.n_z - hz); z++)
            for (uint32_t y = (dir == 1 ? (uint32_t)li : hy); y < (dir == 1 ? (uint32_t)li + 1 : d.n_y - hy); y++)
                for (uint32_t x = (dir == 0 ? (uint32_t)li : hx); x < (dir == 0 ? (uint32_t)li + 1 : d.n_x - hx); x++) {
                    const size_t a = dir == 0 ? y + (size_t)z * d.n_y : dir == 1 ? (size_t)x * d.n_z + z : x + (size_t)y * d.n_x;
                    const uint64_t gx = (uint64_t)((int64_t)x + d.o_x), gy = (uint64_t)((int64_t)y + d.o_y), gz = (uint64_t)((int64_t)z + d.o_z);
                    const size_t g = dir == 0 ? gy + gz * w : dir == 1 ? gx + gz * w : gx + gy * w;
                    out[g] = local[a];
                }
    }
}

// pixel = iron_colormap((value - v_min) / (v_max - v_min)); image row 0 is the highest coordinate of the second in-plane axis
void write_png(Lbm& lbm, int field, int component, uint32_t slice_mode, uint32_t index, float v_min, float v_max, const std::string& path) {
    std::vector<float> plane;
    uint32_t w = 0, h = 0;
    read(lbm, field, component, slice_mode, index, plane, w, h);
    const float inv = v_max != v_min ? 1.0f / (v_max - v_min) : 1.0f;
    std::vector<uint8_t> rgb((size_t)w * h * 3);
    for (uint32_t r = 0; r < h; r++)
        for (uint32_t x = 0; x < w; x++) {
            const uint32_t col = iron_colormap((plane[(size_t)(h - 1 - r) * w + x] - v_min) * inv);
            uint8_t* px = &rgb[((size_t)r * w + x) * 3];
            px[0] = col >> 16; px[1] = (col >> 8) & 0xFF; px[2] = col & 0xFF;
        }
    const std::vector<uint8_t> png = encode_png_rgb(rgb.data(), w, h);
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw IonException(ION_ERR_INVALID, "cannot open \"" + path + "\" for writing");
    const size_t n = fwrite(png.data(), 1, png.size(), f);
    fclose(f);
    if (n != png.size()) throw IonException(ION_ERR_INVALID, "short write to \"" + path + "\"");
}

}  // namespace slice
}  // namespace ionhost
