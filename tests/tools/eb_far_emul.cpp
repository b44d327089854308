// CPU check of the far-slab path of update_e_b_dynamic (ionsolver_b200/csrc/eb_fft_core.cuh: far_accumulate, far_eval, far_reduce_x,
// far_eval_x -- the very functions k_eb_far_tensors / k_eb_combine call): a set of far sources (positions like the coarse pyramid levels
// of slabs two and more below: hundreds of cells away in z), Taylor tensors per FARB^3 block of cells, evaluated at every cell of the
// block (a) in the full form and (b) reduced to a polynomial in x per row, against a double-precision direct sum of
// q r/|r|^3 and w x r/|r|^3 (sim_kernels.cl:957-983).
// usage: eb_far_emul <FARB: 8 | 4> <R_min>     prints one JSON line: largest error relative to the largest far-field magnitude
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../ionsolver_b200/csrc/eb_fft_core.cuh"

using namespace ion::ebfft;

int main(int argc, char** argv) {
    const int FARB = argc > 1 ? atoi(argv[1]) : 8;
    const float rmin = argc > 2 ? (float)atof(argv[2]) : 242.0f;
    std::vector<FarSource> src;
    unsigned seed = 12345u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (float)(seed >> 8) / 16777216.0f; };
    for (int k = 0; k < 76; k++) {  // 64 + 8 + 4 sources: levels 2, 1, 0 of the far slabs at cfg2 on 8 GPUs
        FarSource s;
        s.cx = 256.0f * rnd(); s.cy = 256.0f * rnd(); s.cz = -rmin - 600.0f * rnd();
        s.q = 0.5f + rnd(); s.wx = rnd() - 0.5f; s.wy = rnd() - 0.5f; s.wz = rnd() - 0.5f; s.pad = 0.0f;
        src.push_back(s);
    }
    double worst_full = 0.0, worst_x = 0.0, worst_form = 0.0, scale = 0.0;
    const float c = 0.5f * (float)(FARB - 1);
    for (int blk = 0; blk < 64; blk++) {  // blocks spread over a 256 x 256 x 258 slab, the lowest ones at z = 0 (closest to the sources)
        const int bx = (blk % 4) * (256 / FARB / 4), by = ((blk / 4) % 4) * (256 / FARB / 4), bz = (blk / 16) * (256 / FARB / 4);
        float t[FAR_T];
        for (int i = 0; i < FAR_T; i++) t[i] = 0.0f;
        for (const FarSource& s : src) far_accumulate((float)(bx * FARB) + c, (float)(by * FARB) + c, (float)(bz * FARB) + c, s, t);
        for (int z = 0; z < FARB; z++)
            for (int y = 0; y < FARB; y++) {
                const float dy = (float)y - c, dz = (float)z - c;
                float pa[6], pb[6], pc[6];
                for (int o = 0; o < 6; o++) {
                    const float* p = t + (o / 3) * 30;
                    const int i = o % 3;
                    const float* G = p + 3 + 3 * i;
                    const float* H = p + 12 + 6 * i;
                    far_reduce_x(p[i], G[0], G[1], G[2], H[0], H[1], H[2], H[3], H[4], H[5], dy, dz, pa[o], pb[o], pc[o]);
                }
                for (int x = 0; x < FARB; x++) {
                    const float dx = (float)x - c;
                    double E[3] = {0, 0, 0}, B[3] = {0, 0, 0};
                    const double px = bx * FARB + x, py = by * FARB + y, pz = bz * FARB + z;
                    for (const FarSource& s : src) {
                        const double r[3] = {px - s.cx, py - s.cy, pz - s.cz};
                        const double r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2], inv3 = 1.0 / (r2 * std::sqrt(r2));
                        const double w[3] = {s.wx, s.wy, s.wz};
                        for (int i = 0; i < 3; i++) {
                            const int j = (i + 1) % 3, k = (i + 2) % 3;
                            E[i] += s.q * r[i] * inv3;
                            B[i] += (w[j] * r[k] - w[k] * r[j]) * inv3;
                        }
                    }
                    for (int o = 0; o < 6; o++) {
                        const double want = o < 3 ? E[o] : B[o - 3];
                        const float full = far_eval(t + (o / 3) * 30, o % 3, dx, dy, dz);
                        const float red = far_eval_x(pa[o], pb[o], pc[o], dx);
                        scale = std::fmax(scale, std::fabs(want));
                        worst_full = std::fmax(worst_full, std::fabs((double)full - want));
                        worst_x = std::fmax(worst_x, std::fabs((double)red - want));
                        worst_form = std::fmax(worst_form, std::fabs((double)red - (double)full));
                    }
                }
            }
    }
    printf("{\"farb\": %d, \"r_min\": %.1f, \"sources\": %zu, \"field_scale\": %.4e, \"taylor_err_rel\": %.3e, \"x_reduced_err_rel\": %.3e, \"forms_differ_rel\": %.3e}\n",
           FARB, rmin, src.size(), scale, worst_full / scale, worst_x / scale, worst_form / scale);
    return 0;
}
