// CPU emulation of the polyphase-FFT update_e_b_dynamic kernels (ionsolver_b200/csrc/eb_fft_core.cuh): runs the very phase
// functions the CUDA kernels call, thread by thread, and compares E/B with a direct double-precision evaluation of the
// reference's own-LOD loop (sim_kernels.cl:940-955).  Test infrastructure; built and run by tests/test_host_logic.py.
//   g++ -O2 -std=c++17 -ffp-contract=off -I/usr/local/cuda/include tests/tools/eb_fft_emul.cpp -o eb_fft_emul
//   eb_fft_emul <depth 3|4> <nx> <ny> <nz> <dz> [mirror: 1 = tasks with 2 ox > dsx read their partner's kernel spectra]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include "../../ionsolver_b200/csrc/eb_fft_core.cuh"

using namespace ion::ebfft;

template <int ND> static int run(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t dz, bool mirror) {
    typedef Cfg<ND> C;
    const uint32_t depth = ND == 16 ? 4 : 3;
    uint32_t n_lod_own = 0;
    for (uint32_t i = 0; i <= depth; i++) n_lod_own += 1u << (3 * i);
    Geom g{};
    g.nx = nx; g.ny = ny; g.nz = nz; g.N = (uint64_t)nx * ny * nz;
    g.dsx = nx / ND; g.dsy = ny / ND; g.dsz = nz / ND;
    g.n_lod_own = n_lod_own; g.lo = n_lod_own - ND * ND * ND; g.cz0 = g.lo / (ND * ND);
    g.dx = 1; g.dy = 1; g.dz = dz; g.ke = 1.5f; g.kmu = 0.25f;
    constexpr int NF = ND / 2;
    // second source set: the level D-1 pyramid of the slab below (only when this slab has one: dz > 1 and block sizes commensurate)
    const bool foreign = dz > 1 && nx / NF == 2 * g.dsx && ny / NF == 2 * g.dsy && nz / NF == 2 * g.dsz;
    SourceSet sets[2];
    sets[0] = SourceSet{0u, 0u, -(int32_t)g.cz0, -0.5f * g.dsx, -0.5f * g.dsy, -0.5f * g.dsz};
    sets[1] = SourceSet{1u, n_lod_own, (int32_t)(nz / g.dsz), 0.0f, 0.0f, (float)(nz % g.dsz)};
    const int nsets = foreign ? 2 : 1;
    std::mt19937 rng(1234);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    std::vector<float> lod(4 * ((size_t)n_lod_own + NF * NF * NF));
    for (uint32_t d = n_lod_own; d < n_lod_own + NF * NF * NF; d++) {
        lod[4 * d] = 60.0f * (1.0f + 0.2f * U(rng));
        lod[4 * d + 1] = 0.1f + 0.02f * U(rng);
        lod[4 * d + 2] = 0.01f + 0.02f * U(rng);
        lod[4 * d + 3] = 0.02f * U(rng);
    }
    for (uint32_t d = 0; d < n_lod_own; d++) {
        const bool filled = dz > 1 ? true : d < (uint32_t)(ND * ND * ND);  // single domain: nothing fills the tail (quirk Q5)
        lod[4 * d] = filled ? 8.0f * (1.0f + 0.2f * U(rng)) : 0.0f;
        lod[4 * d + 1] = filled ? 0.1f + 0.02f * U(rng) : 0.0f;
        lod[4 * d + 2] = filled ? 0.01f + 0.02f * U(rng) : 0.0f;
        lod[4 * d + 3] = filled ? 0.02f * U(rng) : 0.0f;
    }
    // tasks
    std::vector<Task> tasks;
    const uint32_t nwz = (nz + ND * g.dsz - 1) / (ND * g.dsz);
    for (uint32_t wz = 0; wz < nwz; wz++)
        for (uint32_t oz = 0; oz < g.dsz; oz++) {
            if (wz * ND * g.dsz + oz >= nz) continue;
            for (uint32_t oy = 0; oy < g.dsy; oy++)
                for (uint32_t ox = 0; ox < g.dsx; ox++) tasks.push_back(Task{(uint16_t)ox, (uint16_t)oy, (uint16_t)oz, (uint16_t)wz});
        }
    const int ntasks = (int)tasks.size();
    std::vector<float2> khat(C::khat_per_task * ntasks), shat(C::shat_count);
    std::vector<float2> S(C::H * C::M * C::ROW);
    std::vector<float2> plane(C::M * C::ROW);
    std::vector<uint8_t> flags(g.N, 0);
    std::vector<float> Es(3 * g.N), Bs(3 * g.N), Ed(3 * g.N, -7.f), Bd(3 * g.N, -7.f);
    for (auto& v : Es) v = U(rng);
    for (auto& v : Bs) v = U(rng);
    flags[5 + 7 * nx + 9 * nx * ny] = 0x01;  // one solid cell: must stay untouched
    std::vector<float2> W((size_t)C::P * C::PLANE);
    std::vector<float4> tw((size_t)C::H * (ND / 2));
    std::vector<float> scratch((size_t)6 * g.N, 0.0f);
    std::vector<float2> S0((size_t)C::P * C::SLOT);
    for (int i = 0; i < C::H * (ND / 2); i++) tw[i] = main_tw4<ND>(i / (ND / 2), i % (ND / 2));
    std::vector<float> accreg((size_t)C::T * 6 * C::XPT);
    std::vector<float2> khat2(nsets == 2 ? C::khat_per_task * ntasks : 0), shat2(nsets == 2 ? C::shat_count : 0);
    std::vector<float2> shat2c(nsets == 2 ? C::shatc_count : 0), S2c((size_t)C::P * 4 * C::CSLOT);
    for (int set = 0; set < nsets; set++) {
        float2* kh = set == 0 ? khat.data() : khat2.data();
        float2* sh = set == 0 ? shat.data() : shat2.data();
        for (int t = 0; t < ntasks; t++)
            for (int c = 0; c < 3; c++) {
                for (int tid = 0; tid < C::T; tid++) khat_phase_x<ND>(tid, C::T, g, sets[set], tasks[t], c, S.data());
                for (int tid = 0; tid < C::T; tid++) khat_phase_y<ND>(tid, C::T, S.data());
                for (int tid = 0; tid < C::T; tid++) khat_phase_z<ND>(tid, C::T, t, c, S.data(), kh);
            }
        for (int j = 0; j < 4; j++)
            for (int kx = 0; kx < C::H; kx++) {
                for (int tid = 0; tid < 128; tid++) src_phase_x<ND>(tid, 128, g, sets[set], lod.data(), kx, j, plane.data());
                for (int tid = 0; tid < 128; tid++) src_phase_y<ND>(tid, 128, plane.data());
                for (int tid = 0; tid < 128; tid++) src_phase_z<ND>(tid, 128, kx, j, plane.data(), sh, set == 1 ? shat2c.data() : nullptr);
            }
    }
    int mirrored = 0;
    for (int t = 0; t < ntasks; t++) {  // one pass: with two source sets their spectra are added before the inverse transform
        int ks = t;  // whose kernel spectra this task reads
        const bool mir = mirror && 2u * tasks[t].ox > g.dsx;
        if (mir) {
            ks = -1;
            for (int u = 0; u < ntasks; u++)
                if (tasks[u].ox == g.dsx - tasks[t].ox && tasks[u].oy == tasks[t].oy && tasks[u].oz == tasks[t].oz && tasks[u].wz == tasks[t].wz) ks = u;
            if (ks < 0) { printf("{\"error\": \"no mirror partner\"}\n"); return 2; }
            mirrored++;
        }
        const float2* kt = khat.data() + C::khat_per_task * ks;
        std::fill(accreg.begin(), accreg.end(), 0.0f);
        for (int kx0 = 0; kx0 < C::H; kx0 += C::P) {
            const int np = C::H - kx0 < C::P ? C::H - kx0 : C::P;
            if (nsets == 2) {
                main_stage2_host<ND>(kt, khat2.data() + C::khat_per_task * ks, shat2c.data(), kx0, np, W.data(), S2c.data());
                for (int tid = 0; tid < C::T; tid++) {
                    if (mir) main_phase_product2_mirror<ND>(tid, shat.data(), S2c.data(), tw.data(), kx0, np, W.data());
                    else main_phase_product2<ND>(tid, shat.data(), S2c.data(), kx0, np, W.data());
                }
            } else {
                main_stage_host<ND>(kt, shat.data(), kx0, np, W.data(), S0.data());
                for (int tid = 0; tid < C::T; tid++) {
                    if (mir) main_phase_product_mirror<ND>(tid, S0.data(), np, W.data());
                    else main_phase_product<ND>(tid, S0.data(), np, W.data());
                }
            }
            for (int tid = 0; tid < C::T; tid++) main_phase_z<ND>(tid, np, W.data());
            for (int tid = 0; tid < C::T; tid++) main_phase_y<ND>(tid, np, W.data());
            for (int tid = 0; tid < C::T; tid++)
                main_phase_accumulate<ND>(tid, kx0, np, W.data(), tw.data(), *reinterpret_cast<float(*)[6][C::XPT]>(&accreg[(size_t)tid * 6 * C::XPT]));
        }
        for (int tid = 0; tid < C::T; tid++)
            main_phase_store<ND>(tid, g, tasks[t], scratch.data(), false, *reinterpret_cast<float(*)[6][C::XPT]>(&accreg[(size_t)tid * 6 * C::XPT]));
    }
    {  // k_eb_combine: one block per row
        std::vector<float> tile((size_t)(nx / ND) * (ND + 1) + ND + 1);
        for (uint32_t z = 0; z < nz; z++)
            for (uint32_t y = 0; y < ny; y++) {
                if (dz > 1 && (z == 0 || z >= nz - 1)) continue;
                for (int c = 0; c < 6; c++) {
                    for (int tid = 0; tid < 128; tid++) combine_load<ND>(tid, 128, g, y, z, c, scratch.data(), tile.data());
                    for (int tid = 0; tid < 128; tid++)
                        combine_write<ND>(tid, 128, g, y, z, c, tile.data(), flags.data(), Es.data(), Bs.data(), Ed.data(), Bd.data());
                }
            }
    }
    // direct reference (double), reference semantics
    double num[2] = {0, 0}, den[2] = {0, 0}, maxerr = 0;
    long untouched_bad = 0, checked = 0;
    for (uint32_t z = 0; z < nz; z++)
        for (uint32_t y = 0; y < ny; y++)
            for (uint32_t x = 0; x < nx; x++) {
                const uint64_t n = x + ((uint64_t)y + (uint64_t)z * ny) * nx;
                const bool halo = dz > 1 && (z == 0 || z >= nz - 1);
                if (halo || (flags[n] & 0x1F) == 0x01) {
                    for (int c = 0; c < 3; c++)
                        if (Ed[c * g.N + n] != -7.f || Bd[c * g.N + n] != -7.f) untouched_bad++;
                    continue;
                }
                const uint32_t ndi = x / g.dsx + (y / g.dsy + z / g.dsz * ND) * ND;
                double e[3] = {0, 0, 0}, b[3] = {0, 0, 0};
                for (uint32_t d = g.lo; d < n_lod_own; d++) {
                    if (d == ndi) continue;
                    const uint32_t t2 = d % (ND * ND);
                    const double cx = (t2 % ND) * (double)g.dsx + 0.5 * g.dsx, cy = (t2 / ND) * (double)g.dsy + 0.5 * g.dsy,
                                 cz = (d / (ND * ND)) * (double)g.dsz + 0.5 * g.dsz;
                    const double rx = x - cx, ry = y - cy, rz = z - cz;
                    const double r2 = rx * rx + ry * ry + rz * rz, inv = 1.0 / (r2 * std::sqrt(r2));
                    const double q = lod[4 * d], vx = lod[4 * d + 1], vy = lod[4 * d + 2], vz = lod[4 * d + 3];
                    const double px = rx * inv, py = ry * inv, pz = rz * inv;
                    e[0] += q * px; e[1] += q * py; e[2] += q * pz;
                    b[0] += q * (vy * pz - vz * py); b[1] += q * (vz * px - vx * pz); b[2] += q * (vx * py - vy * px);
                }
                if (foreign)  // sim.cl:957-983: level D-1 of the slab below, centres shifted by the halo-inclusive slab height (quirk Q8)
                    for (uint32_t f = 0; f < (uint32_t)(NF * NF * NF); f++) {
                        const uint32_t d = n_lod_own + f;
                        const double bsx = nx / NF, bsy = ny / NF, bsz = nz / NF;
                        const double cx = (f % NF) * bsx + 0.5 * bsx, cy = ((f / NF) % NF) * bsy + 0.5 * bsy, cz = (f / (NF * NF)) * bsz + 0.5 * bsz - (double)nz;
                        const double rx = x - cx, ry = y - cy, rz = z - cz;
                        const double r2 = rx * rx + ry * ry + rz * rz, inv = 1.0 / (r2 * std::sqrt(r2));
                        const double q = lod[4 * d], vx = lod[4 * d + 1], vy = lod[4 * d + 2], vz = lod[4 * d + 3];
                        const double px = rx * inv, py = ry * inv, pz = rz * inv;
                        e[0] += q * px; e[1] += q * py; e[2] += q * pz;
                        b[0] += q * (vy * pz - vz * py); b[1] += q * (vz * px - vx * pz); b[2] += q * (vx * py - vy * px);
                    }
                for (int c = 0; c < 3; c++) {
                    const double re = Es[c * g.N + n] + (double)g.ke * e[c], rb = Bs[c * g.N + n] + (double)g.kmu * b[c];
                    const double de = Ed[c * g.N + n] - re, db = Bd[c * g.N + n] - rb;
                    num[0] += de * de; den[0] += re * re; num[1] += db * db; den[1] += rb * rb;
                    maxerr = std::fmax(maxerr, std::fmax(std::fabs(de), std::fabs(db)));
                }
                checked++;
            }
    const double le = std::sqrt(num[0] / den[0]), lb = std::sqrt(num[1] / den[1]);
    printf("{\"nd\": %d, \"source_sets\": %d, \"tasks\": %d, \"mirrored_tasks\": %d, \"cells\": %ld, \"rel_l2_E\": %.3e, \"rel_l2_B\": %.3e, \"max_abs_err\": %.3e, \"untouched_bad\": %ld}\n", ND, nsets, ntasks, mirrored,
           checked, le, lb, maxerr, untouched_bad);
    return (le < 1e-5 && lb < 1e-5 && untouched_bad == 0) ? 0 : 1;
}

int main(int argc, char** argv) {
    const int depth = argc > 1 ? atoi(argv[1]) : 4;
    const uint32_t nx = argc > 2 ? atoi(argv[2]) : 32, ny = argc > 3 ? atoi(argv[3]) : 32, nz = argc > 4 ? atoi(argv[4]) : 32;
    const uint32_t dz = argc > 5 ? atoi(argv[5]) : 1;
    const bool mirror = argc > 6 && atoi(argv[6]) != 0;
    return depth == 4 ? run<16>(nx, ny, nz, dz, mirror) : run<8>(nx, ny, nz, dz, mirror);
}
