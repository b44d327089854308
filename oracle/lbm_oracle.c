/* oracle/lbm_oracle.c -- TEST INFRASTRUCTURE ONLY (checker; never linked into or called by the product path).
 *
 * Plain-C CPU restatement of the reference's device algorithm for the extended-LBM MHD time step, written
 * from /root/reference/src/kernels/sim_kernels.cl ("sim.cl").  Where the reference specialises by `#define`
 * (src/lbm/domain.rs:736-858) this file takes a run-time parameter block, and where the reference unrolls per
 * velocity set this file uses the direction tables of sim.cl:326-365 in loops -- a deliberately different
 * formulation from the CUDA kernels, with the SAME floating-point operation order as the reference so that all
 * deterministic results are bit-identical (compile with -ffp-contract=off; every fused multiply-add below is
 * an explicit fmaf exactly where the reference calls fma()).
 *
 * Pinning: tests/test_oracle_vs_ref.py checks every function below bit-for-bit against the reference's own
 * kernel source compiled for the host (oracle/_ref, built by oracle/build_ref.py from /root/reference), and
 * tests/golden/ holds outputs of that reference build for machines where /root/reference is absent.
 *
 * SUBGRID_ECR (sim.cl:449-462,556-629,827-831) is restated too (stream_collide_cell, ora_initialize).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct OraParams {
    uint32_t nx, ny, nz;          /* DEF_NX.. incl. halos (domain.rs:91-93) */
    uint32_t dx, dy, dz, di;      /* DEF_DX.. DEF_DI */
    int32_t ox, oy, oz;           /* DEF_OX.. */
    uint32_t velocity_set;        /* 0 D2Q9, 1 D3Q15, 2 D3Q19, 3 D3Q27 (types.rs:17-25) */
    uint32_t trt;                 /* 0 SRT, 1 TRT */
    uint32_t float_type;          /* 0 FP16S, 1 FP16C, 2 FP32 (types.rs:76-82) */
    uint32_t eq_boundaries, volume_force, force_field, magneto_hydro, update_fields;
    float w;                      /* DEF_W */
    float ke, kmu, kmu0, kkge, wq;
    uint32_t lod_depth, n_lod, n_lod_own;
    int32_t threads;              /* OpenMP threads, 0 = default */
    uint32_t subgrid_ecr;         /* SUBGRID_ECR (domain.rs:850-854) */
    float kme, kkbme, keabs;      /* DEF_KME, DEF_KKBME, DEF_KEABS */
    float ecrf;                   /* kernel argument "ecrf" (domain.rs:292-296) */
} OraParams;

typedef struct OraBuffers {
    void* fi; float* rho; float* u; uint8_t* flags; float* F;
    float* E_stat; float* B_stat; float* E_dyn; float* B_dyn;
    void* fqi; void* ei; float* Q; float* QU_lod;
    uint8_t* transfer_p; uint8_t* transfer_m;
    float* E_var; void* eti; float* Et;  /* SUBGRID_ECR (domain.rs:200-211) */
} OraBuffers;

#define TYPE_S 0x01
#define TYPE_E 0x02
#define TYPE_C 0x04
#define TYPE_F 0x08
#define TYPE_M 0x10
#define TYPE_BO 0x1F  /* runtime value, domain.rs:828 */
#define DEF_C 0.57735027f
#define QMAX 27

/* ---- velocity sets: sim.cl:326-365, weights domain.rs:757-768 (D3Q27 canonical) ---- */
static const int8_t C9[3][9] = {{0, 1, -1, 0, 0, 1, -1, 1, -1}, {0, 0, 0, 1, -1, 1, -1, -1, 1}, {0, 0, 0, 0, 0, 0, 0, 0, 0}};
static const int8_t C15[3][15] = {{0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1},
                                  {0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, -1, 1, 1, -1},
                                  {0, 0, 0, 0, 0, 1, -1, 1, -1, -1, 1, 1, -1, 1, -1}};
static const int8_t C19[3][19] = {{0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0},
                                  {0, 0, 0, 1, -1, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 1, -1},
                                  {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, -1, 1, -1, 1}};
static const int8_t C27[3][27] = {{0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1},
                                  {0, 0, 0, 1, -1, 0, 0, 1, -1, 0, 0, 1, -1, -1, 1, 0, 0, 1, -1, 1, -1, 1, -1, -1, 1, 1, -1},
                                  {0, 0, 0, 0, 0, 1, -1, 0, 0, 1, -1, 1, -1, 0, 0, -1, 1, -1, 1, 1, -1, -1, 1, 1, -1, 1, -1}};
/* per-face transferred directions, sim.cl:1029-1058 */
static const uint8_t T9[12] = {1, 5, 7, 2, 6, 8, 3, 5, 8, 4, 6, 7};
static const uint8_t T15[30] = {1, 7, 14, 9, 11, 2, 8, 13, 10, 12, 3, 7, 12, 9, 13, 4, 8, 11, 10, 14, 5, 7, 10, 11, 13, 6, 8, 9, 12, 14};
static const uint8_t T19[30] = {1, 7, 13, 9, 15, 2, 8, 14, 10, 16, 3, 7, 14, 11, 17, 4, 8, 13, 12, 18, 5, 9, 16, 11, 18, 6, 10, 15, 12, 17};
static const uint8_t T27[54] = {1, 7, 13, 9,  15, 19, 26, 21, 23, 2, 8,  14, 10, 16, 20, 25, 22, 24, 3, 7, 14, 11, 17, 19, 24, 21, 25,
                                4, 8, 13, 12, 18, 20, 23, 22, 26, 5, 9,  16, 11, 18, 19, 22, 23, 25, 6, 10, 15, 12, 17, 20, 21, 24, 26};

typedef struct Ctx {
    const OraParams* p;
    uint32_t q, dim, transfers;
    uint64_t N;
    int8_t c[3][QMAX];
    float wgt[QMAX];
    float w0;
    const uint8_t* ttab;
} Ctx;

static void make_ctx(Ctx* k, const OraParams* p) {
    k->p = p;
    k->N = (uint64_t)p->nx * p->ny * p->nz;
    float ws, we, wc;
    const int8_t* src;
    switch (p->velocity_set) {
        case 0: k->q = 9; k->dim = 2; k->transfers = 3; src = &C9[0][0]; k->ttab = T9;
            k->w0 = 1.0f / 2.25f; ws = 1.0f / 9.0f; we = 1.0f / 36.0f; wc = 0.0f; break;
        case 1: k->q = 15; k->dim = 3; k->transfers = 5; src = &C15[0][0]; k->ttab = T15;
            k->w0 = 1.0f / 4.5f; ws = 1.0f / 9.0f; we = 0.0f; wc = 1.0f / 72.0f; break;
        case 2: k->q = 19; k->dim = 3; k->transfers = 5; src = &C19[0][0]; k->ttab = T19;
            k->w0 = 1.0f / 3.0f; ws = 1.0f / 18.0f; we = 1.0f / 36.0f; wc = 0.0f; break;
        default: k->q = 27; k->dim = 3; k->transfers = 9; src = &C27[0][0]; k->ttab = T27;
            k->w0 = 1.0f / 3.375f; ws = 1.0f / 13.5f; we = 1.0f / 54.0f; wc = 1.0f / 216.0f; break;
    }
    for (uint32_t a = 0; a < 3; a++)
        for (uint32_t i = 0; i < k->q; i++) k->c[a][i] = src[a * k->q + i];
    for (uint32_t i = 0; i < k->q; i++) {
        const int nz = (k->c[0][i] != 0) + (k->c[1][i] != 0) + (k->c[2][i] != 0);
        k->wgt[i] = nz == 0 ? k->w0 : nz == 1 ? ws : nz == 2 ? we : wc;
    }
}

/* ---- DDF storage codecs: FP16S macros domain.rs:773-776, FP16C sim.cl:79-90, FP32 domain.rs:781-784 ---- */
static inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint16_t f2h_custom(float x) {
    const uint32_t b = as_uint(x) + 0x00000800u;
    const uint32_t e = (b & 0x7F800000u) >> 23;
    const uint32_t m = b & 0x007FFFFFu;
    return (uint16_t)((b & 0x80000000u) >> 16 | (uint32_t)(e > 112u) * ((((e - 112u) << 11) & 0x7800u) | m >> 12) |
                      (uint32_t)((e < 113u) & (e > 100u)) * ((((0x007FF800u + m) >> (124u - e)) + 1u) >> 1));
}
static float h2f_custom(uint16_t x) {
    const uint32_t e = ((uint32_t)x & 0x7800u) >> 11;
    const uint32_t m = ((uint32_t)x & 0x07FFu) << 12;
    const uint32_t v = as_uint((float)m) >> 23;
    return as_float(((uint32_t)x & 0x8000u) << 16 | (uint32_t)(e != 0u) * ((e + 112u) << 23 | m) |
                    (uint32_t)((e == 0u) & (m != 0u)) * ((v - 37u) << 23 | ((m << (150u - v)) & 0x007FF000u)));
}
static inline float ddf_load(const Ctx* k, const void* buf, uint64_t o) {
    switch (k->p->float_type) {
        case 0: return (float)((const _Float16*)buf)[o] * 3.0517578E-5f;
        case 1: return h2f_custom(((const uint16_t*)buf)[o]);
        default: return ((const float*)buf)[o];
    }
}
/* Test-infrastructure hook: how many NaNs were handed to the DDF encoders.  A scene that stores NaN has left the domain where
 * parity is defined: the reference launders NaN through the FP16C bit formula into a finite code that depends on the NaN's sign
 * and payload, i.e. on the hardware that produced it (x86 default NaN 0xFFC00000 -> -1.5, NVIDIA's 0x7FFFFFFF -> -0). */
static uint64_t g_nan_stores = 0;
static inline void ddf_store(const Ctx* k, void* buf, uint64_t o, float x) {
    if (x != x) __atomic_fetch_add(&g_nan_stores, 1, __ATOMIC_RELAXED);
    switch (k->p->float_type) {
        case 0: ((_Float16*)buf)[o] = (_Float16)(x * 32768.0f); break;  /* vstore_half_rte */
        case 1: ((uint16_t*)buf)[o] = f2h_custom(x); break;
        default: ((float*)buf)[o] = x; break;
    }
}
static inline size_t ddf_size(const Ctx* k) { return k->p->float_type == 2 ? 4 : 2; }

/* ---- index math: sim.cl:134-154,248-302 ---- */
static inline void coordinates(const Ctx* k, uint32_t n, uint32_t* x, uint32_t* y, uint32_t* z) {
    const uint32_t nxy = k->p->nx * k->p->ny, t = n % nxy;
    *x = t % k->p->nx; *y = t / k->p->nx; *z = n / nxy;
}
static inline int is_halo(const Ctx* k, uint32_t n) {
    uint32_t x, y, z;
    coordinates(k, n, &x, &y, &z);
    const OraParams* p = k->p;
    return ((p->dx > 1u) & (x == 0u || x >= p->nx - 1u)) || ((p->dy > 1u) & (y == 0u || y >= p->ny - 1u)) ||
           ((p->dz > 1u) & (z == 0u || z >= p->nz - 1u));
}
static inline uint64_t index_f(const Ctx* k, uint32_t n, uint32_t i) { return (uint64_t)i * k->N + (uint64_t)n; }
static void neighbors(const Ctx* k, uint32_t n, uint32_t* j) {
    const OraParams* p = k->p;
    uint32_t x, y, z;
    coordinates(k, n, &x, &y, &z);
    const uint32_t xs[3] = {(x + p->nx - 1u) % p->nx, x, (x + 1u) % p->nx};
    const uint32_t ys[3] = {((y + p->ny - 1u) % p->ny) * p->nx, y * p->nx, ((y + 1u) % p->ny) * p->nx};
    const uint32_t zs[3] = {((z + p->nz - 1u) % p->nz) * p->ny * p->nx, z * p->ny * p->nx, ((z + 1u) % p->nz) * p->ny * p->nx};
    j[0] = n;
    for (uint32_t i = 1; i < k->q; i++) {
        uint32_t v = xs[k->c[0][i] + 1] + ys[k->c[1][i] + 1];
        if (k->dim == 3) v += zs[k->c[2][i] + 1];  /* D2Q9 ignores z, sim.cl:265-268 */
        j[i] = v;
    }
}
static void neighbors_a(const Ctx* k, uint32_t n, uint32_t* j7) {  /* sim.cl:382-389 */
    const OraParams* p = k->p;
    uint32_t x, y, z;
    coordinates(k, n, &x, &y, &z);
    const uint32_t x0 = x, xp = (x + 1u) % p->nx, xm = (x + p->nx - 1u) % p->nx;
    const uint32_t y0 = y * p->nx, yp = ((y + 1u) % p->ny) * p->nx, ym = ((y + p->ny - 1u) % p->ny) * p->nx;
    const uint32_t z0 = z * p->ny * p->nx, zp = ((z + 1u) % p->nz) * p->ny * p->nx, zm = ((z + p->nz - 1u) % p->nz) * p->ny * p->nx;
    j7[0] = n;
    j7[1] = xp + y0 + z0; j7[2] = xm + y0 + z0;
    j7[3] = x0 + yp + z0; j7[4] = x0 + ym + z0;
    j7[5] = x0 + y0 + zp; j7[6] = x0 + y0 + zm;
}

/* ---- Esoteric-Pull, sim.cl:234-247,397-410 ---- */
static void load_ddfs(const Ctx* k, uint32_t q, uint32_t n, float* f, const void* buf, const uint32_t* j, uint64_t t) {
    f[0] = ddf_load(k, buf, index_f(k, n, 0u));
    for (uint32_t i = 1u; i < q; i += 2u) {
        f[i] = ddf_load(k, buf, index_f(k, n, t % 2ul ? i : i + 1u));
        f[i + 1u] = ddf_load(k, buf, index_f(k, j[i], t % 2ul ? i + 1u : i));
    }
}
static void store_ddfs(const Ctx* k, uint32_t q, uint32_t n, const float* f, void* buf, const uint32_t* j, uint64_t t) {
    ddf_store(k, buf, index_f(k, n, 0u), f[0]);
    for (uint32_t i = 1u; i < q; i += 2u) {
        ddf_store(k, buf, index_f(k, j[i], t % 2ul ? i + 1u : i), f[i]);
        ddf_store(k, buf, index_f(k, n, t % 2ul ? i : i + 1u), f[i + 1u]);
    }
}

/* ---- moments / equilibrium / forcing ---- */
static inline float sq(float x) { return x * x; }
static inline float cb(float x) { return x * x * x; }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* sim.cl:155-207: for the odd direction i of each pair, u_i = c_i . (3u) summed x,y,z over the non-zero
 * components (this reproduces u0..u9 of the unrolled reference, e.g. u3 = ux-uy, u9 = -ux+uy+uz); the even
 * partner uses -u_i. */
static void calculate_f_eq(const Ctx* k, float rho, float ux, float uy, float uz, float* feq) {
    const float c3 = -3.0f * (sq(ux) + sq(uy) + sq(uz)), rhom1 = rho - 1.0f;
    ux *= 3.0f; uy *= 3.0f; uz *= 3.0f;
    feq[0] = k->w0 * fmaf(rho, 0.5f * c3, rhom1);
    const float uu[3] = {ux, uy, uz};
    for (uint32_t i = 1; i < k->q; i += 2) {
        float ui = 0.0f;
        int first = 1;
        for (int a = 0; a < 3; a++) {
            const int c = k->c[a][i];
            if (!c) continue;
            const float term = c > 0 ? uu[a] : -uu[a];
            ui = first ? term : ui + term;
            first = 0;
        }
        const float rhow = k->wgt[i] * rho, rhom1w = k->wgt[i] * rhom1;
        feq[i] = fmaf(rhow, fmaf(0.5f, fmaf(ui, ui, c3), ui), rhom1w);
        feq[i + 1] = fmaf(rhow, fmaf(0.5f, fmaf(ui, ui, c3), -ui), rhom1w);
    }
}
/* sim.cl:208-233: per axis, pairs in index order; inside a pair the DDF moving in +axis is added first */
static void calculate_rho_u(const Ctx* k, const float* f, float* rhon, float* uxn, float* uyn, float* uzn) {
    float rho = f[0];
    for (uint32_t i = 1u; i < k->q; i++) rho += f[i];
    rho += 1.0f;
    float m[3] = {0.0f, 0.0f, 0.0f};
    for (int a = 0; a < 3; a++) {
        int first = 1;
        for (uint32_t i = 1; i < k->q; i += 2) {
            const int c = k->c[a][i];
            if (!c) continue;
            const float plus = c > 0 ? f[i] : f[i + 1], minus = c > 0 ? f[i + 1] : f[i];
            m[a] = first ? plus : m[a] + plus;
            m[a] = m[a] - minus;
            first = 0;
        }
    }
    *rhon = rho;
    *uxn = m[0] / rho;
    *uyn = m[1] / rho;
    *uzn = m[2] / rho;  /* D2Q9: uz = 0.0f, 0/rho */
}
/* sim.cl:367-377 */
static void calculate_forcing_terms(const Ctx* k, float ux, float uy, float uz, float fx, float fy, float fz, float* Fin) {
    const float uF = k->dim == 2 ? -0.33333334f * fmaf(ux, fx, uy * fy) : -0.33333334f * fmaf(ux, fx, fmaf(uy, fy, uz * fz));
    Fin[0] = 9.0f * k->w0 * uF;
    for (uint32_t i = 1u; i < k->q; i++) {
        const float cx = (float)k->c[0][i], cy = (float)k->c[1][i], cz = (float)k->c[2][i];
        Fin[i] = 9.0f * k->wgt[i] * fmaf(cx * fx + cy * fy + cz * fz, cx * ux + cy * uy + cz * uz + 0.33333334f, uF);
    }
}
/* sim.cl:390-396 */
static void calculate_a_eq(float Q, float ux, float uy, float uz, float* qeq) {
    const float wsT4 = 0.5f * Q, wsTm1 = 0.125f * (Q - 1.0f);
    qeq[0] = fmaf(0.25f, Q, -0.25f);
    qeq[1] = fmaf(wsT4, ux, wsTm1); qeq[2] = fmaf(wsT4, -ux, wsTm1);
    qeq[3] = fmaf(wsT4, uy, wsTm1); qeq[4] = fmaf(wsT4, -uy, wsTm1);
    qeq[5] = fmaf(wsT4, uz, wsTm1); qeq[6] = fmaf(wsT4, -uz, wsTm1);
}

/* ---- LOD helpers, sim.cl:425-447 ---- */
static uint32_t lod_index(const Ctx* k, uint32_t n, uint32_t d) {
    uint32_t x, y, z;
    coordinates(k, n, &x, &y, &z);
    const uint32_t nd = 1u << d;
    return x / (k->p->nx / nd) + (y / (k->p->ny / nd) + z / (k->p->nz / nd) * nd) * nd;
}
static float lod_s(const Ctx* k, uint32_t d) {
    const uint32_t nd = 1u << d;
    return (float)((k->p->nx / nd) * (k->p->ny / nd) * (k->p->nz / nd));
}
static void lod_coordinates(const Ctx* k, uint32_t n, uint32_t d, float* c) {
    const uint32_t nd = 1u << d;
    const float dsx = (float)(k->p->nx / nd), dsy = (float)(k->p->ny / nd), dsz = (float)(k->p->nz / nd);
    const uint32_t t = n % (nd * nd);
    c[0] = (float)(t % nd) * dsx + (0.5f * dsx);
    c[1] = (float)(t / nd) * dsy + (0.5f * dsy);
    c[2] = (float)(n / (nd * nd)) * dsz + (0.5f * dsz);
}
static inline void atomic_add_f(float* addr, float val) {  /* sim.cl:118-128: order-dependent float sum (quirk Q6) */
#pragma omp atomic
    *addr += val;
}
static inline uint32_t to_d(const Ctx* k, uint32_t x) { return k->dim == 2 ? x * x : x * x * x; }

#define PAR_FOR _Pragma("omp parallel for schedule(static)")
static void set_threads(const OraParams* p) {
#ifdef _OPENMP
    if (p->threads > 0) omp_set_num_threads(p->threads);
#else
    (void)p;
#endif
}

/* =============================== stream_collide, sim.cl:465-758 =============================== */
/* SUBGRID_ECR helpers, sim.cl:449-461 */
static inline float length3(const float* v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }  /* OpenCL length() */
static float mag_v(const Ctx* k, uint32_t n, const float* V) {
    return sqrtf(sq(V[n]) + sq(V[k->N + (uint64_t)n]) + sq(V[2ul * k->N + (uint64_t)n]));
}
static void grad_mag_v(const Ctx* k, uint32_t n, const float* V, float* g) {
    uint32_t x, y, z;
    coordinates(k, n, &x, &y, &z);
    const OraParams* p = k->p;
    if (x == 0 || x == p->nx - 1 || y == 0 || y == p->ny - 1 || z == 0 || z == p->nz - 1) { g[0] = g[1] = g[2] = 0.0f; return; }
    const uint32_t nx = p->nx, nxy = p->nx * p->ny;
    /* `a - b / 2.0f`: the division binds to the second term only (sim.cl:457-459), kept */
    g[0] = mag_v(k, n + 1u, V) - mag_v(k, n - 1u, V) / 2.0f;
    g[1] = mag_v(k, n + nx, V) - mag_v(k, n - nx, V) / 2.0f;
    g[2] = mag_v(k, n + nxy, V) - mag_v(k, n - nxy, V) / 2.0f;
}

static void stream_collide_cell(const Ctx* k, const OraBuffers* b, uint32_t n, uint64_t t, float fx, float fy, float fz) {
    const OraParams* p = k->p;
    if (is_halo(k, n)) return;
    const uint8_t flagsn_bo = b->flags[n] & TYPE_BO;
    if (flagsn_bo == TYPE_S) return;
    uint32_t j[QMAX];
    neighbors(k, n, j);
    float fhn[QMAX];
    load_ddfs(k, k->q, n, fhn, b->fi, j, t);
    const uint64_t nxi = n, nyi = k->N + n, nzi = 2ul * k->N + n;
    float rhon, uxn, uyn, uzn;
    const int type_e = p->eq_boundaries && flagsn_bo == TYPE_E;
    if (type_e) {
        rhon = b->rho[n]; uxn = b->u[nxi]; uyn = b->u[nyi]; uzn = b->u[nzi];
    } else {
        calculate_rho_u(k, fhn, &rhon, &uxn, &uyn, &uzn);
    }
    float fxn = fx, fyn = fy, fzn = fz;
    float Fin[QMAX], feq[QMAX];
    const float w = p->w;
    const float c_tau = fmaf(w, -0.5f, 1.0f);
    if (p->force_field) { fxn += b->F[nxi]; fyn += b->F[nyi]; fzn += b->F[nzi]; }

    if (p->magneto_hydro) {
        const float Bn[3] = {b->B_dyn[nxi], b->B_dyn[nyi], b->B_dyn[nzi]};
        const float En[3] = {b->E_dyn[nxi], b->E_dyn[nyi], b->E_dyn[nzi]};
        /* electron gas part 1 */
        float ehn[QMAX];
        load_ddfs(k, k->q, n, ehn, b->ei, j, t);
        float rhon_e, uxn_e, uyn_e, uzn_e;
        calculate_rho_u(k, ehn, &rhon_e, &uxn_e, &uyn_e, &uzn_e);
        /* gas charge advection 1 */
        uint32_t j7[7];
        neighbors_a(k, n, j7);
        float qhn[7];
        load_ddfs(k, 7u, n, qhn, b->fqi, j7, t);
        float rhon_q = 0.0f;
        for (uint32_t i = 0u; i < 7u; i++) rhon_q += qhn[i];
        rhon_q += 1.0f;
        if (p->subgrid_ecr) {  /* sim.cl:556-629 */
            /* electron temperature 1 */
            float ethn[7];
            load_ddfs(k, 7u, n, ethn, b->eti, j7, t);
            float Etn = 0.0f;
            for (uint32_t i = 0u; i < 7u; i++) Etn += ethn[i];
            Etn += 1.0f;
            /* ECR heating */
            const float Env[3] = {b->E_var[nxi], b->E_var[nyi], b->E_var[nzi]};
            const float f_c = length3(Bn) * (1.0f / (p->kme * 2.0f * 3.14159274101257f));
            /* `0.03125` and `32` are double / int literals: this term is evaluated in double and narrowed by sq(float) */
            const float detune = (float)((double)p->ecrf / (0.03125 * (double)f_c) - 32);
            const float rel_absorbtion = 1.0f / (1.0f + sq(detune));
            (void)rel_absorbtion;  /* computed and printed by the reference, never used (sim.cl:584-585) */
            const float sc = (Env[0] * Bn[0] + Env[1] * Bn[1] + Env[2] * Bn[2]) / sq(length3(Bn));
            const float perp[3] = {Env[0] - sc * Bn[0], Env[1] - sc * Bn[1], Env[2] - sc * Bn[2]};
            const float Env_mag = length3(perp);
            Etn += p->keabs / p->kkbme * (rhon_e + 0.00001f) / p->kkge * sq(Env_mag);
            /* drift of gyrating electrons */
            float gr[3];
            grad_mag_v(k, n, b->B_dyn, gr);
            const float ke_t = p->kkbme * Etn, lb = length3(Bn);
            const float dup[3] = {(ke_t * gr[0]) / lb, (ke_t * gr[1]) / lb, (ke_t * gr[2]) / lb};
            uxn_e += dup[0];
            uyn_e += dup[1];
            uzn_e += dup[2];
            Etn -= length3(dup) / p->kkbme;
            /* electron temperature 2 */
            float eteq[7];
            calculate_a_eq(Etn, uxn_e, uyn_e, uzn_e, eteq);
            if (p->update_fields) b->Et[n] = Etn;
            for (uint32_t i = 0u; i < 7u; i++) ethn[i] = fmaf(1.0f - p->wq, ethn[i], p->wq * eteq[i]);
            store_ddfs(k, 7u, n, ethn, b->eti, j7, t);
            /* ionization */
            const float delta_q_rho = 0.0001f * Etn;
            rhon_e += delta_q_rho;
            rhon_q += delta_q_rho;
        }
        /* gas charge advection 2 */
        b->Q[n] = rhon_q - rhon_e;
        float qeq[7];
        calculate_a_eq(rhon_q, uxn, uyn, uzn, qeq);
        for (uint32_t i = 0u; i < 7u; i++) qhn[i] = fmaf(1.0f - p->wq, qhn[i], p->wq * qeq[i]);
        store_ddfs(k, 7u, n, qhn, b->fqi, j7, t);
        /* electron gas part 2: e_fn = -rho_e * (E + u_e x B), cross() without contraction */
        const float cr[3] = {uyn_e * Bn[2] - uzn_e * Bn[1], uzn_e * Bn[0] - uxn_e * Bn[2], uxn_e * Bn[1] - uyn_e * Bn[0]};
        const float nre = -rhon_e;
        const float e_fn[3] = {nre * (En[0] + cr[0]), nre * (En[1] + cr[1]), nre * (En[2] + cr[2])};
        const float rho2_e = 0.5f / (rhon_e * p->kkge);
        uxn_e = clampf(fmaf(e_fn[0], rho2_e, uxn_e), -DEF_C, DEF_C);
        uyn_e = clampf(fmaf(e_fn[1], rho2_e, uyn_e), -DEF_C, DEF_C);
        uzn_e = clampf(fmaf(e_fn[2], rho2_e, uzn_e), -DEF_C, DEF_C);
        calculate_forcing_terms(k, uxn_e, uyn_e, uzn_e, e_fn[0], e_fn[1], e_fn[2], Fin);
        calculate_f_eq(k, rhon_e, uxn_e, uyn_e, uzn_e, feq);
        for (uint32_t i = 0u; i < k->q; i++) Fin[i] *= c_tau;
        for (uint32_t i = 0u; i < k->q; i++) ehn[i] = type_e ? feq[i] : fmaf(1.0f - w, ehn[i], fmaf(w, feq[i], Fin[i]));
        store_ddfs(k, k->q, n, ehn, b->ei, j, t);
        /* EM force on gas */
        fxn += rhon_q * (En[0] + uyn * Bn[2] - uzn * Bn[1]);
        fyn += rhon_q * (En[1] + uzn * Bn[0] - uxn * Bn[2]);
        fzn += rhon_q * (En[2] + uxn * Bn[1] - uyn * Bn[0]);
        /* LOD construction */
        if (p->lod_depth > 0u) {
            uint32_t off = 0;
            if (p->dx > 1 || p->dy > 1 || p->dz > 1)
                for (uint32_t d = 0; d < p->lod_depth; d++) off += 1u << (d * k->dim);
            const uint32_t ind = (lod_index(k, n, p->lod_depth) + off) * 4;
            const float ils = 1.0f / lod_s(k, p->lod_depth);
            /* quirk Q7: on split axes lod_index divides by the halo-inclusive size and can point past the own finest
             * level; the reference then adds into foreign entries or, past DEF_NUM_LOD, outside the buffer (undefined
             * behaviour).  The in-buffer case is reproduced, the out-of-buffer deposit is dropped. */
            if (ind / 4 < p->n_lod) {
            atomic_add_f(&b->QU_lod[ind + 0], rhon_q - rhon_e);
            atomic_add_f(&b->QU_lod[ind + 1], uxn * ils);
            atomic_add_f(&b->QU_lod[ind + 2], uyn * ils);
            atomic_add_f(&b->QU_lod[ind + 3], uzn * ils);
            }
        }
    }

    if (p->volume_force) {
        const float rho2 = 0.5f / rhon;
        uxn = clampf(fmaf(fxn, rho2, uxn), -DEF_C, DEF_C);
        uyn = clampf(fmaf(fyn, rho2, uyn), -DEF_C, DEF_C);
        uzn = clampf(fmaf(fzn, rho2, uzn), -DEF_C, DEF_C);
        calculate_forcing_terms(k, uxn, uyn, uzn, fxn, fyn, fzn, Fin);
    } else {
        uxn = clampf(uxn, -DEF_C, DEF_C);
        uyn = clampf(uyn, -DEF_C, DEF_C);
        uzn = clampf(uzn, -DEF_C, DEF_C);
        for (uint32_t i = 0u; i < k->q; i++) Fin[i] = 0.0f;
    }
    if (p->update_fields && !type_e) {
        b->rho[n] = rhon; b->u[nxi] = uxn; b->u[nyi] = uyn; b->u[nzi] = uzn;
    }
    calculate_f_eq(k, rhon, uxn, uyn, uzn, feq);
    if (!p->trt) {
        if (p->volume_force)
            for (uint32_t i = 0u; i < k->q; i++) Fin[i] *= c_tau;
        for (uint32_t i = 0u; i < k->q; i++) fhn[i] = type_e ? feq[i] : fmaf(1.0f - w, fhn[i], fmaf(w, feq[i], Fin[i]));
    } else {
        const float wp = w;
        const float wm = 1.0f / (0.1875f / (1.0f / w - 0.5f) + 0.5f);
        if (p->volume_force) {
            const float c_taup = fmaf(wp, -0.25f, 0.5f), c_taum = fmaf(wm, -0.25f, 0.5f);
            float Fib[QMAX];
            Fib[0] = Fin[0];
            for (uint32_t i = 1u; i < k->q; i += 2u) { Fib[i] = Fin[i + 1u]; Fib[i + 1u] = Fin[i]; }
            for (uint32_t i = 0u; i < k->q; i++) Fin[i] = fmaf(c_taup, Fin[i] + Fib[i], c_taum * (Fin[i] - Fib[i]));
        }
        float fhb[QMAX], feb[QMAX];
        fhb[0] = fhn[0];
        feb[0] = feq[0];
        for (uint32_t i = 1u; i < k->q; i += 2u) {
            fhb[i] = fhn[i + 1u]; fhb[i + 1u] = fhn[i];
            feb[i] = feq[i + 1u]; feb[i + 1u] = feq[i];
        }
        for (uint32_t i = 0u; i < k->q; i++)
            fhn[i] = type_e ? feq[i]
                            : fmaf(0.5f * wp, feq[i] - fhn[i] + feb[i] - fhb[i],
                                   fmaf(0.5f * wm, feq[i] - feb[i] - fhn[i] + fhb[i], fhn[i] + Fin[i]));
    }
    store_ddfs(k, k->q, n, fhn, b->fi, j, t);
}

void ora_stream_collide(const OraParams* p, const OraBuffers* b, uint64_t t, float fx, float fy, float fz) {
    Ctx k;
    make_ctx(&k, p);
    set_threads(p);
    PAR_FOR
    for (int64_t n = 0; n < (int64_t)k.N; n++) stream_collide_cell(&k, b, (uint32_t)n, t, fx, fy, fz);
}

/* =============================== initialize, sim.cl:760-832 =============================== */
void ora_initialize(const OraParams* p, const OraBuffers* b) {
    Ctx k;
    make_ctx(&k, p);
    set_threads(p);
    PAR_FOR
    for (int64_t nn = 0; nn < (int64_t)k.N; nn++) {
        const uint32_t n = (uint32_t)nn;
        if (is_halo(&k, n)) continue;
        const uint64_t nxi = n, nyi = k.N + n, nzi = 2ul * k.N + n;
        const uint8_t flagsn_bo = b->flags[n] & TYPE_BO;
        uint32_t j[QMAX];
        neighbors(&k, n, j);
        if (flagsn_bo == TYPE_S) {  /* both branches of sim.cl:784-803 end in the same state */
            b->u[nxi] = 0.0f; b->u[nyi] = 0.0f; b->u[nzi] = 0.0f;
            if (p->magneto_hydro) b->Q[n] = 0.0f;
        }
        float fe_eq[QMAX];
        calculate_f_eq(&k, b->rho[n], b->u[nxi], b->u[nyi], b->u[nzi], fe_eq);
        store_ddfs(&k, k.q, n, fe_eq, b->fi, j, 1ul);
        if (p->magneto_hydro) {
            float qeq[7];
            calculate_a_eq(b->Q[n], b->u[nxi], b->u[nyi], b->u[nzi], qeq);
            uint32_t j7[7];
            neighbors_a(&k, n, j7);
            store_ddfs(&k, 7u, n, qeq, b->fqi, j7, 1ul);
            b->B_dyn[nxi] = b->B_stat[nxi]; b->B_dyn[nyi] = b->B_stat[nyi]; b->B_dyn[nzi] = b->B_stat[nzi];
            b->E_dyn[nxi] = b->E_stat[nxi]; b->E_dyn[nyi] = b->E_stat[nyi]; b->E_dyn[nzi] = b->E_stat[nzi];
            calculate_f_eq(&k, 0.0f, b->u[nxi], b->u[nyi], b->u[nzi], fe_eq);  /* quirk Q10 */
            store_ddfs(&k, k.q, n, fe_eq, b->ei, j, 1ul);
            if (p->subgrid_ecr) {  /* sim.cl:827-831 */
                float eteq[7];
                calculate_a_eq(b->Et[n], b->u[nxi], b->u[nyi], b->u[nzi], eteq);
                store_ddfs(&k, 7u, n, eteq, b->eti, j7, 1ul);
            }
        }
    }
}

/* =============================== update_fields, sim.cl:834-859 =============================== */
void ora_update_fields(const OraParams* p, const OraBuffers* b, uint64_t t) {
    Ctx k;
    make_ctx(&k, p);
    set_threads(p);
    PAR_FOR
    for (int64_t nn = 0; nn < (int64_t)k.N; nn++) {
        const uint32_t n = (uint32_t)nn;
        if (is_halo(&k, n)) continue;
        if ((b->flags[n] & TYPE_BO) == TYPE_S) continue;
        uint32_t j[QMAX];
        neighbors(&k, n, j);
        float fhn[QMAX];
        load_ddfs(&k, k.q, n, fhn, b->fi, j, t);
        float rhon, uxn, uyn, uzn;
        calculate_rho_u(&k, fhn, &rhon, &uxn, &uyn, &uzn);
        b->rho[n] = rhon;
        b->u[n] = clampf(uxn, -DEF_C, DEF_C);
        b->u[k.N + n] = clampf(uyn, -DEF_C, DEF_C);
        b->u[2ul * k.N + n] = clampf(uzn, -DEF_C, DEF_C);
    }
}

/* =============================== LOD kernels, sim.cl:864-895,995-1003 =============================== */
void ora_clear_qu_lod(const OraParams* p, const OraBuffers* b) {  /* global size n_lod (domain.rs:277), guard n > NUM_LOD_OWN */
    for (uint32_t n = 0; n < p->n_lod; n++) {
        if (n > p->n_lod_own) continue;
        for (int c = 0; c < 4; c++) b->QU_lod[n * 4 + c] = 0.0f;
    }
}
void ora_lod_part_2_gather(const OraParams* p, const OraBuffers* b, uint32_t depth) {
    (void)p;
    float* lods = b->QU_lod;
    const uint32_t nd = 1u << depth, nnd = 1u << (depth + 1);
    uint32_t off = 0;
    for (uint32_t d = 0; d < depth; d++) off += 1u << (d * 3);
    const uint32_t base = off + (1u << (depth * 3));
    static const uint8_t ox[8] = {0, 1, 1, 1, 1, 0, 0, 0}, oy[8] = {0, 0, 1, 1, 0, 1, 1, 0}, oz[8] = {0, 0, 0, 1, 1, 0, 1, 1};
    for (uint32_t n = 0; n < (1u << (depth * 3)); n++) {
        const uint32_t t = n % (nd * nd);
        const uint32_t bx = (t % nd) * 2, by = (t / nd) * 2, bz = (n / (nd * nd)) * 2;
        float qs = 0.0f, uxs = 0.0f, uys = 0.0f, uzs = 0.0f;
        for (int i = 0; i < 8; i++) {
            const uint32_t jj = base + (bx + ox[i]) + (by + oy[i]) * nnd + (bz + oz[i]) * nnd * nnd;
            qs += lods[jj * 4 + 0]; uxs += lods[jj * 4 + 1]; uys += lods[jj * 4 + 2]; uzs += lods[jj * 4 + 3];
        }
        lods[(off + n) * 4 + 0] = qs;
        lods[(off + n) * 4 + 1] = (float)(uxs * 0.125);  /* `*0.125` is a double product in OpenCL C; exact either way */
        lods[(off + n) * 4 + 2] = (float)(uys * 0.125);
        lods[(off + n) * 4 + 3] = (float)(uzs * 0.125);
    }
}

/* =============================== update_e_b_dynamic, sim.cl:897-993 =============================== */
static inline int imax(int x, int y) { return x > y ? x : y; }
static inline int imin(int x, int y) { return x < y ? x : y; }
static inline void accumulate(float* e, float* bf, float q_c, const float* v_c, const float* r) {
    /* pre_field = vec_r / cbmagnitude(vec_r); e += q*pre; b += q*cross(v, pre)   (sim.cl:931-935) */
    const float l3 = cb(sqrtf(sq(r[0]) + sq(r[1]) + sq(r[2])));
    const float pre[3] = {r[0] / l3, r[1] / l3, r[2] / l3};
    e[0] += q_c * pre[0]; e[1] += q_c * pre[1]; e[2] += q_c * pre[2];
    bf[0] += q_c * (v_c[1] * pre[2] - v_c[2] * pre[1]);
    bf[1] += q_c * (v_c[2] * pre[0] - v_c[0] * pre[2]);
    bf[2] += q_c * (v_c[0] * pre[1] - v_c[1] * pre[0]);
}
void ora_update_e_b_dynamic(const OraParams* p, const OraBuffers* b) {
    Ctx k;
    make_ctx(&k, p);
    set_threads(p);
    const uint64_t N = k.N;
    PAR_FOR
    for (int64_t nn = 0; nn < (int64_t)N; nn++) {
        const uint32_t n = (uint32_t)nn;
        if (is_halo(&k, n)) continue;
        if ((b->flags[n] & TYPE_BO) == TYPE_S) continue;
        uint32_t cx, cy, cz;
        coordinates(&k, n, &cx, &cy, &cz);
        const float cf[3] = {(float)cx, (float)cy, (float)cz};
        const uint32_t nd = 1u << p->lod_depth;
        /* 1<<nd with nd = 2^depth (quirk Q4); depth <= 4 so the shift stays below 32 */
        const uint32_t dsx = (uint32_t)imax((int)(p->nx / (1u << nd)), 1), dsy = (uint32_t)imax((int)(p->ny / (1u << nd)), 1),
                       dsz = (uint32_t)imax((int)(p->nz / (1u << nd)), 1);
        const uint32_t x_upper = (uint32_t)imin((int)((cx / dsx) * dsx + dsx), (int)(p->dx > 1 ? p->nx - 1 : p->nx));
        const uint32_t y_upper = (uint32_t)imin((int)((cy / dsy) * dsy + dsy), (int)(p->dy > 1 ? p->ny - 1 : p->ny));
        const uint32_t z_upper = (uint32_t)imin((int)((cz / dsz) * dsz + dsz), (int)(p->dz > 1 ? p->nz - 1 : p->nz));
        float e[3] = {0.0f, 0.0f, 0.0f}, bf[3] = {0.0f, 0.0f, 0.0f};
        for (uint32_t x = (uint32_t)imax((int)((cx / dsx) * dsx), p->dx > 1 ? 1 : 0); x < x_upper; x++)
            for (uint32_t y = (uint32_t)imax((int)((cy / dsy) * dsy), p->dy > 1 ? 1 : 0); y < y_upper; y++)
                for (uint32_t z = (uint32_t)imax((int)((cz / dsz) * dsz), p->dz > 1 ? 1 : 0); z < z_upper; z++) {
                    const uint32_t n_c = x + (y + z * p->ny) * p->nx;
                    if (n == n_c) continue;
                    const float q_c = b->Q[n_c];
                    if (q_c == 0.0f) continue;
                    const float v_c[3] = {b->u[n_c], b->u[(uint64_t)n_c + N], b->u[(uint64_t)n_c + N * 2ul]};
                    const float r[3] = {cf[0] - (float)x, cf[1] - (float)y, cf[2] - (float)z};
                    accumulate(e, bf, q_c, v_c, r);
                }
        const uint32_t ndi = lod_index(&k, n, p->lod_depth);
        for (uint32_t d = (uint32_t)imax((int)p->n_lod_own - (int)to_d(&k, 1u << p->lod_depth), 0); d < p->n_lod_own; d++) {
            if (d == ndi) continue;
            float d_c[3];
            lod_coordinates(&k, d, p->lod_depth, d_c);
            const float q_c = b->QU_lod[d * 4 + 0];
            const float v_c[3] = {b->QU_lod[d * 4 + 1], b->QU_lod[d * 4 + 2], b->QU_lod[d * 4 + 3]};
            const float r[3] = {cf[0] - d_c[0], cf[1] - d_c[1], cf[2] - d_c[2]};
            accumulate(e, bf, q_c, v_c, r);
        }
        const uint32_t dxy = p->dx * p->dy;
        const int cdx = (int)((p->di % dxy) % p->dx), cdy = (int)((p->di % dxy) / p->dx), cdz = (int)(p->di / dxy);
        uint32_t offset = p->n_lod_own;
        for (uint32_t d = 0; d < dxy * p->dz; d++) {
            if (d == p->di) continue;
            const int ddx = cdx - (int)((d % dxy) % p->dx), ddy = cdy - (int)((d % dxy) / p->dx), ddz = cdz - (int)(d / dxy);
            const uint32_t dist = (uint32_t)imax(abs(ddx), imax(abs(ddy), abs(ddz)));
            const uint32_t depth = (uint32_t)imax(0, (int)p->lod_depth - (int)dist);
            const uint32_t n_lod_fd = to_d(&k, 1u << depth);
            for (uint32_t l = 0; l < n_lod_fd; l++) {
                float lc[3];
                lod_coordinates(&k, l, depth, lc);
                /* quirk Q8 (halo-inclusive shift) and Q18: `domain_diff.x * DEF_NX` is int * uint = uint in OpenCL C, so a
                 * negative difference wraps to ~4.29e9 before the float conversion (sim.cl:970-972) */
                lc[0] -= (float)((uint32_t)ddx * p->nx);
                lc[1] -= (float)((uint32_t)ddy * p->ny);
                lc[2] -= (float)((uint32_t)ddz * p->nz);
                const float q_c = b->QU_lod[(offset + l) * 4 + 0];
                const float v_c[3] = {b->QU_lod[(offset + l) * 4 + 1], b->QU_lod[(offset + l) * 4 + 2], b->QU_lod[(offset + l) * 4 + 3]};
                const float r[3] = {cf[0] - lc[0], cf[1] - lc[1], cf[2] - lc[2]};
                accumulate(e, bf, q_c, v_c, r);
            }
            offset += n_lod_fd;
        }
        b->E_dyn[n] = b->E_stat[n] + p->ke * e[0];
        b->E_dyn[(uint64_t)n + N] = b->E_stat[(uint64_t)n + N] + p->ke * e[1];
        b->E_dyn[(uint64_t)n + N * 2ul] = b->E_stat[(uint64_t)n + N * 2ul] + p->ke * e[2];
        b->B_dyn[n] = b->B_stat[n] + p->kmu * bf[0];
        b->B_dyn[(uint64_t)n + N] = b->B_stat[(uint64_t)n + N] + p->kmu * bf[1];
        b->B_dyn[(uint64_t)n + N * 2ul] = b->B_stat[(uint64_t)n + N * 2ul] + p->kmu * bf[2];
    }
}

/* =============================== halo transfer kernels, sim.cl:1006-1147 =============================== */
static uint32_t face_cell(const OraParams* p, uint32_t a, uint32_t direction, uint32_t layer) {
    uint32_t x, y, z;
    if (direction == 0u) { x = layer; y = a % p->ny; z = a / p->ny; }
    else if (direction == 1u) { x = a / p->nz; y = layer; z = a % p->nz; }
    else { x = a % p->nx; y = a / p->nx; z = layer; }
    return x + (y + z * p->ny) * p->nx;
}
static uint32_t face_area(const OraParams* p, uint32_t direction) {
    return direction == 0u ? p->ny * p->nz : direction == 1u ? p->nz * p->nx : p->nx * p->ny;
}
static inline void copy_word(size_t s, void* dst, uint64_t di, const void* src, uint64_t si) {
    memcpy((char*)dst + di * s, (const char*)src + si * s, s);  /* fpxx_copy: raw storage words */
}
/* field: 0 fi, 1 rho_u_flags, 2 ei, 3 fqi (TransferField, types.rs:105-111); insert: 0 extract, 1 insert */
void ora_transfer(const OraParams* p, const OraBuffers* b, int field, int insert, uint32_t direction, uint64_t t) {
    Ctx k;
    make_ctx(&k, p);
    const uint32_t A = face_area(p, direction);
    const uint32_t L = direction == 0u ? p->nx : direction == 1u ? p->ny : p->nz;
    const size_t s = ddf_size(&k);
    for (uint32_t a = 0; a < A; a++) {
        for (uint32_t side01 = 0; side01 < 2; side01++) {
            const uint32_t layer = insert ? (side01 == 0 ? L - 1u : 0u) : (side01 == 0 ? L - 2u : 1u);
            const uint32_t n = face_cell(p, a, direction, layer);
            uint8_t* tb = side01 == 0 ? b->transfer_p : b->transfer_m;
            const uint32_t side = 2u * direction + side01;
            if (field == 0 || field == 2) {
                void* fi = field == 0 ? b->fi : b->ei;
                uint32_t j[QMAX];
                neighbors(&k, n, j);
                for (uint32_t bb = 0u; bb < k.transfers; bb++) {
                    const uint32_t i = k.ttab[side * k.transfers + bb];
                    if (!insert) {
                        const uint64_t index = index_f(&k, i % 2u ? j[i] : n, t % 2ul ? (i % 2u ? i + 1u : i - 1u) : i);
                        copy_word(s, tb, (uint64_t)bb * A + a, fi, index);
                    } else {
                        const uint64_t index = index_f(&k, i % 2u ? n : j[i - 1u], t % 2ul ? i : (i % 2u ? i + 1u : i - 1u));
                        copy_word(s, fi, index, tb, (uint64_t)bb * A + a);
                    }
                }
            } else if (field == 3) {
                uint32_t j7[7];
                neighbors_a(&k, n, j7);
                const uint32_t i = side + 1u;
                if (!insert) {
                    const uint64_t index = index_f(&k, i % 2u ? j7[i] : n, t % 2ul ? (i % 2u ? i + 1u : i - 1u) : i);
                    copy_word(s, tb, a, b->fqi, index);
                } else {
                    const uint64_t index = index_f(&k, i % 2u ? n : j7[i - 1u], t % 2ul ? i : (i % 2u ? i + 1u : i - 1u));
                    copy_word(s, b->fqi, index, tb, a);
                }
            } else {
                float* tf = (float*)tb;
                if (!insert) {
                    tf[a] = b->rho[n]; tf[A + a] = b->u[n]; tf[2u * A + a] = b->u[k.N + n]; tf[3u * A + a] = b->u[2ul * k.N + n];
                    tb[16u * (uint64_t)A + a] = b->flags[n];
                } else {
                    b->rho[n] = tf[a]; b->u[n] = tf[A + a]; b->u[k.N + n] = tf[2u * A + a]; b->u[2ul * k.N + n] = tf[3u * A + a];
                    b->flags[n] = tb[16u * (uint64_t)A + a];
                }
            }
        }
    }
}

/* =============================== voxelize_mesh, sim.cl:1150-1231 =============================== */
static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline void cross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
void ora_voxelize_mesh(const OraParams* p, const OraBuffers* b, uint32_t direction, uint8_t flag, const float* p0, const float* p1,
                       const float* p2, const float* bbu, float mpc_x, float mpc_y, float mpc_z) {
    Ctx k;
    make_ctx(&k, p);
    const uint32_t A = face_area(p, direction);
    const uint32_t triangle_number = as_uint(bbu[0]);
    const float x0 = bbu[1], y0 = bbu[2], z0 = bbu[3], x1 = bbu[4], y1 = bbu[5], z1 = bbu[6];
    for (uint32_t a = 0; a < A; a++) {
        uint32_t xyz[3];
        if (direction == 0u) { xyz[0] = (uint32_t)clampi((int)x0 - p->ox, 0, (int)p->nx - 1); xyz[1] = a % p->ny; xyz[2] = a / p->ny; }
        else if (direction == 1u) { xyz[0] = a / p->nz; xyz[1] = (uint32_t)clampi((int)y0 - p->oy, 0, (int)p->ny - 1); xyz[2] = a % p->nz; }
        else { xyz[0] = a % p->nx; xyz[1] = a / p->nx; xyz[2] = (uint32_t)clampi((int)z0 - p->oz, 0, (int)p->nz - 1); }
        const float offset[3] = {0.5f * (float)((int)p->nx + 2 * p->ox) - 0.5f, 0.5f * (float)((int)p->ny + 2 * p->oy) - 0.5f,
                                 0.5f * (float)((int)p->nz + 2 * p->oz) - 0.5f};
        const float pos[3] = {(float)xyz[0] + 0.5f - 0.5f * (float)p->nx, (float)xyz[1] + 0.5f - 0.5f * (float)p->ny,
                              (float)xyz[2] + 0.5f - 0.5f * (float)p->nz};
        const float ro[3] = {pos[0] + offset[0], pos[1] + offset[1], pos[2] + offset[2]};
        const float rd[3] = {(float)(direction == 0u), (float)(direction == 1u), (float)(direction == 2u)};
        uint32_t intersections = 0u, intersections_check = 0u;
        uint16_t distances[64];
        memset(distances, 0, sizeof(distances));
        const int condition = direction == 0u   ? (ro[1] < y0 || ro[2] < z0 || ro[1] >= y1 || ro[2] >= z1)
                              : direction == 1u ? (ro[0] < x0 || ro[2] < z0 || ro[0] >= x1 || ro[2] >= z1)
                                                : (ro[0] < x0 || ro[1] < y0 || ro[0] >= x1 || ro[1] >= y1);
        if (condition) continue;
        for (uint32_t i = 0u; i < triangle_number; i++) {
            const float* a0 = p0 + 3u * i; const float* a1 = p1 + 3u * i; const float* a2 = p2 + 3u * i;
            const float u[3] = {a1[0] - a0[0], a1[1] - a0[1], a1[2] - a0[2]}, v[3] = {a2[0] - a0[0], a2[1] - a0[1], a2[2] - a0[2]},
                        w[3] = {ro[0] - a0[0], ro[1] - a0[1], ro[2] - a0[2]};
            float h[3], q[3];
            cross3(rd, v, h);
            cross3(w, u, q);
            const float f = 1.0f / dot3(u, h), s = f * dot3(w, h), tt = f * dot3(rd, q), d = f * dot3(v, q);
            if (s >= 0.0f && s < 1.0f && tt >= 0.0f && s + tt < 1.0f) {
                if (d > 0.0f) {
                    if (intersections < 64u && d < 65536.0f) distances[intersections] = (uint16_t)d;
                    intersections++;
                } else {
                    intersections_check++;
                }
            }
        }
        for (int i = 1; i < (int)intersections && i < 64; i++) {
            const uint16_t tv = distances[i];
            int jj = i - 1;
            while (jj >= 0 && distances[jj] > tv) { distances[jj + 1] = distances[jj]; jj--; }
            distances[jj + 1] = tv;
        }
        int inside = (intersections % 2u) && (intersections_check % 2u);
        uint32_t intersection = intersections % 2u != intersections_check % 2u;
        const uint32_t h0 = xyz[direction];
        const uint32_t hmax = direction == 0u   ? (uint32_t)clampi((int)x1 - p->ox, 0, (int)p->nx)
                              : direction == 1u ? (uint32_t)clampi((int)y1 - p->oy, 0, (int)p->ny)
                                                : (uint32_t)clampi((int)z1 - p->oz, 0, (int)p->nz);
        const uint32_t last = intersections - 1u < 63u ? intersections - 1u : 63u;
        const uint32_t hmesh = h0 + (uint32_t)distances[last];
        for (uint32_t h = h0; h < hmax; h++) {
            while (intersection < intersections && h > h0 + (uint32_t)distances[intersection < 63u ? intersection : 63u]) {
                inside = !inside;
                intersection++;
            }
            inside = inside && (intersection < intersections && h < hmesh);
            uint32_t c[3] = {xyz[0], xyz[1], xyz[2]};
            c[direction] = h;
            const uint64_t n = c[0] + (c[1] + (uint64_t)c[2] * p->ny) * p->nx;
            if (inside) {
                const uint8_t flagsn = (uint8_t)((b->flags[n] & (uint8_t)~TYPE_BO) | flag);
                if (p->magneto_hydro) {
                    if (flag & TYPE_M) {
                        b->B_dyn[n] = mpc_x; b->B_dyn[k.N + n] = mpc_y; b->B_dyn[2ul * k.N + n] = mpc_z;
                    } else if ((flag & TYPE_F) || (flag & TYPE_C)) {
                        b->B_dyn[n] = mpc_x;
                    }
                }
                b->flags[n] = flagsn;
            }
        }
    }
}

/* =============================== static fields, sim.cl:1234-1300 =============================== */
/* psi_from_mesh: psi (padded grid) -> E_dyn, M <- B_dyn (domain.rs:279-281) */
void ora_psi_from_mesh(const OraParams* p, const OraBuffers* b) {
    Ctx k;
    make_ctx(&k, p);
    set_threads(p);
    const uint32_t lx = p->nx + 2, ly = p->ny + 2, lz = p->nz + 2;
    const int64_t total = (int64_t)lx * ly * lz;
    /* order-preserving list of magnet cells: the reference tests the flag inside its O(N^2) loop; skipping
     * non-magnet cells up front leaves the float sum and its order unchanged */
    uint32_t* src = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)k.N);
    size_t ns = 0;
    for (uint64_t i = 0; i < k.N; i++)
        if (b->flags[i] & TYPE_M) src[ns++] = (uint32_t)i;
    float* psi = b->E_dyn;
    const float* M = b->B_dyn;
    PAR_FOR
    for (int64_t n = 0; n < total; n++) {
        const uint32_t t = (uint32_t)(n % ((int64_t)lx * ly));
        const float c[3] = {(float)(t % lx), (float)(t / lx), (float)(uint32_t)(n / ((int64_t)lx * ly))};
        float psic = 0.0f;
        for (size_t s = 0; s < ns; s++) {
            const uint32_t i = src[s];
            uint32_t x, y, z;
            coordinates(&k, i, &x, &y, &z);
            const float cd[3] = {c[0] - (float)(x + 1u), c[1] - (float)(y + 1u), c[2] - (float)(z + 1u)};
            const float l = sqrtf(cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2]);
            if (!(l == 0.0f)) {
                const float mag[3] = {M[i], M[i + k.N], M[i + k.N * 2]};
                psic += dot3(cd, mag) / cb(l);
            }
        }
        psi[n] = (float)((double)psic / (double)(4.0f * M_PI));  /* 4.0f * M_PI is a double product, sim.cl:1250 */
    }
    free(src);
}
void ora_static_b_from_mesh(const OraParams* p, const OraBuffers* b) {
    Ctx k;
    make_ctx(&k, p);
    const float* psi = b->E_dyn;
    const uint32_t l0 = p->nx + 2, l1 = p->ny + 2;
    for (uint64_t nn = 0; nn < k.N; nn++) {
        const uint32_t n = (uint32_t)nn;
        if (is_halo(&k, n)) continue;
        if ((b->flags[n] & TYPE_S) == TYPE_S) continue;
        uint32_t x, y, z;
        coordinates(&k, n, &x, &y, &z);
        const uint32_t m = (x + 1) + ((y + 1) + (z + 1) * l1) * l0, yo = l0, zo = l0 * l1;
        const float nk = -p->kmu0;
        b->B_stat[n] += nk * ((psi[m + 1] - psi[m - 1]) / 2.0f);
        b->B_stat[n + k.N] += nk * ((psi[m + yo] - psi[m - yo]) / 2.0f);
        b->B_stat[n + 2ul * k.N] += nk * ((psi[m + zo] - psi[m - zo]) / 2.0f);
    }
}
/* static_e_from_mesh: E (E_stat, or any 3N target) += k_e sum_charged q r/|r|^3; charge in the x-plane of B_dyn */
void ora_static_e_from_mesh(const OraParams* p, const OraBuffers* b, float* E) {
    Ctx k;
    make_ctx(&k, p);
    set_threads(p);
    uint32_t* src = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)k.N);
    size_t ns = 0;
    for (uint64_t i = 0; i < k.N; i++)
        if ((b->flags[i] & TYPE_F) || (b->flags[i] & TYPE_C)) src[ns++] = (uint32_t)i;
    const float* C = b->B_dyn;
    PAR_FOR
    for (int64_t nn = 0; nn < (int64_t)k.N; nn++) {
        const uint32_t n = (uint32_t)nn;
        if (is_halo(&k, n)) continue;
        if ((b->flags[n] & TYPE_S) == TYPE_S) continue;
        uint32_t cx, cy, cz;
        coordinates(&k, n, &cx, &cy, &cz);
        float Ec[3] = {0.0f, 0.0f, 0.0f};
        for (size_t s = 0; s < ns; s++) {
            const uint32_t i = src[s];
            uint32_t x, y, z;
            coordinates(&k, i, &x, &y, &z);
            const float cd[3] = {(float)cx - (float)x, (float)cy - (float)y, (float)cz - (float)z};
            const float l = sqrtf(cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2]);
            if (!(l == 0.0f)) {
                const float charge = C[i], l3 = cb(l);
                Ec[0] += cd[0] * charge / l3; Ec[1] += cd[1] * charge / l3; Ec[2] += cd[2] * charge / l3;
            }
        }
        E[n] += Ec[0] * p->ke;
        E[n + k.N] += Ec[1] * p->ke;
        E[n + 2ul * k.N] += Ec[2] * p->ke;
    }
    free(src);
}

/* =============================== small probes used by the unit tests =============================== */
void ora_codec(const OraParams* p, const void* in, void* out, uint64_t count, int dir) {  /* 0: float -> stored, 1: stored -> float */
    Ctx k;
    make_ctx(&k, p);
    if (dir == 0) for (uint64_t o = 0; o < count; o++) ddf_store(&k, out, o, ((const float*)in)[o]);
    else for (uint64_t o = 0; o < count; o++) ((float*)out)[o] = ddf_load(&k, in, o);
}
void ora_neighbors(const OraParams* p, uint32_t n, uint32_t* j) {
    Ctx k;
    make_ctx(&k, p);
    neighbors(&k, n, j);
}
uint32_t ora_lod_index(const OraParams* p, uint32_t n, uint32_t d) {
    Ctx k;
    make_ctx(&k, p);
    return lod_index(&k, n, d);
}
uint64_t ora_nan_stores(int reset) {
    const uint64_t v = __atomic_load_n(&g_nan_stores, __ATOMIC_RELAXED);
    if (reset) __atomic_store_n(&g_nan_stores, 0, __ATOMIC_RELAXED);
    return v;
}
int ora_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
