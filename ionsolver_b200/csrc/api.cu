// api.cu -- the C ABI of libionsolver_b200.so (see include/ionsolver_b200.h for the reference call sites).
//
// One ion_domain_t = one LbmDomain of the reference (/root/reference/src/lbm/domain.rs:20-80): a CUDA device,
// an in-order stream (the OpenCL queue), the buffers of domain.rs:151-211,311-322 and the kernel bindings.
// No CPU fallback exists anywhere in this file: every compute entry point launches sm_100a kernels.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "domain_internal.cuh"

namespace ion {
// launchers implemented in the kernel translation units
template <int VS>
cudaError_t launch_stream_collide_vs(const KArgs& a, int fp, bool mhd, bool trt, bool ecr, uint64_t t, float fx, float fy, float fz,
                                     cudaStream_t s);
template <int VS> cudaError_t launch_update_fields_vs(const KArgs& a, int fp, uint64_t t, cudaStream_t s);
template <int VS> cudaError_t launch_initialize_vs(const KArgs& a, int fp, bool mhd, cudaStream_t s);
cudaError_t launch_update_e_b(const KArgs& a, void* scratch_sources, cudaStream_t s, uint64_t* launches, bool exact);
cudaError_t launch_update_e_b_foreign(const KArgs& a, void* scratch_sources, cudaStream_t s, uint64_t* launches, int skip_domain);
int eb_fft_plan_foreign_domain(const EbFftPlan* p);
bool eb_fft_plan_far_handled(const EbFftPlan* p);
cudaError_t eb_fft_create(const KArgs& a, size_t budget_bytes, cudaStream_t s, EbFftPlan** out, uint64_t* launches);
cudaError_t eb_fft_launch(const EbFftPlan* p, const KArgs& a, cudaStream_t s, uint64_t* launches);
void eb_fft_destroy(EbFftPlan* p);
size_t eb_fft_plan_bytes(const EbFftPlan* p);
uint32_t eb_fft_plan_tasks(const EbFftPlan* p);
cudaError_t launch_lod_fold(const KArgs& a, cudaStream_t s);
size_t lod_source_bytes(uint32_t lod_depth, uint32_t n_lod_own, uint32_t dx, uint32_t dy, uint32_t dz, uint32_t di);
cudaError_t launch_clear_qu_lod(const KArgs& a, cudaStream_t s);
cudaError_t launch_lod_deposit_ordered(const KArgs& a, cudaStream_t s);
cudaError_t launch_lod_gather(const KArgs& a, cudaStream_t s, uint64_t* launches);
cudaError_t launch_transfer(const KArgs& p, int vs, int fp, int transfer_field, int insert, uint32_t direction, uint64_t t,
                            cudaStream_t s);
cudaError_t launch_voxelize(const KArgs& a, uint32_t direction, uint8_t flag, const float* p0, const float* p1, const float* p2,
                            uint32_t tri, const float* bb6, float mx, float my, float mz, int mhd, cudaStream_t s);
size_t compaction_scratch_bytes(uint64_t N);
cudaError_t count_sources(const KArgs& a, uint8_t mask, uint32_t* counts, uint32_t* host_total, cudaStream_t s);
size_t field_source_bytes(uint32_t count);
cudaError_t launch_precompute_b(const KArgs& a, uint32_t* counts, void* table, cudaStream_t s, int fast);
cudaError_t launch_precompute_fft(const KArgs& a, int which, float* E, uint32_t* counts, void* table, uint32_t total, cudaStream_t s, uint64_t* launches,
                                  const char** why);
cudaError_t launch_precompute_e(const KArgs& a, float* E, uint32_t* counts, void* table, cudaStream_t s);

// FP32 issue-peak probe for bench.py's compute roofline (MEASURED_PEAKS.json only carries HBM and BF16 numbers):
// 16 independent FMA chains per thread, scalar FFMA (PACKED = false) or fma.rn.f32x2 (PACKED = true)
template <bool PACKED> __global__ void __launch_bounds__(256) k_fma_peak(float* out, float a, float b, int iters) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 x = make_float2(a, a * 1.0001f), y = make_float2(b, b * 0.9999f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (PACKED) {
                acc[i] = __ffma2_rn(acc[i], x, y);
            } else {
                acc[i].x = fmaf(acc[i].x, x.x, y.x);
                acc[i].y = fmaf(acc[i].y, x.y, y.y);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DDF storage codec probe (test hook, like ref_codec of the oracle driver): dir 0 float -> stored, 1 stored -> float
template <int FP> __global__ void k_codec(const void* in, void* out, uint64_t count, int dir) {
    typedef typename Codec<FP>::store_t S;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (dir == 0) reinterpret_cast<S*>(out)[i] = Codec<FP>::enc(reinterpret_cast<const float*>(in)[i]);
    else reinterpret_cast<float*>(out)[i] = Codec<FP>::dec(reinterpret_cast<const S*>(in)[i]);
}

__global__ void k_fill_f32(float* p, uint64_t n, float v) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
}  // namespace ion

using namespace ion;

namespace {
// frees device memory / destroys events on every exit path of the probe entry points (ION_CUDA returns early on errors)
struct DevMem {
    void* p = nullptr;
    ~DevMem() { if (p) cudaFree(p); }
};
struct Event {
    cudaEvent_t e = nullptr;
    ~Event() { if (e) cudaEventDestroy(e); }
};
}  // namespace

static thread_local char g_err[512] = "";
namespace ion {
std::atomic<uint64_t> g_launches{0};
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}
void set_transfer_ptrs(ion_domain* d) {
    d->k.transfer_p = (uint8_t*)d->buf[ION_FIELD_TRANSFER_P];
    d->k.transfer_m = (uint8_t*)d->buf[ION_FIELD_TRANSFER_M];
}
}  // namespace ion

static int set_q(int vs, int* q, int* dim, int* tr) {
    switch (vs) {
        case ION_D2Q9: *q = 9; *dim = 2; *tr = 3; return 0;
        case ION_D3Q15: *q = 15; *dim = 3; *tr = 5; return 0;
        case ION_D3Q19: *q = 19; *dim = 3; *tr = 5; return 0;
        case ION_D3Q27: *q = 27; *dim = 3; *tr = 9; return 0;
    }
    return -1;
}

extern "C" {

const char* ion_last_error_string(void) { return g_err; }
uint32_t ion_abi_version(void) { return ION_ABI_VERSION; }
uint64_t ion_kernel_launch_count(void) { return g_launches.load(); }

int ion_codec_probe(int device, int float_type, int dir, const void* host_in, void* host_out, uint64_t count) {
    if (!host_in || !host_out) return fail(ION_ERR_INVALID, "NULL argument");
    if (float_type < ION_FP16S || float_type > ION_FP32 || dir < 0 || dir > 1) return fail(ION_ERR_INVALID, "bad float_type / dir");
    ION_CUDA(cudaSetDevice(device));
    const size_t ss = float_type == ION_FP32 ? 4 : 2;
    const size_t in_b = count * (dir == 0 ? 4 : ss), out_b = count * (dir == 0 ? ss : 4);
    DevMem gin, gout;
    ION_CUDA(cudaMalloc(&gin.p, in_b ? in_b : 1));
    ION_CUDA(cudaMalloc(&gout.p, out_b ? out_b : 1));
    void *din = gin.p, *dout = gout.p;
    ION_CUDA(cudaMemcpy(din, host_in, in_b, cudaMemcpyHostToDevice));
    const unsigned grid = (unsigned)((count + 255) / 256);
    if (count) {
        if (float_type == ION_FP32) k_codec<ION_FP32><<<grid, 256>>>(din, dout, count, dir);
        else if (float_type == ION_FP16S) k_codec<ION_FP16S><<<grid, 256>>>(din, dout, count, dir);
        else k_codec<ION_FP16C><<<grid, 256>>>(din, dout, count, dir);
        g_launches++;
    }
    cudaError_t e = cudaMemcpy(host_out, dout, out_b, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "codec probe");
    return ION_OK;
}

int ion_measure_fma_peak(int device, int packed, double* fma_per_s) {
    if (!fma_per_s) return fail(ION_ERR_INVALID, "NULL argument");
    ION_CUDA(cudaSetDevice(device));
    int sms = 0;
    ION_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    DevMem gout;
    ION_CUDA(cudaMalloc(&gout.p, sizeof(float) * blocks * threads));
    float* out = (float*)gout.p;
    Event g0, g1;
    ION_CUDA(cudaEventCreate(&g0.e));
    ION_CUDA(cudaEventCreate(&g1.e));
    cudaEvent_t e0 = g0.e, e1 = g1.e;
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        if (packed) k_fma_peak<true><<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
        else k_fma_peak<false><<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
        g_launches++;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "fma peak probe");
    *fma_per_s = (double)blocks * threads * iters * 32.0 / (best * 1e-3);
    return ION_OK;
}

int ion_device_count(int* count) {
    if (!count) return fail(ION_ERR_INVALID, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount"); }
    int usable = 0;
    for (int d = 0; d < n; d++) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) usable++;
    }
    *count = usable;
    return ION_OK;
}

int ion_domain_create(const IonParams* p, int device, ion_domain_t** out) {
    if (!p || !out) return fail(ION_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (p->abi_version != ION_ABI_VERSION) return fail(ION_ERR_INVALID, "IonParams.abi_version %u != %u", p->abi_version, ION_ABI_VERSION);
    int q, dim, tr;
    if (set_q((int)p->velocity_set, &q, &dim, &tr)) return fail(ION_ERR_INVALID, "unknown velocity_set %u", p->velocity_set);
    if (p->relaxation_time > ION_TRT || p->float_type > ION_FP32) return fail(ION_ERR_INVALID, "unknown relaxation_time/float_type");
    if (!p->nx || !p->ny || !p->nz || !p->dx || !p->dy || !p->dz) return fail(ION_ERR_INVALID, "zero lattice or domain extent");
    const uint64_t n = (uint64_t)p->nx * p->ny * p->nz;
    if (n > 0xFFFFFFFFull) return fail(ION_ERR_UNSUPPORTED, "more than 2^32 cells per domain (cell index is uint, sim_kernels.cl:484)");
    if (p->ny > 65535u || p->nz > 65535u) return fail(ION_ERR_UNSUPPORTED, "ny/nz exceed the CUDA grid limit 65535");
    const bool mhd = p->ext & ION_EXT_MAGNETO_HYDRO, ecr = p->ext & ION_EXT_SUBGRID_ECR;
    if (mhd) {
        if (!(p->ext & ION_EXT_VOLUME_FORCE)) return fail(ION_ERR_UNSUPPORTED, "MAGNETO_HYDRO needs VOLUME_FORCE (c_tau, sim_kernels.cl:519,650)");
        if (dim != 3) return fail(ION_ERR_UNSUPPORTED, "MAGNETO_HYDRO on D2Q9 divides by DEF_NZ/2^depth = 0 in the reference (sim_kernels.cl:431)");
        if (p->lod_depth > 4u) return fail(ION_ERR_UNSUPPORTED, "mhd_lod_depth > 4 shifts by >= 32 bits in the reference (sim_kernels.cl:908)");
        const uint32_t nd = 1u << p->lod_depth;
        if (p->nx < nd || p->ny < nd || p->nz < nd) return fail(ION_ERR_UNSUPPORTED, "domain smaller than 2^lod_depth cells: lod_index divides by zero (sim_kernels.cl:429)");
        if (p->n_lod_own == 0 || p->n_lod < p->n_lod_own) return fail(ION_ERR_INVALID, "bad LOD counts");
        if ((uint64_t)(p->nx + 2u) * (p->ny + 2u) * (p->nz + 2u) > 3ull * n) return fail(ION_ERR_UNSUPPORTED, "padded psi grid does not fit the E_dyn scratch (sim_kernels.cl:1234, domain.rs:280)");
    }
    if (ecr && !mhd) return fail(ION_ERR_UNSUPPORTED, "SUBGRID_ECR needs MAGNETO_HYDRO (its code sits inside the MHD block, sim_kernels.cl:556)");
    const bool deterministic = mhd && (p->ext & ION_EXT_DETERMINISTIC);
    if (deterministic && p->lod_depth > 0u) {
        const uint32_t nd = 1u << p->lod_depth;
        if (p->nx % nd || p->ny % nd || p->nz % nd) return fail(ION_ERR_UNSUPPORTED, "ION_EXT_DETERMINISTIC needs nx, ny, nz (incl. halos) divisible by 2^lod_depth = %u", nd);
    }

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return fail(ION_ERR_NO_DEVICE, "no CUDA device (%s); there is no CPU fallback", cudaGetErrorName(ce));
    if (device < 0 || device >= ndev) return fail(ION_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
    int major = 0;
    ION_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return fail(ION_ERR_NO_DEVICE, "device %d has compute capability %d.x; kernels are built for sm_100a only", device, major);
    ION_CUDA(cudaSetDevice(device));

    ion_domain* d = new (std::nothrow) ion_domain();
    if (!d) return fail(ION_ERR_INVALID, "out of host memory");
    memset(d, 0, sizeof(*d));
    d->params = *p;
    d->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete d; return cuda_fail(e, "cudaStreamCreate"); }

    const size_t s = p->float_type == ION_FP32 ? 4 : 2;
    size_t* B = d->bytes;
    B[ION_FIELD_FI] = n * q * s;  // domain.rs:151-158
    B[ION_FIELD_RHO] = n * 4;
    B[ION_FIELD_U] = n * 12;
    B[ION_FIELD_FLAGS] = n;
    if (p->ext & ION_EXT_FORCE_FIELD) B[ION_FIELD_F] = n * 12;  // domain.rs:168
    if (mhd) {                                                   // domain.rs:172-193
        B[ION_FIELD_E_STAT] = B[ION_FIELD_B_STAT] = B[ION_FIELD_E_DYN] = B[ION_FIELD_B_DYN] = n * 12;
        B[ION_FIELD_FQI] = n * 7 * s;
        B[ION_FIELD_EI] = n * q * s;
        B[ION_FIELD_Q] = n * 4;
        B[ION_FIELD_QU_LOD] = (size_t)p->n_lod * 16;
    }
    if (ecr) {  // domain.rs:200-211
        B[ION_FIELD_E_VAR] = n * 12;
        B[ION_FIELD_ETI] = n * 7 * s;
        B[ION_FIELD_ET] = n * 4;
    }
    size_t a_max = 0;  // domain.rs:311-318
    if (p->dx > 1) a_max = a_max > (size_t)p->ny * p->nz ? a_max : (size_t)p->ny * p->nz;
    if (p->dy > 1) a_max = a_max > (size_t)p->nx * p->nz ? a_max : (size_t)p->nx * p->nz;
    if (p->dz > 1) a_max = a_max > (size_t)p->nx * p->ny ? a_max : (size_t)p->nx * p->ny;
    const size_t per = (size_t)tr * s > 17 ? (size_t)tr * s : 17;
    B[ION_FIELD_TRANSFER_P] = B[ION_FIELD_TRANSFER_M] = a_max * per;

    for (int f = 0; f < ION_FIELD_COUNT; f++) {
        if (!B[f]) continue;
        e = cudaMalloc(&d->buf[f], B[f]);
        if (e == cudaSuccess) e = cudaMemsetAsync(d->buf[f], 0, B[f], d->stream);
        if (e != cudaSuccess) { ion_domain_destroy(d); return cuda_fail(e, "cudaMalloc/cudaMemset of a domain buffer"); }
    }
    k_fill_f32<<<(unsigned)((n + 255) / 256), 256, 0, d->stream>>>((float*)d->buf[ION_FIELD_RHO], n, 1.0f);  // rho = 1, domain.rs:156
    g_launches++;
    if (a_max) {  // spare face buffers: receive side of the device-resident halo exchange
        e = cudaMalloc(&d->alt_p, B[ION_FIELD_TRANSFER_P]);
        if (e == cudaSuccess) e = cudaMalloc(&d->alt_m, B[ION_FIELD_TRANSFER_M]);
        if (e != cudaSuccess) { ion_domain_destroy(d); return cuda_fail(e, "cudaMalloc spare transfer buffers"); }
    }
    e = cudaEventCreateWithFlags(&d->ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->halo_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { ion_domain_destroy(d); return cuda_fail(e, "cudaEventCreate / halo stream"); }
    if (mhd) {
        if (deterministic) {
            e = cudaMalloc((void**)&d->lod_u, n * 12);
            if (e != cudaSuccess) { ion_domain_destroy(d); return cuda_fail(e, "cudaMalloc lod_u"); }
        }
        if (!deterministic && p->lod_depth > 0u) {  // private replicas of the finest LOD level for the deposit (stream_collide.cuh)
            const size_t rep_bytes = (size_t)ION_LOD_REPLICAS * ((size_t)1 << (3u * p->lod_depth)) * 16u;
            e = cudaMalloc((void**)&d->lod_rep, rep_bytes);
            if (e == cudaSuccess) e = cudaMemset(d->lod_rep, 0, rep_bytes);
            if (e != cudaSuccess) { ion_domain_destroy(d); return cuda_fail(e, "cudaMalloc lod_rep"); }
        }
        e = cudaMalloc(&d->lod_sources, lod_source_bytes(p->lod_depth, p->n_lod_own, p->dx, p->dy, p->dz, p->di));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d->cp_counts, compaction_scratch_bytes(n));
        if (e != cudaSuccess) { ion_domain_destroy(d); return cuda_fail(e, "cudaMalloc scratch"); }
    }

    KArgs& k = d->k;
    k.fi = d->buf[ION_FIELD_FI];
    k.rho = (float*)d->buf[ION_FIELD_RHO];
    k.u = (float*)d->buf[ION_FIELD_U];
    k.flags = (uint8_t*)d->buf[ION_FIELD_FLAGS];
    k.F = (const float*)d->buf[ION_FIELD_F];
    k.E_stat = (float*)d->buf[ION_FIELD_E_STAT];
    k.B_stat = (float*)d->buf[ION_FIELD_B_STAT];
    k.E_dyn = (float*)d->buf[ION_FIELD_E_DYN];
    k.B_dyn = (float*)d->buf[ION_FIELD_B_DYN];
    k.fqi = d->buf[ION_FIELD_FQI];
    k.ei = d->buf[ION_FIELD_EI];
    k.Q = (float*)d->buf[ION_FIELD_Q];
    k.QU_lod = (float*)d->buf[ION_FIELD_QU_LOD];
    k.lod_u = d->lod_u;
    k.lod_rep = d->lod_rep;
    k.lod_rep_mask = ION_LOD_REPLICAS - 1u;
    k.lod_rep_entries = d->lod_rep ? (1u << (3u * p->lod_depth)) : 0u;
    d->deterministic = deterministic;
    k.E_var = (const float*)d->buf[ION_FIELD_E_VAR];
    k.eti = d->buf[ION_FIELD_ETI];
    k.Et = (float*)d->buf[ION_FIELD_ET];
    k.transfer_p = (uint8_t*)d->buf[ION_FIELD_TRANSFER_P];
    k.transfer_m = (uint8_t*)d->buf[ION_FIELD_TRANSFER_M];
    k.N = n;
    k.nx = p->nx; k.ny = p->ny; k.nz = p->nz;
    k.dx = p->dx; k.dy = p->dy; k.dz = p->dz; k.di = p->di;
    k.ox = p->ox; k.oy = p->oy; k.oz = p->oz;
    k.ext = p->ext;
    k.w = p->w;
    k.ke = p->ke; k.kmu = p->kmu; k.kmu0 = p->kmu0; k.kkge = p->kkge; k.kme = p->kme; k.wq = p->wq;
    k.kkbme = p->kkbme; k.keabs = p->keabs;
    k.lod_depth = p->lod_depth; k.n_lod = p->n_lod; k.n_lod_own = p->n_lod_own;
    k.ecrf = 0.0f;
    e = cudaStreamSynchronize(d->stream);
    if (e != cudaSuccess) { ion_domain_destroy(d); return cuda_fail(e, "domain initialisation"); }
    *out = d;
    return ION_OK;
}

int ion_domain_destroy(ion_domain_t* d) {
    if (!d) return ION_OK;
    cudaSetDevice(d->device);
    if (d->stream) cudaStreamSynchronize(d->stream);
    for (int f = 0; f < ION_FIELD_COUNT; f++)
        if (d->buf[f]) cudaFree(d->buf[f]);
    if (d->lod_sources) cudaFree(d->lod_sources);
    if (d->eb_plan) eb_fft_destroy(d->eb_plan);
    if (d->cp_counts) cudaFree(d->cp_counts);
    if (d->lod_u) cudaFree(d->lod_u);
    if (d->lod_rep) cudaFree(d->lod_rep);
    if (d->alt_p) cudaFree(d->alt_p);
    if (d->alt_m) cudaFree(d->alt_m);
    if (d->lod_gather) cudaFree(d->lod_gather);
    if (d->halo_stream) { cudaStreamSynchronize(d->halo_stream); cudaStreamDestroy(d->halo_stream); }
    if (d->ev) cudaEventDestroy(d->ev);
    if (d->ev_fork) cudaEventDestroy(d->ev_fork);
    if (d->ev_join) cudaEventDestroy(d->ev_join);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
    return ION_OK;
}

int ion_domain_params(const ion_domain_t* d, IonParams* out) {
    if (!d || !out) return fail(ION_ERR_INVALID, "NULL argument");
    *out = d->params;
    return ION_OK;
}
int ion_domain_stream(const ion_domain_t* d, void** stream) {
    if (!d || !stream) return fail(ION_ERR_INVALID, "NULL argument");
    *stream = (void*)d->stream;
    return ION_OK;
}

static int check_field(const ion_domain_t* d, int field) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    if (field < 0 || field >= ION_FIELD_COUNT) return fail(ION_ERR_INVALID, "field id %d out of range", field);
    if (!d->buf[field]) return fail(ION_ERR_ABSENT, "buffer %d is not allocated for this configuration", field);
    return ION_OK;
}

int ion_buffer_size(const ion_domain_t* d, int field, size_t* bytes) {
    if (!d || !bytes) return fail(ION_ERR_INVALID, "NULL argument");
    if (field < 0 || field >= ION_FIELD_COUNT) return fail(ION_ERR_INVALID, "field id %d out of range", field);
    *bytes = d->bytes[field];
    return ION_OK;
}
int ion_buffer_device_ptr(const ion_domain_t* d, int field, void** dptr) {
    if (!d || !dptr) return fail(ION_ERR_INVALID, "NULL argument");
    if (field < 0 || field >= ION_FIELD_COUNT) return fail(ION_ERR_INVALID, "field id %d out of range", field);
    *dptr = d->buf[field];
    return ION_OK;
}
int ion_buffer_write(ion_domain_t* d, int field, const void* host, size_t off, size_t bytes) {
    int r = check_field(d, field);
    if (r) return r;
    if (!host && bytes) return fail(ION_ERR_INVALID, "NULL host pointer");
    if (off > d->bytes[field] || bytes > d->bytes[field] - off) return fail(ION_ERR_RANGE, "write of %zu bytes at %zu exceeds buffer %d (%zu bytes)", bytes, off, field, d->bytes[field]);
    ION_CUDA(cudaSetDevice(d->device));
    ION_CUDA(cudaMemcpyAsync((char*)d->buf[field] + off, host, bytes, cudaMemcpyHostToDevice, d->stream));
    ION_CUDA(cudaStreamSynchronize(d->stream));
    return ION_OK;
}
int ion_buffer_read(ion_domain_t* d, int field, void* host, size_t off, size_t bytes) {
    int r = check_field(d, field);
    if (r) return r;
    if (!host && bytes) return fail(ION_ERR_INVALID, "NULL host pointer");
    if (off > d->bytes[field] || bytes > d->bytes[field] - off) return fail(ION_ERR_RANGE, "read of %zu bytes at %zu exceeds buffer %d (%zu bytes)", bytes, off, field, d->bytes[field]);
    ION_CUDA(cudaSetDevice(d->device));
    ION_CUDA(cudaMemcpyAsync(host, (const char*)d->buf[field] + off, bytes, cudaMemcpyDeviceToHost, d->stream));
    ION_CUDA(cudaStreamSynchronize(d->stream));
    return ION_OK;
}
// Save-and-load in one call (file.rs:221-268 followed by :118-152 for the next state): every listed buffer is read to host_out[i]
// and then overwritten from host_in[i].  The downloads run on the domain's stream and the uploads on its second stream, each
// upload ordered behind the download of the same bytes only -- so host->device and device->host traffic overlap (PCIe is full
// duplex) instead of running back to back.  Host memory should be pinned.  Blocks until both directions are done.
int ion_buffer_swap(ion_domain_t* d, int n, const int* fields, void* const* host_out, const void* const* host_in, const size_t* bytes) {
    if (!d || !fields || !host_out || !host_in || !bytes || n < 0) return fail(ION_ERR_INVALID, "NULL argument");
    for (int i = 0; i < n; i++) {
        int r = check_field(d, fields[i]);
        if (r) return r;
        if (bytes[i] > d->bytes[fields[i]]) return fail(ION_ERR_RANGE, "swap of %zu bytes exceeds buffer %d (%zu bytes)", bytes[i], fields[i], d->bytes[fields[i]]);
        if ((!host_out[i] || !host_in[i]) && bytes[i]) return fail(ION_ERR_INVALID, "NULL host pointer");
    }
    if (d->halo_active) return fail(ION_ERR_INVALID, "ion_buffer_swap inside a halo fork");
    ION_CUDA(cudaSetDevice(d->device));
    ION_CUDA(cudaEventRecord(d->ev_fork, d->stream));  // the upload stream starts behind everything queued so far
    ION_CUDA(cudaStreamWaitEvent(d->halo_stream, d->ev_fork, 0));
    // in chunks: the upload of a chunk waits for the download of the same chunk only, so the upload direction lags the download
    // direction by 8 MiB instead of by a whole buffer (a wait refers to the event's latest record at the time of the call, so one
    // event serves every chunk)
    const size_t chunk = (size_t)8 << 20;
    for (int i = 0; i < n; i++) {
        for (size_t off = 0; off < bytes[i]; off += chunk) {
            const size_t len = bytes[i] - off < chunk ? bytes[i] - off : chunk;
            ION_CUDA(cudaMemcpyAsync((char*)host_out[i] + off, (const char*)d->buf[fields[i]] + off, len, cudaMemcpyDeviceToHost, d->stream));
            ION_CUDA(cudaEventRecord(d->ev, d->stream));
            ION_CUDA(cudaStreamWaitEvent(d->halo_stream, d->ev, 0));
            ION_CUDA(cudaMemcpyAsync((char*)d->buf[fields[i]] + off, (const char*)host_in[i] + off, len, cudaMemcpyHostToDevice, d->halo_stream));
        }
    }
    ION_CUDA(cudaEventRecord(d->ev_join, d->halo_stream));
    ION_CUDA(cudaStreamWaitEvent(d->stream, d->ev_join, 0));
    ION_CUDA(cudaStreamSynchronize(d->stream));
    return ION_OK;
}
int ion_buffer_copy(ion_domain_t* dst, int df, size_t doff, ion_domain_t* src, int sf, size_t soff, size_t bytes) {
    int r = check_field(dst, df);
    if (r) return r;
    r = check_field(src, sf);
    if (r) return r;
    if (doff > dst->bytes[df] || bytes > dst->bytes[df] - doff || soff > src->bytes[sf] || bytes > src->bytes[sf] - soff)
        return fail(ION_ERR_RANGE, "device copy out of range");
    ION_CUDA(cudaSetDevice(dst->device));
    if (src->stream != dst->stream) {  // order after everything already queued on the source domain
        Event g;
        ION_CUDA(cudaSetDevice(src->device));
        ION_CUDA(cudaEventCreateWithFlags(&g.e, cudaEventDisableTiming));
        ION_CUDA(cudaEventRecord(g.e, src->stream));
        ION_CUDA(cudaSetDevice(dst->device));
        ION_CUDA(cudaStreamWaitEvent(dst->stream, g.e, 0));
    }
    if (src->device == dst->device)
        ION_CUDA(cudaMemcpyAsync((char*)dst->buf[df] + doff, (const char*)src->buf[sf] + soff, bytes, cudaMemcpyDeviceToDevice, dst->stream));
    else
        ION_CUDA(cudaMemcpyPeerAsync((char*)dst->buf[df] + doff, dst->device, (const char*)src->buf[sf] + soff, src->device, bytes, dst->stream));
    return ION_OK;
}

// ---- slice read-back: the data half of the reference's slice view (graphics_kernels.cl:669-706 enumerates the cells of a
// slice as a -> (slice_x, a%NY, a/NY) | (a/NZ, slice_y, a%NZ) | (a%NX, a/NX, slice_z); component 3 = length(), :689) ----
__global__ void k_gather_slice(const void* __restrict__ buf, int is_u8, int planes, uint64_t N, uint32_t nx, uint32_t ny, uint32_t nz,
                               int component, uint32_t direction, uint32_t index, float* __restrict__ out, uint32_t area) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= area) return;
    uint32_t x, y, z;
    if (direction == 0u) { x = index; y = a % ny; z = a / ny; }
    else if (direction == 1u) { x = a / nz; y = index; z = a % nz; }
    else { x = a % nx; y = a / nx; z = index; }
    const uint64_t n = (uint64_t)x + ((uint64_t)y + (uint64_t)z * ny) * nx;
    float v;
    if (is_u8) v = (float)reinterpret_cast<const uint8_t*>(buf)[n];
    else {
        const float* f = reinterpret_cast<const float*>(buf);
        if (planes == 1) v = f[n];
        else if (component < 3) v = f[(uint64_t)component * N + n];
        else v = sqrtf(sq(f[n]) + sq(f[N + n]) + sq(f[2ull * N + n]));
    }
    out[a] = v;
}

int ion_read_slice(ion_domain_t* d, int field, int component, uint32_t direction, uint32_t index, float* host_out) {
    int r = check_field(d, field);
    if (r) return r;
    if (!host_out) return fail(ION_ERR_INVALID, "NULL host pointer");
    int planes = 0, is_u8 = 0;
    switch (field) {
        case ION_FIELD_RHO: case ION_FIELD_Q: case ION_FIELD_ET: planes = 1; break;
        case ION_FIELD_U: case ION_FIELD_F: case ION_FIELD_E_STAT: case ION_FIELD_B_STAT: case ION_FIELD_E_DYN: case ION_FIELD_B_DYN:
        case ION_FIELD_E_VAR: planes = 3; break;
        case ION_FIELD_FLAGS: planes = 1; is_u8 = 1; break;
        default: return fail(ION_ERR_UNSUPPORTED, "field %d is not a per-cell scalar / vector / flag field", field);
    }
    const IonParams& p = d->params;
    if (direction > 2u) return fail(ION_ERR_INVALID, "slice direction %u (0,1,2 = x,y,z)", direction);
    if (planes == 3 && (component < 0 || component > 3)) return fail(ION_ERR_INVALID, "vector component %d (0,1,2 or 3 = magnitude)", component);
    const uint32_t extent = direction == 0u ? p.nx : direction == 1u ? p.ny : p.nz;
    if (index >= extent) return fail(ION_ERR_RANGE, "slice index %u outside the domain (%u cells)", index, extent);
    const uint32_t area = direction == 0u ? p.ny * p.nz : direction == 1u ? p.nx * p.nz : p.nx * p.ny;
    ION_CUDA(cudaSetDevice(d->device));
    float* tmp = nullptr;
    ION_CUDA(cudaMalloc((void**)&tmp, (size_t)area * sizeof(float)));
    k_gather_slice<<<(area + 255u) / 256u, 256, 0, d->stream>>>(d->buf[field], is_u8, planes, d->k.N, p.nx, p.ny, p.nz, component, direction, index, tmp, area);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, tmp, (size_t)area * sizeof(float), cudaMemcpyDeviceToHost, d->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return cuda_fail(e, "read_slice");
    return ION_OK;
}

#define ION_VS_DISPATCH(fn, ...)                                                 \
    switch (d->params.velocity_set) {                                            \
        case ION_D2Q9: e = fn<ION_D2Q9>(__VA_ARGS__); break;                     \
        case ION_D3Q15: e = fn<ION_D3Q15>(__VA_ARGS__); break;                   \
        case ION_D3Q19: e = fn<ION_D3Q19>(__VA_ARGS__); break;                   \
        default: e = fn<ION_D3Q27>(__VA_ARGS__); break;                          \
    }

int ion_enqueue_initialize(ion_domain_t* d) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    ION_CUDA(cudaSetDevice(d->device));
    cudaError_t e;
    const bool mhd = d->params.ext & ION_EXT_MAGNETO_HYDRO;
    ION_VS_DISPATCH(launch_initialize_vs, d->k, (int)d->params.float_type, mhd, d->stream)
    g_launches++;
    if (e != cudaSuccess) return cuda_fail(e, "initialize launch");
    ION_CUDA(cudaStreamSynchronize(d->stream));  // enqueue_initialize ends with queue.finish(), domain.rs:415
    return ION_OK;
}

// z layers [z_begin, z_end) of stream_collide; `finish` = this is the last range of the step: fold the LOD deposits (or sum them in
// order in the deterministic mode), which needs every layer done.  Esoteric-Pull makes the order of the ranges irrelevant: a thread
// reads and writes only its own set of (cell, slot) addresses (sim_kernels.cl:234-247).
static int stream_collide_range(ion_domain_t* d, uint64_t t, float fx, float fy, float fz, uint32_t z_begin, uint32_t z_end, bool finish) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    if (z_begin > z_end || z_end > d->params.nz) return fail(ION_ERR_RANGE, "z range [%u, %u) outside the domain (%u layers)", z_begin, z_end, d->params.nz);
    ION_CUDA(cudaSetDevice(d->device));
    cudaError_t e = cudaSuccess;
    const bool mhd = d->params.ext & ION_EXT_MAGNETO_HYDRO;
    const bool trt = d->params.relaxation_time == ION_TRT;
    const bool ecr = d->params.ext & ION_EXT_SUBGRID_ECR;
    if (z_end > z_begin) {
        KArgs k = d->k;
        k.z_off = z_begin;
        k.z_cnt = (z_begin == 0u && z_end == d->params.nz) ? 0u : z_end - z_begin;
        ION_VS_DISPATCH(launch_stream_collide_vs, k, (int)d->params.float_type, mhd, trt, ecr, t, fx, fy, fz, d->stream)
        g_launches++;
        if (e != cudaSuccess) return cuda_fail(e, "stream_collide launch");
    }
    if (!finish) return ION_OK;
    if (mhd && d->deterministic && d->params.lod_depth > 0u) {  // ordered LOD sums instead of the in-kernel float atomics
        e = launch_lod_deposit_ordered(d->k, d->stream);
        g_launches++;
        if (e != cudaSuccess) return cuda_fail(e, "lod_deposit_ordered launch");
    } else if (mhd && d->lod_rep) {  // replicas of the finest level -> QU_lod (the reference's atomics land there directly)
        e = launch_lod_fold(d->k, d->stream);
        g_launches++;
        if (e != cudaSuccess) return cuda_fail(e, "lod_fold launch");
    }
    return ION_OK;
}
int ion_enqueue_stream_collide(ion_domain_t* d, uint64_t t, float fx, float fy, float fz) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    return stream_collide_range(d, t, fx, fy, fz, 0u, d->params.nz, true);
}
int ion_enqueue_stream_collide_range(ion_domain_t* d, uint64_t t, float fx, float fy, float fz, uint32_t z_begin, uint32_t z_end, int finish) {
    return stream_collide_range(d, t, fx, fy, fz, z_begin, z_end, finish != 0);
}

int ion_enqueue_update_fields(ion_domain_t* d, uint64_t t, float fx, float fy, float fz) {
    (void)fx; (void)fy; (void)fz;  // bound but unused by the reference kernel too (sim_kernels.cl:834-859)
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    ION_CUDA(cudaSetDevice(d->device));
    cudaError_t e;
    ION_VS_DISPATCH(launch_update_fields_vs, d->k, (int)d->params.float_type, t, d->stream)
    g_launches++;
    if (e != cudaSuccess) return cuda_fail(e, "update_fields launch");
    return ION_OK;
}

static int need_mhd(const ion_domain_t* d) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    if (!(d->params.ext & ION_EXT_MAGNETO_HYDRO)) return fail(ION_ERR_ABSENT, "kernel needs ext_magneto_hydro (Option::None in domain.rs:30-35)");
    return ION_OK;
}

int ion_enqueue_update_e_b_dyn(ion_domain_t* d) {
    int r = need_mhd(d);
    if (r) return r;
    ION_CUDA(cudaSetDevice(d->device));
    uint64_t l = 0;
    cudaError_t e;
    // Default path: the own pyramid as a polyphase FFT convolution (eb_fft.cu) when the geometry allows it.  Memory: 24 B per cell
    // of scratch plus the static kernel spectra (102 B per cell) when both fit in three quarters of the free device memory, else the
    // spectra are recomputed per batch of tasks every step (ION_EB_FFT=0 keeps the direct kernels, ION_EB_FFT_MAX_GB overrides the
    // budget).  The deterministic mode always uses the reference-ordered direct kernel.
    if (!d->deterministic && !d->eb_plan_tried) {
        d->eb_plan_tried = true;
        const char* sw = getenv("ION_EB_FFT");
        if (!sw || atoi(sw) != 0) {
            size_t free_b = 0, total_b = 0;
            ION_CUDA(cudaMemGetInfo(&free_b, &total_b));
            size_t budget = free_b / 4 * 3;  // what does not fit is streamed (recomputed per batch of tasks) or left to the direct kernels
            if (const char* gb = getenv("ION_EB_FFT_MAX_GB")) budget = (size_t)(atof(gb) * 1073741824.0);
            e = eb_fft_create(d->k, budget, d->stream, &d->eb_plan, &l);
            if (e != cudaSuccess) return cuda_fail(e, "update_e_b_dynamic: building the kernel spectra");
        }
    }
    if (d->eb_plan && !d->deterministic) {
        e = eb_fft_launch(d->eb_plan, d->k, d->stream, &l);
        // what the FFT pass did not sum: every foreign pyramid when the geometry kept the neighbour off the FFT path, else the far
        // slabs -- unless those went through the Taylor tensors of k_eb_combine
        if (e == cudaSuccess && !eb_fft_plan_far_handled(d->eb_plan))
            e = launch_update_e_b_foreign(d->k, d->lod_sources, d->stream, &l, eb_fft_plan_foreign_domain(d->eb_plan));
    } else {
        e = launch_update_e_b(d->k, d->lod_sources, d->stream, &l, d->deterministic);
    }
    g_launches += l;
    if (e != cudaSuccess) return cuda_fail(e, "update_e_b_dynamic launch");
    return ION_OK;
}
int ion_domain_eb_fft_info(const ion_domain_t* d, uint64_t* spectrum_bytes, uint32_t* tasks) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    if (spectrum_bytes) *spectrum_bytes = eb_fft_plan_bytes(d->eb_plan);
    if (tasks) *tasks = eb_fft_plan_tasks(d->eb_plan);
    return ION_OK;
}
int ion_enqueue_lod_part_2_gather(ion_domain_t* d) {
    int r = need_mhd(d);
    if (r) return r;
    ION_CUDA(cudaSetDevice(d->device));
    uint64_t l = 0;
    cudaError_t e = launch_lod_gather(d->k, d->stream, &l);
    g_launches += l;
    if (e != cudaSuccess) return cuda_fail(e, "lod_part_2_gather launch");
    return ION_OK;
}
int ion_enqueue_clear_qu_lod(ion_domain_t* d) {
    int r = need_mhd(d);
    if (r) return r;
    ION_CUDA(cudaSetDevice(d->device));
    cudaError_t e = launch_clear_qu_lod(d->k, d->stream);
    g_launches++;
    if (e != cudaSuccess) return cuda_fail(e, "clear_qu_lod launch");
    return ION_OK;
}

static int transfer(ion_domain_t* d, int field, int insert, uint32_t direction, uint64_t t) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    if (field < ION_TRANSFER_FI || field > ION_TRANSFER_QI) return fail(ION_ERR_INVALID, "unknown transfer field %d", field);
    const uint32_t dims = d->params.velocity_set == ION_D2Q9 ? 2u : 3u;
    if (direction >= dims) return fail(ION_ERR_INVALID, "direction %u out of range", direction);
    const uint32_t dd = direction == 0 ? d->params.dx : direction == 1 ? d->params.dy : d->params.dz;
    if (dd <= 1 || !d->buf[ION_FIELD_TRANSFER_P]) return fail(ION_ERR_ABSENT, "axis %u is not split: no transfer buffers (domain.rs:311-318)", direction);
    if ((field == ION_TRANSFER_EI || field == ION_TRANSFER_QI) && !(d->params.ext & ION_EXT_MAGNETO_HYDRO))
        return fail(ION_ERR_ABSENT, "transfer kernel needs ext_magneto_hydro (domain.rs:340-367)");
    ION_CUDA(cudaSetDevice(d->device));
    cudaError_t e = launch_transfer(d->k, (int)d->params.velocity_set, (int)d->params.float_type, field, insert, direction, t, xfer_stream(d));
    g_launches++;
    if (e != cudaSuccess) return cuda_fail(e, "transfer launch");
    return ION_OK;
}
int ion_enqueue_transfer_extract(ion_domain_t* d, int field, uint32_t direction, uint64_t t) { return transfer(d, field, 0, direction, t); }
int ion_enqueue_transfer_insert(ion_domain_t* d, int field, uint32_t direction, uint64_t t) { return transfer(d, field, 1, direction, t); }

int ion_voxelize_mesh(ion_domain_t* d, const float* p0, const float* p1, const float* p2, uint32_t triangles, const float bbu[7],
                      uint32_t direction, uint8_t flag, float mx, float my, float mz, uint64_t t) {
    (void)t;  // bound as kernel arg "t" (mesh.rs:317) but never read by voxelize_mesh
    if (!d || !p0 || !p1 || !p2 || !bbu) return fail(ION_ERR_INVALID, "NULL argument");
    if (direction > 2) return fail(ION_ERR_INVALID, "direction %u out of range", direction);
    uint32_t tn;
    memcpy(&tn, &bbu[0], 4);
    if (tn != triangles) return fail(ION_ERR_INVALID, "bbu[0] (bit-cast triangle count %u) disagrees with triangles=%u", tn, triangles);
    ION_CUDA(cudaSetDevice(d->device));
    float* dp = nullptr;  // p0|p1|p2 re-allocated per mesh like mesh.rs:282-287
    const size_t bytes = (size_t)triangles * 3 * sizeof(float);
    ION_CUDA(cudaMallocAsync((void**)&dp, 3 * bytes + 16, d->stream));
    struct AsyncFree {  // the stream-ordered free also happens on the early returns below
        float* p;
        cudaStream_t s;
        ~AsyncFree() { if (p) cudaFreeAsync(p, s); }
    } guard{dp, d->stream};
    ION_CUDA(cudaMemcpyAsync(dp, p0, bytes, cudaMemcpyHostToDevice, d->stream));
    ION_CUDA(cudaMemcpyAsync(dp + 3 * (size_t)triangles, p1, bytes, cudaMemcpyHostToDevice, d->stream));
    ION_CUDA(cudaMemcpyAsync(dp + 6 * (size_t)triangles, p2, bytes, cudaMemcpyHostToDevice, d->stream));
    const int mhd = (d->params.ext & ION_EXT_MAGNETO_HYDRO) ? 1 : 0;
    cudaError_t e = launch_voxelize(d->k, direction, flag, dp, dp + 3 * (size_t)triangles, dp + 6 * (size_t)triangles, triangles, bbu + 1,
                                    mx, my, mz, mhd, d->stream);
    g_launches++;
    if (e != cudaSuccess) return cuda_fail(e, "voxelize_mesh launch");
    ION_CUDA(cudaStreamSynchronize(d->stream));  // host triangle arrays are borrowed only for the duration of the call
    return ION_OK;
}

static int precompute(ion_domain_t* d, int which) {
    int r = need_mhd(d);
    if (r) return r;
    ION_CUDA(cudaSetDevice(d->device));
    const uint8_t mask = which == 0 ? ION_TYPE_M : (ION_TYPE_F | ION_TYPE_C);
    uint32_t total = 0;
    cudaError_t e = count_sources(d->k, mask, d->cp_counts, &total, d->stream);
    g_launches += 2;
    if (e != cudaSuccess) return cuda_fail(e, "source count");
    void* table = nullptr;
    ION_CUDA(cudaMallocAsync(&table, field_source_bytes(total), d->stream));
    if (d->precompute_mode == 2 && total > 0u) {  // FFT convolution (cuFFT transforms); an empty source set takes the direct kernels
        float* E = which == 0 ? nullptr : which == 1 ? d->k.E_stat : (float*)d->buf[ION_FIELD_E_VAR];
        if (which != 0 && !E) { cudaFreeAsync(table, d->stream); return fail(ION_ERR_ABSENT, "E_var needs ext_subgrid_ecr (domain.rs:573)"); }
        const char* why = nullptr;
        uint64_t l = 0;
        e = launch_precompute_fft(d->k, which, E, d->cp_counts, table, total, d->stream, &l, &why);
        g_launches += l;
        cudaFreeAsync(table, d->stream);
        if (e == cudaErrorNotSupported && why) return fail(ION_ERR_ABSENT, "FFT precompute mode: %s", why);
        if (e != cudaSuccess) return cuda_fail(e, "static field precompute (FFT mode)");
        return ION_OK;
    }
    if (which == 0) {
        e = launch_precompute_b(d->k, d->cp_counts, table, d->stream, d->precompute_mode == 1 ? 1 : 0);
        g_launches += 5;
    } else {
        float* E = which == 1 ? d->k.E_stat : (float*)d->buf[ION_FIELD_E_VAR];
        if (!E) { cudaFreeAsync(table, d->stream); return fail(ION_ERR_ABSENT, "E_var needs ext_subgrid_ecr (domain.rs:573)"); }
        e = launch_precompute_e(d->k, E, d->cp_counts, table, d->stream);
        g_launches += 4;
    }
    cudaFreeAsync(table, d->stream);
    if (e != cudaSuccess) return cuda_fail(e, "static field precompute launch");
    return ION_OK;
}
int ion_enqueue_precompute_b(ion_domain_t* d) { return precompute(d, 0); }
int ion_enqueue_precompute_e(ion_domain_t* d) { return precompute(d, 1); }
int ion_enqueue_precompute_e_ecr(ion_domain_t* d) { return precompute(d, 2); }

int ion_domain_set_precompute_mode(ion_domain_t* d, int mode) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    if (mode < 0 || mode > 2) return fail(ION_ERR_INVALID, "precompute mode %d (0 = reference arithmetic and order, 1 = fast direct sum, 2 = FFT convolution)", mode);
    d->precompute_mode = mode;
    return ION_OK;
}

int ion_domain_set_ecr_freq(ion_domain_t* d, float ecrf) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    d->k.ecrf = ecrf;
    return ION_OK;
}

int ion_halo_fork(ion_domain_t* d) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    if (d->halo_active) return ION_OK;
    ION_CUDA(cudaSetDevice(d->device));
    ION_CUDA(cudaEventRecord(d->ev_fork, d->stream));
    ION_CUDA(cudaStreamWaitEvent(d->halo_stream, d->ev_fork, 0));
    d->halo_active = true;
    return ION_OK;
}
int ion_halo_join(ion_domain_t* d) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    if (!d->halo_active) return ION_OK;
    ION_CUDA(cudaSetDevice(d->device));
    ION_CUDA(cudaEventRecord(d->ev_join, d->halo_stream));
    ION_CUDA(cudaStreamWaitEvent(d->stream, d->ev_join, 0));
    d->halo_active = false;
    return ION_OK;
}

int ion_finish(ion_domain_t* d) {
    if (!d) return fail(ION_ERR_INVALID, "NULL domain");
    ION_CUDA(cudaSetDevice(d->device));
    if (d->halo_active) {
        int r = ion_halo_join(d);
        if (r) return r;
    }
    ION_CUDA(cudaStreamSynchronize(d->stream));
    return ION_OK;
}

}  // extern "C"
