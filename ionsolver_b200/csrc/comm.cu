// comm.cu -- device-resident halo and LOD exchange (replaces the host-staged swap of
// /root/reference/src/lbm/mod.rs:371-468 and the buffer reads/writes of src/lbm/domain.rs:504-549).
//
// Two transports, both keep the packed face layout [b*A+a] of sim_kernels.cl:1069 untouched:
//   * one process, several domains  -> pointer swap (same GPU) or cudaMemcpyPeerAsync over NVLink (different GPUs);
//   * one process per GPU (torchrun) -> NCCL send/recv between ring neighbours + all-gather of the LOD pyramids.
// NCCL is bound at run time with dlopen so that the library loads on a single GPU without it; inside a torch
// process the already loaded torch-bundled libnccl.so.2 is the one that resolves.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <new>

#include "domain_internal.cuh"

using namespace ion;

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return ION_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(ION_ERR_UNSUPPORTED, "libnccl.so.2 not found: %s", dlerror());
#define ION_SYM(name)                                                                       \
    *(void**)(&g_nccl.name) = dlsym(h, "nccl" #name);                                       \
    if (!g_nccl.name) return fail(ION_ERR_UNSUPPORTED, "nccl" #name " missing in libnccl")
    ION_SYM(GetUniqueId);
    ION_SYM(CommInitRank);
    ION_SYM(CommDestroy);
    ION_SYM(GroupStart);
    ION_SYM(GroupEnd);
    ION_SYM(Send);
    ION_SYM(Recv);
    ION_SYM(AllGather);
    ION_SYM(GetErrorString);
#undef ION_SYM
    g_nccl.handle = h;
    return ION_OK;
}
int nccl_fail(ncclResult_t r, const char* what) {
    return fail(20000 + (int)r, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
}
#define ION_NCCL(call)                                           \
    do {                                                         \
        ncclResult_t ion_r_ = (call);                            \
        if (ion_r_ != ncclSuccess) return nccl_fail(ion_r_, #call); \
    } while (0)

// make `waiter` wait for everything queued on `signaller` so far
int order_after(ion_domain* waiter, ion_domain* signaller) {  // on the streams the face exchange currently uses
    if (xfer_stream(waiter) == xfer_stream(signaller)) return ION_OK;
    ION_CUDA(cudaSetDevice(signaller->device));
    ION_CUDA(cudaEventRecord(signaller->ev, xfer_stream(signaller)));
    ION_CUDA(cudaSetDevice(waiter->device));
    ION_CUDA(cudaStreamWaitEvent(xfer_stream(waiter), signaller->ev, 0));
    return ION_OK;
}
}  // namespace

struct ion_comm {
    ncclComm_t comm;
    int rank, world, device;
};

extern "C" {

int ion_exchange_transfer(ion_domain_t* d, ion_domain_t* dp, size_t bytes) {
    if (!d || !dp) return fail(ION_ERR_INVALID, "NULL domain");
    if (!d->buf[ION_FIELD_TRANSFER_P] || !dp->buf[ION_FIELD_TRANSFER_M]) return fail(ION_ERR_ABSENT, "no transfer buffers (axis not split)");
    if (bytes > d->bytes[ION_FIELD_TRANSFER_P] || bytes > dp->bytes[ION_FIELD_TRANSFER_M]) return fail(ION_ERR_RANGE, "face payload %zu exceeds the transfer buffers", bytes);
    int r;
    if (d->device == dp->device) {
        // the reference's std::ptr::swap, on device pointers: after the swap each domain's insert kernel reads what
        // the partner's extract kernel wrote, so each stream has to wait for the other one's queued work
        if ((r = order_after(d, dp))) return r;
        if ((r = order_after(dp, d))) return r;
        void* t = d->buf[ION_FIELD_TRANSFER_P];
        d->buf[ION_FIELD_TRANSFER_P] = dp->buf[ION_FIELD_TRANSFER_M];
        dp->buf[ION_FIELD_TRANSFER_M] = t;
        set_transfer_ptrs(d);
        set_transfer_ptrs(dp);
        return ION_OK;
    }
    // different GPUs: d.P -> dp.spare_M on dp's stream, dp.M -> d.spare_P on d's stream (each after the partner's
    // extract), then the spares become current.  The old current buffers turn into spares and are only overwritten
    // by the NEXT exchange, which is queued behind the partner's copy through the same event ordering.
    if ((r = order_after(dp, d))) return r;
    if ((r = order_after(d, dp))) return r;
    ION_CUDA(cudaSetDevice(dp->device));
    ION_CUDA(cudaMemcpyPeerAsync(dp->alt_m, dp->device, d->buf[ION_FIELD_TRANSFER_P], d->device, bytes, xfer_stream(dp)));
    ION_CUDA(cudaSetDevice(d->device));
    ION_CUDA(cudaMemcpyPeerAsync(d->alt_p, d->device, dp->buf[ION_FIELD_TRANSFER_M], dp->device, bytes, xfer_stream(d)));
    // nobody may overwrite a source before the partner's copy has read it
    if ((r = order_after(dp, d))) return r;
    if ((r = order_after(d, dp))) return r;
    void* t = d->buf[ION_FIELD_TRANSFER_P];
    d->buf[ION_FIELD_TRANSFER_P] = d->alt_p;
    d->alt_p = t;
    t = dp->buf[ION_FIELD_TRANSFER_M];
    dp->buf[ION_FIELD_TRANSFER_M] = dp->alt_m;
    dp->alt_m = t;
    set_transfer_ptrs(d);
    set_transfer_ptrs(dp);
    return ION_OK;
}

int ion_copy_lods(ion_domain_t* dst, uint32_t dst_entry, ion_domain_t* src, uint32_t src_entry, uint32_t entries) {
    if (!dst || !src) return fail(ION_ERR_INVALID, "NULL domain");
    if (!dst->buf[ION_FIELD_QU_LOD] || !src->buf[ION_FIELD_QU_LOD]) return fail(ION_ERR_ABSENT, "QU_lod needs ext_magneto_hydro");
    if ((uint64_t)dst_entry + entries > dst->params.n_lod || (uint64_t)src_entry + entries > src->params.n_lod_own)
        return fail(ION_ERR_RANGE, "LOD range out of bounds");
    return ion_buffer_copy(dst, ION_FIELD_QU_LOD, (size_t)dst_entry * 16, src, ION_FIELD_QU_LOD, (size_t)src_entry * 16, (size_t)entries * 16);
}

int ion_neighbor_domains(uint32_t d_x, uint32_t d_y, uint32_t d_z, uint32_t d, uint32_t axis, uint32_t* dp, uint32_t* dm) {
    if (!dp || !dm) return fail(ION_ERR_INVALID, "NULL argument");
    if (!d_x || !d_y || !d_z || d >= d_x * d_y * d_z || axis > 2) return fail(ION_ERR_INVALID, "domain %u / axis %u out of range", d, axis);
    const uint32_t x = (d % (d_x * d_y)) % d_x, y = (d % (d_x * d_y)) / d_x, z = d / (d_x * d_y);  // mod.rs:189-191
    if (axis == 0) { *dp = ((x + 1) % d_x) + (y + z * d_y) * d_x; *dm = ((x + d_x - 1) % d_x) + (y + z * d_y) * d_x; }
    else if (axis == 1) { *dp = x + (((y + 1) % d_y) + z * d_y) * d_x; *dm = x + (((y + d_y - 1) % d_y) + z * d_y) * d_x; }
    else { *dp = x + (y + ((z + 1) % d_z) * d_y) * d_x; *dm = x + (y + ((z + d_z - 1) % d_z) * d_y) * d_x; }
    return ION_OK;
}

int ion_lod_exchange_plan(const IonParams* p, uint32_t dc, uint32_t* src_entry, uint32_t* entries, uint32_t* dst_entry) {
    if (!p || !src_entry || !entries || !dst_entry) return fail(ION_ERR_INVALID, "NULL argument");
    const uint32_t dxy = p->dx * p->dy, dn = dxy * p->dz;
    if (!dn || dc >= dn || p->di >= dn) return fail(ION_ERR_INVALID, "domain %u out of range", dc);
    const int x = (int)((p->di % dxy) % p->dx), y = (int)((p->di % dxy) / p->dx), z = (int)(p->di / dxy);
    const uint32_t dim = p->velocity_set == ION_D2Q9 ? 2u : 3u;
    auto level_start = [&](int depth) {  // get_offset(depth - 1), mod.rs:441-445
        uint32_t cnt = 0;
        for (int i = 0; i < depth; i++) { uint32_t s = 1; for (uint32_t k = 0; k < dim; k++) s *= 1u << i; cnt += s; }
        return cnt;
    };
    uint32_t offset = p->n_lod_own;
    *src_entry = *entries = 0;
    *dst_entry = offset;
    for (uint32_t c = 0; c <= dc; c++) {
        if (c == p->di) continue;
        const int fx = (int)((c % dxy) % p->dx), fy = (int)((c % dxy) / p->dx), fz = (int)(c / dxy);
        int dist = abs(z - fz);
        if (abs(y - fy) > dist) dist = abs(y - fy);
        if (abs(x - fx) > dist) dist = abs(x - fx);
        const int depth = (int)p->lod_depth - dist > 0 ? (int)p->lod_depth - dist : 0;
        const uint32_t rs = level_start(depth), re = level_start(depth + 1);
        if (c == dc) { *src_entry = rs; *entries = re - rs; *dst_entry = offset; }
        offset += re - rs;
    }
    return ION_OK;
}

int ion_comm_unique_id(uint8_t id[ION_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == ION_COMM_ID_BYTES, "ncclUniqueId size");
    if (!id) return fail(ION_ERR_INVALID, "NULL id");
    int r = load_nccl();
    if (r) return r;
    ncclUniqueId u;
    ION_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return ION_OK;
}

int ion_comm_create(const uint8_t id[ION_COMM_ID_BYTES], int rank, int world, int device, ion_comm_t** out) {
    if (!id || !out) return fail(ION_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(ION_ERR_INVALID, "rank %d of %d", rank, world);
    int r = load_nccl();
    if (r) return r;
    ION_CUDA(cudaSetDevice(device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ion_comm* c = new (std::nothrow) ion_comm();
    if (!c) return fail(ION_ERR_INVALID, "out of host memory");
    c->rank = rank;
    c->world = world;
    c->device = device;
    ncclResult_t nr = g_nccl.CommInitRank(&c->comm, world, u, rank);
    if (nr != ncclSuccess) { delete c; return nccl_fail(nr, "ncclCommInitRank"); }
    *out = c;
    return ION_OK;
}

int ion_comm_destroy(ion_comm_t* c) {
    if (!c) return ION_OK;
    if (g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    return ION_OK;
}

int ion_comm_exchange_transfer(ion_comm_t* c, ion_domain_t* d, int rank_p, int rank_m, size_t bytes) {
    if (!c || !d) return fail(ION_ERR_INVALID, "NULL argument");
    if (!d->buf[ION_FIELD_TRANSFER_P]) return fail(ION_ERR_ABSENT, "no transfer buffers (axis not split)");
    if (bytes > d->bytes[ION_FIELD_TRANSFER_P]) return fail(ION_ERR_RANGE, "face payload %zu exceeds the transfer buffers", bytes);
    if (rank_p < 0 || rank_p >= c->world || rank_m < 0 || rank_m >= c->world) return fail(ION_ERR_INVALID, "neighbour rank out of range");
    ION_CUDA(cudaSetDevice(d->device));
    // my +face goes to rank_p (their transfer_m), my -face to rank_m (their transfer_p); the mirror images arrive in
    // the spare buffers, which then become current (the device-side equivalent of mod.rs:383)
    ION_NCCL(g_nccl.GroupStart());
    ION_NCCL(g_nccl.Send(d->buf[ION_FIELD_TRANSFER_P], bytes, ncclUint8, rank_p, c->comm, xfer_stream(d)));
    ION_NCCL(g_nccl.Send(d->buf[ION_FIELD_TRANSFER_M], bytes, ncclUint8, rank_m, c->comm, xfer_stream(d)));
    // receive order m, p: with two ranks rank_p == rank_m, and NCCL pairs the k-th send with the k-th receive of a
    // peer -- the partner's first send is ITS +face, which is my new transfer_m
    ION_NCCL(g_nccl.Recv(d->alt_m, bytes, ncclUint8, rank_m, c->comm, xfer_stream(d)));
    ION_NCCL(g_nccl.Recv(d->alt_p, bytes, ncclUint8, rank_p, c->comm, xfer_stream(d)));
    ION_NCCL(g_nccl.GroupEnd());
    void* t = d->buf[ION_FIELD_TRANSFER_P];
    d->buf[ION_FIELD_TRANSFER_P] = d->alt_p;
    d->alt_p = t;
    t = d->buf[ION_FIELD_TRANSFER_M];
    d->buf[ION_FIELD_TRANSFER_M] = d->alt_m;
    d->alt_m = t;
    set_transfer_ptrs(d);
    return ION_OK;
}

int ion_comm_exchange_lods(ion_comm_t* c, ion_domain_t* d) {
    if (!c || !d) return fail(ION_ERR_INVALID, "NULL argument");
    if (!d->buf[ION_FIELD_QU_LOD]) return fail(ION_ERR_ABSENT, "QU_lod needs ext_magneto_hydro");
    const IonParams& p = d->params;
    if ((int)(p.dx * p.dy * p.dz) != c->world) return fail(ION_ERR_INVALID, "%u domains but %d ranks", p.dx * p.dy * p.dz, c->world);
    ION_CUDA(cudaSetDevice(d->device));
    const size_t own = (size_t)p.n_lod_own * 4;  // floats
    if (!d->lod_gather) ION_CUDA(cudaMalloc((void**)&d->lod_gather, own * sizeof(float) * c->world));
    ION_NCCL(g_nccl.AllGather(d->buf[ION_FIELD_QU_LOD], d->lod_gather, own, ncclFloat, c->comm, d->stream));
    // level selection, mod.rs:448-465 (ion_lod_exchange_plan): foreign domain dc contributes level max(0, depth - dist),
    // appended in ascending dc after the own pyramid
    for (uint32_t dc = 0; dc < p.dx * p.dy * p.dz; dc++) {
        uint32_t src = 0, cnt = 0, dst = 0;
        int r = ion_lod_exchange_plan(&p, dc, &src, &cnt, &dst);
        if (r) return r;
        if (!cnt) continue;
        ION_CUDA(cudaMemcpyAsync((float*)d->buf[ION_FIELD_QU_LOD] + (size_t)dst * 4, d->lod_gather + (size_t)dc * own + (size_t)src * 4,
                                 (size_t)cnt * 16, cudaMemcpyDeviceToDevice, d->stream));
    }
    return ION_OK;
}

}  // extern "C"
