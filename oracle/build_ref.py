"""oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY.

Recipe that compiles the reference's OWN kernel source for the host CPU.  The source is read where it
lies ($IONSOLVER_REF or /root/reference, file src/kernels/sim_kernels.cl) and is never copied into the
repository; only the built shared objects (and the generated translation unit next to them) land in
oracle/_ref/, which is git-ignored but travels to the GPU box.

Pipeline:
  1. text after "EndTempDefines%" (same split as /root/reference/src/opencl.rs:67-75),
  2. one regex: OpenCL vector-literal casts "(float3)(" -> "float3(" (also uint3/int3),
  3. prepend the #define block of get_device_defines (restated in ref_host.device_defines),
  4. wrap in `namespace ref {}` with oracle/ocl_shim.h, append the C driver below (one extern "C"
     entry point per kernel; get_global_id is an OpenMP loop index),
  5. g++ -std=c++17 -O3 -fopenmp -ffp-contract=off  (explicit fma() calls stay fused, nothing else is).

One library per (config, domain) because every parameter is a compile-time #define in the reference.
"""
from __future__ import annotations

import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
CXXFLAGS = ["-std=c++17", "-O3", "-mavx2", "-mfma", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-w"]

DRIVER = r'''
// ---------------------------------------------------------------------------------------------
// driver: one entry point per __kernel; argument order follows the kernel signatures in sim.cl and
// the binding order in /root/reference/src/lbm/domain.rs:217-305
// ---------------------------------------------------------------------------------------------
} // namespace ref
#undef printf
#include <omp.h>
struct RefBuffers {
    void* fi; float* rho; float* u; uchar* flags; float* F; float* E_stat; float* B_stat; float* E_dyn; float* B_dyn;
    void* fqi; void* ei; float* Q; float* QU_lod; float* E_var; void* eti; float* Et;
    uchar* transfer_p; uchar* transfer_m; float* p0; float* p1; float* p2; float* bbu;
};
#define ION_RUN(begin, end, call)                                         \
    _Pragma("omp parallel for schedule(static)")                          \
    for (long long ion_i = (long long)(begin); ion_i < (long long)(end); ion_i++) { ion_shim_gid = (size_t)ion_i; call; }
#ifdef FORCE_FIELD
#define ION_FF_ARG(b) , (b)->F
#else
#define ION_FF_ARG(b)
#endif
#ifdef MAGNETO_HYDRO
#define ION_SC_MHD(b) , (b)->E_dyn, (b)->B_dyn, (fpxx*)(b)->fqi, (fpxx*)(b)->ei, (b)->Q, (b)->QU_lod
#define ION_INIT_MHD(b) , (b)->E_stat, (b)->B_stat, (b)->E_dyn, (b)->B_dyn, (fpxx*)(b)->fqi, (fpxx*)(b)->ei, (b)->Q
#else
#define ION_SC_MHD(b)
#define ION_INIT_MHD(b)
#endif
#ifdef SUBGRID_ECR
#define ION_SC_ECR(b, ecrf) , (b)->E_var, (fpxx*)(b)->eti, (b)->Et, ecrf
#define ION_INIT_ECR(b) , (fpxx*)(b)->eti, (b)->Et
#else
#define ION_SC_ECR(b, ecrf)
#define ION_INIT_ECR(b)
#endif
extern "C" {
void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_has(const char* what) {
    (void)what;
#ifdef MAGNETO_HYDRO
    if (!strcmp(what, "MAGNETO_HYDRO")) return 1;
#endif
#ifdef SUBGRID_ECR
    if (!strcmp(what, "SUBGRID_ECR")) return 1;
#endif
    return 0;
}
void ref_stream_collide(const RefBuffers* b, uint64_t begin, uint64_t end, uint64_t t, float fx, float fy, float fz, float ecrf) {
    (void)ecrf;
    ION_RUN(begin, end, ref::stream_collide((fpxx*)b->fi, b->rho, b->u, b->flags, t, fx, fy, fz ION_FF_ARG(b) ION_SC_MHD(b) ION_SC_ECR(b, ecrf)))
}
void ref_initialize(const RefBuffers* b, uint64_t begin, uint64_t end) {
    ION_RUN(begin, end, ref::initialize((fpxx*)b->fi, b->rho, b->u, b->flags ION_INIT_MHD(b) ION_INIT_ECR(b)))
}
void ref_update_fields(const RefBuffers* b, uint64_t begin, uint64_t end, uint64_t t, float fx, float fy, float fz) {
    ION_RUN(begin, end, ref::update_fields((const fpxx*)b->fi, b->rho, b->u, b->flags, t, fx, fy, fz))
}
void ref_update_e_b_dynamic(const RefBuffers* b, uint64_t begin, uint64_t end) {
#ifdef MAGNETO_HYDRO
    ION_RUN(begin, end, ref::update_e_b_dynamic(b->E_stat, b->B_stat, b->E_dyn, b->B_dyn, b->Q, b->u, b->QU_lod, b->flags))
#endif
}
void ref_clear_qu_lod(const RefBuffers* b, uint64_t count) {
#ifdef MAGNETO_HYDRO
    ION_RUN(0, count, ref::clear_qu_lod(b->QU_lod))
#endif
}
void ref_lod_part_2_gather(const RefBuffers* b, uint32_t depth) {
#ifdef MAGNETO_HYDRO
    ION_RUN(0, (1ull << (depth * 3)), ref::lod_part_2_gather(b->QU_lod, depth))
#endif
}
// field: 0 fi, 1 rho_u_flags, 2 ei, 3 fqi (TransferField, types.rs:105-111); insert: 0 extract, 1 insert
void ref_transfer(const RefBuffers* b, int field, int insert, uint32_t direction, uint64_t t, uint64_t area) {
    if (field == 0 || field == 2) {
        fpxx_copy* f = (fpxx_copy*)(field == 0 ? b->fi : b->ei);
        if (!insert) { ION_RUN(0, area, ref::transfer_extract_fi(direction, t, b->transfer_p, b->transfer_m, f)) }
        else         { ION_RUN(0, area, ref::transfer__insert_fi(direction, t, b->transfer_p, b->transfer_m, f)) }
    } else if (field == 1) {
        if (!insert) { ION_RUN(0, area, ref::transfer_extract_rho_u_flags(direction, t, b->transfer_p, b->transfer_m, b->rho, b->u, b->flags)) }
        else         { ION_RUN(0, area, ref::transfer__insert_rho_u_flags(direction, t, b->transfer_p, b->transfer_m, b->rho, b->u, b->flags)) }
    } else {
#ifdef MAGNETO_HYDRO
        if (!insert) { ION_RUN(0, area, ref::transfer_extract_fqi(direction, t, b->transfer_p, b->transfer_m, (fpxx_copy*)b->fqi)) }
        else         { ION_RUN(0, area, ref::transfer__insert_fqi(direction, t, b->transfer_p, b->transfer_m, (fpxx_copy*)b->fqi)) }
#endif
    }
}
void ref_voxelize_mesh(const RefBuffers* b, uint32_t direction, uint64_t t, uint8_t flag, float mx, float my, float mz, uint64_t area) {
    (void)mx; (void)my; (void)mz;
    ION_RUN(0, area, ref::voxelize_mesh(direction, (fpxx*)b->fi, b->rho, b->u, b->flags, t, flag, b->p0, b->p1, b->p2, b->bbu
#ifdef MAGNETO_HYDRO
        , mx, my, mz, b->B_dyn
#endif
    ))
}
void ref_psi_from_mesh(const RefBuffers* b, uint64_t begin, uint64_t end) {
#ifdef MAGNETO_HYDRO
    ION_RUN(begin, end, ref::psi_from_mesh(b->flags, b->E_dyn, b->B_dyn))   // domain.rs:279-281: psi=E_dyn, M=B_dyn
#endif
}
void ref_static_b_from_mesh(const RefBuffers* b, uint64_t begin, uint64_t end) {
#ifdef MAGNETO_HYDRO
    ION_RUN(begin, end, ref::static_b_from_mesh(b->flags, b->B_stat, b->E_dyn)) // domain.rs:282-284
#endif
}
void ref_static_e_from_mesh(const RefBuffers* b, int target_e_var, uint64_t begin, uint64_t end) {
#ifdef MAGNETO_HYDRO
    float* E = target_e_var ? b->E_var : b->E_stat;                          // domain.rs:558-578
    ION_RUN(begin, end, ref::static_e_from_mesh(b->flags, E, b->B_dyn))
#endif
}
// DDF storage codec of this build (load/store macros, domain.rs:773-784). dir 0: float -> stored, 1: stored -> float
void ref_codec(const void* in, void* out, uint64_t count, int dir) {
    if (dir == 0) { const float* x = (const float*)in; fpxx* p = (fpxx*)out; for (uint64_t o = 0; o < count; o++) { using namespace ref; store(p, o, x[o]); } }
    else          { const fpxx* p = (const fpxx*)in; float* x = (float*)out; for (uint64_t o = 0; o < count; o++) { using namespace ref; x[o] = load(p, o); } }
}
void ref_neighbors(uint32_t n, uint32_t* j) { ref::neighbors(n, j); }
} // extern "C"
namespace ref {
'''


def reference_root() -> str | None:
    for cand in (os.environ.get("IONSOLVER_REF"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "src", "kernels", "sim_kernels.cl")):
            return cand
    return None


def kernel_text() -> str:
    root = reference_root()
    if root is None:
        raise FileNotFoundError("reference kernel source not available (no /root/reference on this machine)")
    src = open(os.path.join(root, "src", "kernels", "sim_kernels.cl"), encoding="utf-8").read()
    body = src.split("EndTempDefines%")[1]                       # opencl.rs:67-75
    return re.sub(r"\((float3|uint3|int3)\)\(", r"\1(", body)    # the one vector-literal rewrite


def translation_unit(defines: str) -> str:
    return ("// GENERATED by oracle/build_ref.py from the reference kernel source -- do not commit\n"
            + defines + '#include "ocl_shim.h"\nnamespace ref {\n' + kernel_text() + DRIVER + "}\n")


def tag_for(defines: str) -> str:
    return hashlib.sha1((defines + DRIVER + " ".join(CXXFLAGS)).encode()).hexdigest()[:16]


def lib_path_for(defines: str) -> str:
    return os.path.join(REF_DIR, f"libref_{tag_for(defines)}.so")


def build(cfg, geometry, force=False) -> str:
    """Build (or reuse) the reference library for one domain of `cfg`; returns its path."""
    from . import ref_host
    defines = ref_host.device_defines(cfg, geometry)
    out = lib_path_for(defines)
    if os.path.isfile(out) and not force:
        return out
    os.makedirs(REF_DIR, exist_ok=True)
    tu = out[:-3] + ".cpp"
    with open(tu, "w") as f:
        f.write(translation_unit(defines))
    cmd = ["g++", *CXXFLAGS, "-I", HERE, tu, "-o", out + ".tmp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-4000:])
        raise RuntimeError("reference kernel build failed: " + " ".join(cmd))
    os.replace(out + ".tmp", out)
    os.remove(tu)  # generated from reference text: never kept
    return out


def available(cfg, geometry) -> bool:
    from . import ref_host
    return os.path.isfile(lib_path_for(ref_host.device_defines(cfg, geometry))) or reference_root() is not None
