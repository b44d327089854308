"""oracle/port.py -- TEST INFRASTRUCTURE ONLY (checker; never imported by the product path).

ctypes driver for oracle/lbm_oracle.c, the plain-C restatement of the reference kernels.  `PortDomain` has
the same interface as `ref_host.RefDomain` (the reference's own kernel source compiled for the host), so
`RefLbm(cfg, backend="port")` runs the host sequence of /root/reference/src/lbm/mod.rs over the restatement.
Unlike the `_ref` libraries (one per #define set, built only where /root/reference exists) the restatement
takes its parameters at run time, so it serves every configuration on the GPU box.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

from . import ref_host as rh

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lbm_oracle.c")
LIB = os.path.join(HERE, "liblbm_oracle.so")
# -ffp-contract=off: nothing is fused except the explicit fmaf() calls (same rule as oracle/_ref)
CFLAGS = ["-std=gnu11", "-O3", "-mavx2", "-mfma", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-w"]


def build(force=False) -> str:
    if os.path.isfile(LIB) and not force and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", *CFLAGS, SRC, "-o", LIB + ".tmp", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-4000:])
        raise RuntimeError("oracle build failed: " + " ".join(cmd))
    os.replace(LIB + ".tmp", LIB)
    return LIB


class OraParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("nx", "ny", "nz", "dx", "dy", "dz", "di")] + \
               [(n, ctypes.c_int32) for n in ("ox", "oy", "oz")] + \
               [(n, ctypes.c_uint32) for n in ("velocity_set", "trt", "float_type", "eq_boundaries", "volume_force",
                                               "force_field", "magneto_hydro", "update_fields")] + \
               [(n, ctypes.c_float) for n in ("w", "ke", "kmu", "kmu0", "kkge", "wq")] + \
               [(n, ctypes.c_uint32) for n in ("lod_depth", "n_lod", "n_lod_own")] + [("threads", ctypes.c_int32)] + \
               [("subgrid_ecr", ctypes.c_uint32)] + [(n, ctypes.c_float) for n in ("kme", "kkbme", "keabs", "ecrf")]


class OraBuffers(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("fi", "rho", "u", "flags", "F", "E_stat", "B_stat", "E_dyn", "B_dyn", "fqi",
                                               "ei", "Q", "QU_lod", "transfer_p", "transfer_m", "E_var", "eti", "Et")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        c = ctypes
        P, B = c.POINTER(OraParams), c.POINTER(OraBuffers)
        L.ora_stream_collide.argtypes = [P, B, c.c_uint64, c.c_float, c.c_float, c.c_float]
        L.ora_initialize.argtypes = [P, B]
        L.ora_update_fields.argtypes = [P, B, c.c_uint64]
        L.ora_clear_qu_lod.argtypes = [P, B]
        L.ora_lod_part_2_gather.argtypes = [P, B, c.c_uint32]
        L.ora_update_e_b_dynamic.argtypes = [P, B]
        L.ora_transfer.argtypes = [P, B, c.c_int, c.c_int, c.c_uint32, c.c_uint64]
        L.ora_voxelize_mesh.argtypes = [P, B, c.c_uint32, c.c_uint8, c.c_void_p, c.c_void_p, c.c_void_p, c.c_void_p,
                                        c.c_float, c.c_float, c.c_float]
        L.ora_psi_from_mesh.argtypes = [P, B]
        L.ora_static_b_from_mesh.argtypes = [P, B]
        L.ora_static_e_from_mesh.argtypes = [P, B, c.c_void_p]
        L.ora_codec.argtypes = [P, c.c_void_p, c.c_void_p, c.c_uint64, c.c_int]
        L.ora_neighbors.argtypes = [P, c.c_uint32, c.c_void_p]
        L.ora_lod_index.argtypes = [P, c.c_uint32, c.c_uint32]
        L.ora_lod_index.restype = c.c_uint32
        L.ora_max_threads.restype = c.c_int
        L.ora_nan_stores.argtypes = [c.c_int]
        L.ora_nan_stores.restype = c.c_uint64
        for n in ("ora_stream_collide", "ora_initialize", "ora_update_fields", "ora_clear_qu_lod", "ora_lod_part_2_gather",
                  "ora_update_e_b_dynamic", "ora_transfer", "ora_voxelize_mesh", "ora_psi_from_mesh",
                  "ora_static_b_from_mesh", "ora_static_e_from_mesh", "ora_codec", "ora_neighbors"):
            getattr(L, n).restype = None
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().ora_max_threads())


def nan_stores(reset=False) -> int:
    """NaNs handed to the DDF encoders since the last reset (test hook: a scene that stores NaN is outside the domain where
    parity is defined, see lbm_oracle.c)."""
    return int(lib().ora_nan_stores(1 if reset else 0))


VS_ID = {"D2Q9": 0, "D3Q15": 1, "D3Q19": 2, "D3Q27": 3}
FP_ID = {"FP16S": 0, "FP16C": 1, "FP32": 2}


def params_for(cfg: rh.RefConfig, g: rh.Geometry, threads=0) -> OraParams:
    """The run-time twin of ref_host.device_defines (get_device_defines, domain.rs:736-858)."""
    f32 = np.float32
    p = OraParams()
    p.nx, p.ny, p.nz = g.n_x, g.n_y, g.n_z
    p.dx, p.dy, p.dz, p.di = cfg.d_x, cfg.d_y, cfg.d_z, g.d_i
    p.ox, p.oy, p.oz = g.o_x, g.o_y, g.o_z
    p.velocity_set = VS_ID[cfg.velocity_set]
    p.trt = int(cfg.relaxation_time == "TRT")
    p.float_type = FP_ID[cfg.float_type]
    p.eq_boundaries = int(cfg.ext_equilibrium_boudaries)
    p.volume_force = int(cfg.ext_volume_force)
    p.force_field = int(cfg.ext_force_field)
    p.magneto_hydro = int(cfg.ext_magneto_hydro)
    p.update_fields = int(cfg.graphics_active)
    p.w = float(f32(1.0) / f32(f32(3.0) * f32(cfg.nu) + f32(0.5)))
    u = cfg.units
    p.ke = float(u.ke_lu())
    p.kmu0 = float(u.mu_0_lu())
    p.kmu = float(f32(u.mu_0_lu() / f32(f32(4.0) * f32(np.pi))))
    p.kkge = float(u.kkge_lu())
    p.wq = float(f32(f32(1.0) / f32(f32(f32(2.0) * u.k_charge_expansion_lu()) + f32(0.5))))
    p.lod_depth, p.n_lod, p.n_lod_own = cfg.mhd_lod_depth, g.n_lod, g.n_lod_own
    p.threads = threads
    p.subgrid_ecr = int(cfg.ext_subgrid_ecr)
    p.kme, p.kkbme, p.keabs = float(u.kme_lu()), float(u.kkBme_lu()), float(u.keabs_lu())
    p.ecrf = float(f32(cfg.ecr_freq))
    return p


class PortDomain(rh.RefDomain):
    """LbmDomain (domain.rs:20-80) over the plain-C restatement."""

    def _load(self, lib_path=None):
        self.lib = lib()
        self.lib_path = LIB
        self.p = params_for(self.cfg, self.g, self.threads)

    def _pb(self):
        b = OraBuffers()
        P = rh._ptr
        b.fi, b.rho, b.u, b.flags, b.F = P(self.fi), P(self.rho), P(self.u), P(self.flags), P(self.f)
        b.E_stat, b.B_stat, b.E_dyn, b.B_dyn = P(self.e_stat), P(self.b_stat), P(self.e_dyn), P(self.b_dyn)
        b.fqi, b.ei, b.Q, b.QU_lod = P(self.fqi), P(self.ei), P(self.qc), P(self.qu_lod)
        b.transfer_p, b.transfer_m = P(self.transfer_p), P(self.transfer_m)
        b.E_var, b.eti, b.Et = P(self.e_var), P(self.eti), P(self.et)
        return ctypes.byref(self.p), ctypes.byref(b)

    def enqueue_initialize(self):
        self.lib.ora_initialize(*self._pb())

    def enqueue_stream_collide(self, begin=0, end=None):
        self.lib.ora_stream_collide(*self._pb(), self.t, self.fx, self.fy, self.fz)

    def enqueue_update_fields(self):
        self.lib.ora_update_fields(*self._pb(), self.t)

    def enqueue_update_e_b_dyn(self, begin=0, end=None):
        self.lib.ora_update_e_b_dynamic(*self._pb())

    def enqueue_lod_part_2_gather(self):
        for i in reversed(range(0, self.cfg.mhd_lod_depth)):
            self.lib.ora_lod_part_2_gather(*self._pb(), i)

    def enqueue_clear_qu_lod(self):
        self.lib.ora_clear_qu_lod(*self._pb())

    def enqueue_transfer_extract_field(self, field, direction):
        self.lib.ora_transfer(*self._pb(), field, 0, direction, self.t)

    def enqueue_transfer_insert_field(self, field, direction):
        self.lib.ora_transfer(*self._pb(), field, 1, direction, self.t)

    def enqueue_precompute_b(self):
        self.lib.ora_psi_from_mesh(*self._pb())
        self.lib.ora_static_b_from_mesh(*self._pb())

    def enqueue_precompute_e(self):
        self.lib.ora_static_e_from_mesh(*self._pb(), self.e_stat.ctypes.data)

    def enqueue_precompute_e_ecr(self):
        self.lib.ora_static_e_from_mesh(*self._pb(), self.e_var.ctypes.data)

    def _voxelize(self, direction, flag, mpc):
        self.lib.ora_voxelize_mesh(*self._pb(), direction, flag, self.p0.ctypes.data, self.p1.ctypes.data,
                                   self.p2.ctypes.data, self.bbu.ctypes.data, mpc[0], mpc[1], mpc[2])

    # probes
    def codec(self, arr, direction):
        p, _ = self._pb()
        if direction == 0:
            a = np.ascontiguousarray(arr, np.float32)
            out = np.empty(a.size, np.float32 if self.cfg.float_type == "FP32" else np.uint16)
        else:
            a = np.ascontiguousarray(arr)
            out = np.empty(a.size, np.float32)
        self.lib.ora_codec(p, a.ctypes.data, out.ctypes.data, a.size, direction)
        return out

    def neighbors(self, n):
        p, _ = self._pb()
        j = np.zeros(self.q, np.uint32)
        self.lib.ora_neighbors(p, int(n), j.ctypes.data)
        return j
