"""oracle/ref_host.py -- TEST INFRASTRUCTURE ONLY (checker; never imported by the product path).

CPU restatement of the reference's *host* logic that surrounds the hot path, used to drive the
shim-compiled reference kernels (oracle/_ref, see build_ref.py) exactly like the Rust host does:

  * Units                 -> /root/reference/src/lbm/units.rs:41-197  (f32 arithmetic reproduced with numpy.float32)
  * RefConfig             -> /root/reference/src/lbm/mod.rs:46-135    (LbmConfig + defaults)
  * domain_geometry       -> /root/reference/src/lbm/domain.rs:88-126 (sizes, offsets, LOD counts)
  * device_defines        -> /root/reference/src/lbm/domain.rs:736-858 (get_device_defines)
  * RefLbm                -> /root/reference/src/lbm/mod.rs:166-272,371-468 (new/initialize/do_time_step/communicate_*)
                             /root/reference/src/lbm/domain.rs:412-578 (enqueue_* launch sizes and argument order)
  * Mesh / voxelise       -> /root/reference/src/mesh.rs:175-343

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes
import dataclasses
import math
import os
import struct

import numpy as np

f32 = np.float32

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

# velocity set -> (dimensions, velocity_set, transfers)   types.rs:37-46
SET_VALUES = {"D2Q9": (2, 9, 3), "D3Q15": (3, 15, 5), "D3Q19": (3, 19, 5), "D3Q27": (3, 27, 9)}
FLOAT_SIZE = {"FP16S": 2, "FP16C": 2, "FP32": 4}  # types.rs:85-92
# TransferField discriminants, types.rs:105-111
TF_FI, TF_RHO_U_FLAGS, TF_EI, TF_QI = 0, 1, 2, 3


def _sq(x):
    return f32(x * x)


def _cb(x):
    return f32(f32(x * x) * x)


def _to4(x):
    return f32(f32(f32(x * x) * x) * x)


class Units:
    """units.rs:14-27; all members f32."""

    def __init__(self):
        self.m = f32(1.0)
        self.kg = f32(1.0)
        self.s = f32(1.0)
        self.a = f32(1.0)
        self.k = f32(1.0)
        self.prop_atom_mass = 1.6735575e-27  # Propellant::H default, units.rs:238-247

    def set(self, lbm_length, lbm_velocity, lbm_rho, lbm_charge, lbm_temp, si_length, si_velocity, si_rho,
            si_charge, si_temp):  # units.rs:41-59
        L = [f32(v) for v in (lbm_length, lbm_velocity, lbm_rho, lbm_charge, lbm_temp, si_length, si_velocity,
                              si_rho, si_charge, si_temp)]
        lbm_length, lbm_velocity, lbm_rho, lbm_charge, lbm_temp, si_length, si_velocity, si_rho, si_charge, si_temp = L
        self.m = f32(si_length / lbm_length)
        self.kg = f32(f32(si_rho / lbm_rho) * _cb(self.m))
        self.s = f32(self.m / f32(si_velocity / lbm_velocity))
        self.a = f32(f32(si_charge / lbm_charge) / self.s)
        self.k = f32(si_temp / lbm_temp)

    # si -> lu (units.rs:102-147)
    def len_si_lu(self, l):
        return f32(f32(l) / self.m)

    def nu_si_lu(self, nu):
        return f32(f32(nu) / f32(_sq(self.m) / self.s))

    def charge_si_lu(self, q):
        return f32(f32(q) / f32(self.a * self.s))

    def mag_flux_si_lu(self, b):
        return f32(f32(b) / f32(self.kg / f32(self.a * _sq(self.s))))

    def e_field_si_lu(self, e):
        return f32(f32(e) / f32(f32(self.kg * self.m) / f32(self.a * _cb(self.s))))

    def magnetization_si_lu(self, m):
        return f32(f32(m) / f32(self.a / self.m))

    def time_lu_si(self, t):
        return f32(f32(t) * self.s)

    def epsilon_0_lu(self):  # units.rs:149-153
        return f32(f32(8.8541878128E-12) / f32(f32(_sq(self.a) * _to4(self.s)) / f32(self.kg * _cb(self.m))))

    def ke_lu(self):  # units.rs:155-161
        return f32(f32(1.0) / f32(f32(f32(4.0) * f32(math.pi)) * self.epsilon_0_lu()))

    def mu_0_lu(self):  # units.rs:163-167
        return f32(f32(1.256637062E-6) / f32(f32(self.kg * self.m) / f32(_sq(self.a) * _sq(self.s))))

    def k_charge_expansion_lu(self):  # units.rs:169-173
        return f32(1.0)

    def kkge_lu(self):  # units.rs:175-177
        return f32(f32(9.1093837139E-31 / -1.602176634E-19) / f32(self.kg / f32(self.a * self.s)))

    def kimg_lu(self):  # units.rs:179-181
        return f32((1.0 / (self.prop_atom_mass * 1e20)) / float(self.kg))

    def kveV_lu(self):  # units.rs:183-185
        return f32(9.1093837139E-31 / (2.0 * 1.602176634E-19) / float(self.kg))

    def kkBme_lu(self):  # units.rs:187-189
        m, s, k = float(self.m), float(self.s), float(self.k)
        return f32(-22734499.72063751808909449412 / ((m * m) / ((s * s) * k)))

    def keabs_lu(self):  # units.rs:191-193
        a, s, kg = float(self.a), float(self.s), float(self.kg)
        return f32(1.40897016100511360652E-8 / ((a * a) * (s * s) / kg))

    def kme_lu(self):  # units.rs:195-197
        return f32(5.68563006E-12 / (float(self.kg) / (float(self.a) * float(self.s))))


@dataclasses.dataclass
class RefConfig:
    """LbmConfig, mod.rs:46-135 (graphics reduced to the one switch that touches the hot path)."""
    velocity_set: str = "D2Q9"
    relaxation_time: str = "SRT"
    float_type: str = "FP16S"
    units: Units = dataclasses.field(default_factory=Units)
    n_x: int = 1
    n_y: int = 1
    n_z: int = 1
    d_x: int = 1
    d_y: int = 1
    d_z: int = 1
    nu: float = float(f32(1.0) / f32(6.0))
    f_x: float = 0.0
    f_y: float = 0.0
    f_z: float = 0.0
    ext_equilibrium_boudaries: bool = False
    ext_volume_force: bool = False
    ext_force_field: bool = False
    ext_magneto_hydro: bool = False
    ext_subgrid_ecr: bool = False
    mhd_lod_depth: int = 4
    ecr_freq: float = 0.0
    graphics_active: bool = False  # GraphicsConfig::graphics_active (defaults to true in the reference, graphics.rs:160)
    d3q27_patched_weights: bool = True  # quirk Q3: the reference cannot build D3Q27 (no DEF_WC); canonical weights


def c_float_literal(v) -> str:
    """Rust `{:?}` prints the shortest round-trip f32; any round-tripping decimal gives the same f32 bits."""
    v = f32(v)
    if np.isnan(v):
        return "NAN"
    if np.isinf(v):
        return "INFINITY" if v > 0 else "(-INFINITY)"
    s = "%.9g" % float(v)
    if "e" not in s and "." not in s:
        s += ".0"
    return s + "f"


def domain_coords(d, d_x, d_y):  # mod.rs:189-191
    return (d % (d_x * d_y)) % d_x, (d % (d_x * d_y)) // d_x, d // (d_x * d_y)


@dataclasses.dataclass
class Geometry:
    n_x: int
    n_y: int
    n_z: int
    n: int
    o_x: int
    o_y: int
    o_z: int
    n_lod: int
    n_lod_own: int
    d_i: int


def domain_geometry(cfg: RefConfig, x: int, y: int, z: int, i: int) -> Geometry:
    """domain.rs:91-126"""
    n_x = cfg.n_x // cfg.d_x + 2 * (cfg.d_x > 1)
    n_y = cfg.n_y // cfg.d_y + 2 * (cfg.d_y > 1)
    n_z = cfg.n_z // cfg.d_z + 2 * (cfg.d_z > 1)
    o_x = (x * cfg.n_x // cfg.d_x) - (cfg.d_x > 1)
    o_y = (y * cfg.n_y // cfg.d_y) - (cfg.d_y > 1)
    o_z = (z * cfg.n_z // cfg.d_z) - (cfg.d_z > 1)
    dim = SET_VALUES[cfg.velocity_set][0]
    c = 1
    for k in range(cfg.mhd_lod_depth):
        c += (1 << (k + 1)) ** dim
    n_lod_own = c
    d_n = cfg.d_x * cfg.d_y * cfg.d_z
    for d in range(d_n):
        dx, dy, dz = domain_coords(d, cfg.d_x, cfg.d_y)
        dist = max(abs(z - dz), abs(y - dy), abs(x - dx))
        if dist != 0:
            c += (1 << max(cfg.mhd_lod_depth - dist, 0)) ** dim
    return Geometry(n_x, n_y, n_z, n_x * n_y * n_z, o_x, o_y, o_z, c, n_lod_own, i)


def device_defines(cfg: RefConfig, g: Geometry) -> str:
    """get_device_defines, domain.rs:736-858 (graphics-only DEF_DOMAIN_OFFSET_* omitted: unused by sim kernels)."""
    dim, q, transfers = SET_VALUES[cfg.velocity_set]
    L = []
    A = L.append
    A(f"#define DEF_NX {g.n_x}u")
    A(f"#define DEF_NY {g.n_y}u")
    A(f"#define DEF_NZ {g.n_z}u")
    A(f"#define DEF_N  {g.n}ul")
    A(f"#define DEF_DX {cfg.d_x}u")
    A(f"#define DEF_DY {cfg.d_y}u")
    A(f"#define DEF_DZ {cfg.d_z}u")
    A(f"#define DEF_DI {g.d_i}u")
    A(f"#define DEF_OX {g.o_x}")
    A(f"#define DEF_OY {g.o_y}")
    A(f"#define DEF_OZ {g.o_z}")
    A(f"#define DEF_AX {g.n_y * g.n_z}u")
    A(f"#define DEF_AY {g.n_z * g.n_x}u")
    A(f"#define DEF_AZ {g.n_x * g.n_y}u")
    A(f"#define D{dim}Q{q}")
    A(f"#define DEF_VELOCITY_SET {q}u")
    A(f"#define DEF_DIMENSIONS {dim}u")
    A(f"#define DEF_TRANSFERS {transfers}u")
    A("#define DEF_C 0.57735027f")
    w = f32(1.0) / f32(f32(3.0) * f32(cfg.nu) + f32(0.5))
    A(f"#define DEF_W {c_float_literal(w)}")
    if cfg.velocity_set == "D2Q9":
        L += ["#define DEF_W0 (1.0f/2.25f)", "#define DEF_WS (1.0f/9.0f)", "#define DEF_WE (1.0f/36.0f)"]
    elif cfg.velocity_set == "D3Q15":
        L += ["#define DEF_W0 (1.0f/4.5f)", "#define DEF_WS (1.0f/9.0f)", "#define DEF_WC (1.0f/72.0f)"]
    elif cfg.velocity_set == "D3Q19":
        L += ["#define DEF_W0 (1.0f/3.0f)", "#define DEF_WS (1.0f/18.0f)", "#define DEF_WE (1.0f/36.0f)"]
    else:
        if cfg.d3q27_patched_weights:  # quirk Q3: canonical D3Q27 weights, reference is unbuildable
            L += ["#define DEF_W0 (1.0f/3.375f)", "#define DEF_WS (1.0f/13.5f)", "#define DEF_WE (1.0f/54.0f)",
                  "#define DEF_WC (1.0f/216.0f)"]
        else:  # verbatim domain.rs:766-768 (does not compile: DEF_WC missing)
            L += ["#define DEF_W0 (1.0f/3.0f)", "#define DEF_WS (1.0f/18.0f)", "#define DEF_WE (1.0f/36.0f)"]
    A("#define SRT" if cfg.relaxation_time == "SRT" else "#define TRT")
    L += ["#define TYPE_S  0x01", "#define TYPE_E  0x02", "#define TYPE_C  0x04", "#define TYPE_F  0x08",
          "#define TYPE_M  0x10", "#define TYPE_G  0x20", "#define TYPE_X  0x40", "#define TYPE_Y  0x80",
          "#define TYPE_MS 0x03", "#define TYPE_BO 0b00011111"]
    if cfg.float_type == "FP16S":
        L += ["#define fpxx half", "#define fpxx_copy ushort", "#define load(p,o) vload_half(o,p)*3.0517578E-5f",
              "#define store(p,o,x) vstore_half_rte((x)*32768.0f,o,p)"]
    elif cfg.float_type == "FP16C":
        L += ["#define fpxx ushort", "#define fpxx_copy ushort", "#define load(p,o) half_to_float_custom(p[o])",
              "#define store(p,o,x) p[o]=float_to_half_custom(x)"]
    else:
        L += ["#define fpxx float", "#define fpxx_copy float", "#define load(p,o) p[o]", "#define store(p,o,x) p[o]=x"]
    if cfg.ext_equilibrium_boudaries:
        A("#define EQUILIBRIUM_BOUNDARIES")
    if cfg.ext_volume_force:
        A("#define VOLUME_FORCE")
    if cfg.ext_magneto_hydro:
        u = cfg.units
        A("#define MAGNETO_HYDRO")
        A(f"#define DEF_KE {c_float_literal(u.ke_lu())}")
        A(f"#define DEF_KMU {c_float_literal(f32(u.mu_0_lu() / f32(f32(4.0) * f32(math.pi))))}")
        A(f"#define DEF_KMU0 {c_float_literal(u.mu_0_lu())}")
        A(f"#define DEF_KKGE {c_float_literal(u.kkge_lu())}")
        A(f"#define DEF_KIMG {c_float_literal(u.kimg_lu())}")
        A(f"#define DEF_KVEV {c_float_literal(u.kveV_lu())}")
        A(f"#define DEF_KME {c_float_literal(u.kme_lu())}")
        A(f"#define DEF_LOD_DEPTH {cfg.mhd_lod_depth}u")
        A(f"#define DEF_NUM_LOD {g.n_lod}u")
        A(f"#define DEF_NUM_LOD_OWN {g.n_lod_own}u")
        wq = f32(f32(1.0) / f32(f32(f32(2.0) * u.k_charge_expansion_lu()) + f32(0.5)))
        A(f"#define DEF_WQ {c_float_literal(wq)}")
    if cfg.ext_subgrid_ecr:
        A("#define SUBGRID_ECR")
        A(f"#define DEF_KKBME {c_float_literal(cfg.units.kkBme_lu())}")
        A(f"#define DEF_KEABS {c_float_literal(cfg.units.keabs_lu())}")
    if cfg.ext_force_field:
        A("#define FORCE_FIELD")
    if cfg.graphics_active:
        A("#define UPDATE_FIELDS")
    return "\n".join(L) + "\n"


# ------------------------------------------------------------------------------------------------
# ctypes view of the driver appended by build_ref.py
# ------------------------------------------------------------------------------------------------
class RefBuffers(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "fi", "rho", "u", "flags", "F", "E_stat", "B_stat", "E_dyn", "B_dyn", "fqi", "ei", "Q", "QU_lod", "E_var",
        "eti", "Et", "transfer_p", "transfer_m", "p0", "p1", "p2", "bbu")]


def _ptr(a):
    return None if a is None else a.ctypes.data


class RefDomain:
    """LbmDomain (domain.rs:20-80) over one shim-compiled library; buffers are numpy arrays."""

    def __init__(self, cfg: RefConfig, x, y, z, i, lib_path=None, threads=0):
        from . import build_ref
        self.cfg = cfg
        self.g = g = domain_geometry(cfg, x, y, z, i)
        self.threads = threads
        self._load(lib_path)
        dim, q, transfers = SET_VALUES[cfg.velocity_set]
        self.q, self.transfers = q, transfers
        n = g.n
        ddf_t = np.float32 if cfg.float_type == "FP32" else np.uint16
        self.t = 0
        self.fx, self.fy, self.fz = cfg.f_x, cfg.f_y, cfg.f_z
        self.fi = np.zeros(n * q, ddf_t)                      # domain.rs:151-158
        self.rho = np.ones(n, np.float32)
        self.u = np.zeros(3 * n, np.float32)
        self.flags = np.zeros(n, np.uint8)
        self.p0 = np.zeros(1, np.float32)
        self.p1 = np.zeros(1, np.float32)
        self.p2 = np.zeros(1, np.float32)
        self.bbu = np.zeros(16, np.float32)
        self.f = np.zeros(3 * n, np.float32) if cfg.ext_force_field else None
        mhd = cfg.ext_magneto_hydro
        self.e_stat = np.zeros(3 * n, np.float32) if mhd else None
        self.e_dyn = np.zeros(3 * n, np.float32) if mhd else None
        self.b_stat = np.zeros(3 * n, np.float32) if mhd else None
        self.b_dyn = np.zeros(3 * n, np.float32) if mhd else None
        self.fqi = np.zeros(7 * n, ddf_t) if mhd else None
        self.ei = np.zeros(q * n, ddf_t) if mhd else None
        self.qc = np.zeros(n, np.float32) if mhd else None  # `q` in domain.rs:191
        self.qu_lod = np.zeros(4 * g.n_lod, np.float32) if mhd else None
        ecr = cfg.ext_subgrid_ecr
        self.e_var = np.zeros(3 * n, np.float32) if ecr else None
        self.eti = np.zeros(7 * n, ddf_t) if ecr else None
        self.et = np.zeros(n, np.float32) if ecr else None
        a_max = 0                                             # domain.rs:311-322
        if cfg.d_x > 1:
            a_max = max(a_max, g.n_y * g.n_z)
        if cfg.d_y > 1:
            a_max = max(a_max, g.n_x * g.n_z)
        if cfg.d_z > 1:
            a_max = max(a_max, g.n_x * g.n_y)
        tsize = a_max * max(17, transfers * FLOAT_SIZE[cfg.float_type])
        self.transfer_p = np.zeros(max(tsize, 1), np.uint8)
        self.transfer_m = np.zeros(max(tsize, 1), np.uint8)
        self.transfer_lod_host = np.zeros(4 * g.n_lod_own, np.float32) if mhd else None

    def _load(self, lib_path=None):
        from . import build_ref
        self.lib_path = lib_path or build_ref.build(self.cfg, self.g)
        self.lib = ctypes.CDLL(self.lib_path)
        self._declare()
        if self.threads:
            self.lib.ref_set_threads(self.threads)

    def _declare(self):
        L, c = self.lib, ctypes
        P = c.POINTER(RefBuffers)
        L.ref_set_threads.argtypes = [c.c_int]
        L.ref_stream_collide.argtypes = [P, c.c_uint64, c.c_uint64, c.c_uint64, c.c_float, c.c_float, c.c_float,
                                         c.c_float]
        L.ref_initialize.argtypes = [P, c.c_uint64, c.c_uint64]
        L.ref_update_fields.argtypes = [P, c.c_uint64, c.c_uint64, c.c_uint64, c.c_float, c.c_float, c.c_float]
        L.ref_update_e_b_dynamic.argtypes = [P, c.c_uint64, c.c_uint64]
        L.ref_clear_qu_lod.argtypes = [P, c.c_uint64]
        L.ref_lod_part_2_gather.argtypes = [P, c.c_uint32]
        L.ref_transfer.argtypes = [P, c.c_int, c.c_int, c.c_uint32, c.c_uint64, c.c_uint64]
        L.ref_voxelize_mesh.argtypes = [P, c.c_uint32, c.c_uint64, c.c_uint8, c.c_float, c.c_float, c.c_float,
                                        c.c_uint64]
        L.ref_psi_from_mesh.argtypes = [P, c.c_uint64, c.c_uint64]
        L.ref_static_b_from_mesh.argtypes = [P, c.c_uint64, c.c_uint64]
        L.ref_static_e_from_mesh.argtypes = [P, c.c_int, c.c_uint64, c.c_uint64]
        L.ref_codec.argtypes = [c.c_void_p, c.c_void_p, c.c_uint64, c.c_int]
        L.ref_neighbors.argtypes = [c.c_uint32, c.c_void_p]
        L.ref_has.argtypes = [c.c_char_p]
        L.ref_has.restype = c.c_int
        for name in ("ref_stream_collide", "ref_initialize", "ref_update_fields", "ref_update_e_b_dynamic",
                     "ref_clear_qu_lod", "ref_lod_part_2_gather", "ref_transfer", "ref_voxelize_mesh",
                     "ref_psi_from_mesh", "ref_static_b_from_mesh", "ref_static_e_from_mesh", "ref_codec",
                     "ref_neighbors", "ref_set_threads"):
            getattr(L, name).restype = None

    def bufs(self):
        b = RefBuffers()
        b.fi, b.rho, b.u, b.flags = _ptr(self.fi), _ptr(self.rho), _ptr(self.u), _ptr(self.flags)
        b.F = _ptr(self.f)
        b.E_stat, b.B_stat, b.E_dyn, b.B_dyn = _ptr(self.e_stat), _ptr(self.b_stat), _ptr(self.e_dyn), _ptr(self.b_dyn)
        b.fqi, b.ei, b.Q, b.QU_lod = _ptr(self.fqi), _ptr(self.ei), _ptr(self.qc), _ptr(self.qu_lod)
        b.E_var, b.eti, b.Et = _ptr(self.e_var), _ptr(self.eti), _ptr(self.et)
        b.transfer_p, b.transfer_m = _ptr(self.transfer_p), _ptr(self.transfer_m)
        b.p0, b.p1, b.p2, b.bbu = _ptr(self.p0), _ptr(self.p1), _ptr(self.p2), _ptr(self.bbu)
        return ctypes.byref(b)

    # ---- enqueue_* (domain.rs:412-578); global sizes as in domain.rs:217-287 ----
    def enqueue_initialize(self):
        self.lib.ref_initialize(self.bufs(), 0, self.g.n)

    def enqueue_stream_collide(self, begin=0, end=None):
        self.lib.ref_stream_collide(self.bufs(), begin, self.g.n if end is None else end, self.t, self.fx, self.fy,
                                    self.fz, self.cfg.ecr_freq)

    def enqueue_update_fields(self):
        self.lib.ref_update_fields(self.bufs(), 0, self.g.n, self.t, self.fx, self.fy, self.fz)

    def enqueue_update_e_b_dyn(self, begin=0, end=None):
        self.lib.ref_update_e_b_dynamic(self.bufs(), begin, self.g.n if end is None else end)

    def enqueue_lod_part_2_gather(self):  # domain.rs:453-462
        for i in reversed(range(0, self.cfg.mhd_lod_depth)):
            self.lib.ref_lod_part_2_gather(self.bufs(), i)

    def enqueue_clear_qu_lod(self):  # global size n_lod (domain.rs:277), guard n>NUM_LOD_OWN (quirk Q12)
        self.lib.ref_clear_qu_lod(self.bufs(), self.g.n_lod)

    def get_area(self, direction):  # domain.rs:475-482
        g = self.g
        return (g.n_y * g.n_z, g.n_x * g.n_z, g.n_x * g.n_y)[direction]

    def enqueue_transfer_extract_field(self, field, direction):  # domain.rs:484-513
        self.lib.ref_transfer(self.bufs(), field, 0, direction, self.t, self.get_area(direction))

    def enqueue_transfer_insert_field(self, field, direction):  # domain.rs:516-543
        self.lib.ref_transfer(self.bufs(), field, 1, direction, self.t, self.get_area(direction))

    def read_lods(self):  # domain.rs:547-549
        self.transfer_lod_host[:] = self.qu_lod[:4 * self.g.n_lod_own]

    def enqueue_precompute_b(self):  # domain.rs:551-556
        g = self.g
        self.lib.ref_psi_from_mesh(self.bufs(), 0, (g.n_x + 2) * (g.n_y + 2) * (g.n_z + 2))
        self.lib.ref_static_b_from_mesh(self.bufs(), 0, g.n)

    def enqueue_precompute_e(self):  # domain.rs:558-567
        self.lib.ref_static_e_from_mesh(self.bufs(), 0, 0, self.g.n)

    def enqueue_precompute_e_ecr(self):  # domain.rs:569-578
        self.lib.ref_static_e_from_mesh(self.bufs(), 1, 0, self.g.n)

    def voxelize_mesh_on_device(self, mesh: "Mesh", ctype: str, value):  # mesh.rs:281-343
        u = self.cfg.units
        self.p0 = np.ascontiguousarray(mesh.p0.reshape(-1), np.float32)
        self.p1 = np.ascontiguousarray(mesh.p1.reshape(-1), np.float32)
        self.p2 = np.ascontiguousarray(mesh.p2.reshape(-1), np.float32)
        two = f32(2.0)
        x0, y0, z0 = (f32(mesh.p_min[k] - two) for k in range(3))
        x1, y1, z1 = (f32(mesh.p_max[k] + two) for k in range(3))
        self.bbu[:] = 0
        self.bbu[0] = np.array([mesh.triangle_number], np.uint32).view(np.float32)[0]
        self.bbu[1:7] = [x0, y0, z0, x1, y1, z1]
        c = [f32(f32(y1 - y0) * f32(z1 - z0)), f32(f32(z1 - z0) * f32(x1 - x0)), f32(f32(x1 - x0) * f32(y1 - y0))]
        direction = 0 if (c[0] < c[1] and c[0] < c[2]) else (1 if c[1] < c[2] else 2)
        flag = {"Solid": 0b00000001, "Magnet": 0b00010001, "Charged": 0b00001001, "ChargedECR": 0b00000101}[ctype]
        mpc = [0.0, 0.0, 0.0]
        if self.cfg.ext_magneto_hydro:
            if ctype == "Magnet":
                mpc = [u.magnetization_si_lu(v) for v in value]
            elif ctype in ("Charged", "ChargedECR"):
                mpc[0] = u.charge_si_lu(value)
        self._voxelize(direction, flag, mpc)
        return direction, flag, mpc

    # probes
    def codec(self, arr, direction):  # 0: float -> stored, 1: stored -> float
        if direction == 0:
            a = np.ascontiguousarray(arr, np.float32)
            out = np.empty(a.size, np.float32 if self.cfg.float_type == "FP32" else np.uint16)
        else:
            a = np.ascontiguousarray(arr)
            out = np.empty(a.size, np.float32)
        self.lib.ref_codec(a.ctypes.data, out.ctypes.data, a.size, direction)
        return out

    def neighbors(self, n):
        j = np.zeros(self.q, np.uint32)
        self.lib.ref_neighbors(int(n), j.ctypes.data)
        return j

    def _voxelize(self, direction, flag, mpc):
        self.lib.ref_voxelize_mesh(self.bufs(), direction, self.t + 1, flag, mpc[0], mpc[1], mpc[2],
                                   self.get_area(direction))


class Mesh:
    """mesh.rs:69-220 (binary STL import, f32 arithmetic)."""

    def __init__(self, p0, p1, p2, center):
        self.p0, self.p1, self.p2 = p0, p1, p2
        self.triangle_number = len(p0)
        self.center = np.asarray(center, np.float32)
        self.update_bounds()

    def update_bounds(self):  # mesh.rs:93-108
        allp = np.stack([self.p0, self.p1, self.p2])
        self.p_min = allp.min(axis=(0, 1)).astype(np.float32)
        self.p_max = allp.max(axis=(0, 1)).astype(np.float32)

    def translate(self, t):  # mesh.rs:120-129
        t = np.asarray(t, np.float32)
        self.p0 = (self.p0 + t).astype(np.float32)
        self.p1 = (self.p1 + t).astype(np.float32)
        self.p2 = (self.p2 + t).astype(np.float32)
        self.center = (self.center + t).astype(np.float32)
        self.p_min = (self.p_min + t).astype(np.float32)
        self.p_max = (self.p_max + t).astype(np.float32)

    @staticmethod
    def _rotm_around_v(v, r):  # mesh.rs:58-66 (f32 sin/cos)
        r = f32(r)
        sr, cr = f32(np.sin(r)), f32(np.cos(r))
        x, y, z = (f32(c) for c in v)
        one = f32(1.0)
        sq = lambda a: f32(a * a)
        omc = f32(one - cr)
        return np.array([
            [f32(sq(x) + f32(f32(one - sq(x)) * cr)), f32(f32(f32(x * y) * omc) - f32(z * sr)), f32(f32(f32(x * z) * omc) + f32(y * sr))],
            [f32(f32(f32(x * y) * omc) + f32(z * sr)), f32(sq(y) + f32(f32(one - sq(y)) * cr)), f32(f32(f32(y * z) * omc) - f32(x * sr))],
            [f32(f32(f32(x * z) * omc) - f32(y * sr)), f32(f32(f32(y * z) * omc) + f32(x * sr)), f32(sq(z) + f32(f32(one - sq(z)) * cr))],
        ], np.float32)

    @staticmethod
    def _matmul(a, b):  # mesh.rs:418-428 (left-to-right f32 sums)
        r = np.zeros((3, 3), np.float32)
        for i in range(3):
            for j in range(3):
                r[i, j] = f32(f32(f32(a[i, 0] * b[0, j]) + f32(a[i, 1] * b[1, j])) + f32(a[i, 2] * b[2, j]))
        return r

    @staticmethod
    def rotation_matrix(rx, ry, rz):  # mesh.rs:52-56 ; radians
        m = Mesh._matmul(Mesh._rotm_around_v((1, 0, 0), rx), Mesh._rotm_around_v((0, 1, 0), ry))
        return Mesh._matmul(m, Mesh._rotm_around_v((0, 0, 1), rz))

    @staticmethod
    def _apply(rot, p):  # mesh.rs:392-402
        p = p.astype(np.float32)
        out = np.empty_like(p)
        for i in range(3):
            out[:, i] = ((rot[i, 0] * p[:, 0]).astype(np.float32) + (rot[i, 1] * p[:, 1]).astype(np.float32)).astype(
                np.float32) + (rot[i, 2] * p[:, 2]).astype(np.float32)
        return out.astype(np.float32)

    @staticmethod
    def read_stl_raw(data: bytes, reposition, box_size, center, rotation, size):  # mesh.rs:175-220
        tn = struct.unpack_from("<I", data, 80)[0]
        if not (tn > 0 and len(data) == 84 + 50 * tn):
            raise ValueError("Mesh import failed")
        rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=tn,
                            offset=84)
        v = rec["v"].astype(np.float32)
        p0, p1, p2 = (Mesh._apply(rotation, v[:, k, :]) for k in range(3))
        center = np.asarray(center, np.float32)
        mesh = Mesh(p0, p1, p2, center)
        ext = (mesh.p_max - mesh.p_min).astype(np.float32)
        if size == 0.0:
            bs = np.asarray(box_size, np.float32)
            scale = f32(min(f32(bs[0] / ext[0]), min(f32(bs[1] / ext[1]), f32(bs[2] / ext[2]))))
        elif size > 0.0:
            scale = f32(f32(size) / f32(max(ext[0], max(ext[1], ext[2]))))
        else:
            scale = f32(-f32(size))
        if reposition:
            offset = (f32(-0.5) * (mesh.p_min + mesh.p_max).astype(np.float32)).astype(np.float32)
        else:
            offset = np.zeros(3, np.float32)
        for name in ("p0", "p1", "p2"):
            p = getattr(mesh, name)
            setattr(mesh, name, (center + (scale * (offset + p).astype(np.float32)).astype(np.float32)).astype(
                np.float32))
        mesh.update_bounds()
        return mesh


class RefLbm:
    """Lbm (mod.rs:152-495) over RefDomain objects."""

    def __init__(self, cfg: RefConfig, threads=0, backend="ref"):
        """backend "ref": the reference's own kernel source compiled for the host (oracle/_ref);
        backend "port": the plain-C restatement (oracle/lbm_oracle.c)."""
        cfg = dataclasses.replace(cfg)
        cfg.n_x = (cfg.n_x // cfg.d_x) * cfg.d_x  # mod.rs:167-179
        cfg.n_y = (cfg.n_y // cfg.d_y) * cfg.d_y
        cfg.n_z = (cfg.n_z // cfg.d_z) * cfg.d_z
        self.config = cfg
        self.domains = []
        for d in range(cfg.d_x * cfg.d_y * cfg.d_z):
            x, y, z = domain_coords(d, cfg.d_x, cfg.d_y)
            if backend == "port":
                from .port import PortDomain
                self.domains.append(PortDomain(cfg, x, y, z, d, threads=threads))
            else:
                self.domains.append(RefDomain(cfg, x, y, z, d, threads=threads))
        self.meshes = []
        self.initialized = False

    def get_d_n(self):
        return len(self.domains)

    # mod.rs:214-231
    def initialize(self):
        self.increment_timestep(1)
        self.communicate_rho_u_flags()
        for d in self.domains:
            d.enqueue_initialize()
        self.communicate_rho_u_flags()
        self.communicate_fi()
        if self.config.ext_magneto_hydro:
            self.communicate_fqi()
            self.communicate_ei()
            self.communicate_qu_lods()
            self.update_e_b_dynamic()
        for d in self.domains:
            d.t = 0
        self.initialized = True

    def run(self, steps):  # mod.rs:235-245
        if not self.initialized:
            self.initialize()
        for _ in range(steps):
            self.do_time_step()

    def do_time_step(self):  # mod.rs:250-272
        mhd = self.config.ext_magneto_hydro
        if mhd:
            for d in self.domains:
                d.enqueue_clear_qu_lod()
        for d in self.domains:
            d.enqueue_stream_collide()
        if self.config.graphics_active:
            self.communicate_rho_u_flags()
        self.communicate_fi()
        if mhd:
            if len(self.domains) > 1:
                for d in self.domains:
                    d.enqueue_lod_part_2_gather()
            self.communicate_fqi()
            self.communicate_ei()
            self.communicate_qu_lods()
            self.update_e_b_dynamic()
        self.increment_timestep(1)

    def update_e_b_dynamic(self):
        for d in self.domains:
            d.enqueue_update_e_b_dyn()

    def increment_timestep(self, s):
        for d in self.domains:
            d.t += s

    def communicate_field(self, field):  # mod.rs:371-407
        c = self.config
        dxyz = (c.d_x, c.d_y, c.d_z)
        d_n = self.get_d_n()
        for axis in range(3):
            if dxyz[axis] <= 1:
                continue
            for d in self.domains:
                d.enqueue_transfer_extract_field(field, axis)
            for d in range(d_n):
                x, y, z = domain_coords(d, c.d_x, c.d_y)
                if axis == 0:
                    dp = ((x + 1) % c.d_x) + (y + z * c.d_y) * c.d_x
                elif axis == 1:
                    dp = x + (((y + 1) % c.d_y) + z * c.d_y) * c.d_x
                else:
                    dp = x + (y + ((z + 1) % c.d_z) * c.d_y) * c.d_x
                a, b = self.domains[d], self.domains[dp]
                a.transfer_p, b.transfer_m = b.transfer_m, a.transfer_p  # ptr::swap, mod.rs:383
            for d in self.domains:
                d.enqueue_transfer_insert_field(field, axis)

    def communicate_fi(self):
        self.communicate_field(TF_FI)

    def communicate_rho_u_flags(self):
        self.communicate_field(TF_RHO_U_FLAGS)

    def communicate_fqi(self):
        self.communicate_field(TF_QI)

    def communicate_ei(self):
        self.communicate_field(TF_EI)

    def communicate_qu_lods(self):  # mod.rs:436-468
        c = self.config
        d_n = self.get_d_n()
        dim = SET_VALUES[c.velocity_set][0]

        def get_offset(depth):
            return sum((1 << i) ** dim for i in range(0, depth + 1))

        if d_n > 1:
            for d in self.domains:
                d.read_lods()
            for d in range(d_n):
                x, y, z = domain_coords(d, c.d_x, c.d_y)
                offset = self.domains[d].g.n_lod_own
                for dc in range(d_n):
                    if d != dc:
                        dx, dy, dz = domain_coords(dc, c.d_x, c.d_y)
                        dist = max(abs(z - dz), abs(y - dy), abs(x - dx))
                        depth = max(0, c.mhd_lod_depth - dist)
                        rs, re = get_offset(depth - 1), get_offset(depth)
                        self.domains[d].qu_lod[offset * 4:(offset + re - rs) * 4] = \
                            self.domains[dc].transfer_lod_host[rs * 4:re * 4]
                        offset += re - rs

    # mesh.rs:233-279
    def import_mesh(self, data: bytes, scale, ox, oy, oz, rx, ry, rz):
        rot = Mesh.rotation_matrix(f32(f32(f32(rx) * f32(math.pi)) / f32(180.0)),
                                   f32(f32(f32(ry) * f32(math.pi)) / f32(180.0)),
                                   f32(f32(f32(rz) * f32(math.pi)) / f32(180.0)))
        scale_lu = self.config.units.len_si_lu(scale)
        self.meshes.append(Mesh.read_stl_raw(data, False, (1.0, 1.0, 1.0), (ox, oy, oz), rot, -abs(float(scale_lu))))

    def import_mesh_reposition(self, data: bytes, cx, cy, cz, rx, ry, rz, size):
        c = self.config
        rot = Mesh.rotation_matrix(f32(f32(f32(rx) * f32(math.pi)) / f32(180.0)),
                                   f32(f32(f32(ry) * f32(math.pi)) / f32(180.0)),
                                   f32(f32(f32(rz) * f32(math.pi)) / f32(180.0)))
        self.meshes.append(Mesh.read_stl_raw(data, True, (c.n_x, c.n_y, c.n_z), (cx, cy, cz), rot, size))

    def voxelise_mesh(self, index, ctype, value=None):
        for d in self.domains:
            d.voxelize_mesh_on_device(self.meshes[index], ctype, value)

    def precompute_B(self):  # mod.rs:284-297
        for d in self.domains:
            d.enqueue_precompute_b()

    def precompute_E(self):
        for d in self.domains:
            d.enqueue_precompute_e()

    def precompute_E_ECR(self):  # mod.rs:319-331
        for d in self.domains:
            d.enqueue_precompute_e_ecr()
