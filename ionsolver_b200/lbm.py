"""Python face of the host layer: `LbmConfig`, `Units`, `Lbm`, `LbmDomain` with the names, fields and method
meaning of the reference's Rust host (/root/reference/src/lbm/mod.rs, domain.rs, units.rs, types.rs, mesh.rs,
setup.rs, file.rs).  Everything here is a thin ctypes veneer over include/ionsolver_b200_host.h; the host logic
itself is the C++ in ionsolver_b200/csrc/host/, and all compute is CUDA behind include/ionsolver_b200.h.  There is no
CPU fallback: importing works without a GPU (so configs, params, JSON and file parsing can be tested), but creating
an `Lbm` needs an sm_100 device.
"""
from __future__ import annotations

import ctypes
import dataclasses

import numpy as np

from . import capi
from .capi import IonLbmConfig, check


class VelocitySet:  # types.rs:17-25
    D2Q9, D3Q15, D3Q19, D3Q27 = 0, 1, 2, 3


class RelaxationTime:  # types.rs:55-60
    Srt, Trt = 0, 1


class FloatType:  # types.rs:76-82
    FP16S, FP16C, FP32 = 0, 1, 2

    @staticmethod
    def size_of(ft):
        return 4 if ft == FloatType.FP32 else 2


class TransferField:  # types.rs:105-111
    Fi, RhoUFlags, Ei, Qi = 0, 1, 2, 3


class ModelType:  # mesh.rs:9-14
    Solid, Magnet, Charged, ChargedECR = 0, 1, 2, 3


_UNIT_FNS = {
    "len_si_lu": 0, "nu_si_lu": 1, "charge_si_lu": 2, "mag_flux_si_lu": 3, "e_field_si_lu": 4, "magnetization_si_lu": 5,
    "time_lu_si": 6, "time_si_lu": 7, "speed_si_lu": 8, "len_lu_si": 9, "speed_lu_si": 10, "charge_lu_si": 11,
    "mag_flux_lu_si": 12, "e_field_lu_si": 13,
    "epsilon_0_lu": 32, "ke_lu": 33, "mu_0_lu": 34, "kkge_lu": 35, "kimg_lu": 36, "kveV_lu": 37, "kkBme_lu": 38,
    "keabs_lu": 39, "kme_lu": 40,
}


@dataclasses.dataclass
class Units:
    """units.rs:14-27; conversions are evaluated by the C++ host (f32 arithmetic of units.rs)."""
    m: float = 1.0
    kg: float = 1.0
    s: float = 1.0
    a: float = 1.0
    k: float = 1.0
    prop: int = 0

    def _c(self):
        c = IonLbmConfig()
        capi.load().ion_lbm_config_default(c)
        c.unit_m, c.unit_kg, c.unit_s, c.unit_a, c.unit_k, c.propellant = self.m, self.kg, self.s, self.a, self.k, self.prop
        return c

    def set(self, lbm_length, lbm_velocity, lbm_rho, lbm_charge, lbm_temp, si_length, si_velocity, si_rho, si_charge,
            si_temp):
        c = self._c()
        capi.load().ion_units_set(c, lbm_length, lbm_velocity, lbm_rho, lbm_charge, lbm_temp, si_length, si_velocity,
                                  si_rho, si_charge, si_temp)
        self.m, self.kg, self.s, self.a, self.k = c.unit_m, c.unit_kg, c.unit_s, c.unit_a, c.unit_k

    def __getattr__(self, name):
        if name in _UNIT_FNS:
            fn = _UNIT_FNS[name]
            return lambda v=0.0: float(capi.load().ion_units_eval(self._c(), fn, v))
        raise AttributeError(name)


@dataclasses.dataclass
class GraphicsConfig:
    graphics_active: bool = True  # graphics.rs:160


@dataclasses.dataclass
class LbmConfig:
    """mod.rs:46-135 (same field names, same defaults)."""
    velocity_set: int = VelocitySet.D2Q9
    relaxation_time: int = RelaxationTime.Srt
    float_type: int = FloatType.FP16S
    units: Units = dataclasses.field(default_factory=Units)
    n_x: int = 1
    n_y: int = 1
    n_z: int = 1
    d_x: int = 1
    d_y: int = 1
    d_z: int = 1
    nu: float = float(np.float32(1.0) / np.float32(6.0))
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
    ecr_field_strength: float = 0.0
    graphics_config: GraphicsConfig = dataclasses.field(default_factory=GraphicsConfig)
    run_steps: int = 0
    deterministic: bool = False  # not in the reference: reproducible, reference-ordered E/B path (ION_EXT_DETERMINISTIC)

    def to_c(self) -> IonLbmConfig:
        c = IonLbmConfig()
        c.velocity_set, c.relaxation_time, c.float_type = self.velocity_set, self.relaxation_time, self.float_type
        u = self.units
        c.unit_m, c.unit_kg, c.unit_s, c.unit_a, c.unit_k, c.propellant = u.m, u.kg, u.s, u.a, u.k, u.prop
        c.n_x, c.n_y, c.n_z, c.d_x, c.d_y, c.d_z = self.n_x, self.n_y, self.n_z, self.d_x, self.d_y, self.d_z
        c.nu, c.f_x, c.f_y, c.f_z = self.nu, self.f_x, self.f_y, self.f_z
        c.ext_equilibrium_boudaries = int(self.ext_equilibrium_boudaries)
        c.ext_volume_force = int(self.ext_volume_force)
        c.ext_force_field = int(self.ext_force_field)
        c.ext_magneto_hydro = int(self.ext_magneto_hydro)
        c.ext_subgrid_ecr = int(self.ext_subgrid_ecr)
        c.mhd_lod_depth = self.mhd_lod_depth
        c.graphics_active = int(self.graphics_config.graphics_active)
        c.deterministic = int(self.deterministic)
        c.ecr_freq, c.ecr_field_strength, c.run_steps = self.ecr_freq, self.ecr_field_strength, self.run_steps
        return c

    @staticmethod
    def from_c(c: IonLbmConfig) -> "LbmConfig":
        return LbmConfig(
            velocity_set=c.velocity_set, relaxation_time=c.relaxation_time, float_type=c.float_type,
            units=Units(c.unit_m, c.unit_kg, c.unit_s, c.unit_a, c.unit_k, c.propellant),
            n_x=c.n_x, n_y=c.n_y, n_z=c.n_z, d_x=c.d_x, d_y=c.d_y, d_z=c.d_z, nu=c.nu, f_x=c.f_x, f_y=c.f_y, f_z=c.f_z,
            ext_equilibrium_boudaries=bool(c.ext_equilibrium_boudaries), ext_volume_force=bool(c.ext_volume_force),
            ext_force_field=bool(c.ext_force_field), ext_magneto_hydro=bool(c.ext_magneto_hydro),
            ext_subgrid_ecr=bool(c.ext_subgrid_ecr), mhd_lod_depth=c.mhd_lod_depth, ecr_freq=c.ecr_freq,
            ecr_field_strength=c.ecr_field_strength, graphics_config=GraphicsConfig(bool(c.graphics_active)),
            run_steps=c.run_steps, deterministic=bool(c.deterministic))

    def make_params(self, d=0) -> capi.IonParams:
        """get_device_defines (domain.rs:736-858) for domain d; needs no GPU."""
        p = capi.IonParams()
        check(capi.load().ion_lbm_make_params(self.to_c(), d, ctypes.byref(p)))
        return p

    def to_json(self) -> str:  # file::write_config, file.rs:323
        out = ctypes.c_void_p()
        check(capi.load().ion_config_to_json(self.to_c(), ctypes.byref(out)))
        try:
            return ctypes.string_at(out).decode()
        finally:
            capi.load().ion_free(out)

    @staticmethod
    def from_json(text: str) -> "LbmConfig":  # file::read_config, file.rs:310
        c = IonLbmConfig()
        check(capi.load().ion_config_from_json(text.encode(), c))
        return LbmConfig.from_c(c)


class LbmDomain(capi.Domain):
    """&lbm.domains[i]: buffer access and per-domain enqueue_* (domain.rs:412-578) on a borrowed handle."""

    def __init__(self, handle, d_i):
        super().__init__(borrowed=handle)
        self.d_i = d_i
        self.n_x, self.n_y, self.n_z = self.params.nx, self.params.ny, self.params.nz
        self.o_x, self.o_y, self.o_z = self.params.ox, self.params.oy, self.params.oz
        self.n_lod, self.n_lod_own = self.params.n_lod, self.params.n_lod_own


def _devs(devices):
    if not devices:
        return None, 0
    arr = (ctypes.c_int * len(devices))(*devices)
    return arr, len(devices)


class Lbm:
    """mod.rs:152-495."""

    def __init__(self, config: LbmConfig = None, devices=None, _handle=None):
        self.lib = capi.load()
        self.handle = ctypes.c_void_p()
        if _handle is not None:
            self.handle = _handle
        else:
            arr, n = _devs(devices)
            check(self.lib.ion_lbm_create(config.to_c(), arr, n, ctypes.byref(self.handle)))
        self._refresh()

    def _refresh(self):
        c = IonLbmConfig()
        check(self.lib.ion_lbm_get_config(self.handle, c))
        self.config = LbmConfig.from_c(c)
        cnt = ctypes.c_uint32()
        check(self.lib.ion_lbm_local_domains(self.handle, ctypes.byref(cnt)))
        self.domains = []
        for i in range(cnt.value):
            h, di = ctypes.c_void_p(), ctypes.c_uint32()
            check(self.lib.ion_lbm_domain(self.handle, i, ctypes.byref(h), ctypes.byref(di)))
            self.domains.append(LbmDomain(h, di.value))

    @classmethod
    def new_distributed(cls, config: LbmConfig, rank, world, device, comm_id: bytes):
        """One process per GPU: this process owns domain `rank` (see bench.py for the torch.distributed bootstrap)."""
        lib = capi.load()
        h = ctypes.c_void_p()
        buf = (ctypes.c_uint8 * capi.COMM_ID_BYTES).from_buffer_copy(comm_id)
        check(lib.ion_lbm_create_distributed(config.to_c(), rank, world, device, buf, ctypes.byref(h)))
        return cls(_handle=h)

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (ctypes.c_uint8 * capi.COMM_ID_BYTES)()
        check(capi.load().ion_comm_unique_id(buf))
        return bytes(buf)

    # ---- scene constructors (setup.rs) ----
    @classmethod
    def setup_taylor_green(cls, n, d_z=1, velocity_set=VelocitySet.D3Q19, float_type=FloatType.FP32, graphics_active=False,
                           devices=None):
        h = ctypes.c_void_p()
        arr, nd = _devs(devices)
        check(capi.load().ion_setup_taylor_green(n, d_z, velocity_set, float_type, int(graphics_active), arr, nd, ctypes.byref(h)))
        return cls(_handle=h)

    @classmethod
    def setup_lid_driven_cavity(cls, n, devices=None):
        h = ctypes.c_void_p()
        arr, nd = _devs(devices)
        check(capi.load().ion_setup_lid_driven_cavity(n, arr, nd, ctypes.byref(h)))
        return cls(_handle=h)

    @classmethod
    def setup_charged_fluid(cls, nx, ny, nz, velocity_set=VelocitySet.D3Q19, float_type=FloatType.FP32, lod_depth=4,
                            magnet_stl=None, devices=None):
        h = ctypes.c_void_p()
        arr, nd = _devs(devices)
        check(capi.load().ion_setup_charged_fluid(nx, ny, nz, velocity_set, float_type, lod_depth,
                                                  magnet_stl.encode() if magnet_stl else None, arr, nd, ctypes.byref(h)))
        return cls(_handle=h)

    @classmethod
    def setup_scene(cls, name, stl_dir="stl", scale=1.0, subgrid_ecr=False, first_step=False, devices=None):
        """One of the reference's scene functions by name (setup.rs:22-64): setup_verification, setup_field_vis, setup_ecr_test,
        setup_mesh_test, setup_mesh_field_test, setup_deeva_test, setup_taylor_green, setup_domain_test."""
        h = ctypes.c_void_p()
        arr, nd = _devs(devices)
        flags = (1 if subgrid_ecr else 0) | (2 if first_step else 0)
        check(capi.load().ion_setup_scene(name.encode(), str(stl_dir).encode(), float(scale), flags, arr, nd, ctypes.byref(h)))
        return cls(_handle=h)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            for d in self.domains:
                d.handle = ctypes.c_void_p()
            self.lib.ion_lbm_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_d_n(self):
        return self.config.d_x * self.config.d_y * self.config.d_z

    def get_time_step(self):
        t = ctypes.c_uint64()
        check(self.lib.ion_lbm_get_time_step(self.handle, ctypes.byref(t)))
        return t.value

    def set_time_step(self, t):
        check(self.lib.ion_lbm_set_time_step(self.handle, t))

    def initialize(self):
        check(self.lib.ion_lbm_initialize(self.handle))

    def run(self, steps):
        check(self.lib.ion_lbm_run(self.handle, steps))

    def do_time_step(self):
        check(self.lib.ion_lbm_do_time_step(self.handle))

    def finish_queues(self):
        check(self.lib.ion_lbm_finish_queues(self.handle))

    def precompute_B(self):
        check(self.lib.ion_lbm_precompute_b(self.handle))

    def precompute_E(self):
        check(self.lib.ion_lbm_precompute_e(self.handle))

    def precompute_E_ECR(self):
        check(self.lib.ion_lbm_precompute_e_ecr(self.handle))

    def communicate_fi(self):
        check(self.lib.ion_lbm_communicate_field(self.handle, TransferField.Fi))

    def communicate_rho_u_flags(self):
        check(self.lib.ion_lbm_communicate_field(self.handle, TransferField.RhoUFlags))

    def communicate_ei(self):
        check(self.lib.ion_lbm_communicate_field(self.handle, TransferField.Ei))

    def communicate_fqi(self):
        check(self.lib.ion_lbm_communicate_field(self.handle, TransferField.Qi))

    def communicate_qu_lods(self):
        check(self.lib.ion_lbm_communicate_qu_lods(self.handle))

    def import_mesh(self, path, scale, ox, oy, oz, rx, ry, rz):
        check(self.lib.ion_lbm_import_mesh(self.handle, str(path).encode(), scale, ox, oy, oz, rx, ry, rz))

    def import_mesh_reposition(self, path, cx, cy, cz, rx, ry, rz, size):
        check(self.lib.ion_lbm_import_mesh_reposition(self.handle, str(path).encode(), cx, cy, cz, rx, ry, rz, size))

    def voxelise_mesh(self, index, model_type=ModelType.Solid, value=None):
        v = [0.0, 0.0, 0.0]
        if model_type == ModelType.Magnet:
            v = list(value)
        elif model_type in (ModelType.Charged, ModelType.ChargedECR):
            v[0] = float(value)
        check(self.lib.ion_lbm_voxelise_mesh(self.handle, index, model_type, v[0], v[1], v[2]))

    def mesh(self, index):
        tn = ctypes.c_uint32()
        pmin, pmax = np.zeros(3, np.float32), np.zeros(3, np.float32)
        check(self.lib.ion_lbm_mesh_info(self.handle, index, ctypes.byref(tn), pmin.ctypes.data, pmax.ctypes.data))
        p = [np.zeros((tn.value, 3), np.float32) for _ in range(3)]
        check(self.lib.ion_lbm_mesh_triangles(self.handle, index, p[0].ctypes.data, p[1].ctypes.data, p[2].ctypes.data))
        return {"triangle_number": tn.value, "p_min": pmin, "p_max": pmax, "p0": p[0], "p1": p[1], "p2": p[2]}

    def mesh_translate(self, index, t):
        check(self.lib.ion_lbm_mesh_translate(self.handle, index, t[0], t[1], t[2]))

    def set_taylor_green(self, periodicity=1):
        check(self.lib.ion_lbm_set_taylor_green(self.handle, periodicity))

    def setup_velocity_field(self, velocity, density):
        check(self.lib.ion_lbm_setup_velocity_field(self.handle, velocity[0], velocity[1], velocity[2], density))

    # ---- file.rs ----
    def encode(self, reference_compatible=False) -> bytes:
        data, ln = ctypes.c_void_p(), ctypes.c_size_t()
        check(self.lib.ion_lbm_encode(self.handle, int(reference_compatible), ctypes.byref(data), ctypes.byref(ln)))
        try:
            return ctypes.string_at(data, ln.value)
        finally:
            self.lib.ion_free(data)

    @classmethod
    def decode(cls, data: bytes, config: LbmConfig = None, reference_compatible=False, devices=None):
        c = (config or LbmConfig()).to_c()
        h = ctypes.c_void_p()
        arr, nd = _devs(devices)
        check(capi.load().ion_lbm_decode(data, len(data), c, int(reference_compatible), arr, nd, ctypes.byref(h)))
        return cls(_handle=h)

    def write(self, path):
        check(self.lib.ion_lbm_write_file(self.handle, str(path).encode()))

    @classmethod
    def read(cls, path, config: LbmConfig = None):
        c = (config or LbmConfig()).to_c()
        h = ctypes.c_void_p()
        check(capi.load().ion_lbm_read_file(str(path).encode(), c, ctypes.byref(h)))
        return cls(_handle=h)

    def read_slice(self, field, slice_mode, index, component=3):
        """One plane of a field over the whole lattice, halo layers removed (SliceMode 1,2,3 = X,Y,Z of graphics.rs:124-130;
        component 0,1,2 or 3 = magnitude for vector fields).  Returns a float32 array [height, width]:
        X -> [z, y], Y -> [z, x], Z -> [y, x]."""
        cfg = self.config
        w = cfg.n_y if slice_mode == 1 else cfg.n_x
        h = cfg.n_y if slice_mode == 3 else cfg.n_z
        out = np.empty(w * h, np.float32)
        cw, ch = ctypes.c_uint32(), ctypes.c_uint32()
        check(self.lib.ion_lbm_read_slice(self.handle, int(field), int(component), int(slice_mode), int(index), out.ctypes.data, out.size,
                                          ctypes.byref(cw), ctypes.byref(ch)))
        return out.reshape(ch.value, cw.value)

    def write_slice_png(self, path, field, slice_mode, index, v_min=0.0, v_max=1.0, component=3):
        """Colour-mapped slice as an RGB PNG, one pixel per cell (the reference saves rendered frames, graphics.rs:328-373)."""
        check(self.lib.ion_lbm_write_slice_png(self.handle, int(field), int(component), int(slice_mode), int(index), float(v_min),
                                               float(v_max), str(path).encode()))

    def dump_cell(self, local_index, cell) -> str:
        out = ctypes.c_void_p()
        check(self.lib.ion_lbm_dump_cell(self.handle, local_index, cell, ctypes.byref(out)))
        try:
            return ctypes.string_at(out).decode()
        finally:
            self.lib.ion_free(out)
