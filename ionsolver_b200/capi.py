"""ctypes binding of include/ionsolver_b200.h (libionsolver_b200.so).

This is the raw C-ABI layer: what the reference's Rust host would bind with `extern "C"` in place of the
`ocl`/`ocl-macros` crates.  It never falls back to a CPU implementation: if the CUDA
library is missing, `load()` raises; if there is no sm_100 device, ion_domain_create fails.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ION_LIB") or os.path.join(_HERE, "libionsolver_b200.so")  # ION_LIB: A/B builds of the same library

ION_ABI_VERSION = 1
# enum IonVelocitySet / IonRelaxationTime / IonFloatType (wire values of src/lbm/types.rs)
D2Q9, D3Q15, D3Q19, D3Q27 = 0, 1, 2, 3
SRT, TRT = 0, 1
FP16S, FP16C, FP32 = 0, 1, 2
VELOCITY_SETS = {"D2Q9": D2Q9, "D3Q15": D3Q15, "D3Q19": D3Q19, "D3Q27": D3Q27}
RELAXATION_TIMES = {"SRT": SRT, "TRT": TRT}
FLOAT_TYPES = {"FP16S": FP16S, "FP16C": FP16C, "FP32": FP32}
SET_VALUES = {D2Q9: (2, 9, 3), D3Q15: (3, 15, 5), D3Q19: (3, 19, 5), D3Q27: (3, 27, 9)}
# enum IonTransferField
TRANSFER_FI, TRANSFER_RHO_U_FLAGS, TRANSFER_EI, TRANSFER_QI = 0, 1, 2, 3
# enum IonExt
EXT_EQUILIBRIUM_BOUNDARIES, EXT_VOLUME_FORCE, EXT_FORCE_FIELD = 1, 2, 4
EXT_MAGNETO_HYDRO, EXT_SUBGRID_ECR, EXT_UPDATE_FIELDS, EXT_DETERMINISTIC = 8, 16, 32, 64
# enum IonField
(FIELD_FI, FIELD_RHO, FIELD_U, FIELD_FLAGS, FIELD_F, FIELD_E_STAT, FIELD_B_STAT, FIELD_E_DYN, FIELD_B_DYN, FIELD_FQI,
 FIELD_EI, FIELD_Q, FIELD_QU_LOD, FIELD_E_VAR, FIELD_ETI, FIELD_ET, FIELD_TRANSFER_P, FIELD_TRANSFER_M) = range(18)
FIELD_COUNT = 18
FIELD_NAMES = ["fi", "rho", "u", "flags", "f", "e_stat", "b_stat", "e_dyn", "b_dyn", "fqi", "ei", "q", "qu_lod", "e_var",
               "eti", "et", "transfer_p", "transfer_m"]
ION_ERR_INVALID, ION_ERR_UNSUPPORTED, ION_ERR_NO_DEVICE, ION_ERR_ABSENT, ION_ERR_RANGE = 10001, 10002, 10003, 10004, 10005

# every symbol include/ionsolver_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "ion_device_count", "ion_last_error_string", "ion_abi_version", "ion_domain_create", "ion_domain_destroy",
    "ion_domain_params", "ion_buffer_size", "ion_buffer_write", "ion_buffer_read", "ion_buffer_device_ptr",
    "ion_buffer_copy", "ion_enqueue_initialize", "ion_enqueue_stream_collide", "ion_enqueue_update_fields",
    "ion_enqueue_stream_collide_range", "ion_buffer_swap", "ion_enqueue_update_e_b_dyn", "ion_domain_eb_fft_info", "ion_domain_set_precompute_mode", "ion_enqueue_lod_part_2_gather", "ion_enqueue_clear_qu_lod",
    "ion_enqueue_transfer_extract", "ion_enqueue_transfer_insert", "ion_voxelize_mesh", "ion_enqueue_precompute_b",
    "ion_enqueue_precompute_e", "ion_enqueue_precompute_e_ecr", "ion_domain_set_ecr_freq", "ion_finish",
    "ion_kernel_launch_count", "ion_domain_stream",
    "ion_exchange_transfer", "ion_copy_lods", "ion_comm_unique_id", "ion_comm_create", "ion_comm_destroy",
    "ion_comm_exchange_transfer", "ion_comm_exchange_lods", "ion_neighbor_domains", "ion_lod_exchange_plan",
    "ion_measure_fma_peak", "ion_halo_fork", "ion_halo_join", "ion_codec_probe", "ion_read_slice",
]
# every symbol include/ionsolver_b200_host.h declares
HOST_SYMBOLS = [
    "ion_lbm_config_default", "ion_units_set", "ion_units_eval", "ion_lbm_make_params", "ion_lbm_create",
    "ion_lbm_create_distributed", "ion_lbm_destroy", "ion_lbm_get_config", "ion_lbm_local_domains", "ion_lbm_domain",
    "ion_lbm_initialize", "ion_lbm_run", "ion_lbm_do_time_step", "ion_lbm_finish_queues", "ion_lbm_get_time_step",
    "ion_lbm_precompute_b", "ion_lbm_precompute_e", "ion_lbm_precompute_e_ecr", "ion_lbm_communicate_field",
    "ion_lbm_communicate_qu_lods", "ion_lbm_set_time_step", "ion_lbm_import_mesh", "ion_lbm_import_mesh_reposition",
    "ion_lbm_voxelise_mesh", "ion_lbm_mesh_info", "ion_lbm_mesh_triangles", "ion_lbm_mesh_translate",
    "ion_lbm_set_taylor_green", "ion_lbm_setup_velocity_field", "ion_setup_taylor_green", "ion_setup_lid_driven_cavity",
    "ion_setup_charged_fluid", "ion_setup_scene", "ion_lbm_encode", "ion_lbm_decode", "ion_lbm_write_file", "ion_lbm_read_file",
    "ion_config_to_json", "ion_config_from_json", "ion_lbm_dump_cell", "ion_lbm_read_slice", "ion_lbm_write_slice_png",
    "ion_iron_colormap", "ion_write_png_rgb", "ion_free",
]
COMM_ID_BYTES = 128


class IonParams(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_uint32),
        ("nx", ctypes.c_uint32), ("ny", ctypes.c_uint32), ("nz", ctypes.c_uint32),
        ("dx", ctypes.c_uint32), ("dy", ctypes.c_uint32), ("dz", ctypes.c_uint32),
        ("di", ctypes.c_uint32),
        ("ox", ctypes.c_int32), ("oy", ctypes.c_int32), ("oz", ctypes.c_int32),
        ("velocity_set", ctypes.c_uint32), ("relaxation_time", ctypes.c_uint32), ("float_type", ctypes.c_uint32),
        ("ext", ctypes.c_uint32),
        ("w", ctypes.c_float),
        ("ke", ctypes.c_float), ("kmu", ctypes.c_float), ("kmu0", ctypes.c_float), ("kkge", ctypes.c_float),
        ("kimg", ctypes.c_float), ("kvev", ctypes.c_float), ("kme", ctypes.c_float),
        ("wq", ctypes.c_float),
        ("kkbme", ctypes.c_float), ("keabs", ctypes.c_float),
        ("lod_depth", ctypes.c_uint32), ("n_lod", ctypes.c_uint32), ("n_lod_own", ctypes.c_uint32),
    ]


class IonLbmConfig(ctypes.Structure):
    """IonLbmConfig of include/ionsolver_b200_host.h (LbmConfig + Units, mod.rs:46-101, units.rs:14-27)."""
    _fields_ = [
        ("velocity_set", ctypes.c_uint32), ("relaxation_time", ctypes.c_uint32), ("float_type", ctypes.c_uint32),
        ("unit_m", ctypes.c_float), ("unit_kg", ctypes.c_float), ("unit_s", ctypes.c_float), ("unit_a", ctypes.c_float),
        ("unit_k", ctypes.c_float),
        ("propellant", ctypes.c_uint32),
        ("n_x", ctypes.c_uint32), ("n_y", ctypes.c_uint32), ("n_z", ctypes.c_uint32),
        ("d_x", ctypes.c_uint32), ("d_y", ctypes.c_uint32), ("d_z", ctypes.c_uint32),
        ("nu", ctypes.c_float),
        ("f_x", ctypes.c_float), ("f_y", ctypes.c_float), ("f_z", ctypes.c_float),
        ("ext_equilibrium_boudaries", ctypes.c_uint8), ("ext_volume_force", ctypes.c_uint8),
        ("ext_force_field", ctypes.c_uint8), ("ext_magneto_hydro", ctypes.c_uint8), ("ext_subgrid_ecr", ctypes.c_uint8),
        ("mhd_lod_depth", ctypes.c_uint8), ("graphics_active", ctypes.c_uint8), ("deterministic", ctypes.c_uint8),
        ("ecr_freq", ctypes.c_float), ("ecr_field_strength", ctypes.c_float),
        ("run_steps", ctypes.c_uint64),
    ]


class IonError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ionsolver_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> ctypes.CDLL:
    """Load libionsolver_b200.so (fails loudly when the CUDA extension has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    c = ctypes
    D = c.c_void_p
    L.ion_device_count.argtypes = [c.POINTER(c.c_int)]
    L.ion_last_error_string.restype = c.c_char_p
    L.ion_abi_version.restype = c.c_uint32
    L.ion_domain_create.argtypes = [c.POINTER(IonParams), c.c_int, c.POINTER(D)]
    L.ion_domain_destroy.argtypes = [D]
    L.ion_domain_params.argtypes = [D, c.POINTER(IonParams)]
    L.ion_buffer_size.argtypes = [D, c.c_int, c.POINTER(c.c_size_t)]
    L.ion_buffer_write.argtypes = [D, c.c_int, c.c_void_p, c.c_size_t, c.c_size_t]
    L.ion_buffer_read.argtypes = [D, c.c_int, c.c_void_p, c.c_size_t, c.c_size_t]
    L.ion_buffer_device_ptr.argtypes = [D, c.c_int, c.POINTER(c.c_void_p)]
    L.ion_buffer_copy.argtypes = [D, c.c_int, c.c_size_t, D, c.c_int, c.c_size_t, c.c_size_t]
    L.ion_enqueue_initialize.argtypes = [D]
    L.ion_enqueue_stream_collide.argtypes = [D, c.c_uint64, c.c_float, c.c_float, c.c_float]
    L.ion_enqueue_update_fields.argtypes = [D, c.c_uint64, c.c_float, c.c_float, c.c_float]
    L.ion_buffer_swap.argtypes = [D, c.c_int, c.POINTER(c.c_int), c.POINTER(c.c_void_p), c.POINTER(c.c_void_p), c.POINTER(c.c_size_t)]
    L.ion_enqueue_stream_collide_range.argtypes = [D, c.c_uint64, c.c_float, c.c_float, c.c_float, c.c_uint32, c.c_uint32, c.c_int]
    L.ion_enqueue_update_e_b_dyn.argtypes = [D]
    L.ion_enqueue_lod_part_2_gather.argtypes = [D]
    L.ion_domain_eb_fft_info.argtypes = [D, c.POINTER(c.c_uint64), c.POINTER(c.c_uint32)]
    L.ion_domain_set_precompute_mode.argtypes = [D, c.c_int]
    L.ion_enqueue_clear_qu_lod.argtypes = [D]
    L.ion_enqueue_transfer_extract.argtypes = [D, c.c_int, c.c_uint32, c.c_uint64]
    L.ion_enqueue_transfer_insert.argtypes = [D, c.c_int, c.c_uint32, c.c_uint64]
    L.ion_voxelize_mesh.argtypes = [D, c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint32, c.c_void_p, c.c_uint32, c.c_uint8,
                                    c.c_float, c.c_float, c.c_float, c.c_uint64]
    L.ion_enqueue_precompute_b.argtypes = [D]
    L.ion_enqueue_precompute_e.argtypes = [D]
    L.ion_enqueue_precompute_e_ecr.argtypes = [D]
    L.ion_domain_set_ecr_freq.argtypes = [D, c.c_float]
    L.ion_finish.argtypes = [D]
    L.ion_kernel_launch_count.restype = c.c_uint64
    L.ion_domain_stream.argtypes = [D, c.POINTER(c.c_void_p)]
    L.ion_exchange_transfer.argtypes = [D, D, c.c_size_t]
    L.ion_copy_lods.argtypes = [D, c.c_uint32, D, c.c_uint32, c.c_uint32]
    L.ion_comm_unique_id.argtypes = [c.c_void_p]
    L.ion_comm_create.argtypes = [c.c_void_p, c.c_int, c.c_int, c.c_int, c.POINTER(c.c_void_p)]
    L.ion_comm_destroy.argtypes = [c.c_void_p]
    L.ion_comm_exchange_transfer.argtypes = [c.c_void_p, D, c.c_int, c.c_int, c.c_size_t]
    L.ion_comm_exchange_lods.argtypes = [c.c_void_p, D]
    U32P = c.POINTER(c.c_uint32)
    L.ion_neighbor_domains.argtypes = [c.c_uint32, c.c_uint32, c.c_uint32, c.c_uint32, c.c_uint32, U32P, U32P]
    L.ion_lod_exchange_plan.argtypes = [c.POINTER(IonParams), c.c_uint32, U32P, U32P, U32P]
    L.ion_halo_fork.argtypes = [D]
    L.ion_halo_join.argtypes = [D]
    L.ion_codec_probe.argtypes = [c.c_int, c.c_int, c.c_int, c.c_void_p, c.c_void_p, c.c_uint64]
    L.ion_measure_fma_peak.argtypes = [c.c_int, c.c_int, c.POINTER(c.c_double)]
    # ---- host layer (include/ionsolver_b200_host.h) ----
    CFG = c.POINTER(IonLbmConfig)
    H = c.c_void_p
    PI = c.POINTER(c.c_int)
    L.ion_lbm_config_default.argtypes = [CFG]
    L.ion_lbm_config_default.restype = None
    L.ion_units_set.argtypes = [CFG] + [c.c_float] * 10
    L.ion_units_set.restype = None
    L.ion_units_eval.argtypes = [CFG, c.c_int, c.c_float]
    L.ion_units_eval.restype = c.c_float
    L.ion_lbm_make_params.argtypes = [CFG, c.c_uint32, c.POINTER(IonParams)]
    L.ion_lbm_create.argtypes = [CFG, PI, c.c_int, c.POINTER(H)]
    L.ion_lbm_create_distributed.argtypes = [CFG, c.c_int, c.c_int, c.c_int, c.c_void_p, c.POINTER(H)]
    L.ion_lbm_destroy.argtypes = [H]
    L.ion_lbm_get_config.argtypes = [H, CFG]
    L.ion_lbm_local_domains.argtypes = [H, c.POINTER(c.c_uint32)]
    L.ion_lbm_domain.argtypes = [H, c.c_uint32, c.POINTER(D), c.POINTER(c.c_uint32)]
    for n in ("ion_lbm_initialize", "ion_lbm_do_time_step", "ion_lbm_finish_queues", "ion_lbm_precompute_b",
              "ion_lbm_precompute_e", "ion_lbm_precompute_e_ecr", "ion_lbm_communicate_qu_lods"):
        getattr(L, n).argtypes = [H]
    L.ion_lbm_run.argtypes = [H, c.c_uint64]
    L.ion_lbm_get_time_step.argtypes = [H, c.POINTER(c.c_uint64)]
    L.ion_lbm_set_time_step.argtypes = [H, c.c_uint64]
    L.ion_lbm_communicate_field.argtypes = [H, c.c_int]
    L.ion_lbm_import_mesh.argtypes = [H, c.c_char_p] + [c.c_float] * 7
    L.ion_lbm_import_mesh_reposition.argtypes = [H, c.c_char_p] + [c.c_float] * 7
    L.ion_lbm_voxelise_mesh.argtypes = [H, c.c_uint32, c.c_int, c.c_float, c.c_float, c.c_float]
    L.ion_lbm_mesh_info.argtypes = [H, c.c_uint32, c.POINTER(c.c_uint32), c.c_void_p, c.c_void_p]
    L.ion_lbm_mesh_triangles.argtypes = [H, c.c_uint32, c.c_void_p, c.c_void_p, c.c_void_p]
    L.ion_lbm_mesh_translate.argtypes = [H, c.c_uint32, c.c_float, c.c_float, c.c_float]
    L.ion_lbm_set_taylor_green.argtypes = [H, c.c_uint32]
    L.ion_lbm_setup_velocity_field.argtypes = [H, c.c_float, c.c_float, c.c_float, c.c_float]
    L.ion_setup_taylor_green.argtypes = [c.c_uint32, c.c_uint32, c.c_int, c.c_int, c.c_int, PI, c.c_int, c.POINTER(H)]
    L.ion_setup_lid_driven_cavity.argtypes = [c.c_uint32, PI, c.c_int, c.POINTER(H)]
    L.ion_setup_charged_fluid.argtypes = [c.c_uint32, c.c_uint32, c.c_uint32, c.c_int, c.c_int, c.c_uint32, c.c_char_p, PI,
                                          c.c_int, c.POINTER(H)]
    L.ion_setup_scene.argtypes = [c.c_char_p, c.c_char_p, c.c_float, c.c_uint32, PI, c.c_int, c.POINTER(H)]
    L.ion_lbm_encode.argtypes = [H, c.c_int, c.POINTER(c.c_void_p), c.POINTER(c.c_size_t)]
    L.ion_lbm_decode.argtypes = [c.c_void_p, c.c_size_t, CFG, c.c_int, PI, c.c_int, c.POINTER(H)]
    L.ion_lbm_write_file.argtypes = [H, c.c_char_p]
    L.ion_lbm_read_file.argtypes = [c.c_char_p, CFG, c.POINTER(H)]
    L.ion_config_to_json.argtypes = [CFG, c.POINTER(c.c_void_p)]
    L.ion_config_from_json.argtypes = [c.c_char_p, CFG]
    L.ion_lbm_dump_cell.argtypes = [H, c.c_uint32, c.c_uint64, c.POINTER(c.c_void_p)]
    L.ion_lbm_read_slice.argtypes = [H, c.c_int, c.c_int, c.c_uint32, c.c_uint32, c.c_void_p, c.c_size_t, c.POINTER(c.c_uint32), c.POINTER(c.c_uint32)]
    L.ion_lbm_write_slice_png.argtypes = [H, c.c_int, c.c_int, c.c_uint32, c.c_uint32, c.c_float, c.c_float, c.c_char_p]
    L.ion_iron_colormap.argtypes = [c.c_float]
    L.ion_iron_colormap.restype = c.c_uint32
    L.ion_write_png_rgb.argtypes = [c.c_char_p, c.c_void_p, c.c_uint32, c.c_uint32]
    L.ion_read_slice.argtypes = [D, c.c_int, c.c_int, c.c_uint32, c.c_uint32, c.c_void_p]
    L.ion_free.argtypes = [c.c_void_p]
    L.ion_free.restype = None
    _lib = L
    return L


def check(code: int):
    if code != 0:
        raise IonError(code, load().ion_last_error_string().decode(errors="replace"))


def device_count() -> int:
    n = ctypes.c_int(0)
    check(load().ion_device_count(ctypes.byref(n)))
    return n.value


def kernel_launch_count() -> int:
    return int(load().ion_kernel_launch_count())


class Domain:
    """Thin RAII wrapper around ion_domain_t* with numpy-based buffer access."""

    def __init__(self, params: IonParams = None, device: int = 0, borrowed=None):
        self.lib = load()
        self.owned = borrowed is None
        if borrowed is None:
            self.params = params
            self.handle = ctypes.c_void_p()
            check(self.lib.ion_domain_create(ctypes.byref(params), device, ctypes.byref(self.handle)))
        else:  # &lbm.domains[i]: owned by the Lbm
            self.handle = borrowed
            self.params = IonParams()
            check(self.lib.ion_domain_params(self.handle, ctypes.byref(self.params)))
        self.ddf_dtype = np.float32 if self.params.float_type == FP32 else np.uint16
        self.n = self.params.nx * self.params.ny * self.params.nz

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            if self.owned:
                self.lib.ion_domain_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dtype_of(self, field):
        if field in (FIELD_FI, FIELD_FQI, FIELD_EI, FIELD_ETI):
            return self.ddf_dtype
        if field in (FIELD_FLAGS, FIELD_TRANSFER_P, FIELD_TRANSFER_M):
            return np.uint8
        return np.float32

    def size(self, field) -> int:
        n = ctypes.c_size_t(0)
        check(self.lib.ion_buffer_size(self.handle, field, ctypes.byref(n)))
        return n.value

    def write(self, field, array, offset_bytes=0):
        a = np.ascontiguousarray(array)
        check(self.lib.ion_buffer_write(self.handle, field, a.ctypes.data, offset_bytes, a.nbytes))

    def read(self, field, offset_bytes=0, nbytes=None, dtype=None):
        dtype = np.dtype(dtype or self.dtype_of(field))
        if nbytes is None:
            nbytes = self.size(field) - offset_bytes
        out = np.empty(nbytes // dtype.itemsize, dtype)
        check(self.lib.ion_buffer_read(self.handle, field, out.ctypes.data, offset_bytes, out.nbytes))
        return out

    def device_ptr(self, field) -> int:
        p = ctypes.c_void_p()
        check(self.lib.ion_buffer_device_ptr(self.handle, field, ctypes.byref(p)))
        return p.value or 0

    def stream(self) -> int:
        p = ctypes.c_void_p()
        check(self.lib.ion_domain_stream(self.handle, ctypes.byref(p)))
        return p.value or 0

    def enqueue_initialize(self):
        check(self.lib.ion_enqueue_initialize(self.handle))

    def enqueue_stream_collide(self, t, fx=0.0, fy=0.0, fz=0.0):
        check(self.lib.ion_enqueue_stream_collide(self.handle, t, fx, fy, fz))

    def enqueue_stream_collide_range(self, t, z_begin, z_end, finish, fx=0.0, fy=0.0, fz=0.0):
        """stream_collide on the z layers [z_begin, z_end); finish = last range of the step (folds the LOD deposits)."""
        check(self.lib.ion_enqueue_stream_collide_range(self.handle, t, fx, fy, fz, z_begin, z_end, 1 if finish else 0))

    def enqueue_update_fields(self, t, fx=0.0, fy=0.0, fz=0.0):
        check(self.lib.ion_enqueue_update_fields(self.handle, t, fx, fy, fz))

    def enqueue_update_e_b_dyn(self):
        check(self.lib.ion_enqueue_update_e_b_dyn(self.handle))

    def set_precompute_mode(self, mode):
        """0 = reference arithmetic and order (default), 1 = fast psi_from_mesh (rounding-level differences), 2 = psi / static E as
        FFT convolutions (cuFFT transforms; ~1e-5 relative L2)."""
        check(self.lib.ion_domain_set_precompute_mode(self.handle, int(mode)))

    def eb_fft_info(self):
        """(bytes of static kernel spectra, polyphase problems per step) of the FFT field update; (0, 0) = direct kernels."""
        b, t = ctypes.c_uint64(0), ctypes.c_uint32(0)
        check(self.lib.ion_domain_eb_fft_info(self.handle, ctypes.byref(b), ctypes.byref(t)))
        return int(b.value), int(t.value)

    def enqueue_lod_part_2_gather(self):
        check(self.lib.ion_enqueue_lod_part_2_gather(self.handle))

    def enqueue_clear_qu_lod(self):
        check(self.lib.ion_enqueue_clear_qu_lod(self.handle))

    def enqueue_transfer_extract(self, field, direction, t):
        check(self.lib.ion_enqueue_transfer_extract(self.handle, field, direction, t))

    def enqueue_transfer_insert(self, field, direction, t):
        check(self.lib.ion_enqueue_transfer_insert(self.handle, field, direction, t))

    def voxelize_mesh(self, p0, p1, p2, bbu, direction, flag, mpc, t):
        p0 = np.ascontiguousarray(p0, np.float32).reshape(-1)
        p1 = np.ascontiguousarray(p1, np.float32).reshape(-1)
        p2 = np.ascontiguousarray(p2, np.float32).reshape(-1)
        bbu = np.ascontiguousarray(bbu, np.float32)
        check(self.lib.ion_voxelize_mesh(self.handle, p0.ctypes.data, p1.ctypes.data, p2.ctypes.data, p0.size // 3,
                                         bbu.ctypes.data, direction, flag, mpc[0], mpc[1], mpc[2], t))

    def enqueue_precompute_b(self):
        check(self.lib.ion_enqueue_precompute_b(self.handle))

    def enqueue_precompute_e(self):
        check(self.lib.ion_enqueue_precompute_e(self.handle))

    def finish(self):
        check(self.lib.ion_finish(self.handle))


def neighbor_domains(d_x, d_y, d_z, d, axis):
    """(dp, dm): ring neighbours of domain d along axis (ion_neighbor_domains; mod.rs:386-404).  Needs no GPU."""
    dp, dm = ctypes.c_uint32(), ctypes.c_uint32()
    check(load().ion_neighbor_domains(d_x, d_y, d_z, d, axis, ctypes.byref(dp), ctypes.byref(dm)))
    return dp.value, dm.value


def lod_exchange_plan(params: IonParams, dc):
    """(src_entry, entries, dst_entry) of foreign domain dc for the domain described by params (mod.rs:448-465)."""
    a, b, c_ = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    check(load().ion_lod_exchange_plan(ctypes.byref(params), dc, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c_)))
    return a.value, b.value, c_.value


def measure_fma_peak(device=0, packed=False) -> float:
    """FP32 FMA/s of the device measured with scalar FFMA (or fma.rn.f32x2): compute roofline for update_e_b_dynamic."""
    v = ctypes.c_double()
    check(load().ion_measure_fma_peak(device, int(packed), ctypes.byref(v)))
    return v.value


def codec_probe(float_type, arr, direction, device=0):
    """DDF storage codec on the device: direction 0 float32 -> stored words, 1 stored words -> float32."""
    import numpy as np
    if direction == 0:
        a = np.ascontiguousarray(arr, np.float32)
        out = np.empty(a.size, np.float32 if float_type == 2 else np.uint16)
    else:
        a = np.ascontiguousarray(arr, np.float32 if float_type == 2 else np.uint16)
        out = np.empty(a.size, np.float32)
    check(load().ion_codec_probe(device, float_type, direction, a.ctypes.data, out.ctypes.data, a.size))
    return out
