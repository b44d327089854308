"""C-ABI surface: libionsolver_b200.so loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

from ionsolver_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"ION_API\s+[^;()]*?\b(ion_\w+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = capi.load()
    names = declared("ionsolver_b200.h") + declared("ionsolver_b200_host.h")
    assert len(names) > 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"symbols declared in include/ but not exported: {missing}"


def test_binding_lists_match_headers():
    assert sorted(capi.SYMBOLS) == declared("ionsolver_b200.h")
    assert sorted(capi.HOST_SYMBOLS) == declared("ionsolver_b200_host.h")


def test_abi_version_and_struct_layout():
    lib = capi.load()
    assert lib.ion_abi_version() == capi.ION_ABI_VERSION
    # IonParams: 15 u32/i32 + 11 floats + 3 u32 = 29 four-byte members
    assert ctypes.sizeof(capi.IonParams) == 29 * 4
    assert ctypes.sizeof(capi.IonLbmConfig) == 104
    c = capi.IonLbmConfig()
    lib.ion_lbm_config_default(c)  # LbmConfig::new, mod.rs:103-135: the last member proves the layout
    assert (c.velocity_set, c.float_type, c.mhd_lod_depth, c.graphics_active, c.run_steps, c.d_z) == (0, 0, 4, 1, 0, 1)


def test_no_cpu_fallback_without_device():
    """Without a usable sm_100 device, creating a domain must fail loudly (ION_ERR_NO_DEVICE), not compute on the CPU."""
    try:
        n = capi.device_count()
    except capi.IonError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present; the failure path is exercised on the CPU box")
    from ionsolver_b200 import lbm as L
    cfg = L.LbmConfig(velocity_set=L.VelocitySet.D3Q19, float_type=L.FloatType.FP32, n_x=8, n_y=8, n_z=8)
    with pytest.raises(capi.IonError) as e:
        L.Lbm(cfg)
    assert e.value.code in (capi.ION_ERR_NO_DEVICE, 100, 35, 38)  # no device / cudaErrorNoDevice / insufficient driver
    p = cfg.make_params(0)
    h = ctypes.c_void_p()
    rc = capi.load().ion_domain_create(ctypes.byref(p), 0, ctypes.byref(h))
    assert rc != 0 and not h.value


def test_argument_validation_needs_no_gpu():
    lib = capi.load()
    assert lib.ion_domain_create(None, 0, None) == capi.ION_ERR_INVALID
    assert b"NULL" in lib.ion_last_error_string()
    assert lib.ion_finish(None) == capi.ION_ERR_INVALID
    assert lib.ion_enqueue_stream_collide(None, 0, 0.0, 0.0, 0.0) == capi.ION_ERR_INVALID
    assert lib.ion_lbm_initialize(None) == capi.ION_ERR_INVALID
