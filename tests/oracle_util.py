"""Helpers shared by the oracle tests and the GPU parity tests."""
import hashlib

import numpy as np


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def same_bits(x, y):
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    return x.dtype == y.dtype and x.shape == y.shape and x.tobytes() == y.tobytes()


def buffer_names(cfg):
    names = ["fi", "rho", "u", "flags"]
    if cfg.d_x * cfg.d_y * cfg.d_z > 1:
        names += ["transfer_p", "transfer_m"]
    if cfg.ext_magneto_hydro:
        names += ["ei", "fqi", "qc", "e_dyn", "b_dyn", "qu_lod"]
    if cfg.ext_subgrid_ecr:
        names += ["eti", "et"]
    return names


def check_against_golden(lbm, cfg, gold_domains, where):
    bad = []
    for d, g in zip(lbm.domains, gold_domains):
        for n in buffer_names(cfg):
            if sha(getattr(d, n)) != g[n]["sha256"]:
                bad.append(f"{where}: domain {d.g.d_i} buffer {n}")
    return bad


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))
