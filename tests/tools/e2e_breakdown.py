#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its time: host->device upload of the four state sections from pinned memory,
Lbm::initialize, Lbm::do_time_step, device->host download.  One JSON line.  usage: python tests/tools/e2e_breakdown.py"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import argparse
    import torch
    import bench
    from ionsolver_b200 import capi
    args = argparse.Namespace(lod_depth=4, gpus=1)
    lbm = bench.build_scene(args, 0, 1, 0)
    dom = lbm.domains[0]
    n = dom.n
    host = {name: torch.empty(sz, dtype=dt, pin_memory=True) for name, sz, dt in
            (("flags", n, torch.uint8), ("rho", n, torch.float32), ("u", 3 * n, torch.float32), ("q", n, torch.float32))}
    ids = {"flags": 3, "rho": 1, "u": 2, "q": 11}
    for name, t in host.items():
        t.numpy()[:] = dom.read(ids[name])
    lib = capi.load()

    def io(fn):
        for name, t in host.items():
            capi.check(fn(dom.handle, ids[name], ctypes.c_void_p(t.data_ptr()), 0, t.numel() * t.element_size()))

    out = {}
    for rep in range(3):
        lbm.finish_queues()
        t0 = time.perf_counter(); io(lib.ion_buffer_write); lbm.finish_queues()
        t1 = time.perf_counter(); lbm.initialize(); lbm.finish_queues()
        t2 = time.perf_counter(); lbm.do_time_step(); lbm.finish_queues()
        t3 = time.perf_counter(); io(lib.ion_buffer_read); lbm.finish_queues()
        t4 = time.perf_counter()
        out = {"upload_ms": (t1 - t0) * 1e3, "initialize_ms": (t2 - t1) * 1e3, "step_ms": (t3 - t2) * 1e3, "download_ms": (t4 - t3) * 1e3}
    nbytes = sum(t.numel() * t.element_size() for t in host.values())
    out["bytes_each_way"] = nbytes
    out["h2d_gbs"] = nbytes / out["upload_ms"] / 1e6
    out["d2h_gbs"] = nbytes / out["download_ms"] / 1e6
    print(json.dumps(out))
    lbm.close()


if __name__ == "__main__":
    main()
