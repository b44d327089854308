"""Generates the synthetic binary STL meshes used by the tests and by bench.py (tests/golden/stl/*.stl).

/root/reference/stl/ does not exist on the GPU box and reference assets are not copied into this repository, so the
scenes of BASELINE.json configs 2-3 use procedurally generated stand-ins with the same shapes and SI dimensions as
the reference's thruster parts (a disk magnet of radius 0.05 m and thickness 0.01 m along y, a ring magnet, a tube and
two electrode plates; extents read off `stl/*.stl` with numpy).  Deterministic: re-running reproduces the files byte for
byte.  Run: python tests/golden/make_stl.py
"""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "stl")


def write_stl(path, tris, name):
    tris = np.asarray(tris, np.float32)
    with open(path, "wb") as f:
        f.write(name.encode().ljust(80, b" "))
        f.write(struct.pack("<I", len(tris)))
        for t in tris:
            n = np.cross(t[1] - t[0], t[2] - t[0])
            ln = np.linalg.norm(n)
            n = n / ln if ln > 0 else n
            f.write(struct.pack("<12fH", *n.astype(np.float32), *t.reshape(-1), 0))


def ring(r_out, r_in, y0, y1, seg):
    """Annulus (r_in > 0) or disk (r_in == 0) extruded along y, outward-facing triangles."""
    a = np.linspace(0.0, 2.0 * np.pi, seg, endpoint=False)
    co, si = np.cos(a), np.sin(a)
    tris = []

    def p(r, i, y):
        return [r * co[i % seg], y, r * si[i % seg]]

    for i in range(seg):
        j = i + 1
        tris += [[p(r_out, i, y0), p(r_out, i, y1), p(r_out, j, y1)], [p(r_out, i, y0), p(r_out, j, y1), p(r_out, j, y0)]]
        if r_in > 0:
            tris += [[p(r_in, i, y0), p(r_in, j, y1), p(r_in, i, y1)], [p(r_in, i, y0), p(r_in, j, y0), p(r_in, j, y1)]]
            tris += [[p(r_in, i, y1), p(r_in, j, y1), p(r_out, j, y1)], [p(r_in, i, y1), p(r_out, j, y1), p(r_out, i, y1)]]
            tris += [[p(r_in, i, y0), p(r_out, j, y0), p(r_in, j, y0)], [p(r_in, i, y0), p(r_out, i, y0), p(r_out, j, y0)]]
        else:
            tris += [[[0, y1, 0], p(r_out, j, y1), p(r_out, i, y1)], [[0, y0, 0], p(r_out, i, y0), p(r_out, j, y0)]]
    return tris


def box(lo, hi):
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    v = [[x0, y0, z0], [x1, y0, z0], [x1, y1, z0], [x0, y1, z0], [x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1]]
    q = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (0, 4, 7, 3)]
    tris = []
    for a, b, c, d in q:
        tris += [[v[a], v[b], v[c]], [v[a], v[c], v[d]]]
    return tris


def main():
    os.makedirs(OUT, exist_ok=True)
    write_stl(os.path.join(OUT, "disk_magnet.stl"), ring(0.05, 0.0, -0.005, 0.005, 64), "ionsolver_b200 synthetic disk magnet")
    write_stl(os.path.join(OUT, "ring_magnet.stl"), ring(0.05, 0.03, -0.005, 0.005, 32), "ionsolver_b200 synthetic ring magnet")
    write_stl(os.path.join(OUT, "tube.stl"), ring(0.025, 0.022, 0.03, 0.13, 32), "ionsolver_b200 synthetic quartz tube")
    write_stl(os.path.join(OUT, "plate1.stl"), box((-0.015, 0.045, 0.024), (0.015, 0.095, 0.028)), "ionsolver_b200 synthetic e-plate 1")
    write_stl(os.path.join(OUT, "plate2.stl"), box((-0.015, 0.045, -0.028), (0.015, 0.095, -0.024)), "ionsolver_b200 synthetic e-plate 2")
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
