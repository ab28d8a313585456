"""AMReX plotfile writer (iamrx_ns_write_plotfile): the directory layout and text syntax of an AMReX plotfile / VisMF
(Header, Level_0/Cell_H, Level_0/Cell_D_*), read back with an independent minimal parser and compared with the state."""
import os
import re

import numpy as np

import iamr_b200 as ix

BOX = r"\(\((-?\d+),(-?\d+),(-?\d+)\) \((-?\d+),(-?\d+),(-?\d+)\) \(0,0,0\)\)"


def read_plotfile(path):
    """Minimal reader of a single-level AMReX plotfile: returns (header dict, {box: array (ncomp, nz, ny, nx)})."""
    lines = open(os.path.join(path, "Header")).read().split("\n")
    it = iter(lines)
    hdr = {"type": next(it)}
    ncomp = int(next(it))
    hdr["names"] = [next(it) for _ in range(ncomp)]
    hdr["dim"] = int(next(it)); hdr["time"] = float(next(it)); hdr["finest"] = int(next(it))
    hdr["prob_lo"] = [float(x) for x in next(it).split()]
    hdr["prob_hi"] = [float(x) for x in next(it).split()]
    next(it)   # refinement ratios
    hdr["domain"] = tuple(int(x) for x in re.match(BOX, next(it).strip()).groups())
    hdr["steps"] = int(next(it).split()[0])
    hdr["dx"] = [float(x) for x in next(it).split()]
    hdr["coord"] = int(next(it)); hdr["bwidth"] = int(next(it))
    lev, ngrids, t = next(it).split()
    hdr["ngrids"] = int(ngrids)
    next(it)
    hdr["grid_loc"] = [[tuple(float(x) for x in next(it).split()) for _ in range(3)] for _ in range(hdr["ngrids"])]
    hdr["mf"] = next(it)
    # VisMF header
    vl = open(os.path.join(path, hdr["mf"] + "_H")).read().split("\n")
    vi = iter(vl)
    assert next(vi) == "1" and next(vi) == "0"
    assert int(next(vi)) == ncomp and next(vi) == "0"
    nb = int(re.match(r"\((\d+) 0", next(vi)).group(1))
    boxes = [tuple(int(x) for x in re.match(BOX, next(vi)).groups()) for _ in range(nb)]
    assert next(vi) == ")"
    assert int(next(vi)) == nb
    fod = []
    for _ in range(nb):
        m = re.match(r"FabOnDisk: (\S+) (\d+)", next(vi))
        fod.append((m.group(1), int(m.group(2))))
    next(vi)
    mm = []
    for _ in range(2):
        assert next(vi) == f"{nb},{ncomp}"
        mm.append([[float(x) for x in next(vi).rstrip(",").split(",")] for _ in range(nb)])
        next(vi)
    data = {}
    for b, (fname, off) in zip(boxes, fod):
        with open(os.path.join(path, os.path.dirname(hdr["mf"]), fname), "rb") as f:
            f.seek(off)
            head = f.readline().decode()
            m = re.match(r"FAB \(\(8, \(64 11 52 0 1 12 0 1023\)\),\(8, \(8 7 6 5 4 3 2 1\)\)\)" + BOX + r" (\d+)\n", head)
            assert m, head
            assert tuple(int(x) for x in m.groups()[:6]) == b and int(m.group(7)) == ncomp
            shape = (ncomp, b[5] - b[2] + 1, b[4] - b[1] + 1, b[3] - b[0] + 1)
            data[b] = np.frombuffer(f.read(8 * int(np.prod(shape))), dtype="<f8").reshape(shape)
    return hdr, data, mm


def test_plotfile_round_trip(backend, tmp_path):
    lib, dev = backend
    n = (16, 16, 8)
    boxes = [((0, 0, 0), (7, 15, 7)), ((8, 0, 0), (15, 15, 7))]
    lev = ix.Level(lib, ix.Geom.make(n, (0.0, 0.0, 0.0), (1.0, 1.0, 0.5)), boxes)
    ns = ix.NavierStokes(lib, lev, dev, visc_coef=1e-3, cfl=0.7)
    ns.init_prob(100, [1.0, 1.0, 1.0, 1.0, 1.0])
    ns.post_init()
    ns.step()
    path = tmp_path / "plt00001"
    ns.write_plotfile(path)
    assert sorted(os.listdir(path)) == ["Header", "Level_0", "job_info"]
    assert sorted(os.listdir(path / "Level_0")) == ["Cell_D_00000", "Cell_H"]
    hdr, data, (mn, mx) = read_plotfile(str(path))
    assert hdr["type"] == "NavierStokes-V1.1"            # NavierStokesBase.cpp:3349
    assert hdr["names"] == ["x_velocity", "y_velocity", "z_velocity", "density", "tracer", "gradpx", "gradpy", "gradpz"]
    assert hdr["dim"] == 3 and hdr["finest"] == 0 and hdr["steps"] == 1 and hdr["ngrids"] == 2 and hdr["mf"] == "Level_0/Cell"
    assert hdr["domain"] == (0, 0, 0, 15, 15, 7) and np.allclose(hdr["dx"], [1 / 16, 1 / 16, 1 / 16]) and hdr["time"] == ns.time
    assert hdr["prob_hi"] == [1.0, 1.0, 0.5] and hdr["grid_loc"][1][0] == (0.5, 1.0)
    for il, (lo, hi) in enumerate(boxes):
        b = lo + hi
        S = ns.field(0, il).cpu().numpy()
        G = ns.field(2, il).cpu().numpy()
        want = np.concatenate([S, G], axis=0)
        assert np.array_equal(data[b], want)           # raw fp64: bit exact
        assert np.allclose(mn[il], want.min(axis=(1, 2, 3)), rtol=1e-15) and np.allclose(mx[il], want.max(axis=(1, 2, 3)), rtol=1e-15)
    ns.close(); lev.close()
