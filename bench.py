#!/usr/bin/env python
"""bench.py -- cells-advanced/sec (fp64) of the IAMR hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one NavierStokes::advance (predict_velocity -> MAC projection -> velocity and
scalar advection -> Crank-Nicolson tensor diffusion -> nodal projection) of the Taylor-Green
problem of Tutorials/TaylorGreen/inputs.3d.taylorgreen on a 256^3 box per GPU
(BASELINE.json configs[1] at N=1; weak scaling: one 256^3 box per rank).  Prints ONE JSON line.

  value  : whole-job cells/s with the state resident in HBM (CUDA events, max over ranks)
  e2e    : the same through the host-buffer entry iamrx_ns_step_host (pinned host state in,
           new state out, both copies inside the timed region)
  roofline : the ABec GSRB smoother colour pass (the dominant kernel), algorithmic bytes /
           CUDA-event time per launch, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the CPU oracle ("port": restatement of the IAMR path, not the AMReX build --
           that cannot be built here) on the host cores, bounded sample
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cells-advanced/sec (fp64)"
UNIT = "cells/s"
NU, CFL = 1.0e-4, 0.7
TG = [1.0, 1.0, 0.0, 1.0, 1.0]   # prob.a, b, c, velocity_factor, density_ic (inputs.3d.taylorgreen:107-109)
# weak scaling decomposes the domain into z slabs, one 256^3 box per rank: every box then spans the periodic domain in x
# and y (neighbours wrap inside the kernels, no ghost traffic) and exchanges contiguous z planes with two neighbours
RANK_GRID = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 1, 4), 8: (1, 1, 8)}


def ncu_traffic(kernel, box):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json)."""
    try:
        e = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]
        return e["dram_bytes_per_launch"] if e["box"] == box else None, e["source"]
    except Exception:  # noqa: BLE001
        return None, None


# kernel class -> entry of profiles/ncu_traffic.json (ncu --set full summaries of the shipping kernels)
NCU_KEY = {"nodal_gs": "gs_sweep_kernel", "compute_aofs": "aofs_tile_kernel", "extrap_vel": "ev_kernels",
           "abec_apply": "apply2_kernel", "nodal_adotx": "adotx_march_kernel"}


def ncu_entry(key):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(key)
    except Exception:  # noqa: BLE001
        return None


def fp64_peak():
    """Measured DFMA-pipe rate (profiles/fp64_peak.json, written by scripts/fp64_peak.py on the B200), else None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))
    except Exception:  # noqa: BLE001
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


BLOCK_GRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def domain_for(nranks, nbox, decomp="slabs"):
    gx, gy, gz = (RANK_GRID if decomp == "slabs" else BLOCK_GRID)[nranks]
    ncell = (gx * nbox, gy * nbox, gz * nbox)
    boxes, owners = [], []
    r = 0
    for k in range(gz):
        for j in range(gy):
            for i in range(gx):
                lo = (i * nbox, j * nbox, k * nbox)
                boxes.append((lo, tuple(l + nbox - 1 for l in lo)))
                owners.append(r)
                r += 1
    prob_hi = tuple(float(g) for g in (gx, gy, gz))   # dx = 1/nbox in every direction for every N
    return ncell, boxes, owners, prob_hi


# ---------------------------------------------------------------------------
# CPU arm: the oracle timed on the host cores (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def cpu_oracle_run(n, steps, warmup, budget_s=None):
    """TaylorGreen n^3 on the CPU oracle with ALL host threads (set explicitly: torchrun exports OMP_NUM_THREADS=1).
    With a time budget the number of timed steps actually run is reduced (never below 1) so that the run ends in a few
    minutes; the count is returned and reported."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    orc.lib().orc_set_num_threads(host_threads())
    t_start = time.perf_counter()
    o = orc.OracleNS((n, n, n), visc_coef=NU, cfl=CFL)
    o.init_prob(11, TG)
    o.post_init()
    t_step = None
    wdone = 0
    for _ in range(warmup):
        t0 = time.perf_counter()
        o.step()
        t_step = time.perf_counter() - t0
        wdone += 1
        if budget_s is not None and (time.perf_counter() - t_start) + 2 * t_step > budget_s:
            break   # keep room for at least one timed step
    if budget_s is not None and t_step is not None:
        left = budget_s - (time.perf_counter() - t_start)
        steps = max(1, min(steps, int(left / t_step)))
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step()
    dt = time.perf_counter() - t0
    cores = orc.lib().orc_num_threads()
    o.close()
    return n ** 3 * steps / dt, dt / steps * 1e3, cores, steps, wdone


def workload_config(nbox, world, ncell, nboxes, decomp="slabs"):
    par = f"one {nbox}^3 box per rank x{world}"
    if world > 1:
        par += (", z slabs: x/y wrapped in-kernel, z ghost planes over NCCL, coarse multigrid levels replicated" if decomp == "slabs"
                else ", 3-D block decomposition: ghost shells over NCCL per smoother colour, coarse multigrid levels replicated")
    return {"workload": f"TaylorGreen 3D {nbox}^3 per GPU single-level (Tutorials/TaylorGreen/inputs.3d.taylorgreen, "
                        f"nu=1e-4, cfl=0.7, periodic), BASELINE.json configs[1]" + ("" if world == 1 else " weak-scaled"),
            "n_cell": list(ncell), "boxes": nboxes, "box": nbox, "parallelism": par,
            "l2": "working set per kernel >> 126 MB L2 (one fp64 cell array = 134 MB)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_n if args.cpu_n > 0 else args.n    # the SAME box as the GPU arm (256^3) unless overridden
    budget = float(os.environ.get("IAMRX_REF_BUDGET_S", "300"))
    val, ms, cores, steps_run, warm_run = cpu_oracle_run(n, args.steps, args.warmup, budget)
    sample = (f"TaylorGreen {n}^3 single box = the GPU arm's per-GPU workload; {steps_run} timed steps (asked {args.steps}) after init + "
              f"{warm_run} warm-up (asked {args.warmup}), bounded to ~{budget:.0f} s of host time; OpenMP on {cores} host threads; "
              f"CPU restatement of the IAMR path (oracle/), not the AMReX build: AMReX/AMReX-Hydro are not vendored in the reference")
    world = args.gpus
    ncell, boxes, _, _ = domain_for(world, args.n, args.decomp)
    cfg = workload_config(args.n, world, ncell, len(boxes), args.decomp)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_run,
        "warmup": warm_run, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def verify_multirank(lib, ix, dev, rank, world, decomp):
    """Untimed parity leg of the N>1 run: a z-DEPENDENT variable-density field (probtype 100, c = 1) advanced on the
    bench decomposition (one box per rank) and on the single-rank layout (the same boxes, all owned by rank 0); the two
    must agree to rounding.  A wrong ghost plane / gather / reduction shows up here."""
    import torch
    import torch.distributed as dist
    nb = 32
    ncell, boxes, owners, prob_hi = domain_for(world, nb, decomp)
    g = ix.Geom.make(ncell, (0.0, 0.0, 0.0), prob_hi)
    pp = [1.0, 1.0, 1.0, 1.0, 1.0]
    kw = dict(visc_coef=1e-3, cfl=0.7, gravity=-0.5)
    states, dts = [], []
    for own in (owners, [0] * len(boxes)):
        lev = ix.Level(lib, g, boxes, own)
        ns = ix.NavierStokes(lib, lev, dev, **kw)
        ns.init_prob(100, pp)
        d = [ns.post_init()] + [ns.step() for _ in range(2)]
        torch.cuda.synchronize()
        dts.append(d)
        states.append([ns.field(0, il).clone() for il in range(lev.num_local())])
        ns.close(); lev.close()
    mine = states[0][0].contiguous()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    err = 0.0
    if rank == 0:
        for r in range(world):
            err = max(err, float((gathered[r] - states[1][r]).abs().max()))
    return {"linf_state_vs_single_rank_layout": err, "field": "probtype 100 (z-dependent velocity and density), 2 steps after post_init",
            "n_cell": list(ncell), "dt_equal": bool(max(abs(a - b) for a, b in zip(dts[0], dts[1])) <= 1e-12 * dts[1][0])}


def hit_forcedata(L, nmodes=4, seed=111397, array_size=33):
    """A TurbulentForcing::forcedata table (17 x 33^3) with the structure TurbulentForcing_def.H:141-230 builds for
    inputs.3d.forced (turb.nmodes = 4, mode_start 0, div_free_force, spectrum_type 2, moderate_zero_modes); the random
    numbers come from numpy, not DepRand -- the table is an INPUT of the forcing kernel."""
    import numpy as np
    rng = np.random.default_rng(seed)
    fd = np.zeros((17, array_size, array_size, array_size))
    lmin = min(L)
    step = [int(l / lmin + 0.5) for l in L]
    kmax = nmodes / lmin + 1e-8

    def fill(kx, ky, kz):
        kappa = math.sqrt((kx / L[0]) ** 2 + (ky / L[1]) ** 2 + (kz / L[2]) ** 2)
        if kappa > kmax:
            return
        fd[0, kz, ky, kx] = (1.0 + rng.random()) * 2 * math.pi
        fd[1:5, kz, ky, kx] = rng.random(4) * 2 * math.pi
        fd[8:17, kz, ky, kx] = rng.random(9) * 2 * math.pi
        th, ph = rng.random() * 2 * math.pi, rng.random() * math.pi
        p = np.array([math.cos(th) * math.sin(ph), math.sin(th) * math.sin(ph), math.cos(ph)])
        if kappa < 1e-6:
            return
        e = 1.0 / kappa ** 3
        for kk in (kx, ky, kz):
            if kk == 0:
                e /= 2.0
        fd[5:8, kz, ky, kx] = p * e / float(p @ p)
    for kz in range(0, nmodes * step[2] + 1, step[2]):
        for ky in range(0, nmodes * step[1] + 1, step[1]):
            for kx in range(0, nmodes * step[0] + 1, step[0]):
                fill(kx, ky, kz)
    for kz in range(1, step[2]):
        for ky in range(0, nmodes * step[1] + 1):
            for kx in range(0, nmodes * step[0] + 1):
                fill(kx, ky, kz)
    return fd


def bind_to_gpu_numa_node(torch, local):
    """Multi-rank runs: pin this rank's host threads (and hence its first-touch pinned staging buffers) to the NUMA node its GPU
    hangs off, so that the e2e leg's host<->device copies of all ranks do not funnel through one socket.  Best effort: returns
    the node or None."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import iamr_b200 as ix

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; iamr_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    lib = ix.load()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = C.create_string_buffer(128)
            lib.check(lib.iamrx_comm_unique_id(buf))
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.to(dev)
        dist.broadcast(uid, 0)
        lib.check(lib.iamrx_comm_init(rank, world, bytes(uid.cpu().numpy().tobytes())))
    if world not in RANK_GRID:
        raise SystemExit(f"unsupported --gpus {world}")

    verify = verify_multirank(lib, ix, dev, rank, world, args.decomp) if (world > 1 and not args.no_verify) else None

    nbox = args.n
    hit = args.problem == "hit"
    rt = args.problem == "rt"
    bs = 1 if args.bottom_solver == "bicgstab" else 0
    dsl = args.problem == "dsl2d"
    if dsl:
        # BASELINE.json configs[2] on ONE level: DoubleShearLayer 2-D 1024^2 (inputs.2d.double_shear_layer-rotate: [-1,1]^2 periodic,
        # inviscid, cfl 0.5), run as a two-layer 3-D problem (tests/test_twod.py): n = (1024, 1024, 2), z extent = x extent
        if world != 1:
            raise SystemExit("--problem dsl2d is a single-GPU configuration")
        ncell = (4 * nbox, 4 * nbox, 2)
        boxes, owners = [((0, 0, 0), (ncell[0] - 1, ncell[1] - 1, 1))], [0]
        g = ix.Geom.make(ncell, (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))
    elif rt:
        # BASELINE.json configs[3] geometry on ONE level: RayleighTaylor 256^2 x 512 (regtest.3d.rayleightaylor: periodic x / y, slip
        # walls in z, gravity -1, inviscid, rho 1 -> 2), the fixed domain split into N z slabs (strong scaling: N must divide 512)
        ncell = (nbox, nbox, 2 * nbox)
        nzs = ncell[2] // world
        boxes = [((0, 0, r * nzs), (nbox - 1, nbox - 1, (r + 1) * nzs - 1)) for r in range(world)]
        owners = list(range(world))
        g = ix.Geom.make(ncell, (0.0, 0.0, 0.0), (0.5, 0.5, 1.0), periodic=(1, 1, 0))
    else:
        ncell, boxes, owners, prob_hi = domain_for(world, nbox, args.decomp)
        g = ix.Geom.make(ncell, (0.0, 0.0, 0.0), prob_hi)
    lev = ix.Level(lib, g, boxes, owners)
    if dsl:
        # mac_tol / proj_tol 1e-10: at 1024 cells per direction the fp64 floor of the residual (1.4e-12 of the right-hand side here,
        # profiles/r02_notes.md) lies above IAMR's default 1e-12, as for HIT 512^3
        ns = ix.NavierStokes(lib, lev, dev, visc_coef=0.0, cfl=0.5, mac_tol=1e-10, proj_tol=1e-10, bottom_solver=bs)
        ns.init_prob(5, [1.0, 1.0, 0.0, 0.0, 0.0, 0.4])
    elif rt:
        ns = ix.NavierStokes(lib, lev, dev, lo_bc=(0, 0, 4), hi_bc=(0, 0, 4), visc_coef=0.0, cfl=CFL, gravity=-1.0, bottom_solver=bs)
        ns.init_prob(10, [1.0, 2.0, 1.0, 0.0, 0.01, 0.005])
    elif hit:   # BASELINE.json configs[4]: Tutorials/HIT initial field, nu = 1e-4, proj_tol 1e-10 (inputs.3d.forced:129), synthetic variable density.
        # mac_tol 1e-10: at 512^3 the fp64 round-off floor of the MAC residual is eps / (h k)^2 ~ 4e-12 of the right-hand side for
        # domain-scale modes (profiles/r02_notes.md: residual stalls at 3.7e-12 relative), above IAMR's default 1e-12 -- the
        # reference's MLMG would abort the same way; a site running this case sets mac_proj.mac_tol as here.
        ns = ix.NavierStokes(lib, lev, dev, visc_coef=NU, cfl=CFL, proj_tol=1e-10, mac_tol=1e-10, bottom_solver=bs)
        if not args.no_forcing:   # USE_TURBULENT_FORCING = TRUE (Tutorials/HIT/GNUmakefile), turb.nmodes = 4 (inputs.3d.forced:110)
            ns.set_turbulent_forcing(4, 0, 1, hit_forcedata(prob_hi))
        ns.init_prob(20, [1.0, 1.0, 0.5])
    else:
        ns = ix.NavierStokes(lib, lev, dev, visc_coef=NU, cfl=CFL, bottom_solver=bs)
        ns.init_prob(11, TG)
    ns.post_init()
    cells_total = ncell[0] * ncell[1] * (1 if dsl else ncell[2])   # dsl2d: 2-D cells (the second layer is a copy)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident steps: the headline pass runs WITHOUT per-launch event profiling ----------------------------
    for _ in range(args.warmup):
        ns.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.iamrx_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.cudart().cudaProfilerStart()   # ncu --profile-from-start off profiles only the timed steps
    e0.record()
    iters = []
    for _ in range(args.steps):
        ns.step()
        iters.append(ns.last_iters())
    e1.record()
    barrier()
    torch.cuda.cudart().cudaProfilerStop()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.iamrx_launch_count()
    clocks = sampler.stop() if rank == 0 else None
    value = cells_total * args.steps / (ms * 1e-3)

    # ---- roofline pass: a few MORE steps of the same run with CUDA events around every launch of the tracked kernel
    #      classes on the two finest multigrid levels (on the launching stream) --------------------------------------------
    lib.iamrx_prof_reset()
    lib.iamrx_prof_enable(1, (nbox // 2) ** 3)
    barrier()
    for _ in range(args.prof_steps):
        ns.step()
    barrier()
    lib.iamrx_prof_enable(0, 0)
    prof = {}
    for name, k in (("abec_gsrb", 0), ("nodal_gs", 1), ("compute_aofs", 2), ("extrap_vel", 3), ("abec_apply", 4), ("nodal_adotx", 5)):
        t, nl, by = C.c_double(), C.c_int64(), C.c_double()
        lib.check(lib.iamrx_prof_report(k, C.byref(t), C.byref(nl), C.byref(by)))
        prof[name] = (t.value, nl.value, by.value)
    lib.iamrx_prof_reset()

    # ---- end-to-end steps through the host-buffer entry ---------------------------
    nloc = lev.num_local()
    mine = [b for b, o in zip(boxes, owners) if o == rank]
    shp = [(5,) + tuple(hi[d] - lo[d] + 1 for d in (2, 1, 0)) for lo, hi in mine]
    e2e_value, state_bytes, checksum = None, sum(8 * math.prod(q) for q in shp), None
    if args.e2e_steps > 0:
        hin = [torch.empty(q, dtype=torch.float64).pin_memory() for q in shp]
        hout = [torch.empty(q, dtype=torch.float64).pin_memory() for q in shp]
        for il in range(nloc):
            hin[il].copy_(ns.field(0, il))
        ns.step_host(hin, hout)      # warm-up of the staging path
        hin, hout = hout, hin
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ns.step_host(hin, hout)  # synchronous: returns when the new state is in host memory
            hin, hout = hout, hin
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e_value = cells_total * args.e2e_steps / e2e_s
        checksum = float(hin[0][0].abs().max())

    if args.kernel_table:   # diagnostic: per-kernel device time of ONE extra step (rank 0, stderr); not part of any reported number
        barrier()
        lib.iamrx_prof_all(1)
        t0 = time.perf_counter()
        ns.step()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        lib.iamrx_prof_all(0)
        buf = C.create_string_buffer(1 << 16)
        lib.iamrx_prof_dump(buf, len(buf))
        if rank == 0:
            rows = []
            for ln in buf.value.decode().splitlines():
                name, cnt, kms = ln.rsplit(' ', 2)
                rows.append((float(kms), int(cnt), name.replace(' ', '')))
            rows.sort(reverse=True)
            tot = sum(r[0] for r in rows)
            print(f"# kernel table, 1 step, {world} rank(s): wall {wall_ms:.2f} ms, sum of kernel times on rank 0 {tot:.2f} ms, "
                  f"launches {sum(r[1] for r in rows)}", file=sys.stderr)
            for kms, cnt, name in rows[:40]:
                print(f"{name:28s} {cnt:7d} {kms:10.3f} ms {100 * kms / tot:5.1f}%", file=sys.stderr)

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic, traffic_src = ncu_traffic("gsrb_kernel", nbox)
        t, nl, by = prof["abec_gsrb"]
        achieved = (by / 1e9) / (t * 1e-3) if t > 0 else 0.0
        fp64 = fp64_peak()
        secondary = {}
        for name in ("nodal_gs", "compute_aofs", "extrap_vel", "abec_apply", "nodal_adotx"):
            tt, nn, bb = prof[name]
            if tt > 0:
                secondary[name] = {"achieved_gbs": (bb / 1e9) / (tt * 1e-3), "frac": (bb / 1e9) / (tt * 1e-3) / peak,
                                   "launches": nn, "ms_total": tt, "ms_per_step": tt / max(1, args.prof_steps)}
                e = ncu_entry(NCU_KEY[name])
                if e:   # committed ncu --set full summary of the same kernel: DRAM traffic and, for the Godunov kernels, the
                    secondary[name]["ncu"] = e   # FP64-pipe utilisation (their second roofline; fp64_peak = measured DFMA rate)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cn = args.cpu_n if args.cpu_n > 0 else 128
            cpu_val, cpu_ms, cores, srun, wrun = cpu_oracle_run(cn, 2, 1)
            cpu = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"TaylorGreen {cn}^3 (a bounded sample of the 256^3 workload; --impl reference runs 256^3): {srun} timed steps after init + {wrun} warm-up of "
                             f"the CPU oracle (restatement of the IAMR path; the AMReX build cannot be produced offline), "
                             f"OpenMP on {cores} host threads"}
        cfg = workload_config(nbox, world, ncell, len(boxes), args.decomp)
        if hit:
            cfg["workload"] = (f"HIT 3D {nbox}^3 per GPU single-level variable-density (Tutorials/HIT/prob_init.cpp:100-131 field, nu=1e-4, "
                               f"proj_tol 1e-10, mac_tol 1e-10 (fp64 floor at 512^3), rho = 1 + 0.5 sin sin sin, turbulent forcing {'off' if args.no_forcing else 'on: turb.nmodes 4, div-free, synthetic mode table'}), BASELINE.json configs[4]" + ("" if world == 1 else " weak-scaled"))
        if dsl:
            cfg["workload"] = (f"DoubleShearLayer 2D {ncell[0]}^2 SINGLE level (BASELINE.json configs[2] without the fine level; "
                               f"inputs.2d.double_shear_layer-rotate: periodic, inviscid, cfl 0.5; mac_tol = proj_tol = 1e-10), run as a two-layer 3-D box "
                               f"{ncell[0]}x{ncell[1]}x2 (z-uniform; value counts 2-D cells)")
            cfg["parallelism"] = "one box; multigrid semi-coarsens x / y only"
        if rt:
            cfg["workload"] = (f"RayleighTaylor 3D {ncell[0]}x{ncell[1]}x{ncell[2]} SINGLE level (BASELINE.json configs[3] geometry without the fine level: "
                               f"periodic x/y, slip walls z, gravity, inviscid, rho 1->2), fixed domain in {world} z slab(s)")
            cfg["parallelism"] = f"{world} z slab(s) of {ncell[2] // world} planes; x/y wrapped in-kernel, z ghost planes over NCCL, walls mirrored in the kernels"
        cfg["mg_iters_last_step"] = {"mac": iters[-1][0], "visc": iters[-1][1], "nodal": iters[-1][2]}
        cfg["bottom_solver"] = args.bottom_solver
        if numa is not None:
            cfg["host_numa_node_rank0"] = numa
        cfg["timing"] = "headline pass without per-launch events; roofline from a separate pass of %d steps" % args.prof_steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if rt else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes,
                    "steps": args.e2e_steps, "api": "iamrx_ns_step_host (pinned host state in, new state out)",
                    "checksum_max_abs_u": checksum},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "gsrb_kernel (ABec red-black colour pass, finest two MG levels)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_note": f"DRAM bytes per finest-level launch ({nbox}^3, algorithmic {48 * nbox ** 3}); {traffic_src}",
                         "peak_source": peak_src, "launches_timed": nl, "ms_total": t, "ms_per_step": t / max(1, args.prof_steps),
                         "algorithmic_bytes": "per colour pass and cell: 48 B (a=0), 56 with a variable alpha (SURVEY.md 8d); 24 (+8 with alpha) when the face "
                                              "coefficients are constants taken from the kernel arguments (viscous solves); 8 less for the first pass of a "
                                              "correction (zero initial guess: phi is not read).  achieved = sum of these per launch / sum of launch times",
                         "fp64_peak": fp64, "other_kernels": secondary},
            "clocks": clocks,
        }
        if verify is not None:
            line["verify"] = verify
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    ns.close(); lev.close()
    if world > 1:
        lib.iamrx_comm_finalize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=256, help="box size per GPU")
    ap.add_argument("--cpu-n", type=int, default=0, help="box size of the CPU runs (0: --impl reference uses --n, the cpu_baseline leg 128)")
    ap.add_argument("--decomp", default="slabs", choices=["slabs", "blocks"], help="weak-scaling decomposition: z slabs or 3-D blocks")
    ap.add_argument("--no-forcing", action="store_true", help="--problem hit without the tutorial's turbulent forcing")
    ap.add_argument("--bottom-solver", default="smoother", choices=["smoother", "bicgstab"],
                    help="multigrid bottom solver: smoother sweeps (default) or BiCGStab (IAMR's bicgcg)")
    ap.add_argument("--problem", default="tg", choices=["tg", "hit", "rt", "dsl2d"],
                    help="tg: BASELINE configs[1]; hit: configs[4] (HIT, variable density); rt: configs[3] geometry on one level (walls, gravity; strong-scaled); dsl2d: configs[2] on one level (2-D as two layers)")
    ap.add_argument("--prof-steps", type=int, default=3, help="steps of the separate roofline pass")
    ap.add_argument("--no-verify", action="store_true", help="skip the untimed multi-rank parity leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-table", action="store_true", help="diagnostic per-kernel table of one extra step on stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
