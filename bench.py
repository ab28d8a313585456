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


def domain_for(nranks, nbox):
    gx, gy, gz = RANK_GRID[nranks]
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
def cpu_oracle_run(n, steps, warmup):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    o = orc.OracleNS((n, n, n), visc_coef=NU, cfl=CFL)
    o.init_prob(11, TG)
    o.post_init()
    for _ in range(warmup):
        o.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step()
    dt = time.perf_counter() - t0
    cores = orc.lib().orc_num_threads()
    o.close()
    return n ** 3 * steps / dt, dt / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_n
    val, ms, cores = cpu_oracle_run(n, args.steps, args.warmup)
    sample = (f"TaylorGreen {n}^3 single box (bounded sample of the 256^3 workload), {args.steps} timed steps after "
              f"init + {args.warmup} warm-up, OpenMP on {cores} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "TaylorGreen 3D single-level (inputs.3d.taylorgreen), CPU sample of the 256^3/GPU workload",
                   "n_cell": [n, n, n], "note": "CPU restatement of the IAMR path (oracle/), not the AMReX build: "
                   "AMReX/AMReX-Hydro are not vendored in the reference and cannot be built offline"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
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

    nbox = args.n
    ncell, boxes, owners, prob_hi = domain_for(world, nbox)
    g = ix.Geom.make(ncell, (0.0, 0.0, 0.0), prob_hi)
    lev = ix.Level(lib, g, boxes, owners)
    ns = ix.NavierStokes(lib, lev, dev, visc_coef=NU, cfl=CFL)
    ns.init_prob(11, TG)
    ns.post_init()
    cells_total = ncell[0] * ncell[1] * ncell[2]

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

    # ---- device-resident steps ------------------------------------------------
    for _ in range(args.warmup):
        ns.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.iamrx_launch_count_reset()
    lib.iamrx_prof_reset()
    lib.iamrx_prof_enable(1, (nbox // 2) ** 3)   # time launches on the two finest multigrid levels
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
    lib.iamrx_prof_enable(0, 0)
    prof = {}
    for name, k in (("abec_gsrb", 0), ("nodal_gs", 1), ("compute_aofs", 2), ("extrap_vel", 3), ("abec_apply", 4), ("nodal_adotx", 5)):
        t, nl, by = C.c_double(), C.c_int64(), C.c_double()
        lib.check(lib.iamrx_prof_report(k, C.byref(t), C.byref(nl), C.byref(by)))
        prof[name] = (t.value, nl.value, by.value)
    lib.iamrx_prof_reset()
    value = cells_total * args.steps / (ms * 1e-3)

    # ---- end-to-end steps through the host-buffer entry ---------------------------
    nloc = lev.num_local()
    e2e_value, state_bytes, checksum = None, 5 * nbox ** 3 * 8 * nloc, None
    if args.e2e_steps > 0:
        hin = [torch.empty((5, nbox, nbox, nbox), dtype=torch.float64).pin_memory() for _ in range(nloc)]
        hout = [torch.empty((5, nbox, nbox, nbox), dtype=torch.float64).pin_memory() for _ in range(nloc)]
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
            for kms, cnt, name in rows[:32]:
                print(f"{name:28s} {cnt:7d} {kms:10.3f} ms {100 * kms / tot:5.1f}%", file=sys.stderr)

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic, traffic_src = ncu_traffic("gsrb_kernel", nbox)
        t, nl, by = prof["abec_gsrb"]
        achieved = (by / 1e9) / (t * 1e-3) if t > 0 else 0.0
        secondary = {}
        for name in ("nodal_gs", "compute_aofs", "extrap_vel", "abec_apply", "nodal_adotx"):
            tt, nn, bb = prof[name]
            if tt > 0:
                secondary[name] = {"achieved_gbs": (bb / 1e9) / (tt * 1e-3), "frac": (bb / 1e9) / (tt * 1e-3) / peak,
                                   "launches": nn, "ms_total": tt}
        cpu_val, cpu_ms, cores = (None, None, None)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_val, cpu_ms, cores = cpu_oracle_run(args.cpu_n, 2, 1)
            cpu = {"value": cpu_val, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"TaylorGreen {args.cpu_n}^3 (a bounded sample of the 256^3 workload; BASELINE.json configs[0] is 64^3): 2 timed steps after init + 1 warm-up of "
                             f"the CPU oracle (restatement of the IAMR path; the AMReX build cannot be produced offline), "
                             f"OpenMP on {cores} host threads"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"TaylorGreen 3D {nbox}^3 per GPU single-level (Tutorials/TaylorGreen/inputs.3d.taylorgreen, "
                                   f"nu=1e-4, cfl=0.7, periodic), BASELINE.json configs[1]" + ("" if world == 1 else " weak-scaled"),
                       "n_cell": list(ncell), "boxes": len(boxes), "box": nbox, "parallelism": f"one {nbox}^3 box per rank x{world}" + ("" if world == 1 else ", z slabs: x/y wrapped in-kernel, z ghost planes over NCCL, coarse multigrid levels replicated"),
                       "l2": "working set per kernel >> 126 MB L2 (one fp64 cell array = 134 MB)",
                       "mg_iters_last_step": {"mac": iters[-1][0], "visc": iters[-1][1], "nodal": iters[-1][2]}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes,
                    "steps": args.e2e_steps, "api": "iamrx_ns_step_host (pinned host state in, new state out)",
                    "checksum_max_abs_u": checksum},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "gsrb_kernel (ABec red-black colour pass, finest two MG levels)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_note": f"DRAM bytes per finest-level launch ({nbox}^3, algorithmic {48 * nbox ** 3}); {traffic_src}",
                         "peak_source": peak_src, "launches_timed": nl, "ms_total": t,
                         "algorithmic_bytes": "48 B/cell/colour pass (a=0), 56 with alpha (SURVEY.md 8d)",
                         "other_kernels": secondary},
            "clocks": clocks,
        }
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
    ap.add_argument("--cpu-n", type=int, default=128, help="box size of the CPU sample (128^3: ~10-30 s of host work)")
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
