#!/usr/bin/env python
"""bench.py — Guacho hydro/MHD time-step throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): MHD zone-updates/s (HLLD + flux-CD, FP64); one zone-update = one
full tstep (both stages, boundaries, CFL) of one physical cell.  Workload at N=1:
BASELINE.json configs[1], the 3-D Orszag-Tang vortex on 256^3 (OT initial conditions
extruded along z, periodic, minmod, cfl 0.2), synthetic data generated on the host.
N>1: weak scaling, the same 256^3 block per GPU, z-slab decomposition, NCCL halo exchange.

One JSON line on stdout (rank 0).  `value` = device-resident throughput (CUDA events on
the solver's stream, max over ranks); `e2e` = the same metric through the C ABI with HOST
buffers: every step uploads u (pinned host, reference layout), runs get_timestep + tstep
and downloads u again, all inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from guacho_b200.config import Params, ot_3d  # noqa: E402
from guacho_b200 import problems  # noqa: E402

METRIC = "MHD zone-updates/s (HLLD + flux-CD, FP64)"
UNIT = "zone-updates/s"
BYTES_PER_ZONE = 320.0      # algorithmic: 5*neq doubles, neq = 8


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (sampled from before the
    warm-up; only the samples whose host timestamp falls inside the timed window are kept)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [ln for (t, ln) in self.lines if t0 <= t <= t1 + 0.03]
        note = None
        if not inside:       # window shorter than the sampling period: nearest samples after the start
            inside = [ln for (t, ln) in self.lines if t >= t0][:2] or [ln for (_t, ln) in self.lines[-2:]]
            note = "timed window shorter than the sampling period; nearest samples used"
        sm, smax, reasons, power = [], [], set(), []
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
               "samples": len(sm), "reasons": sorted(reasons)}
        if note:
            out["note"] = note
        return out


SOLVERS = {"hll": (1, False), "hllc": (2, False), "hlle": (3, True), "hlld": (4, True)}   # name -> (SOLVER_*, MHD)


def workload(n: int, world: int, strong: bool = False, solver: str = "hlld") -> Params:
    if strong:   # strong scaling: n^3 in total, z-slabs of n/world planes (BASELINE configs[2] "512^3 ... strong")
        p = ot_3d(n)
    else:        # weak scaling (default): the same n^3 block per GPU, stacked along z
        p = ot_3d(n, nztot=n * world, zmax=1.0 * world)
    sv, mhd = SOLVERS[solver]
    if solver != "hlld":   # BASELINE configs[4], the solver sweep: HLL/HLLC are hydro (neq = 5), HLLE is MHD + flux-CD
        p = p.replace(riemann_solver=sv, mhd=mhd, enable_flux_cd=mhd)
    return p


CPU_RATE_GUESS = 1.2e7      # zone-updates/s of the oracle on ~16 host threads (sizes the bounded sample; the measured rate is what is reported)


def cpu_reference(p_block: Params, steps: int, warmup: int, threads: int, budget_s: float = 90.0):
    """The reference's CPU path (C++ restatement, oracle/, -O3 -ffp-contract=off): one block per host thread, split along x
    like the reference's MPI ranks.  The whole grid when `steps + warmup` steps of it fit the time budget, otherwise a bounded
    sample: a z-slab of the same initial conditions (OT is z-invariant, so the per-zone work is identical)."""
    from tests.oracle_lib import Oracle
    # the largest block count that divides the grid and leaves every block at least 4 cells wide (the host's core count need not divide 256)
    threads = max(t for t in range(1, max(1, threads) + 1) if p_block.nxtot % t == 0 and p_block.nxtot // t >= 4)
    plane = p_block.nxtot * p_block.nytot
    nzs = int(budget_s * CPU_RATE_GUESS * (threads / 16.0) / (plane * max(1, steps + warmup)))
    nzs = max(4, min(p_block.nztot, nzs))
    if nzs < p_block.nztot:
        nzs = max(4, nzs // 4 * 4)
    full = nzs == p_block.nztot
    p = p_block.replace(nztot=nzs, zmax=p_block.zmax * nzs / p_block.nztot, MPI_NBX=threads, MPI_NBY=1, MPI_NBZ=1)
    o = Oracle(p, fast=True, threads=threads)
    g = problems.orszag_tang(p.replace(MPI_NBX=1), (0, 0, 0))
    o.scatter_u(g)
    del g
    o.start()
    o.iter = 11                                   # past the 10-step CFL ramp (hydro_core.f90:677-682)
    if warmup > 0:
        o.run_timed(warmup)
    sec = o.run_timed(steps)
    zones = p.nxtot * p.nytot * p.nztot
    what = "the whole grid" if full else f"{p.nxtot}x{p.nytot}x{nzs} slab of the same OT field (bounded sample)"
    sample = f"{what}, {threads} blocks x 1 thread (x-split like the reference's MPI), {steps} steps after {warmup} warm-up"
    return zones * steps / sec, sec / steps * 1e3, sample, threads, full


def workload_name(n: int, strong: bool, solver: str, flux_cd: bool) -> str:
    return ((f"3-D Orszag-Tang {n}^3 in total (BASELINE configs[2], strong scaling)" if strong else f"3-D Orszag-Tang {n}^3 per GPU (BASELINE configs[1])")
            + ", " + solver.upper() + (" + flux-CD" if flux_cd else "") + ", minmod, periodic, cfl 0.2")


def exo_cpu_reference(steps: int, warmup: int, threads: int, scale: float = 1.0):
    """EXO as shipped on the oracle (its restatement of EXO/user_mod.f90 + EXO/exoplanet.f90 + src/cooling_h.f90), one block per
    host thread along x; `scale` < 1 shrinks the grid for a bounded sample (same dx ratios, fewer cells)."""
    import ctypes as C
    from tests.oracle_lib import Oracle
    from guacho_b200.exo import exo_params, Scalings
    nx, ny, nz = (max(8, int(round(n * scale)) // 4 * 4) for n in (400, 100, 400))
    threads = max(t for t in range(1, max(1, threads) + 1) if nx % t == 0 and nx // t >= 4)
    p = exo_params(nx, ny, nz, MPI_NBX=threads)
    sc = Scalings.of(p)
    o = Oracle(p, fast=True, threads=threads)
    o.L.orc_init_exo(o.h, *[C.c_double(v) for v in (sc.rsc, sc.rhosc, sc.Tempsc, sc.vsc2, sc.tsc, sc.bsc)])
    o.L.orc_exo_initial_conditions(o.h)
    o.start()
    o.iter = 11
    if warmup > 0:
        o.run_timed(warmup)
    sec = o.run_timed(steps)
    zones = nx * ny * nz
    sample = f"EXO {nx}x{ny}x{nz}" + (" (the shipped grid)" if scale == 1.0 else " (bounded sample of the shipped 400x100x400)") + f", {threads} blocks x 1 thread, {steps} steps after {warmup} warm-up"
    return zones * steps / sec, sec / steps * 1e3, sample, threads


EXO_METRIC = "MHD zone-updates/s (EXO as shipped: HLLD + flux-CD + 2 passives + EOS_H_RATE + COOL_H + wind-sphere BC + gravity source + eta, FP64)"
EXO_WORKLOAD = "EXO/ exoplanet wind as shipped (BASELINE configs[3]): 400x100x400, HLLD MHD + flux-CD, 2 passives, EOS_H_RATE, COOL_H, outflow + user BC, gravity, eta 0.01, cfl 0.4"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: no Fortran compiler / MPI in this image)
    on the host cores, same workload, metric and unit; --steps / --warmup are honoured.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    if args.problem == "exo":
        v, ms, sample, threads = exo_cpu_reference(steps, warmup, threads)
        print(json.dumps({"impl": "reference", "metric": EXO_METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": EXO_WORKLOAD, "sample": sample},
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return 0
    p = workload(args.n, 1, False, args.solver)
    v, ms, sample, threads, full = cpu_reference(p, steps, warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n, False, args.solver, p.enable_flux_cd), "grid_total": [p.nxtot, p.nytot, p.nztot], "neq": p.neq,
                   "sample": sample, "whole_grid": full},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is Fortran+MPI and cannot be built in this image (no Fortran compiler/MPI); this is the line-faithful C++ restatement in oracle/ "
                "(ms_per_step is per step of the sampled grid; value is zone-updates/s and does not depend on the sample size)",
    }
    print(json.dumps(line), flush=True)
    return 0


def run_exo(args, rank, local_rank, world):
    """BASELINE configs[3]: EXO as shipped on one GPU through the same C ABI (wind spheres, gravity, bc hook as device functors)."""
    import torch
    from guacho_b200.exo import Exo, exo_params
    from guacho_b200.solver import Block
    if world != 1:
        raise SystemExit("--problem exo runs on one GPU (the shipped problem is 400x100x400)")
    p = exo_params(strict_fp=args.strict, device=local_rank)
    ex = Exo(p)
    sampler = ClockSampler(local_rank)
    sampler.start()
    with Block(p) as blk:
        ex.attach(blk, 0.0)
        u0 = ex.initial_conditions()
        blk.set_state(u0)
        tsim, it, _ = blk.run(max(3, args.warmup), 0.0, 1)
        torch.cuda.synchronize()
        l0, tw0 = blk.launch_count, time.time()
        tsim, it, last_dt = blk.run(args.steps, tsim, it)
        tw1, l1 = time.time(), blk.launch_count
        ms = blk.last_elapsed_ms
        clocks = sampler.stop(tw0, tw1)
        zones = p.nx * p.ny * p.nz
        value = zones * args.steps / (ms * 1e-3)
        blk.set_profiling(True)
        nprof = min(args.steps, 3)
        tsim, it, _ = blk.run(nprof, tsim, it)
        ktimes = blk.kernel_times()
        blk.set_profiling(False)
        # end to end: host buffers every step
        e2e_steps = args.e2e_steps if args.e2e_steps is not None else 3
        pinned = torch.empty(int(np.prod(p.block_shape())), dtype=torch.float64, pin_memory=True)
        uh = pinned.numpy().reshape(p.block_shape(), order="F")
        uh[...] = blk.get_state()
        t_e, it_e = tsim, it
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            blk.set_time(t_e); blk.set_state(uh); dt, _d = blk.get_timestep(it_e, 10, t_e, 1e300); blk.tstep(dt); blk.get_state_into(uh); t_e += dt; it_e += 1
        torch.cuda.synchronize()
        e2e_sec = time.perf_counter() - t0
        finite = bool(np.isfinite(uh).all())
        fused = ktimes["stage2"][1] > 0
    peak_gbs, peak_src = measured_peaks()
    step_ms = ms / args.steps
    bytes_zone = 40.0 * p.neq + 16.0 * p.neq          # 5*neq doubles + the eta pass (read up, write u): 400 + 160 B
    if fused:
        dom_key, dom_name, dom_bytes = "stage2", "k_stage<HLLD,minmod,ORDER=2,fluxCD,2 passives,H_RATE,gravity>", 8.0 * (p.neq + (p.neq - 3) * 2 + 3)
    else:
        dom_key, dom_name, dom_bytes = "flux", "k_flux<HLLD,minmod> (pass-per-routine path: 3 launches per stage)", 2 * 8.0 * p.neq * 3
    dom_ms = ktimes[dom_key][0] / nprof if ktimes[dom_key][1] else None
    cpu = None
    if not args.no_cpu_baseline:
        v, _ms, sample, threads = exo_cpu_reference(1, 1, os.cpu_count() or 1, scale=0.5)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    line = {"metric": EXO_METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": EXO_WORKLOAD, "grid_total": [p.nxtot, p.nytot, p.nztot], "neq": p.neq, "path": "fused stage kernels" if fused else "pass-per-routine kernels",
                       "kernels": "strict (-fmad=false)" if args.strict else "fast (-fmad=true)", "l2": "working set exceeds the 126 MB L2; no flush needed"},
            "clocks": clocks,
            "e2e": {"value": (zones * e2e_steps / e2e_sec if e2e_steps else None), "unit": UNIT, "h2d_bytes_per_step": int(uh.nbytes), "d2h_bytes_per_step": int(uh.nbytes) + 16,
                    "steps": e2e_steps, "ms_per_step": (e2e_sec / e2e_steps * 1e3 if e2e_steps else None)},
            "gpu_launches": int(l1 - l0),
            "roofline": {"bound": "hbm", "achieved": (dom_bytes * zones / (dom_ms * 1e-3) / 1e9 if dom_ms else None), "peak": peak_gbs, "unit": "GB/s",
                         "frac": (dom_bytes * zones / (dom_ms * 1e-3) / 1e9 / peak_gbs if dom_ms else None), "traffic": None, "peak_source": peak_src, "kernel": dom_name,
                         "ms_per_step": dom_ms, "algorithmic_bytes_per_zone": dom_bytes,
                         "whole_step": {"achieved": bytes_zone * zones / (step_ms * 1e-3) / 1e9, "frac": bytes_zone * zones / (step_ms * 1e-3) / 1e9 / peak_gbs,
                                        "algorithmic_bytes_per_zone": bytes_zone, "definition": "40*neq + 16*neq B per zone-update (eta != 0: extra up -> u pass), neq = 10"},
                         "kernel_ms_per_step": {k: v[0] / nprof for k, v in ktimes.items() if v[1]}},
            "cpu_baseline": cpu, "finite": finite, "last_dt": last_dt}
    print(json.dumps(line), flush=True)
    return 0


def run_tcond(args, rank, local_rank, world):
    """the headline workload with the thermal-conduction operator on (src/thermal_cond.f90: isotropic Spitzer
    conduction with saturation, super-time-stepping), one GPU.  Reports the whole step and the operator's own HBM fraction."""
    import torch
    from guacho_b200.config import TC_ISOTROPIC
    from guacho_b200.solver import Block
    if world != 1:
        raise SystemExit("--problem tcond runs on one GPU")
    mu, Rg, gamma, T0, rsc, rhosc = 0.6, 8.3145e7, 5.0 / 3.0, 1.0e6, 1e10, 5e-16
    vsc2 = gamma * Rg * T0 / mu
    p = workload(args.n, 1, False, "hlld").replace(strict_fp=args.strict, device=local_rank, th_cond=TC_ISOTROPIC, tc_saturation=True, rsc=rsc, rhosc=rhosc,
                                                    vsc2=vsc2, tsc=rsc / np.sqrt(vsc2), bsc=float(np.sqrt(4 * np.pi * rhosc * vsc2)), mu=mu, Tempsc=T0 * gamma)
    sampler = ClockSampler(local_rank)
    sampler.start()
    with Block(p) as blk:
        blk.set_state(problems.orszag_tang(p, (0, 0, 0)))
        tsim, it, _ = blk.run(max(3, args.warmup), 0.0, 11)           # past the CFL ramp: the hydro step is long against the conduction time scale
        torch.cuda.synchronize()
        l0, tw0 = blk.launch_count, time.time()
        tsim, it, last_dt = blk.run(args.steps, tsim, it)
        tw1, l1 = time.time(), blk.launch_count
        ms = blk.last_elapsed_ms
        clocks = sampler.stop(tw0, tw1)
        dt_cond, nsub = blk.tc_info()
        blk.set_profiling(True)
        nprof = min(args.steps, 3)
        sub = 0
        for _ in range(nprof):
            tsim, it, _ = blk.run(1, tsim, it)
            sub += blk.tc_info()[1]
        ktimes = blk.kernel_times()
        blk.set_profiling(False)
        finite = bool(np.isfinite(blk.get_state()).all())
    peak_gbs, peak_src = measured_peaks()
    zones = p.nx * p.ny * p.nz
    step_ms = ms / args.steps
    tc_ms = ktimes["tcond"][0] / nprof
    # one block, isotropic: ONE marching kernel per substep reads the dynamic variables once and writes u(5): 8*(neqdyn + 1) B
    tc_bytes_zone = 8.0 * (p.neqdyn + 1)
    tc_bytes = tc_bytes_zone * zones * (sub / nprof)
    line = {"metric": "MHD zone-updates/s with thermal conduction (HLLD + flux-CD + isotropic saturated Spitzer conduction, super-time-stepping, FP64)",
            "value": zones * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D Orszag-Tang {args.n}^3 HLLD+flux-CD + th_cond=ISOTROPIC, tc_saturation (thermal_cond.f90), periodic box", "grid_total": [p.nxtot, p.nytot, p.nztot],
                       "substeps_last_step": nsub, "dt_cond_s": dt_cond, "l2": "working set exceeds the 126 MB L2; no flush needed"},
            "clocks": clocks, "gpu_launches": int(l1 - l0),
            "roofline": {"bound": "hbm", "achieved": tc_bytes / (tc_ms * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": tc_bytes / (tc_ms * 1e-3) / 1e9 / peak_gbs, "traffic": None,
                         "peak_source": peak_src, "kernel": "k_tc_march (one launch per thermal-conduction substep) + k_tc_prim once per step (dt_cond)", "ms_per_step": tc_ms,
                         "substeps_per_step": sub / nprof, "algorithmic_bytes_per_zone_per_substep": tc_bytes_zone,
                         "kernel_ms_per_step": {k: v[0] / nprof for k, v in ktimes.items() if v[1]}},
            "finite": finite, "last_dt": last_dt}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", dest="n", type=int, default=256, help="cells per side of the per-GPU block (with --strong: of the whole domain)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strict", action="store_true", help="use the -fmad=false bit-comparison kernels")
    ap.add_argument("--solver", default="hlld", choices=sorted(SOLVERS), help="Riemann solver (BASELINE configs[4] sweep); the headline metric is hlld")
    ap.add_argument("--strong", action="store_true", help="strong scaling: --n is the TOTAL grid side, split into z-slabs over the GPUs")
    ap.add_argument("--problem", default="ot", choices=["ot", "exo", "tcond"],
                    help="ot: the headline workload; exo: EXO/ as shipped (BASELINE configs[3]), 400x100x400, one GPU; tcond: ot with the thermal-conduction operator on")
    ap.add_argument("--no-extras", action="store_true", help="skip the 512^3 lines (extra.grid512 at N=1; extra.weak512 / extra.strong512 at N>1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from guacho_b200.distributed import init_process_group, make_rank_block
    from guacho_b200.decomp import coords_of, halo_bytes_per_step

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the step has no CPU fallback (use --impl reference for the CPU path)")
    args.warmup = max(args.warmup, 3)
    rank, local_rank, world = init_process_group()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)

    if args.problem == "exo":
        return run_exo(args, rank, local_rank, world)
    if args.problem == "tcond":
        return run_tcond(args, rank, local_rank, world)
    if args.strong and args.n % world:
        raise SystemExit(f"--strong: {args.n} planes do not split over {world} GPUs")
    p = workload(args.n, world, args.strong, args.solver).replace(strict_fp=args.strict)
    nb = (1, 1, world)
    blk = make_rank_block(p, rank, world, local_rank, nb=nb)
    pb = blk.p
    coords = coords_of(rank, nb)
    u0 = problems.orszag_tang(pb, coords)
    zones_rank = pb.nx * pb.ny * pb.nz
    zones_total = zones_rank * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    blk.set_state(u0)
    tsim, it = 0.0, 1
    tsim, it, _ = blk.run(args.warmup, tsim, it)
    barrier()
    l0 = blk.launch_count
    tw0 = time.time()
    tsim, it, last_dt = blk.run(args.steps, tsim, it)
    tw1 = time.time()
    ms = blk.last_elapsed_ms
    l1 = blk.launch_count
    barrier()
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    ms = max_over_ranks(ms)
    value = zones_total * args.steps / (ms * 1e-3)

    # per-kernel-class device times (separate pass, CUDA events around each launch on the solver's stream)
    blk.set_profiling(True)
    tsim, it, _ = blk.run(min(args.steps, 5), tsim, it)
    ktimes = blk.kernel_times()
    nprof = min(args.steps, 5)
    blk.set_profiling(False)

    # ---------------- end to end through the C ABI with host buffers ----------------
    e2e_steps = args.e2e_steps if args.e2e_steps is not None else max(3, min(args.steps, 10))
    pinned = torch.empty(int(np.prod(pb.block_shape())), dtype=torch.float64, pin_memory=True)
    uh = pinned.numpy().reshape(pb.block_shape(), order="F")
    uh[...] = blk.get_state()
    t_e, it_e = tsim, it
    for _ in range(2 if e2e_steps > 0 else 0):   # warm-up
        blk.set_state(uh); dt, _d = blk.get_timestep(it_e, 10, t_e, 1e300); blk.tstep(dt); blk.get_state_into(uh); t_e += dt; it_e += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        blk.set_state(uh)                                   # H2D: u in reference layout (with ghosts)
        dt, _d = blk.get_timestep(it_e, 10, t_e, 1e300)     # D2H: CFL dt
        blk.tstep(dt)
        blk.get_state_into(uh)                              # D2H: updated u
        t_e += dt; it_e += 1
    torch.cuda.synchronize()
    e2e_sec = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = zones_total * e2e_steps / e2e_sec if e2e_steps > 0 else None
    state_bytes = int(uh.nbytes)
    finite = bool(np.isfinite(uh).all())
    # the loop a Fortran host runs between outputs (main.f90:94-125): state stays resident, only
    # dt / dump_flag cross the C ABI every step (wall clock, host-driven, one sync per step)
    res_steps = max(3, min(args.steps, 20))
    barrier()
    t0 = time.perf_counter()
    for _ in range(res_steps):
        dt, _d = blk.get_timestep(it_e, 10, t_e, 1e300)
        blk.tstep(dt)
        t_e += dt; it_e += 1
    torch.cuda.synchronize()
    res_sec = max_over_ranks(time.perf_counter() - t0)
    barrier()

    # ---------------- BASELINE configs[2] / the north_star's target grid, same run: 512^3 ----------------
    peak_gbs, peak_src = measured_peaks()

    def measure_grid(n: int, strong: bool, steps: int = 5, warmup: int = 3):
        """Device-resident zone-updates/s of an n^3 block per GPU (weak) or n^3 in total (strong), same kernels and timing rules."""
        try:
            if strong and n % world:
                return {"skipped": f"{n} planes do not split over {world} GPUs"}
            p2 = workload(n, world, strong, args.solver).replace(strict_fp=args.strict)
            b2 = make_rank_block(p2, rank, world, local_rank, nb=nb)
            try:
                u2 = problems.orszag_tang(b2.p, coords)
                b2.set_state(u2)
                del u2
                t2, i2, _ = b2.run(warmup, 0.0, 1)
                barrier()
                t2, i2, _ = b2.run(steps, t2, i2)
                ms2 = max_over_ranks(b2.last_elapsed_ms)
                barrier()
                zr = b2.p.nx * b2.p.ny * b2.p.nz
                v2 = zr * world * steps / (ms2 * 1e-3)
                return {"value": v2, "unit": UNIT, "ms_per_step": ms2 / steps, "steps": steps, "warmup": warmup, "n_gpus": world,
                        "scaling": "strong" if strong else "weak", "grid_total": [b2.p.nxtot, b2.p.nytot, b2.p.nztot],
                        "whole_step_frac": 40.0 * b2.p.neq * zr * steps / (ms2 * 1e-3) / 1e9 / peak_gbs,
                        "workload": workload_name(n, strong, args.solver, b2.p.enable_flux_cd)}
            finally:
                b2.close()
        except Exception as e:      # an extra line never takes the headline down with it
            return {"error": f"{type(e).__name__}: {e}"}

    extra = {}
    if not args.no_extras and args.n == 256 and not args.strong and args.solver == "hlld":
        if world == 1:
            extra["grid512"] = measure_grid(512, False)
        else:
            extra["weak512"] = measure_grid(512, False)
            extra["strong512"] = measure_grid(512, True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    step_ms = ms / args.steps
    tot_prof = sum(v[0] for v in ktimes.values())
    fused = ktimes["stage2"][1] > 0
    if fused:
        # dominant kernel: the fused 2nd-order stage k_stage<HLLD,minmod,2,fluxcd>.  Its algorithmic traffic per zone:
        # read U* (neq) and the non-B part of U^n (5), write the non-B part of U^{n+1} (5) and E (3) = 21 doubles with
        # flux-CD (the B part of the stage belongs to k_bupdate); 3*neq doubles without flux-CD.
        dom_doubles = (pb.neq + 5 + 5 + 3) if pb.enable_flux_cd else 3 * pb.neq
        dom_name, dom_key, dom_bytes_zone = f"k_stage<{args.solver.upper()},minmod,ORDER=2{',fluxCD' if pb.enable_flux_cd else ''}> (fused prim+3 sweeps+E+update)", "stage2", 8 * dom_doubles
    else:
        dom_name, dom_key, dom_bytes_zone = "k_flux<HLLD,minmod> (3 launches per stage, unfused path)", "flux", 2 * 8 * pb.neq * 3
    # time of the dominant kernel per step = everything its class launched in a step (the overlap path of N > 1 launches a
    # stage three times: two boundary slabs and the interior), against the bytes of the whole block
    dom_ms, dom_n = ktimes[dom_key]
    dom_step_ms = dom_ms / nprof if dom_n else None
    dom_launches_per_step = dom_n / nprof if nprof else None
    dom_gbs = dom_bytes_zone * zones_rank / (dom_step_ms * 1e-3) / 1e9 if dom_step_ms else None
    traffic = None
    try:       # DRAM bytes of that kernel per launch from the committed ncu --set full capture: only for the captured launch geometry
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj = json.load(f)
        if fused and tj.get("zones") == zones_rank and args.solver == "hlld" and world == 1 and dom_launches_per_step == 1:
            traffic = tj["stage2_dram_bytes_per_launch"]
    except Exception:
        pass
    bytes_zone = 40.0 * pb.neq                        # 5*neq doubles: 320 B (MHD), 200 B (hydro)
    step_gbs = bytes_zone * zones_rank / (step_ms * 1e-3) / 1e9
    # second ceiling: FP64 instructions per zone-update (ncu, committed profile) against the measured DFMA
    # issue rate of this GPU (profiles/peaks_r1.json) and all instructions against the issue slots at the sampled clock
    fp64 = None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj2 = json.load(f)
        with open(os.path.join(ROOT, "profiles", "peaks_r1.json")) as f:
            pk = json.load(f)
        if fused and args.solver == "hlld":
            ceil_fp64 = pk["dfma_per_s"] / tj2["fp64_thread_inst_per_zone_update"]
            sm_hz = 1e6 * (clocks or {}).get("sm_mhz", 1965.0) if (clocks or {}).get("sm_mhz") else 1.965e9
            ceil_issue = pk["sms"] * 4 * 32 * sm_hz / tj2["thread_inst_per_zone_update"]
            fp64 = {"fp64_inst_per_zone_update": tj2["fp64_thread_inst_per_zone_update"], "inst_per_zone_update": tj2["thread_inst_per_zone_update"],
                    "dfma_per_s_measured": pk["dfma_per_s"], "fp64_ceiling_zone_updates_per_s": ceil_fp64, "frac_of_fp64_ceiling": value / world / ceil_fp64,
                    "issue_ceiling_zone_updates_per_s": ceil_issue, "frac_of_issue_ceiling": value / world / ceil_issue,
                    "note": "informational: the step is bound by FP64 issue, not HBM; instruction counts from the committed ncu capture"}
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "achieved": dom_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": (dom_gbs / peak_gbs if dom_gbs else None), "traffic": traffic,
        "peak_source": peak_src, "kernel": dom_name, "ms_per_step": dom_step_ms, "launches_per_step": dom_launches_per_step,
        "algorithmic_bytes_per_zone": dom_bytes_zone,
        "definition": "algorithmic bytes of the dominant kernel (read U* 8 + U^n 5, write U^{n+1} 5 + E 3 = 21 doubles per zone with flux-CD) x zones per GPU / the device time its launches take per step (CUDA events on the solver's stream)",
        "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak_gbs, "algorithmic_bytes_per_zone": bytes_zone,
                       "definition": "40*neq B per zone-update (5*neq doubles: 320 B MHD, 200 B hydro) x zones per GPU / whole-step device time"},
        "fp64_note": "the kernel is FP64-pipe/issue bound, not HBM bound: see profiles/ (sm__inst_executed_pipe_fp64 ~47%, issue ~56%, dram ~16%) and DESIGN.md",
        "fp64_issue": fp64,
        "kernel_share": {k: (v[0] / tot_prof if tot_prof > 0 else None) for k, v in ktimes.items() if v[1]},
        "kernel_ms_per_step": {k: v[0] / nprof for k, v in ktimes.items() if v[1]},
    }
    cpu = None
    if not args.no_cpu_baseline and world == 1:        # reported baseline: rank 0 at N = 1 only (bounded sample, a few seconds)
        threads = os.cpu_count() or 1
        v, _ms, sample, threads, _full = cpu_reference(workload(args.n, 1), 2, 1, threads, budget_s=12.0)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    line = {
        "metric": METRIC if args.solver == "hlld" else f"{'MHD' if pb.mhd else 'hydro'} zone-updates/s ({args.solver.upper()}{' + flux-CD' if pb.enable_flux_cd else ''}, FP64)",
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.n, args.strong, args.solver, pb.enable_flux_cd),
                   "grid_total": [pb.nxtot, pb.nytot, pb.nztot], "blocks": list(nb), "neq": pb.neq,
                   "kernels": "strict (-fmad=false)" if args.strict else "fast (-fmad=true)",
                   "l2": f"working set (u, up, E: {(2 * pb.neq + 3) * 8 * (pb.nx + 32) * (pb.ny + 4) * (pb.nz + 4) / 1e9:.1f} GB per GPU) exceeds the 126 MB L2; no flush needed",
                   "halo_bytes_per_step_per_gpu": halo_bytes_per_step(pb)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": state_bytes + 16,
                "steps": e2e_steps, "ms_per_step": (e2e_sec / e2e_steps * 1e3 if e2e_steps > 0 else None),
                "path": "gx_set_state(host u) -> gx_get_timestep -> gx_tstep -> gx_get_state(host u) through libguacho_gx.so"},
        "e2e_resident": {"value": zones_total * res_steps / res_sec, "unit": UNIT, "steps": res_steps, "ms_per_step": res_sec / res_steps * 1e3,
                         "path": "gx_get_timestep -> gx_tstep per step, state resident on the device (the reference host's loop between outputs)"},
        "gpu_launches": int(l1 - l0),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "finite": finite, "last_dt": last_dt,
        "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
