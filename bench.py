#!/usr/bin/env python
"""bench.py — throughput of the PICSP particle loop (deposit + solve + E field + gather/push).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): particle-steps/s over the full step, both species.
Workload: BASELINE.json config "1024x1024 periodic spectral, 1e9 particles" (two-stream
electrons + cold ions, synthetic Maxwellian state from the bench-only device loader); the
1e9 particles are sharded across the N ranks by index range (strong scaling), the grid is
replicated and the per-rank partial charge densities are summed with one NCCL all-reduce.

One JSON line on stdout (rank 0).  `value` = particle-steps/s with the state resident in
HBM, timed with CUDA events on the library's stream (max over ranks).  `e2e` = the same
metric through the C ABI with HOST buffers: upload of the particle state from pinned host
memory, then --e2e-periods dump periods of 50 steps (the reference's diagnostic cadence,
main.cpp:507), each followed by the dump of everything the reference writes (phase space of
both species, den.i, den.e, phi, KE) through picsp_dump_begin / picsp_dump_wait, i.e. copied
out while the next period steps; the timed region ends when the last dump has landed.  `roofline` is for the
dominant kernel (the mover fused with the next step's deposit): 64 algorithmic bytes per
particle-step.  `cpu_baseline` / `--impl reference` time the UNMODIFIED reference
translation unit (oracle/_ref) on one host core (the reference is single-threaded).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALGO_BYTES_PER_PARTICLE_STEP = 64   # read + write {x, y, vx, vy} f64, deposit fused into the mover
METRIC = "particle_steps_per_sec"
UNIT = "particle-steps/s"


def physical_normalisation():
    """The reference's normalisation (main.cpp:279-291) of its shipped physical values (input.ini)."""
    EPS_un, K, EV_TO_K = 8.85418782e-12, 1.38065e-23, 11604.52
    q, me, mi, n0, vthE, vthI, driftE = 1.602e-19, 9.109e-31, 1.673e-27, 1e12, 0.9, 0.026, 0.2
    omega_pe = np.sqrt((q * q * n0) / (me * EPS_un))
    lambda_d = np.sqrt((EPS_un * K * vthE * EV_TO_K) / (n0 * q * q))
    return dict(dt=float(1e-10 * omega_pe), dx=float(1.2e-4 / lambda_d), mass_i=mi / me,
                vth_i=vthI / vthE, vth_e=1.0, drift_e=driftE / vthE)


def ncu_traffic_bytes_per_particle():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full
    capture (profiles/ncu_traffic.json), per particle; None if no capture is recorded."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return float(json.load(f)["k_tile_mover<0>"]["dram_bytes_per_particle"])
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        bits = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                0x80: "hw_power_brake_slowdown"}
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, name in bits.items():
                    if r & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference TU on one host core
# ------------------------------------------------------------------------------------------
def synthetic_state(n, seed, xl, vth, xdrift):
    """The bench workload's particle distribution (uniform positions, v = vth*sqrt(2)*(r1+r2+r3-1.5), +-xdrift
    alternating in x: what picsp_species_fill_synthetic produces on the device), drawn with numpy for the CPU arm."""
    rng = np.random.default_rng(seed)
    x, y = rng.random(n) * xl, rng.random(n) * xl
    sgn = np.where(np.arange(n) & 1, -1.0, 1.0)
    vx = vth * np.sqrt(2.0) * (rng.random(n) + rng.random(n) + rng.random(n) - 1.5) + sgn * xdrift
    vy = vth * np.sqrt(2.0) * (rng.random(n) + rng.random(n) + rng.random(n) - 1.5)
    return x, y, vx, vy


def run_reference_cpu(cells, n_per_species, steps, warmup):
    """Times the reference's own loop body (main.cpp:481-504) on a bounded sample of the workload: the same grid,
    the same particle distribution (two-stream electrons + cold ions, uniform in the box), fewer particles
    (throughput is per particle).

    Returns (value, info).  value = particle-steps/s over deposit + rho + EF + push, i.e. the
    reference's own functions; the Poisson solve is timed but reported separately because
    FFTW3 is absent from this image and the shim FFT that stands in for it is not FFTW (at
    the full workload the solve is <0.1% of the reference's step).  The dead velocity-moment
    deposit the reference also runs every step (scatterSpeciesVel, outputs never consumed) is
    timed and reported but NOT charged to the reference."""
    from oracle import oracle as orc
    nm = physical_normalisation()
    kind = "reference" if orc.have_reference() else "port"
    t0 = time.perf_counter()
    xl = cells * nm["dx"]
    ions = synthetic_state(n_per_species, 1, xl, nm["vth_i"], 0.0)
    electrons = synthetic_state(n_per_species, 2, xl, nm["vth_e"], nm["drift_e"])
    if kind == "reference":
        r = orc.Reference(cells, cells, nm["dx"], nm["dt"], nm["mass_i"], n_per_species, n_per_species,
                          vth_i=nm["vth_i"], vth_e=nm["vth_e"], solver=1)
        r.L.picsp_ref_set_fft_mode(3)
        r.set_species(0, *ions); r.set_species(1, *electrons)
        r.bootstrap()
        if warmup:
            r.step(warmup, with_dead_vel=True)
        r.phase_seconds(reset=True)
        r.step(steps, with_dead_vel=True)
        ph = r.phase_seconds(reset=True)
        r.close()
    else:
        o = orc.Oracle(cells, cells, nm["dx"], nm["dt"], nm["mass_i"], n_per_species, n_per_species,
                       vth_i=nm["vth_i"], vth_e=nm["vth_e"], solver=1)
        orc.Oracle.lib().oracle_set_fft_mode(3)
        o.set_species(0, *ions); o.set_species(1, *electrons)
        o.bootstrap()
        ph = dict(deposit=0.0, dead_vel_deposit=0.0, rho=0.0, solve=0.0, ef=0.0, push=0.0)

        def timed(key, fn, *a):
            t = time.perf_counter(); fn(*a); ph[key] += time.perf_counter() - t
        for it in range(warmup + steps):
            if it == warmup:
                for k in ph:
                    ph[k] = 0.0
            timed("deposit", o.scatterSpecies, 0); timed("deposit", o.scatterSpecies, 1)
            timed("rho", o.computeRho); timed("solve", o.solve); timed("ef", o.computeEF)
            timed("push", o.pushSpecies, 0); timed("push", o.pushSpecies, 1)
    path_s = ph["deposit"] + ph["rho"] + ph["ef"] + ph["push"]
    psteps = 2.0 * n_per_species * steps
    value = psteps / path_s
    info = {
        "value": value, "unit": UNIT, "cores": 1, "kind": kind,
        "sample": (f"bounded sample of the workload: same {cells}x{cells}-cell grid, same particle distribution (two-stream "
                   f"electrons + cold ions, uniform positions), {n_per_species} particles/species instead of the full count "
                   f"(throughput is per particle); {steps} steps after {warmup} warm-up, 1 of {os.cpu_count()} host cores "
                   f"(the reference is single-threaded), g++ -O2 (the reference's own makefile uses -O0, ~3x slower); value "
                   f"counts deposit+rho+EF+push; per-step seconds: deposit {ph['deposit'] / steps:.3f}, push {ph['push'] / steps:.3f}, "
                   f"rho+EF {(ph['rho'] + ph['ef']) / steps:.4f}; excluded: Poisson solve {ph['solve'] / steps:.3f} s/step "
                   f"(shim FFT, not FFTW) and the reference's dead velocity deposit {ph['dead_vel_deposit'] / steps:.3f} s/step"),
        "ms_per_step": 1e3 * path_s / steps, "wall_s": time.perf_counter() - t0,
        "sample_particles_per_species": int(n_per_species),
    }
    return value, info


def main_reference(args, rank):
    if rank != 0:
        return 0
    value, info = run_reference_cpu(args.cells, args.cpu_particles, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": info["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, None),
                       sample=f"each step is a bounded sample of this workload: {info['sample_particles_per_species']} particles per species "
                              f"of the same distribution on the same grid, on 1 host core (see cpu_baseline.sample)",
                       sample_particles_total=2 * info["sample_particles_per_species"]),
        "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


def workload_config(args, n_local):
    load = ("synthetic two-stream electrons + cold ions uniform in the box (bench-only device loader)" if args.load == "synthetic" else
            "the reference's own loadType-2 two-stream load (main.cpp:597-615, host loader picsp_host_loader_fill): every particle "
            "on the domain diagonal")
    bounded = bool(getattr(args, "walls", False))
    cfg = {
        "workload": ((f"{load}, {args.cells}x{args.cells} cells periodic "
                      f"({args.cells + 1}^2 nodes), spectral solver, {args.particles:.3g} particles total "
                      f"({args.particles // 2} per species), BASELINE.json configs[3]") if not bounded else
                     (f"{load}, {args.cells}x{args.cells} cells BOUNDED ({args.cells + 1}^2 nodes): absorbing walls, phi = 0 on the walls, "
                      f"red-black Gauss-Seidel/SOR to an L2 residual of 1e-12, density cleared every step, {args.particles:.3g} particles "
                      f"total; BASELINE.json configs[2] as an EXTENSION WITHOUT REFERENCE SEMANTICS (the reference is periodic-only)")),
        "load": args.load,
        "cells": args.cells, "particles_total": int(args.particles),
        "solver": "red-black SOR, Dirichlet walls (k_rb_sor, one cooperative launch)" if bounded else
                  "spectral (2-D DFT of the node array, k-space Green multiply, inverse DFT: spectralPotentialSolver, main.cpp:960-1058)",
        "sharding": f"particles by index range over {args.gpus} rank(s), grid replicated, the partial rho summed once per step "
                    f"({getattr(args, 'rho_reduction', 'n/a')})",
        "l2_policy": "inputs exceed L2 (particle state per rank >> 126 MB); no explicit flush",
    }
    if n_local is not None:
        cfg["particles_per_rank_per_species"] = int(n_local)
    return cfg


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def main_ours(args, rank, world, local_rank):
    import torch
    from picsp_b200 import ELECTRON, ION, Params, Simulation
    from picsp_b200.sim import shard_range

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: picsp_b200 has no CPU path"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    nm = physical_normalisation()
    n_species = int(args.particles) // 2
    lo, hi = shard_range(n_species, rank, world)
    n_local = hi - lo
    def make_sim():
        sim = Simulation(Params(args.cells, args.cells, nm["dx"], nm["dt"], nm["mass_i"], n_species, n_species,
                                solverType=1, device=local_rank, capacity=(n_local, n_local), flags=args.flags,
                                parts=args.parts))
        # tuning switches first: the first binning happens inside the fill / upload
        if args.sort_period_e > 0:
            sim.set_sort_period(ELECTRON, args.sort_period_e)
        if args.sort_period_i > 0:
            sim.set_sort_period(ION, args.sort_period_i)
        if args.cell_period_e >= 0:
            sim.set_cell_sort_period(ELECTRON, args.cell_period_e)
        if args.cell_period_i >= 0:
            sim.set_cell_sort_period(ION, args.cell_period_i)
        if args.bank_order_e is not None:
            sim.set_bank_order(ELECTRON, args.bank_order_e)
        if args.bank_order_i is not None:
            sim.set_bank_order(ION, args.bank_order_i)
        if args.load == "synthetic":
            sim.fill_synthetic(ION, n_local, first_index=lo, seed=1, vth=nm["vth_i"], xdrift=0.0)
            sim.fill_synthetic(ELECTRON, n_local, first_index=lo, seed=2, vth=nm["vth_e"], xdrift=nm["drift_e"])
        else:
            # the reference's loadType 2 through the product's host loader: a sequential recurrence over ALL particles
            # (ions first, the stale x carried into the electrons); every rank generates it and keeps its range
            from picsp_b200 import host
            from picsp_b200.lib import CRunConfig
            cfg = CRunConfig()
            cfg.numxCells = cfg.numyCells = args.cells
            cfg.nParticlesI = cfg.nParticlesE = n_species
            cfg.loadType, cfg.solverType = 2, 1
            cfg.stepSize, cfg.timeStep = nm["dx"], nm["dt"]
            cfg.vthI, cfg.vthE, cfg.driftI, cfg.driftE = nm["vth_i"], nm["vth_e"], 0.0, nm["drift_e"]
            t_load = time.perf_counter()
            loaded = host.load_species(cfg, seed=0)
            for s_ in (ION, ELECTRON):
                sim.set_species(s_, *(a[lo:hi] for a in loaded[s_]))
            del loaded
            print(f"[bench] reference loadType-2 load generated and uploaded in {time.perf_counter() - t_load:.1f} s", file=sys.stderr)
        if world > 1:
            u = [Simulation.comm_unique_id() if rank == 0 else None]     # one NCCL communicator per context
            dist.broadcast_object_list(u, src=0)
            sim.comm_attach(u[0], rank, world)
        if args.agg is not None:
            sim.set_deposit_aggregation(ION, args.agg); sim.set_deposit_aggregation(ELECTRON, args.agg)
        return sim

    sim = make_sim()
    args.spectral_engine = sim.spectral_engine()
    args.store_parts = sim.parts()
    args.rho_reduction = ("single rank: none" if world == 1 else
                          ("own reduce-scatter + all-gather kernel over NVLink peer memory (CUDA IPC)" if sim.comm_peer_reduction() else "ncclAllReduce"))
    sim.bootstrap()
    sim.profile_enable(True)
    sim.step(args.warmup)
    sim.sync()
    sim.profile_reset()

    sampler = ClockSampler(local_rank)
    launches0 = sim.kernel_launches()
    barrier()
    sampler.start()
    t_wall = time.perf_counter()
    sim.step(args.steps)
    sim.sync()
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.result()
    launches = sim.kernel_launches() - launches0
    prof = sim.profile()
    ms_total = max_over_ranks(prof["step"][0])
    ms_per_step = ms_total / args.steps
    value = float(args.particles) * args.steps / (ms_total * 1e-3)

    # roofline of the dominant kernel (fused mover+deposit): per-launch average over the timed region
    push_ms, push_calls = prof["push"]
    peak, peak_src = measured_peak_gbs()
    per_launch_s = push_ms * 1e-3 / max(push_calls, 1)
    nparts = sim.parts()                       # a species split into parts is pushed by one launch per part
    n_launch = n_local / nparts
    achieved = ALGO_BYTES_PER_PARTICLE_STEP * n_launch / per_launch_s / 1e9
    traffic = args.traffic_bytes_per_launch
    if traffic is None and ncu_traffic_bytes_per_particle() is not None:
        traffic = ncu_traffic_bytes_per_particle() * n_launch     # per launch, like `achieved`
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": "k_tile_mover<0> (leapfrog mover + CIC gather from a TMA-staged E tile + next-step CIC deposit); "
                          "the average includes the re-binning launches k_tile_mover<4> (every 8th electron launch, "
                          "73 B of DRAM traffic per particle instead of 64), charged at the same 64 algorithmic bytes",
                "traffic_source": "profiles/ncu_traffic.json: ncu --set full dram bytes per particle x particles per launch",
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PARTICLE_STEP * n_launch, "particles_per_launch": n_launch,
                "avg_launch_ms": per_launch_s * 1e3, "launches_timed": push_calls, "peak_source": peak_src,
                "share_of_step": push_ms / prof["step"][0] if prof["step"][0] else None}
    phases_ms = {k: v[0] / args.steps for k, v in prof.items()}

    # ---- parity probe: state after bootstrap + warm-up + timed steps.  The device loader is keyed by the GLOBAL
    # particle index, so the state is the same for every rank count up to summation order: the N = 1/2/4/8 lines can be
    # compared with each other (agreement expected to ~1e-12 relative; chaotic growth over these few steps is negligible)
    rho_g, phi_g = sim.grid("rho"), sim.grid("phi")
    parity_probe = {"steps_total": args.warmup + args.steps, "sum_abs_rho": float(np.abs(rho_g).sum()),
                    "l2_phi": float(np.sqrt((phi_g * phi_g).sum())), "ke_i": sim.computeKE(ION), "ke_e": sim.computeKE(ELECTRON),
                    "tolerance": "lines of different rank counts agree to 1e-10 relative (summation order of the partial densities "
                                 "and of the KE partial sums only)"}

    # ---- end to end through the C ABI with host buffers ---------------------------------
    # The e2e run starts from a FRESH context: in the reference's semantics the density is never cleared,
    # which makes boxes of many Debye lengths unstable; at this noise level velocities run away after ~510 steps in
    # total (profiles/r02_long_run_instability.md), so the e2e periods must not be stacked on top of the steps above.
    e2e = None
    e2e_skipped = None
    if not args.no_e2e:
        # the end-to-end run holds the whole phase space in (page-locked) host memory: 64 bytes per particle of a species
        try:
            import psutil
            need, avail = 64 * n_local, psutil.virtual_memory().available
            if need > 0.9 * avail:
                e2e_skipped = f"host buffers of {need / 1e9:.0f} GB against {avail / 1e9:.0f} GB of available host memory"
        except ImportError:
            pass
    if not args.no_e2e and e2e_skipped is None:
        sim.close()
        sim = make_sim()
        sim.bootstrap()
        sim.step(2); sim.sync()
        e2e = run_e2e(sim, args, n_local, barrier, max_over_ranks, torch, rank)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, info = run_reference_cpu(args.cells, args.cpu_particles, 2, 1)
        cpu_baseline = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}

    sim.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, n_local),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "spectral_engine": {"own": "the library's own shared-memory DFT (fft_kernels.cuh: two direct coprime factors, or a prime-factor split + Bluestein on 2^k)",
                                "cufft": "cuFFT D2Z / Z2D"}.get(getattr(args, "spectral_engine", ""), None),
            "e2e_skipped": e2e_skipped,
            "store_parts_per_species": int(getattr(args, "store_parts", 1)),     # picsp_params::parts resolved (> 1: one shared spare buffer set)
            "parity_probe": parity_probe,
            "clocks": clocks, "phases_ms_per_step": phases_ms, "wall_ms_per_step": 1e3 * t_wall / args.steps,
        }
        emit(line)
    return 0


def run_e2e(sim, args, n_local, barrier, max_over_ranks, torch, rank=0):
    """Upload the particle state from pinned host memory, then run --e2e-periods dump periods: 50 steps followed by
    the dump of what the reference writes every 50 steps (phase space of both species as [n][4] rows, den.i, den.e,
    phi, the two kinetic energies; main.cpp:507-527), all through the C ABI.  A dump is snapshot on the device and
    copied out on a second stream while the next period steps (picsp_dump_begin / picsp_dump_wait); the timed region
    ends when the last dump has landed in host memory.  With N ranks den.i / den.e are reduced to rank 0."""
    import ctypes as C
    from picsp_b200.lib import check
    dp = C.POINTER(C.c_double)
    nn = sim.nix * sim.niy
    pinned = True
    try:
        # one page-locked block of 4*n doubles per species: four arrays for the upload, [n][4] rows for the dumps
        blocks = [torch.empty(4 * n_local, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        grids = [torch.empty(nn, dtype=torch.float64, pin_memory=True) for _ in range(3)]
        ke = torch.empty(2, dtype=torch.float64, pin_memory=True)
    except RuntimeError:
        pinned = False
        blocks = [torch.empty(4 * n_local, dtype=torch.float64) for _ in range(2)]
        grids = [torch.empty(nn, dtype=torch.float64) for _ in range(3)]
        ke = torch.empty(2, dtype=torch.float64)
    ptr = lambda t: C.cast(t.data_ptr(), dp)  # noqa: E731
    quarters = [[b[k * n_local:(k + 1) * n_local] for k in range(4)] for b in blocks]
    for s in range(2):   # untimed: seed the host buffers with the current device state
        check(sim.L.picsp_species_download(sim.ctx, s, *(ptr(t) for t in quarters[s])))
        sim.computeKE(s)   # untimed: NCCL sets up the small-message path of the KE all-reduce lazily on its first call (one-off, ~1 s at 8 ranks)
    steps, periods = args.e2e_steps, args.e2e_periods
    root = rank == 0
    null = dp()
    # untimed: one dump with nothing else running (allocates the device snapshot; measures the copies alone)
    sim.sync(); barrier()
    ta = time.perf_counter()
    check(sim.L.picsp_dump_begin(sim.ctx, ptr(blocks[0]), ptr(blocks[1]), ptr(grids[0]) if root else null,
                                 ptr(grids[1]) if root else null, ptr(grids[2]) if root else null, ptr(ke)))
    check(sim.L.picsp_dump_wait(sim.ctx))
    dump_alone = max_over_ranks(time.perf_counter() - ta)
    for s in range(2):   # the dump overwrote the blocks with rows: seed them again
        check(sim.L.picsp_species_download(sim.ctx, s, *(ptr(t) for t in quarters[s])))
    barrier()
    t0 = time.perf_counter()
    for s in range(2):
        check(sim.L.picsp_species_upload(sim.ctx, s, *(ptr(t) for t in quarters[s]), n_local))
    t1 = time.perf_counter()
    wait_s = 0.0
    for p in range(periods):
        check(sim.L.picsp_step(sim.ctx, steps))            # enqueued; returns at once
        tw = time.perf_counter()
        check(sim.L.picsp_dump_wait(sim.ctx))              # the previous dump must have landed before its buffers are reused
        wait_s += time.perf_counter() - tw
        check(sim.L.picsp_dump_begin(sim.ctx, ptr(blocks[0]), ptr(blocks[1]), ptr(grids[0]) if root else null,
                                     ptr(grids[1]) if root else null, ptr(grids[2]) if root else null, ptr(ke)))
    t2 = time.perf_counter()
    check(sim.L.picsp_dump_wait(sim.ctx))
    sim.sync()
    t3 = time.perf_counter()
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    upload = max_over_ranks(t1 - t0)
    steady = max_over_ranks(t3 - t1)
    h2d = 2 * 4 * 8 * n_local
    d2h = periods * (2 * 4 * 8 * n_local + (3 * 8 * nn if root else 0) + 16)
    total_steps = steps * periods
    return {"value": float(args.particles) * total_steps / dt, "unit": UNIT,
            "h2d_bytes_per_step": h2d / total_steps, "d2h_bytes_per_step": d2h / total_steps,
            "steps_per_dump": steps, "dump_periods": periods, "seconds": dt, "pinned_host_memory": pinned,
            "upload_seconds": upload, "steady_state_seconds": steady, "one_dump_alone_seconds": dump_alone,
            "steady_state_value": float(args.particles) * total_steps / steady,
            "seconds_rank0": {"upload": t1 - t0, "periods_enqueue_and_dump_waits": t2 - t1, "waiting_for_dumps": wait_s,
                              "last_dump_drain": t3 - t2},
            "what": f"picsp_species_upload x2 (one-off, inside the timed region) -> {periods} x [picsp_step({steps}) -> picsp_dump_begin "
                    "(rows of both species, den.i, den.e, phi, KE; copied out while the next period steps)] -> picsp_dump_wait; "
                    "bytes are per rank, amortised over all steps; steady_state_* excludes the one-off upload",
            "ke_finite": bool(np.isfinite(ke.numpy()).all())}


_REAL_STDOUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else that writes to fd 1 from here on (NCCL's version
    banner, library chatter, the reference's own banner in the CPU arm) goes to stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--cells", type=int, default=1024)
    ap.add_argument("--particles", type=float, default=1e9, help="total particles (both species, all ranks)")
    ap.add_argument("--e2e-steps", type=int, default=50, help="steps per dump period (the reference dumps every 50 steps, main.cpp:507)")
    ap.add_argument("--e2e-periods", type=int, default=8, help="dump periods of the end-to-end measurement")
    ap.add_argument("--load", choices=["synthetic", "ref2"], default="synthetic",
                    help="particle load: bench-only device loader (uniform two-stream) or the reference's loadType 2 (diagonal) from the host loader")
    ap.add_argument("--walls", action="store_true", help="BASELINE.json configs[2]: bounded domain (PICSP_FLAG_WALLS | PICSP_FLAG_CLEAR_DENSITY): "
                    "absorbing walls, Dirichlet red-black SOR; extension WITHOUT reference semantics")
    ap.add_argument("--agg", type=int, default=None, choices=[-1, 0, 1], help="warp-aggregated deposit: -1 automatic (library default), 0 off, 1 on")
    ap.add_argument("--cell-period-e", type=int, default=-1, help="steps between electron cell orderings (-1: library default, 0: never)")
    ap.add_argument("--cell-period-i", type=int, default=-1, help="steps between ion cell orderings (-1: library default, 0: never)")
    ap.add_argument("--bank-order-e", type=int, default=None, choices=[-1, 0, 1], help="bank order inside the electron chunks after a re-binning: -1 automatic (library default: off), 0 off, 1 on")
    ap.add_argument("--bank-order-i", type=int, default=None, choices=[-1, 0, 1], help="the same for ions (library default: on)")
    ap.add_argument("--sort-period-e", type=int, default=0, help="steps between electron tile sorts (0: library default)")
    ap.add_argument("--sort-period-i", type=int, default=0, help="steps between ion tile sorts (0: library default)")
    ap.add_argument("--parts", type=int, default=0, help="picsp_params::parts: stores per species (0 = automatic: 1 unless the device memory asks for more; "
                    "BASELINE config 5, 4e9 particles, on ONE GPU runs with 8)")
    ap.add_argument("--flags", type=int, default=0, help="PICSP_FLAG_* bits for A/B runs (16: stand-alone re-sort instead of the re-binning mover)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-particles", type=int, default=4_000_000, help="particles/species of the CPU sample")
    ap.add_argument("--traffic-bytes-per-launch", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from the committed ncu --set full capture")
    args = ap.parse_args()
    args.particles = int(args.particles)
    if args.walls:
        args.flags |= 64 | 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return main_reference(args, rank)
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-launch under torchrun
            import subprocess
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    return main_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
