#!/usr/bin/env python
"""bench.py -- ms per EC3D timestep on the synthetic plate(N) grid.

Contract: ``python bench.py --gpus N --steps K --warmup W`` (under torchrun for N > 1) prints ONE
JSON line from rank 0.  A "step" is one pass of the hot path -- source scatter, inertial sources,
BiCGSTABwr solve with the matrix-free operator, history update (reference EC3D.f90:275-433) --
over the same grid at every N (strong scaling: the grid is z-slab partitioned across the ranks,
one NCCL halo exchange per SpMV and scalar all-reduces for the dot products).

  value / ms_per_step : device time (CUDA events on the library's stream, max over ranks) of K
                        steps with Uaf/Jaf resident in HBM, divided by K.
  e2e                 : the same K steps through the C ABI with host buffers: per-step source
                        scalars H2D from pinned memory and the full Uaf/Jaf fields D2H into pinned
                        memory after every step (what the Fortran host needs for its VTK output).
  roofline            : matrix-free SpMV (k_spmv_tma, the TMA-staged stencil kernel fused with
                        (As,s),(As,As)), algorithmic bytes 16 n + 5 nC over the
                        CUDA-event launch time measured here; peak from MEASURED_PEAKS.json.
  cpu_baseline        : the CPU oracle (oracle/, a port of the reference -- no Fortran compiler
                        exists here) timed on this host, 1 thread like the reference, on a bounded
                        sample; rank 0, N == 1 only.

``--impl reference`` times that CPU port alone (the reference arm).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ms per timestep (fp64 BiCGSTABwr, matrix-free A-U operator, synthetic plate(N))"
UNIT = "ms/step"
ITERS_FILE = os.path.join(ROOT, "profiles", "bench_iters.json")


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def oracle_ms_per_iter(grid: int, iters: int):
    """Bounded CPU sample: assemble plate(grid) with the oracle, then run `iters` BiCGSTABwr
    iterations of timestep 1 (itmax = iters-1 -> the reference runs itmax+1 iterations)."""
    from eddy_currents_3d_b200 import plate
    from oracle import oracle
    p = plate(grid, "A")
    t0 = time.perf_counter()
    run = oracle.OracleRun(p)
    t_asm = time.perf_counter() - t0
    f, v = p.source_scalars(0.0)
    run.step(f, v, solve=False)           # builds the step-1 right-hand side
    b = run.rhs
    x = np.zeros_like(b)
    # silence the reference's PRINT of the residual norm on iter > itmax
    sys.stdout.flush()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        t0 = time.perf_counter()
        it = oracle.bicgstabwr(run.A.valA, run.A.irow, run.A.jcol, b, x, 0.0, iters - 1)
        dt = time.perf_counter() - t0
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    return 1e3 * dt / max(it, 1), it, t_asm, p.nCells


def decks_section(nsteps: int = 10):
    """The three shipped decks (the only configurations the reference itself can run), first `nsteps`
    timesteps, measured for real -- no extrapolation: (a) GPU-resident stepping through ec3d_step; (b) the
    strict drop-in: the reference's time loop (CPU, oracle port standing in for the Fortran host) calling
    sprsbcgstabwr_ on its own CSR arrays with host buffers, solver calls timed; (c) the CPU baseline: the
    oracle port end to end, 1 thread, on this host.  ms per step = mean over steps 1..nsteps-1."""
    from eddy_currents_3d_b200 import lib
    from eddy_currents_3d_b200.problem import load_problem_npz
    from oracle import oracle
    out = {}
    for deck in ("compare_to_Elmer", "ec_src_move_hole", "LIM"):
        p = load_problem_npz(os.path.join(ROOT, "tests", "golden", deck + ".npz"))
        scal = [p.source_scalars(s * p.dt) for s in range(nsteps)]
        h = lib.Handle(p, device=0)
        its, ms = [], []
        for s in range(nsteps):
            t0 = time.perf_counter(); its.append(h.step(*scal[s])); ms.append(1e3 * (time.perf_counter() - t0))
        h.close()
        gpu = {"ms_per_step": float(np.mean(ms[1:])), "iters": its, "us_per_iteration": 1e3 * float(np.sum(ms[1:])) / max(sum(its[1:]), 1)}
        A = oracle.Assembled(p)
        tsolve = []

        def gpu_solver(valA, irow, jcol, n, b, x, tol, itmax):
            t0 = time.perf_counter()
            it = lib.sprsBCGstabWR(valA, irow, jcol, n, b, x, tol, itmax)
            tsolve.append(1e3 * (time.perf_counter() - t0))
            return it
        host = oracle.OracleRun(p, A, solver=gpu_solver)
        its_d = [host.step(*scal[s]) for s in range(nsteps)]
        lib.load().ec3d_csr_cache_clear()
        drop = {"solver_ms_per_step": float(np.mean(tsolve[1:])), "iters": its_d,
                "us_per_iteration": 1e3 * float(np.sum(tsolve[1:])) / max(sum(its_d[1:]), 1),
                "note": "CSR SpMV on the reference's arrays; b and x cross PCIe every call; first call uploads the matrix"}
        cpu = oracle.OracleRun(p, A)
        its_c, ms_c = [], []
        for s in range(nsteps):
            t0 = time.perf_counter(); its_c.append(cpu.step(*scal[s])); ms_c.append(1e3 * (time.perf_counter() - t0))
        out[deck] = {"unknowns": p.nCellsGlob, "steps": nsteps, "gpu_resident": gpu, "gpu_dropin": drop,
                     "cpu_oracle_1thread": {"ms_per_step": float(np.mean(ms_c[1:])), "iters": its_c,
                                            "ms_per_iteration": float(np.sum(ms_c[1:])) / max(sum(its_c[1:]), 1)},
                     "speedup_resident_vs_cpu": float(np.mean(ms_c[1:]) / np.mean(ms[1:]))}
    return out


def recorded_iters(grid: int, first: int = 0, count: int = 0):
    """Mean BiCGSTABwr iterations per timestep of the GPU arm on plate(grid), steps [first, first+count)
    when the record has them (the iteration counts are deterministic: same for every GPU count)."""
    try:
        with open(ITERS_FILE) as fh:
            d = json.load(fh)[str(grid)]
        by_step = d.get("iters_by_step")
        if by_step and count > 0 and first + count <= len(by_step):
            sel = by_step[first:first + count]
            return float(np.mean(sel)), f"recorded GPU-arm iteration counts of timesteps {first}..{first + count - 1}: {sel}"
        return float(d["mean_iters_per_step"]), d.get("source", "")
    except Exception:
        return None, ""


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (oracle port; the Fortran cannot be built
    here) on the host cores, same metric and unit, each step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    grid = args.grid
    sample_grid = min(args.ref_grid, grid)
    per = []
    t_asm = 0.0
    for s in range(args.warmup + args.steps):
        ms_it, it, t_asm, ncell = oracle_ms_per_iter(sample_grid, args.ref_iters)
        if s >= args.warmup:
            per.append(ms_it)
    ms_iter_sample = float(np.mean(per))
    scale = (grid / sample_grid) ** 3            # CSR SpMV + vector passes are linear in the cell count
    mean_it, src = recorded_iters(grid, args.warmup, args.steps)
    if mean_it is None:
        mean_it, src = 100.0, "assumed 100 iterations/step (no recorded GPU run)"
    ms_step = ms_iter_sample * scale * mean_it
    sample = (f"{args.ref_iters} BiCGSTABwr iterations of timestep 1 on plate({sample_grid}) with the "
              f"oracle's CSR (1 thread, like the serial reference); ms/step extrapolated: x{scale:.0f} "
              f"cells to plate({grid}), x{mean_it:.1f} iterations/step ({src}); assembly {t_asm:.1f} s not included")
    line = {
        "impl": "reference", "metric": METRIC, "value": ms_step, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"plate({grid}) variant A, tol 5e-3, itmax 10000", "grid": grid},
        "cpu_baseline": {"value": ms_step, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "ms_per_iteration_sample": ms_iter_sample},
        "e2e": {"value": ms_step, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=int(os.environ.get("EC3D_BENCH_GRID", "512")))
    ap.add_argument("--variant", default="A")
    ap.add_argument("--ref-grid", type=int, default=160, help="grid of the bounded CPU sample (reference arm)")
    ap.add_argument("--ref-iters", type=int, default=15)
    ap.add_argument("--cpu-grid", type=int, default=160, help="grid of the in-line cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-decks", action="store_true", help="skip the shipped-deck section (N = 1 only)")
    ap.add_argument("--record-iters", action="store_true", help="write profiles/bench_iters.json")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from eddy_currents_3d_b200 import lib, plate

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(lib.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    grid = args.grid
    p = plate(grid, args.variant)
    t0 = time.perf_counter()
    h = lib.Handle(p, nranks=world, rank=rank, nccl_id=nccl_id, device=local)
    t_create = time.perf_counter() - t0
    n, nC = p.nCellsGlob, p.nCells
    nsteps_total = args.warmup + args.steps
    scal = [p.source_scalars(s * p.dt) for s in range(nsteps_total)]
    # pinned host buffers: per-step scalars in, full fields out
    fsrc = torch.zeros(max(p.numfun, 1), dtype=torch.float64).pin_memory()
    # full-length host fields in the reference layout (every rank fills its owned entries).  Pinned, unless
    # all ranks together would pin more than 64 GB of host memory (plate(768) on 8 ranks): then pageable,
    # lazily committed memory -- only the owned ranges are ever touched
    pin_fields = 16.0 * n * world < 64e9
    U_host = torch.empty(n, dtype=torch.float64)
    J_host = torch.empty(n, dtype=torch.float64)
    if pin_fields:
        U_host, J_host = U_host.pin_memory(), J_host.pin_memory()

    def do_step(s):
        fsrc[:p.numfun] = torch.from_numpy(scal[s][0])
        return h.step_raw(fsrc.data_ptr(), 0)

    step = 0
    warm_iters = []
    for _ in range(args.warmup):
        warm_iters.append(do_step(step)); step += 1
        h.get_fields_raw(U_host.data_ptr(), J_host.data_ptr())
    # ---- ONE timed region, two clocks over the same K steps ----
    #   value : device time of the hot path (CUDA events recorded by ec3d_step on the library's stream
    #           around scatter + RHS + solve + history update), fields resident in HBM
    #   e2e   : wall clock around [step through the C ABI with pinned host scalars] + [D2H of the full
    #           Uaf / Jaf into pinned host memory], i.e. what the Fortran host sees per timestep
    c0 = h.counters()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    iters, dev_ms = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        iters.append(do_step(step)); step += 1
        dev_ms.append(h.counters()["last_step_ms"])
        h.get_fields_raw(U_host.data_ptr(), J_host.data_ptr())
    torch.cuda.synchronize()
    e2e_ms_total = 1e3 * (time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    c1 = h.counters()
    ms_total = float(sum(dev_ms))
    tt = torch.tensor([ms_total, e2e_ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total = float(tt[0]), float(tt[1])
    ms_step = ms_total / args.steps
    e2e_step = e2e_ms_total / args.steps

    # ---- per-kernel roofline: every rank times its own slab, the slowest rank of each kernel is reported ----
    peak, peak_kind = measured_peak_gbs()
    n_own = h.n_owned
    nC_own = (h.k1 - h.k0) * p.sdx * p.sdy
    kdefs = lib.KERNELS                      # (which, name, bytes per unknown, bytes per cell, reference lines)
    mine = [h.bench_kernel(k[0], 3, 20) for k in kdefs] + [float(n_own), float(nC_own)]
    tk = torch.tensor(mine, dtype=torch.float64, device="cuda")
    if world > 1:
        allk = [torch.zeros_like(tk) for _ in range(world)]
        dist.all_gather(allk, tk)
        allk = torch.stack(allk).cpu().numpy()
    else:
        allk = tk.cpu().numpy()[None, :]
    mean_it = float(np.mean(iters)) if iters else 0.0
    solve_ms_per_iter = (ms_total / max(sum(iters), 1))
    if rank != 0:
        h.close()
        if world > 1:
            dist.destroy_process_group()
        return
    kernels = {}
    for q, (which, name, bpn, bpc, what) in enumerate(kdefs):
        r = int(np.argmax(allk[:, q]))                       # the slowest rank sets the pace
        ms_k, nn, cc = float(allk[r, q]), float(allk[r, -2]), float(allk[r, -1])
        byt = bpn * nn + bpc * cc
        kernels[name] = {"ms": ms_k, "bytes": byt, "GBps": byt / (ms_k * 1e-3) / 1e9, "frac": byt / (ms_k * 1e-3) / 1e9 / peak,
                         "slowest_rank": r, "replaces": what}
        if world > 1:
            kernels[name]["ms_by_rank"] = [round(float(v), 4) for v in allk[:, q]]
    it_bytes_rank0 = 152.0 * n_own + 10.0 * nC_own
    # whole iteration against SURVEY 8d's UNCHANGED figure (19 passes + 2 map reads), slab of the largest rank
    it_bytes = max(152.0 * float(a[-2]) + 10.0 * float(a[-1]) for a in allk)
    dom = max(kernels, key=lambda k: kernels[k]["ms"])         # dominant kernel = largest share of an iteration
    kd = kernels[dom]
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            tj = json.load(fh)
        if world == 1 and str(grid) in tj and tj[str(grid)].get(dom):
            traffic = float(tj[str(grid)][dom])
            traffic_src = "committed ncu --set full capture (profiles/ncu_traffic.json), not measured in this run"
    except Exception:
        traffic = None
    line = {
        "metric": METRIC, "value": ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"plate({grid}) variant {args.variant}: {grid}^3 cells, n = {n} unknowns, "
                               f"tol 5e-3, itmax 10000, z-slab partition over {world} GPU(s)",
                   "grid": grid, "unknowns": n, "parallelism": f"zslab{world}",
                   "l2_policy": "inputs larger than L2 (every Krylov vector is %.1f GB per GPU)" % (8.0 * n_own / 1e9),
                   "iters_per_step": iters, "ms_per_iteration": solve_ms_per_iter,
                   "timing": "value = sum of the per-step CUDA-event times recorded by ec3d_step on the library's "
                             "stream (max over ranks); e2e = wall clock of the same steps incl. H2D scalars and "
                             "D2H of the full fields after every step",
                   "create_s": t_create},
        "e2e": {"value": e2e_step, "unit": UNIT, "h2d_bytes_per_step": 8 * (p.numfun + p.numMech),
                "d2h_bytes_per_step": 16 * n_own + 16, "iters_per_step": iters,
                "same_steps_as_value": True, "host_fields_pinned": bool(pin_fields)},
        "gpu_launches": int(c1["launches"] - c0["launches"]),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": kd["GBps"], "peak": peak, "unit": "GB/s", "frac": kd["frac"],
                     "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind,
                     "kernel": f"{dom} ({kd['replaces']}) -- the kernel with the largest share of an iteration",
                     "algorithmic_bytes": kd["bytes"], "ms_per_launch": kd["ms"], "slowest_rank": kd["slowest_rank"]},
        "roofline_iteration": {"bound": "hbm", "achieved": it_bytes / (solve_ms_per_iter * 1e-3) / 1e9,
                               "peak": peak, "unit": "GB/s",
                               "frac": it_bytes / (solve_ms_per_iter * 1e-3) / 1e9 / peak,
                               "algorithmic_bytes": it_bytes,
                               "note": "SURVEY 8d figure 152 n + 10 nC (19 vector passes + 2 map reads) of the largest slab "
                                       "over the measured time per BiCGSTABwr iteration of the timed solves; the kernels "
                                       "make fewer passes than that figure assumes (see DESIGN.md section 3)"},
        "kernels": kernels,
        "owned_unknowns_by_rank": [int(a[-2]) for a in allk],
        "sum_kernel_ms": float(sum(k["ms"] for k in kernels.values())),
    }
    if world == 1:
        try:
            d = {}
            if os.path.exists(ITERS_FILE):
                d = json.load(open(ITERS_FILE))
            d[str(grid) if args.variant == "A" else f"{grid}{args.variant}"] = {"mean_iters_per_step": float(np.mean(warm_iters + iters)),
                            "iters_by_step": warm_iters + iters, "warmup": args.warmup,
                            "source": "iteration counts of the GPU arm by timestep (deterministic; identical for 1/2/4/8 GPUs; "
                                      "equal to the oracle's where the oracle was run, see tests)"}
            os.makedirs(os.path.dirname(ITERS_FILE), exist_ok=True)
            out = os.path.join(ROOT, "gpurun_out", "bench_iters.json")
            os.makedirs(os.path.dirname(out), exist_ok=True)
            json.dump(d, open(out, "w"), indent=1)
        except Exception as e:  # noqa: BLE001
            print("record-iters failed:", e, file=sys.stderr)
    if world == 1 and not args.no_cpu:
        try:
            from oracle import oracle
            oracle.build()
            cg = min(args.cpu_grid, grid)
            ms_it, it, t_asm, _ = oracle_ms_per_iter(cg, 30)
            scale = (grid / cg) ** 3
            line["cpu_baseline"] = {
                "value": ms_it * scale * mean_it, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"30 BiCGSTABwr iterations on plate({cg}) with the oracle's CSR, 1 thread; extrapolated "
                          f"x{scale:.0f} cells and x{mean_it:.1f} iterations/step (this run's GPU count); "
                          f"oracle assembly {t_asm:.1f} s not included",
                "ms_per_iteration_sample": ms_it}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port", "sample": f"failed: {e}"}
    h.close()
    if world == 1 and not args.no_decks and not args.no_cpu:
        try:
            line["decks"] = decks_section()
        except Exception as e:  # noqa: BLE001
            line["decks"] = {"error": str(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
