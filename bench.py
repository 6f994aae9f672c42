#!/usr/bin/env python
"""bench.py -- fp64 particle-steps/s, MilkyWayPotential fixed-step "Leapfrog" (BASELINE.json metric).

One bench "step" = one pass of the hot path over one batch: every particle of the batch is integrated
from t = 0 to 1 Gyr with dt = 0.1 Myr (10 000 SemiImplicitEuler steps, the reference's "leapfrog",
SURVEY.md 8a-11) in MilkyWayPotential, final state saved.  The batch is C1's synthetic initial-condition
distribution scaled up to fill one B200 (C1's own 10^4 particles occupy 2 warps per SM; its number is
reported beside the headline as ``c1_exact``).  Weak scaling: every rank owns N_PER_GPU particles; the
only collective is the all-gather of the final states.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nproc-per-node 8 ... bench.py --gpus 8 --steps 5 --warmup 3
  python bench.py --impl reference     # the CPU restatement (oracle port) on the host cores

Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "fp64 particle-steps/s, MilkyWayPotential Leapfrog (SemiImplicitEuler dt=0.1 Myr, 1 Gyr)"
UNIT = "particle-steps/s"
N_PER_GPU = 148 * 8192  # 1 212 416 particles: 64 CTAs of 128 threads per SM
N_STEPS = 10_000  # dt0 = 0.1 Myr over 1 Gyr
T1, DT0 = 1000.0, 0.1
FLOP_PER_STEP = 280.0  # canonical weighted fp64 flop per MilkyWayPotential fixed step
# FP64 thread-instructions the kernel really executes per particle-step, counted by ncu on this very launch
# (profiles/ncu_k_integrate_fixed_r1f.txt: smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on / (N x steps))
NCU_FP64_INSTR_PER_STEP = {"dfma": 39.07, "dmul": 23.03, "dadd": 7.01}
NCU_FP64_PIPE_ACTIVE_PCT = 82.9
CPU_SAMPLE_PER_CORE = 8192  # cpu_baseline leg: ~10 s of CPU work at ~9e6 particle-steps/s/core
REF_ARM_SAMPLE_PER_CORE = 2048  # --impl reference: ~2.5 s per bench step


def workload_config(n_gpus: int, n: int = N_PER_GPU) -> dict:
    return {
        "workload": "C1-scaled: MilkyWayPotential, SemiImplicitEuler dt0=0.1 Myr, t=0..1 Gyr (10000 steps), "
        f"{n} particles per GPU, C1 synthetic ICs (r~U(4,20) kpc, |v|=v_c*U(0.6,1)), final state saved",
        "particles_per_gpu": n,
        "integrator_steps": N_STEPS,
        "parallelism": f"particle-sharded x{n_gpus}, all-gather of final states",
        "l2": "L2 flushed between timed iterations (256 MiB write)",
    }


def host_ics(n: int, seed: int):
    """C1 synthetic initial conditions on the host (numpy), v_c from a tabulated rotation curve of the model."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(4.0, 20.0, n)

    def iso(m):
        v = rng.normal(size=(m, 3))
        return v / np.linalg.norm(v, axis=1, keepdims=True)

    q = iso(n) * r[:, None]
    return q, r, iso(n), rng.uniform(0.6, 1.0, n)


# --------------------------------------------------------------------------------------------- reference arm


def run_reference(args) -> None:
    """The reference's CPU path for this metric.  galax (JAX/diffrax) cannot be installed in this image, so this
    times the repo's C restatement of the same algorithm (oracle/, OpenMP over particles) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cref
    from oracle import potentials as op

    cref.build()
    opot = op.milky_way_potential()
    cores = cref.use_all_cores()
    n = REF_ARM_SAMPLE_PER_CORE * cores
    q, r, vdir, f = host_ics(n, seed=1)
    p = vdir * (op.circular_velocity(opot, r) * f)[:, None]

    def step():
        cref.integrate_fixed(opot, q, p, 0.0, T1, DT0, [T1])

    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = n * N_STEPS / dt
    sample = f"{n} particles x {N_STEPS} steps per bench step (same ICs/potential/dt as the GPU arm)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "galax/JAX/diffrax are not installable offline; 'port' = oracle/galax_oracle.c (plain C, OpenMP)",
    }  # fmt: skip
    emit(line)


# --------------------------------------------------------------------------------------------- GPU arm


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, index: int):
        self.index = index
        self.rows: list[list[str]] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": sorted(reasons)}  # fmt: skip


def run_gpu(args) -> None:
    import torch
    import torch.distributed as dist

    import galax_b200
    import galax_b200.dynamics as gd
    import galax_b200.potential as gp
    from galax_b200 import _lib
    from galax_b200 import distributed as gdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()  # fails loudly if libgalax_b200.so is missing
    pot = gp.MilkyWayPotential()
    P = pot.c_struct()
    n = args.particles_per_gpu

    # ---- synthetic C1 initial conditions (rank-specific seed), host-pinned for the e2e path
    q_h, r_h, vdir, f = host_ics(n, seed=1 + rank)
    xr = np.stack([r_h, np.zeros_like(r_h), np.zeros_like(r_h)], axis=1)
    vc = np.sqrt(r_h * pot.gradient(xr)[:, 0])  # rotation curve from the product's own gradient kernel
    p_h = vdir * (vc * f)[:, None]
    q_pin = torch.from_numpy(q_h).pin_memory()
    p_pin = torch.from_numpy(p_h).pin_memory()
    q_d, p_d = q_pin.to(dev), p_pin.to(dev)
    ts_d = torch.tensor([T1], dtype=torch.float64, device=dev)
    q_out = torch.empty((n, 1, 3), dtype=torch.float64, device=dev)
    p_out = torch.empty_like(q_out)
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    gathered_q = torch.empty((world * n, 1, 3), dtype=torch.float64, device=dev) if world > 1 else None
    gathered_p = torch.empty_like(gathered_q) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    launches = 0
    kernel_ms: list[float] = []

    def device_step(record: bool):
        nonlocal launches
        flush.zero_()  # evict L2 between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.gx_integrate_fixed(C.byref(P), q_d.data_ptr(), p_d.data_ptr(), n, 0.0, T1, DT0, ts_d.data_ptr(), 1,
                                  _lib.SCHEME_SIE, -1, _lib.LAYOUT_NT3, q_out.data_ptr(), p_out.data_ptr(),
                                  status.data_ptr(), stream)  # fmt: skip
        e1.record()
        _lib.check(rc, "gx_integrate_fixed")
        launches += 1
        if world > 1:  # the job's only collective: gather the result shards
            dist.all_gather_into_tensor(gathered_q, q_out)
            dist.all_gather_into_tensor(gathered_p, p_out)
        if record:
            return e0, e1
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(args.warmup):
        device_step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches = 0
    t0 = time.perf_counter()
    evs = [device_step(True) for _ in range(args.steps)]
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    assert int((status != 0).sum()) == 0, "integration reported failures"
    gpu_launches = launches
    el = torch.tensor([elapsed], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed = float(el)
    value = world * n * N_STEPS * args.steps / elapsed

    # ---- end to end through the public API: pinned host arrays in, host arrays out
    solver = gd.OrbitSolver(solver=gd.SemiImplicitEuler(), stepsize_controller=gd.ConstantStepSize(), max_steps=None)

    def e2e_step():
        sol = solver.solve(pot, (q_pin, p_pin), 0.0, T1, dt0=DT0)  # H2D copies, launch, status check, D2H copies
        return sol.ys[0], sol.ys[1]

    qf = pf = None
    for _ in range(args.warmup):  # hold the previous result like the timed loop does (pinned-block recycling)
        qf, pf = e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_ms = []
    for _ in range(args.steps):
        ts0 = time.perf_counter()
        qf, pf = e2e_step()
        e2e_ms.append((time.perf_counter() - ts0) * 1e3)
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    if rank == 0 and os.environ.get("GX_BENCH_DEBUG"):
        print("e2e per-step ms:", [round(v, 2) for v in e2e_ms], file=sys.stderr)
    el = torch.tensor([e2e_elapsed], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    e2e_value = world * n * N_STEPS * args.steps / float(el)
    assert not qf.is_cuda and qf.shape == (n, 1, 3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rank 0 extras: FP64 peak (live), C1 exact, energy drift, CPU baseline
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    nf = C.c_int64()
    best = None
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        L.gx_bench_dfma(148 * 8, 256, 20000, sink.data_ptr(), C.byref(nf), stream)
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) * 1e-3
        best = t if best is None else min(best, t)
    dfma_peak = 2.0 * nf.value * 148 * 8 * 256 / best / 1e12
    best3 = None
    for _ in range(3):  # same chains with three distinct register operands per DFMA: the register-file-fed rate
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        L.gx_bench_dfma(-148 * 8, 256, 20000, sink.data_ptr(), C.byref(nf), stream)
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) * 1e-3
        best3 = t if best3 is None else min(best3, t)
    dfma_peak_3reg = 2.0 * nf.value * 148 * 8 * 256 / best3 / 1e12

    k_ms = float(np.mean(kernel_ms))
    per_gpu_rate = n * N_STEPS / (k_ms * 1e-3)
    achieved = per_gpu_rate * FLOP_PER_STEP / 1e12
    roofline = {
        "bound": "fp64", "achieved": achieved, "peak": dfma_peak, "unit": "TFLOP/s", "frac": achieved / dfma_peak,
        "traffic": 78.8e6 if n == N_PER_GPU else None,
        "traffic_source": "ncu --set full, profiles/ncu_k_integrate_fixed_r1f.txt: dram read 59.8 MB + write 19.0 MB per "
                          "launch (algorithmic: 58.2 MB in + 58.2 MB out + 4.8 MB status; L2 absorbs part of the writes)",
        "kernel": "k_integrate_fixed_seg<MW> (SemiImplicitEuler, run-length time grid)", "kernel_ms": k_ms,
        "algorithmic_flop_per_particle_step": FLOP_PER_STEP,
        "peak_three_register_operands": dfma_peak_3reg,
        "peak_source": "measured live: gx_bench_dfma (8 independent DFMA chains/thread), best of 3; "
                       "MEASURED_PEAKS.json has no FP64 entry",
        "note": "canonical weighted flop (div/sqrt=18, log1p=56) per SURVEY.md 8d; the kernel issues fewer real "
                "instructions than that (MUFU-seeded rcp/rsqrt): see 'executed' for the ncu-counted rate",
    }  # fmt: skip
    n_instr = sum(NCU_FP64_INSTR_PER_STEP.values())
    n_flop = 2.0 * NCU_FP64_INSTR_PER_STEP["dfma"] + NCU_FP64_INSTR_PER_STEP["dmul"] + NCU_FP64_INSTR_PER_STEP["dadd"]
    roofline["executed"] = {
        "fp64_instr_per_particle_step": NCU_FP64_INSTR_PER_STEP,
        "tflops": per_gpu_rate * n_flop / 1e12,                         # dfma x 2 + dmul + dadd, as ncu counts flop
        "frac_of_dfma_peak": per_gpu_rate * n_flop / 1e12 / dfma_peak,
        "fp64_issue_frac": per_gpu_rate * n_instr / (dfma_peak * 1e12 / 2.0),  # FP64 instructions / FP64 issue slots
        "ncu_fp64_pipe_active_pct": NCU_FP64_PIPE_ACTIVE_PCT,
        "source": "profiles/ncu_k_integrate_fixed_r1f.txt (same kernel, same launch shape)",
    }

    # The HBM-bound leg of the path (C5): acceleration + Hessian on 2e7 points, against the measured copy peak
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        hbm_peak, hbm_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json (of measured)"
    else:
        hbm_peak, hbm_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    npts = 20_000_000
    xs = torch.randn(npts, 3, dtype=torch.float64, device=dev) * 10
    acc_o = torch.empty((npts, 3), dtype=torch.float64, device=dev)
    hes_o = torch.empty((npts, 9), dtype=torch.float64, device=dev)
    k1 = []
    for i in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        L.gx_potential_eval(C.byref(P), xs.data_ptr(), 0.0, npts, _lib.ACC | _lib.HESS, None, None, acc_o.data_ptr(),
                            hes_o.data_ptr(), stream)
        b.record()
        torch.cuda.synchronize()
        if i:
            k1.append(a.elapsed_time(b) * 1e-3)
    k1_gbs = npts * 120 / float(np.mean(k1)) / 1e9
    roofline_k1 = {"bound": "hbm", "achieved": k1_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k1_gbs / hbm_peak,
                   "traffic": None, "kernel": "k_potential_eval<MW> (acc + Hessian, TMA-staged tiles)",
                   "algorithmic_bytes_per_point": 120, "points": npts, "peak_source": hbm_src}  # fmt: skip
    del xs, acc_o, hes_o

    # C1 exactly as stated: 10^4 particles
    qc, pc = q_d[:10_000].contiguous(), p_d[:10_000].contiguous()
    qo, po = torch.empty((10_000, 1, 3), dtype=torch.float64, device=dev), torch.empty((10_000, 1, 3), dtype=torch.float64, device=dev)
    c1 = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        L.gx_integrate_fixed(C.byref(P), qc.data_ptr(), pc.data_ptr(), 10_000, 0.0, T1, DT0, ts_d.data_ptr(), 1,
                             _lib.SCHEME_SIE, -1, _lib.LAYOUT_NT3, qo.data_ptr(), po.data_ptr(), None, stream)
        b.record()
        torch.cuda.synchronize()
        if i:
            c1.append(a.elapsed_time(b) * 1e-3)
    c1_rate = 10_000 * N_STEPS / min(c1)

    E0 = gd._energy(pot, q_d, p_d)
    E1 = gd._energy(pot, q_out[:, 0], p_out[:, 0])
    drift = (E1 / E0 - 1).abs()
    energy = {"median": float(drift.median()), "p99": float(drift.quantile(0.99)), "max": float(drift.max())}

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import cref
        from oracle import potentials as op

        cref.build()
        opot = op.milky_way_potential()
        ns = min(n, CPU_SAMPLE_PER_CORE * cref.use_all_cores())
        cref.integrate_fixed(opot, q_h[:256], p_h[:256], 0.0, 10.0, DT0, [10.0])  # warm
        t0 = time.perf_counter()
        qr, pr, st, _ = cref.integrate_fixed(opot, q_h[:ns], p_h[:ns], 0.0, T1, DT0, [T1])
        dt = time.perf_counter() - t0
        e = np.linalg.norm(q_out[:ns, 0].cpu().numpy() - qr[:, 0], axis=1) / np.linalg.norm(qr[:, 0], axis=1)
        cpu = {"value": ns * N_STEPS / dt, "unit": UNIT, "cores": cref.num_threads(), "kind": "port",
               "sample": f"first {ns} particles of rank 0's batch x {N_STEPS} steps, oracle/galax_oracle.c (OpenMP)",
               "parity_vs_gpu": {"median_rel": float(np.median(e)), "p99_rel": float(np.quantile(e, 0.99)),
                                 "frac_le_1e-12": float(np.mean(e <= 1e-12))}}  # fmt: skip

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(world, n), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 24 * world,
                "d2h_bytes_per_step": (2 * n * 24 + n * 4) * world,
                "api": "galax_b200.dynamics.OrbitSolver(SemiImplicitEuler, ConstantStepSize).solve(pot, (q, p), 0, 1000, dt0=0.1) "
                       "with pinned host tensors"},
        "gpu_launches": gpu_launches, "roofline": roofline, "roofline_hbm_leg": roofline_k1, "cpu_baseline": cpu,
        "c1_exact": {"particles": 10_000, "value": c1_rate, "unit": UNIT, "ms": min(c1) * 1e3},
        "energy_drift": energy, "fp64_peak_tflops_measured": dfma_peak,
        "fp64_peak_tflops_measured_3reg_operands": dfma_peak_3reg,
    }  # fmt: skip
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line: dict) -> None:
    """The one JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # stdout carries exactly one JSON line.  Libraries write banners there (NCCL prints "NCCL version ..." on the first
    # communicator when NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for the
    # duration of the run and the line goes to a saved copy of the original descriptor.
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="galax_b200", choices=["galax_b200", "reference"])
    ap.add_argument("--particles-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
