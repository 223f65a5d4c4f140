#!/usr/bin/env python
"""bench.py — headline benchmark of the AutoBZCore hot path on B200 (contract: see the build prompt).

Workload (config.workload): BASELINE.json config 4 — synthetic Wannier Hamiltonian norb=32, R in [-8,8]^3
(M=17 per dimension), PTR 256^3 k-grid, 128-point frequency sweep, integrand tr[(w + i eta - H(k))^-1].
One "step" = each rank evaluates its k3 slab of `--planes` planes (default 32 = 256/8, so that N=8 ranks
cover the whole 256^3 grid; weak scaling) for all 128 frequencies: Fourier stages 3/2/1 + resolvent + weighted
k-sum.  metric = k-points/sec (H(k)+resolvent), whole-job aggregate over ranks.

  value : device-resident (H_R already in HBM; rule built), CUDA events on the library's stream
  e2e   : the same through the public API (IntegralSolver + batchsolve) with HOST buffers: H_R uploaded,
          rule built, sums, results read back, partial sums all-reduced — every step
  --impl reference : the CPU restatement of the reference path (oracle, OpenMP over k3 planes like the
          reference's Threads.@threads loop) on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NORB, RMAX, NPT, NW = 32, 8, 256, 128
FUSED_DRAM_BYTES_PER_KPOINT = 1177.5    # (71.92 MB read + 5.25 MB written) / 65 536 k-points of one launch, profiles/r02_ncu_full_fused_raw.csv
METRIC = "k-points/sec (H(k)+resolvent)"


def workload_inputs():
    import autobz_b200 as ab
    H, lo = ab.synthetic.wannier_hamiltonian(NORB, RMAX)
    ext = ab.synthetic.band_extent(H)
    # the spectrum of the synthetic H(k) lies well inside [-ext, ext]; sweep the central part, eta = 1e-2 * bandwidth scale
    bw = 0.25 * ext
    omegas = np.linspace(-bw, bw, NW)
    eta = 1e-2 * 2 * bw
    return H, lo, omegas, eta


def flops_per_node(n=NORB, M=2 * RMAX + 1, N=NPT, nw=NW):
    f_four = 8.0 * n * n * M * (1 + M / N + (M / N) ** 2)
    f_res = 8.0 * n ** 3 * nw
    return f_four, f_res


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); out["sm_max_mhz"] = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            busy = [x for x in sm if x > 0.5 * max(sm)]
            out["sm_mhz"] = float(np.median(busy))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def measure_fp64_peak():
    """cuBLAS ZGEMM 4096^3 burst (the FP64 denominator MEASURED_PEAKS.json lacks)."""
    import torch
    a = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda")
    b = torch.randn(4096, 4096, dtype=torch.complex128, device="cuda")
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    torch.cuda.empty_cache()
    return 8.0 * 4096 ** 3 / best * 1e-9


def cpu_sample(threads=None, target_s=15.0):
    """CPU restatement on a bounded sample: `threads` k3 planes (one per thread, as the reference threads the
    outermost k3 loop, src/fourier.jl:156) x r k2-rows x 256 k1 x 128 frequencies."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    orc.build()
    H, lo, omegas, eta = workload_inputs()
    S = orc.Series(H, lo)
    z = omegas + 1j * eta
    cores = threads or os.cpu_count() or 1
    planes = min(cores, NPT)
    # calibrate on a small sub-sample (first 16 frequencies, one row per plane)
    t0 = time.time()
    orc.ptr_sum(S, NPT, z[:16], k3_lo=0, k3_hi=planes, k2_lo=0, k2_hi=1, nthreads=cores)
    t_cal = time.time() - t0
    per_row = t_cal * (NW / 16.0)
    rows = int(max(1, min(NPT, target_s / max(per_row, 1e-9))))
    t0 = time.time()
    orc.ptr_sum(S, NPT, z, k3_lo=0, k3_hi=planes, k2_lo=0, k2_hi=rows, nthreads=cores)
    t = time.time() - t0
    nodes = planes * rows * NPT
    return {"value": nodes / t, "unit": "k-points/s", "cores": cores, "kind": "port",
            "sample": f"{planes} k3-planes x {rows} k2-rows x {NPT} k1 = {nodes} k-points x {NW} freqs of the C4 workload in {t:.1f} s "
                      f"(C/OpenMP restatement of the AutoBZCore 0.3.8 path, not Julia)",
            "lapack_cross_check": lapack_resolvent_rate(orc, S, z, cores)}, t


def lapack_resolvent_rate(orc, S, z, cores, seconds=2.0):
    """The per-(k, omega) work of the reference's integrand with LAPACK itself, as Julia's `inv(::Matrix)` does it (zgetrf + zgetri
    through numpy.linalg.inv), on H(k) of the workload: one core, then scaled to the box's cores - printed beside the port's own
    unblocked LU so that the CPU baseline is not flattered by the port's LU."""
    try:
        rng = np.random.default_rng(0)
        Hk = np.moveaxis(orc.eval_points(S, rng.random((16, 3))), 2, 0)          # [16, n, n]
        A = z[:, None, None, None] * np.eye(NORB)[None, None] - Hk[None]          # [nw, 16, n, n]
        np.trace(np.linalg.inv(A[:4]), axis1=2, axis2=3)
        cnt, t0 = 0, time.time()
        while time.time() - t0 < seconds:
            np.trace(np.linalg.inv(A), axis1=2, axis2=3)
            cnt += A.shape[0] * A.shape[1]
        r = cnt / (time.time() - t0)
        return {"inverse_traces_per_s_one_core": r, "kpoints_per_s_scaled_to_cores": r / NW * cores,
                "note": "numpy.linalg.inv (LAPACK zgesv-class) on 32x32 H(k) of the workload, Fourier interpolation not included"}
    except Exception as e:          # a cross-check must never cost the baseline
        return {"error": repr(e)}


def oracle_plane_sum(orc, So, N, z, k3):
    """the oracle's PTR sum over ONE k3 plane, k2 rows dealt to host threads (ctypes releases the GIL)"""
    from concurrent.futures import ThreadPoolExecutor
    nt = max(1, min(os.cpu_count() or 1, N))
    bounds = [(i * N // nt, (i + 1) * N // nt) for i in range(nt)]
    with ThreadPoolExecutor(nt) as ex:
        parts = list(ex.map(lambda b: orc.ptr_sum(So, N, z, k3_lo=k3, k3_hi=k3 + 1, k2_lo=b[0], k2_hi=b[1], nthreads=1), bounds))
    return sum(parts)


def parity_check(ctx, S, H, lo, z, k3):
    """Parity evidence carried by the bench line itself: one whole k3 plane of the timed workload (same series, same full-grid
    rule interface, kernels and frequency count as the timed steps), 4 of the 128 frequencies compared, GPU against the CPU oracle."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    from autobz_b200 import _lib as L
    orc.build()
    zi = [0, NW // 3, (2 * NW) // 3, NW - 1]
    z4 = np.ascontiguousarray(z[zi])
    Rc = L.DeviceRule(ctx, S, NPT, k3_lo=k3, k3_hi=k3 + 1)
    # the GPU runs the plane with ALL frequencies - the timed step's kernel configuration (the fused kernel serves nw >= 8) - and the
    # oracle checks 4 of them
    got = Rc.resolvent_sum(np.ascontiguousarray(z), scale=1.0 / NPT ** 3)[zi]
    Rc.close()
    t0 = time.time()
    ref = oracle_plane_sum(orc, orc.Series(H, lo), NPT, z4, k3)
    err = float(np.max(np.abs(got - ref) / np.abs(ref)))
    return {"rel_err_vs_cpu": err, "tolerance": 1e-10, "what": f"k3 plane {k3} of the timed slab ({NPT * NPT} k-points, full-grid rule interface) at frequencies "
            f"{zi} of the sweep, GPU sum vs CPU oracle (pivoted LU)", "gpu": [[float(v.real), float(v.imag)] for v in got],
            "cpu": [[float(v.real), float(v.imag)] for v in ref], "cpu_seconds": round(time.time() - t0, 2)}


def other_configs(ctx, shard=None, world=1):
    """The other BASELINE.json configs (C2, C3, C5) through the public API, a few seconds in total; reported beside the headline
    under "other_configs" (parity for them lives in tests/, CPU baselines in tools/run_configs.py).  With world > 1 every rank
    takes part (k3 planes / outermost IAI nodes dealt to the ranks, one small allreduce per rule or outer step): the SAME problems
    on N GPUs, i.e. strong scaling - the driver's scaling file then carries them for every N."""
    import autobz_b200 as ab
    out = []
    kw = {} if shard is None else {"shard": shard}
    d = np.load(os.path.join(ROOT, "tests", "golden", "svo_hr.npz"))
    Hs, los, A = np.asfortranarray(d["H_R"]), tuple(int(x) for x in d["lo"]), d["A"]
    fs = ab.FourierSeries(Hs, period=1.0, lo=los, norb=3)
    ibz = ab.load_bz(ab.CubicSymIBZ(), A)

    reps_ms = []

    def best(fn, reps=3):
        t_best, r = 1e30, None
        reps_ms.clear()
        for _ in range(reps):
            t = time.perf_counter(); r = fn(); dt = time.perf_counter() - t
            reps_ms.append(round(1e3 * dt, 3))
            t_best = min(t_best, dt)
        return r, t_best

    f2 = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=1e-2)
    solver = ab.IntegralSolver(f2, ibz, ab.PTR(npt=400), **kw)
    ws = [{"omega": w} for w in np.linspace(11.0, 14.0, 64)]
    # the CPU baseline leaves the GPU idle for seconds: bring the clocks back before timing millisecond-sized solves
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 0.75:
        ab.batchsolve(solver, ws)
    _, t = best(lambda: ab.batchsolve(solver, ws))
    nn = len(solver.cache.cacheval["rule"])
    out.append({"config": "C2 SrVO3 Green's-function trace, PTR npt=400 on CubicSymIBZ, 64 freqs, eta=1e-2", "n_gpus": world, "irreducible_kpoints": nn,
                "ms": 1e3 * t, "kpoints_per_s": nn / t, "k_omega_per_s": 64 * nn / t, "fbz_equivalent_kpoints_per_s": 400 ** 3 / t})
    sol, t = best(lambda: ab.solve(ab.IntegralProblem(f2, ibz, {"omega": 12.5}), ab.EvalCounter(ab.AutoPTR(a=1e-2, nmin=50, nmax=1000)), abstol=1e-3, **kw), 6)
    out.append({"config": "C2 SrVO3 AutoPTR(a=eta=1e-2) on CubicSymIBZ, omega=12.5, abstol=1e-3 (rule construction included)", "n_gpus": world,
                "numevals": sol.numevals, "ms": 1e3 * t, "kpoints_per_s": sol.numevals / t, "ms_all_repetitions": list(reps_ms)})
    f3 = ab.FourierIntegrand(ab.dos_integrand, fs, 1e-4)
    sol, t = best(lambda: ab.solve(ab.IntegralProblem(f3, ibz, 12.0), ab.EvalCounter(ab.IAI()), abstol=1e-3, **kw), 2)
    out.append({"config": "C3 SrVO3 DOS via IAI, eta=1e-4, omega=12.0, abstol=1e-3", "n_gpus": world, "numevals": sol.numevals, "s": t, "evals_per_s": sol.numevals / t})
    H5, lo5 = ab.synthetic.wannier_hamiltonian(64, 4, cubic=True)
    f5 = ab.FourierIntegrand(ab.EigenIntegrand("fermi_energy"), ab.FourierSeries(H5, period=1.0, lo=lo5, norb=64), 0.0, 0.5)
    ibz5 = ab.load_bz(ab.CubicSymIBZ(), 2 * np.pi * np.eye(3))
    cache = ab.init(ab.IntegralProblem(f5, ibz5), ab.PTR(npt=96), **kw)
    ab.solve_(cache)
    _, t = best(lambda: ab.solve_(cache))
    nn = len(cache.cacheval["rule"])
    out.append({"config": "C5 norb=64 band-energy integrand (Hermitian eigenvalues) on CubicSymIBZ, PTR npt=96", "n_gpus": world, "irreducible_kpoints": nn,
                "ms": 1e3 * t, "kpoints_per_s": nn / t, "fbz_equivalent_kpoints_per_s": 96 ** 3 / t,
                "device_eig_ms": ctx.last_timings()[1], "device_eval_ms": ctx.last_timings()[0]})
    cache.cacheval["rule"].close()
    # BASELINE config 5 as written: AutoPTR on the symmetry-reduced IBZ, grids 48 -> 96 -> 144 (rule construction included)
    alg5 = ab.EvalCounter(ab.AutoPTR(a=1.0, nmin=48, nmax=1000, n0=48.0, dn=48.0))
    sol, t = best(lambda: ab.solve(ab.IntegralProblem(f5, ibz5), alg5, reltol=1e-12, maxiters=80000, **kw), 2)
    out.append({"config": "C5 norb=64 band-energy integrand, AutoPTR 48 -> 96 -> 144 on CubicSymIBZ (rule construction included)", "n_gpus": world,
                "numevals": sol.numevals, "ms": 1e3 * t, "kpoints_per_s": sol.numevals / t})
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    vals, times = [], []
    cb = None
    for i in range(args.warmup + args.steps):
        cb, t = cpu_sample(target_s=args.ref_seconds)
        if i >= args.warmup:
            vals.append(cb["value"]); times.append(t)
    v = float(np.mean(vals))
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "k-points/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C4: synthetic Wannier H norb={NORB}, R in [-{RMAX},{RMAX}]^3, PTR {NPT}^3, {NW} freqs; bounded CPU sample per step"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "k-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--planes", type=int, default=NPT // 8, help="k3 planes per rank per step")
    ap.add_argument("--ref-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--algo", type=int, default=0, help="resolvent algorithm: 0 auto, 1 generic, 2 DMMA")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the GPU-vs-CPU parity check of one plane of the timed workload")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import autobz_b200 as ab
    from autobz_b200 import _lib as L

    for attempt in range(4):            # a freshly vacated GPU has been seen to refuse cuInit once ("CUDA driver initialization failed")
        try:
            torch.cuda.set_device(local)
            break
        except RuntimeError:
            if attempt == 3:
                raise
            time.sleep(5.0)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = ab.default_context(local)
    ctx.set_option(L.OPT_RESOLVENT_ALGO, args.algo)
    H, lo, omegas, eta = workload_inputs()
    z = omegas + 1j * eta
    planes = args.planes
    k3_lo = (rank * planes) % NPT
    k3_hi = k3_lo + planes
    nodes_rank = planes * NPT * NPT
    fp64_peak = measure_fp64_peak() if rank == 0 else None

    # ---------------- device-resident arm: H_R in HBM, rule built, time the sums
    S = L.DeviceSeries(ctx, H, lo, (1.0,) * 3)
    R = L.DeviceRule(ctx, S, NPT, k3_lo=k3_lo, k3_hi=k3_hi)
    # the step's one exchange: partial sums of the ranks' slabs meet in a 2 KB allreduce (NCCL over NVLink), inside the timed loop
    step_allreduce = ab.torch_allreduce(torch.device("cuda", local)) if world > 1 else (lambda a: a)
    for _ in range(args.warmup):
        step_allreduce(R.resolvent_sum(z, scale=1.0 / NPT ** 3))
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launch_count
    ev_eval = ev_mat = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        part = R.resolvent_sum(z, scale=1.0 / NPT ** 3)
        a, b = ctx.last_timings()
        ev_eval += a; ev_mat += b
        total = step_allreduce(part)
    barrier()
    t_dev = time.perf_counter() - t0
    launches = ctx.launch_count - l0
    clocks = sampler.stop() if sampler else None
    tt = torch.tensor([t_dev, (ev_eval + ev_mat) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev_max, t_events_max = float(tt[0]), float(tt[1])
    # ---- the same step with the opt-in frequency-sweep path (ABZ_OPT_RESOLVENT_ALGO = 3: one Householder tridiagonalisation per
    # k, then p'(z)/p(z) per frequency; Hermitian H(k), scalar self-energy only) - reported beside the headline, never as it
    sweep = None
    if rank == 0 and world == 1 and args.algo == 0:
        try:
            ctx.set_option(L.OPT_RESOLVENT_ALGO, 3)
            fast = R.resolvent_sum(z, scale=1.0 / NPT ** 3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                fast = R.resolvent_sum(z, scale=1.0 / NPT ** 3)
            t_fast = (time.perf_counter() - t0) / args.steps
            ev_f, mat_f = ctx.last_timings()
            sweep = {"value": nodes_rank / t_fast, "unit": "k-points/s", "ms_per_step": 1e3 * t_fast, "eval_ms_per_step": ev_f,
                     "matfun_ms_per_step": mat_f, "max_rel_diff_vs_lu_path": float(np.max(np.abs(fast - part) / np.abs(part))),
                     "note": "opt-in algorithm 3: tr (z-H)^-1 = p'(z)/p(z) of the tridiagonalised H(k); needs Hermitian H(k) and a scalar "
                             "self-energy, so it is NOT the headline (the LU path serves a matrix Sigma(omega))"}
        except Exception as e:
            sweep = {"error": repr(e)}
        finally:
            ctx.set_option(L.OPT_RESOLVENT_ALGO, args.algo)
    check = None
    if rank == 0 and not args.no_check:
        check = parity_check(ctx, S, H, lo, z, k3_lo + planes // 2)
    R.close(); S.close()

    # ---------------- end-to-end arm through the public API with host buffers
    shard = ab.Shard(rank, world, ab.torch_allreduce(torch.device("cuda", local))) if world > 1 else ab.Shard()
    # weak scaling: this run's grid is the slab set [0, world*planes) of the 256^3 grid; expressed through the public API
    # as a PTR rule whose k3 range is sharded over ranks.  The full-grid API shards npt planes over nranks, so use a
    # virtual world of NPT/planes ranks of which the first `world` exist.
    vworld = max(1, NPT // planes)

    class _VShard(ab.Shard):
        def __init__(self):
            super().__init__(rank, vworld, None)

        def allreduce(self, arr):
            return shard.allreduce(arr)

    bz = ab.load_bz(ab.FBZ(), 2 * np.pi * np.eye(3))
    plist = [{"omega": float(w)} for w in omegas]

    e2e_phases = []

    def e2e_step():
        t0 = time.perf_counter()
        fs = ab.FourierSeries(H, period=1.0, lo=lo, norb=NORB)            # host buffer -> uploaded inside
        f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=eta)
        solver = ab.IntegralSolver(f, bz, ab.PTR(npt=NPT), shard=_VShard())
        t1 = time.perf_counter()
        out = ab.batchsolve(solver, plist)
        t2 = time.perf_counter()
        solver.cache.cacheval["rule"].close()
        fs.drop_device()
        e2e_phases.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (time.perf_counter() - t2)) + tuple(ctx.last_timings()))
        return out

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g = e2e_step()
    barrier()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e_max = float(te[0])

    others = None
    if not args.no_other_configs:
        try:                            # collective at world > 1: every rank takes part
            others = other_configs(ctx, shard if world > 1 else None, world)
        except Exception as e:          # never lose the headline line to a secondary measurement
            others = [{"error": repr(e)}]
    if rank == 0:
        total_nodes = nodes_rank * world * args.steps
        value = total_nodes / t_dev_max
        e2e_val = total_nodes / t_e2e_max
        f_four, f_res = flops_per_node()
        # dominant kernel: the resolvent (matfun phase).  One launch = one chunk of nodes x 128 frequencies.
        mat_s = ev_mat * 1e-3 / args.steps
        achieved = f_res * nodes_rank / mat_s * 1e-12
        fused = os.environ.get("ABZ_FUSED_MMA", "1") != "0" and args.algo in (0, 2)
        roof = {"bound": "tensor", "kernel": "resolvent_mma_fused_kernel (FP64, DMMA/DFMA pipe; stage 1 of the Fourier evaluation folded in)" if fused
                else "resolvent (FP64, DMMA/DFMA pipe)", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak,
                # dram__bytes_read+write of one resolvent launch, from the ncu --set full capture in profiles/, scaled to this run's
                # launch size.  Fused: the C1 rows (M n^2 16 B per (k2,k3) row = 1.09 kB per k-point at N = 256) are read once, H(k) never
                # exists in HBM.  Unfused: 16.68 kB per k-point = the 16 n^2 B of H(k) read once for all 128 frequencies.
                "traffic": (FUSED_DRAM_BYTES_PER_KPOINT * nodes_rank) if fused else 16678.0 * min(nodes_rank, 262144),
                "traffic_source": "profiles/r02_ncu_full_fused_raw.csv" if fused else "profiles/r01_ncu_full_resolvent_mma_raw.csv",
                "peak_source": "cuBLAS ZGEMM 4096^3 burst measured in this run (MEASURED_PEAKS.json has no FP64 entry; a 4 s sustained ZGEMM "
                               "measures the same 36.8 TFLOP/s, profiles/r01_fp64_peaks.json)",
                "algorithmic_flops_per_kpoint": {"fourier": f_four, "resolvent": f_res},
                "eval_ms_per_step": ev_eval / args.steps, "matfun_ms_per_step": ev_mat / args.steps}
        cb = None
        if not args.no_cpu_baseline:
            cb, _ = cpu_sample(target_s=args.ref_seconds)
        line = {"metric": METRIC, "value": value, "unit": "k-points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_dev_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C4: synthetic Wannier H norb={NORB}, R in [-{RMAX},{RMAX}]^3 (M=17), PTR {NPT}^3 grid, {NW}-point frequency sweep; "
                                       f"step = ONE k3 slab of {planes} planes per rank = {planes}/{NPT} of the grid per GPU per step (x{world} ranks; 8 ranks = the whole "
                                       f"grid per step; the N=1 line is a 1/8-slab rate, equal to the full-grid rate by linearity of the k-sum)",
                           "kpoints_per_step": nodes_rank * world, "k_omega_evals_per_sec": value * NW,
                           "l2": ("inputs larger than L2 (per step the stage-2 output C1, 2.3 GB for 32 planes, is written and streamed once by the fused kernel; "
                                  "H(k) itself is never in HBM)") if os.environ.get("ABZ_FUSED_MMA", "1") != "0" and args.algo in (0, 2)
                           else "inputs larger than L2 (H(k) chunk 1-4 GB per pass)", "parallelism": f"k3-slab x{world}", "exchange": "one allreduce of 128 complex partial sums per step inside the timed loop" if world > 1 else "none (1 rank)",
                           "resolvent_algo": args.algo, "device_event_ms_per_step": 1e3 * t_events_max / args.steps},
                "roofline": roof, "cpu_baseline": cb,
                "e2e": {"value": e2e_val, "unit": "k-points/s", "h2d_bytes_per_step": int(H.nbytes + z.nbytes), "d2h_bytes_per_step": int(NW * 16),
                        "ms_per_step": 1e3 * t_e2e_max / args.steps,
                        "phases_ms_per_timed_step": [dict(zip(("upload_and_rule", "batchsolve", "teardown", "device_eval", "device_matfun"),
                                                              [round(x, 2) for x in ph])) for ph in e2e_phases[-args.steps:]]},
                "gpu_launches": int(launches), "clocks": clocks, "other_configs": others, "frequency_sweep_fast_path": sweep,
                "check": dict(check or {}, G_first=[float(g[0].real), float(g[0].imag)])}
        if check is not None and not (check["rel_err_vs_cpu"] <= check["tolerance"]):
            print(f"bench.py: parity check FAILED, no bench line printed: {json.dumps(check)}", file=sys.stderr, flush=True)
            sys.exit(3)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
