#!/usr/bin/env python
"""bench.py — EM cell-iterations/s of the robust multi-sample GMM fit (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3]

A "step" is one complete fit (initial M-step + EM loop until convergence, em_iters = 50) of the
workload's synthetic simulate_qcs-style data.  `value` = cells x executed EM iterations x steps / device
time with the cell matrix resident in HBM; `e2e` is the same metric through the reference-facing call
(fit_sampleqc_robust_cpp -> sqc_fit_robust) with pinned HOST buffers, H2D and D2H copies inside the timed
region.  `--impl reference` times the CPU restatement of the reference (oracle/) with all the host threads the
reference can use: a fit is single-threaded (its src/ has no OpenMP pragma), the R package parallelises over
independent sample groups, so one fit per host core runs concurrently, each on a bounded run of samples of the
same workload.  The `cpu_baseline` leg of the default run times ONE such fit (one core).
Under torchrun (N > 1) cells are sharded by sample, one rank per GPU, and the fit is ONE mixture over all ranks'
cells (global parameters agreed through the in-kernel NVLink exchange).  Default `--scaling weak`: every GPU holds
one workload-sized shard (N x 10 M cells x N x 1000 samples in total; per-GPU work fixed); `--scaling strong`
splits the one workload over the GPUs instead (latency-bound at 10 M cells).
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

METRIC = "EM cell-iters/s"
EM_ITERS, MCD_ALPHA, MCD_ITERS, SEED = 50, 0.5, 50, 0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [v.strip() for v in r.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def workload(name: str, scale: float = 1.0):
    """the BASELINE config's inputs; cached under /tmp so warm reruns skip the numpy generator"""
    from sampleqc_b200.simulate import CONFIGS, make_config
    cache = f"/tmp/sqc_bench_{name}_{scale:g}.npz"
    if os.path.exists(cache):
        d = np.load(cache)
        return np.asfortranarray(d["x"]), d["groups"], d["init_z"], {k: int(d[k]) for k in ("D", "J", "K", "N")}
    sim, iz, _, prm = make_config(name, scale=scale)
    try:                                                          # ranks may race: write aside, then rename
        tmp = f"{cache}.{os.getpid()}.npz"
        np.savez(tmp, x=sim.x, groups=sim.groups, init_z=iz, **{k: prm[k] for k in ("D", "J", "K", "N")})
        os.replace(tmp, cache)
    except Exception:
        pass
    return sim.x, sim.groups, iz, {k: prm[k] for k in ("D", "J", "K", "N")}


def shard_by_sample(groups: np.ndarray, J: int, world: int):
    """contiguous sample ranges with balanced cell counts: returns [(j0, j1, i0, i1)] per rank"""
    bounds = np.searchsorted(groups, np.arange(J + 1))
    N = groups.shape[0]
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(bounds, N * r / world)))
    cuts.append(J)
    cuts = [min(max(c, 0), J) for c in cuts]
    return [(cuts[r], cuts[r + 1], int(bounds[cuts[r]]), int(bounds[cuts[r + 1]])) for r in range(world)]


def cpu_reference(x, groups, init_z, prm, budget_cells=1_000_000, em_iters=8):
    """the oracle (CPU restatement of the reference) on a bounded sample: the first samples of the workload
    holding about `budget_cells` cells, `em_iters` EM iterations at most."""
    from oracle import oracle as O
    J = prm["J"]
    bounds = np.searchsorted(groups, np.arange(J + 1))
    j1 = max(int(np.searchsorted(bounds, min(budget_cells, groups.shape[0]))), 2)
    j1 = min(j1, J)
    n = int(bounds[j1])
    xs, gs, zs = np.asfortranarray(x[:n]), groups[:n], init_z[:n]
    t0 = time.perf_counter()
    fit = O.fit_sampleqc_robust_cpp(xs, zs, gs, prm["D"], j1, prm["K"], n, em_iters, MCD_ALPHA, MCD_ITERS, SEED, rng_mode=0)
    dt = time.perf_counter() - t0
    return {"value": n * fit["n_iter"] / dt, "unit": METRIC, "cores": 1, "kind": "port",
            "sample": f"first {j1} samples / {n} cells of the workload, {fit['n_iter']} EM iterations, {dt:.1f} s, g++ -O2 single thread "
                      "(the reference's src/ has no OpenMP pragma)"}, dt


_PAR = {}


def _par_worker(w):
    """one process of cpu_reference_all_cores: the oracle on its own run of samples"""
    from oracle import oracle as O
    x, groups, init_z, prm, per, em_iters = _PAR["args"]
    J = prm["J"]
    bounds = np.searchsorted(groups, np.arange(J + 1))
    j0 = (w * per) % max(J - per, 1)
    j1 = j0 + per
    i0, i1 = int(bounds[j0]), int(bounds[j1])
    xs, gs, zs = np.asfortranarray(x[i0:i1]), (groups[i0:i1] - j0).astype(np.int32), init_z[i0:i1]
    t0 = time.perf_counter()
    try:
        fit = O.fit_sampleqc_robust_cpp(xs, zs, gs, prm["D"], per, prm["K"], i1 - i0, em_iters, MCD_ALPHA, MCD_ITERS, SEED, rng_mode=0)
        work = (i1 - i0) * fit["n_iter"]
    except Exception:                                             # (a slice whose initial labelling misses a cluster: counts as no work)
        work = 0
    return work, time.perf_counter() - t0


def cpu_reference_all_cores(x, groups, init_z, prm, cells_per_proc=300_000, em_iters=5):
    """The reference's CPU path with all the host threads it can use.  One fit is single-threaded (no OpenMP pragma
    in the reference's src/); what the R package parallelises is independent sample groups, one fit per core
    (BiocParallel, R/fit_sampleqc.R:77-92).  So: one oracle fit per core, each on its own ~cells_per_proc-cell run of
    samples of the workload, all running at once; value = all cells x iterations / wall time."""
    import multiprocessing as mp
    from oracle import oracle as O
    O.build()
    cores = len(os.sched_getaffinity(0))
    J, N = prm["J"], prm["N"]
    per = max(2, min(J // 2, int(round(cells_per_proc / (N / J)))))
    _PAR["args"] = (x, groups, init_z, prm, per, em_iters)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_par_worker_warm, range(cores))                  # processes up, library loaded
        t0 = time.perf_counter()
        res = pool.map(_par_worker, range(cores), chunksize=1)
        dt = time.perf_counter() - t0
    work = sum(r[0] for r in res)
    return {"value": work / dt, "unit": METRIC, "cores": cores, "kind": "port",
            "sample": f"{cores} concurrent single-threaded fits (one per core, as fit_sampleqc(n_cores=) runs sample groups), each on {per} samples / "
                      f"~{int(per * N / J)} cells of the workload, {em_iters} EM iterations, {dt:.1f} s wall, g++ -O2"}, dt


def _par_worker_warm(w):
    from oracle import oracle as O
    O.lib()
    return 0


_REAL_STDOUT = None


def guard_stdout():
    """stdout carries exactly ONE JSON line: anything libraries write to fd 1 (NCCL prints its version banner there
    when a box-level nccl.conf asks for it) is sent to stderr instead; emit() writes to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--rng", default="r_compat", choices=["r_compat", "counter"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = every GPU holds one copy-sized shard of the workload (N x the cells and samples), strong = the workload split over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    guard_stdout()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(a.warmup, 0)

    x, groups, init_z, prm = workload(a.workload, a.scale)
    D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
    weak = world > 1 and a.scaling == "weak"
    N_job, J_job = (world * N, world * J) if weak else (N, J)      # the whole job's cells / samples
    config = {"workload": f"{a.workload}: {N} cells x {J} samples, D={D}, K={K}, robust EM (hard-assignment EM + FastMCD), "
                          f"em_iters={EM_ITERS} mcd_alpha={MCD_ALPHA} mcd_iters={MCD_ITERS} seed={SEED}"
                          + (f"; weak scaling: one such shard per GPU, the fit is ONE mixture over {N_job} cells x {J_job} samples" if weak else ""),
              "rng_mode": a.rng, "sharding": f"by sample over {world} GPU(s)", "l2": "inputs larger than L2 (x + residual columns = %.0f MB per GPU)" % (2 * N * D * 8 / 1e6 / (1 if weak or world == 1 else world))}

    if a.impl == "reference":
        if rank != 0:
            return
        vals = []
        for i in range(W + a.steps):
            cb, dt = cpu_reference_all_cores(x, groups, init_z, prm)
            if i >= W:
                vals.append((cb, dt))
        v = float(np.mean([c["value"] for c, _ in vals]))
        cb = dict(vals[-1][0]); cb["value"] = v
        emit(({"impl": "reference", "metric": METRIC, "value": v, "unit": "cell-iters/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": W,
                          "ms_per_step": 1e3 * float(np.mean([d for _, d in vals])), "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "cell-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    from sampleqc_b200 import api
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng_mode = api.SQC_RNG_R_COMPAT if a.rng == "r_compat" else api.SQC_RNG_COUNTER

    # ---- shard by sample --------------------------------------------------------------------------
    if weak:
        # every rank holds the workload's recipe once: samples rank*J .. (rank+1)*J of the job.  Ranks > 0 perturb the
        # cells slightly so that no two cells of the job are identical (same mixture, distinct data)
        j0, j1, i0, i1 = rank * J, (rank + 1) * J, rank * N, (rank + 1) * N
        xs = np.asfortranarray(x if rank == 0 else x + 1e-3 * np.random.default_rng(4000 + rank).standard_normal(x.shape))
        gs, zs = groups.astype(np.int32), init_z
    else:
        j0, j1, i0, i1 = shard_by_sample(groups, J, world)[rank]
        xs, gs, zs = np.asfortranarray(x[i0:i1]), (groups[i0:i1] - j0).astype(np.int32), init_z[i0:i1]
    kw = dict(variant=api.SQC_VARIANT_ROBUST, init_z=zs, mcd_alpha=MCD_ALPHA, mcd_iters=MCD_ITERS, seed=SEED, rng_mode=rng_mode, device=local)
    if world > 1:
        from sampleqc_b200 import multigpu
        kw.update(multigpu.exchange_kwargs(dist, local, rank, world, i0, N_job))
    sess = api.Session(xs, gs, D, j1 - j0, K, i1 - i0, EM_ITERS, **kw)
    sess.profile(True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n_warm = W if os.environ.get("SQC_BENCH_PROFILING") else max(W, 3)     # profiling runs may warm up less
    for _ in range(n_warm):
        sess.run()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); time.sleep(0.25)
    barrier()
    t0 = time.time()
    loop_ms, fam_ms, fam_n, launches, n_iter, mcd_bytes = 0.0, {}, {}, 0, 0, 0.0
    for _ in range(a.steps):
        sess.run()
        st = sess.stats()
        loop_ms += st["loop_ms"] + st["init_ms"]
        launches += st["n_gpu_launches"]; n_iter = st["n_iter"]
        for k, v in st["family_ms"].items():
            fam_ms[k] = fam_ms.get(k, 0.0) + v
            fam_n[k] = fam_n.get(k, 0) + st["family_launches"][k]
        fam_bytes = st["family_bytes"]
        mcd_rounds, mcd_csteps = st["mcd_passes"], st["mcd_csteps"]
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    tt = torch.tensor([loop_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms = float(tt.item())
    value = N_job * n_iter * a.steps / (dev_ms / 1e3)

    # ---- dominant kernel: roofline ------------------------------------------------------------------
    hbm, peak_src = peaks()
    top = max(("mcd", "estep", "partition"), key=lambda k: fam_ms.get(k, 0.0))
    n_top = max(fam_n.get(top, 1), 1)
    top_ms = fam_ms[top] / n_top
    top_bytes = fam_bytes[top] / max(n_top // a.steps, 1)       # family_bytes is per run
    achieved = top_bytes / (top_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": {"mcd": "k_mcd (persistent FastMCD M-step)", "estep": "k_estep", "partition": "k_partition"}[top],
                "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None, "peak_source": peak_src,
                "avg_launch_ms": top_ms, "algorithmic_bytes_per_launch": top_bytes,
                "share_of_step": fam_ms[top] / (dev_ms if dev_ms else 1.0),
                "families": {k: {"ms_per_step": fam_ms[k] / a.steps, "launches_per_step": fam_n[k] // a.steps,
                                 "GBps": (fam_bytes[k] / (fam_ms[k] / a.steps / 1e3) / 1e9) if fam_ms[k] > 0 and fam_bytes.get(k, 0) else None}
                             for k in fam_ms if fam_n[k]},
                "mcd_rounds_per_step": mcd_rounds, "mcd_csteps_per_step": mcd_csteps}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get(top)
        except Exception:
            pass
    resident_fit = sess.fetch()                                   # outside the timed region: for the validity checks below
    sess.close()

    # ---- e2e: the reference-facing call with pinned host buffers -------------------------------------
    e2e = None
    if not a.no_e2e:
        def pinned(shape, dtype=np.float64, order="C"):
            shape = (shape,) if np.isscalar(shape) else tuple(shape)
            n = int(np.prod(shape)) if shape else 1
            t = torch.empty(max(n, 1), dtype={np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8}[np.dtype(dtype)], pin_memory=True)
            arr = t.numpy()[:n]
            arr[...] = 0
            pinned.keep.append(t)
            return arr.reshape(shape, order=order)
        pinned.keep = []
        xp = pinned((i1 - i0, D), order="F"); xp[...] = xs
        gp = pinned(i1 - i0, dtype=np.int32); gp[...] = gs
        zp = pinned(i1 - i0, dtype=np.int32); zp[...] = zs
        h2d = xp.nbytes + gp.nbytes + zp.nbytes
        e2e_iters = 0

        from sampleqc_b200 import _abi as A
        obufs = A.FitBuffers(D, j1 - j0, K, i1 - i0, EM_ITERS, A.SQC_VARIANT_ROBUST, False, alloc=pinned)   # caller-owned, reused every step

        def one():
            nonlocal e2e_iters
            if world > 1:
                from sampleqc_b200 import multigpu
                fit = multigpu.fit_robust_sharded(dist, local, rank, world, xp, zp, gp, D, j1 - j0, K, i1 - i0, i0, N_job, EM_ITERS, MCD_ALPHA, MCD_ITERS, SEED, rng_mode, out=obufs)
            else:
                fit = api.fit_sampleqc_robust_cpp(xp, zp, gp, D, J, K, N, EM_ITERS, MCD_ALPHA, MCD_ITERS, SEED, rng_mode=rng_mode, device=local, out=obufs)
            e2e_iters = fit["n_iter"]
            return fit
        for _ in range(2):
            fit = one()
        barrier()
        ts = time.perf_counter()
        n_e2e = max(min(a.steps, 5), 1)
        for _ in range(n_e2e):
            fit = one()
        barrier()
        te = torch.tensor([time.perf_counter() - ts], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        d2h = sum(fit[k].nbytes for k in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk", "like_1", "like_2")) + (i1 - i0) * 4 + 625 * 4
        if dist is not None:                                      # whole-job bytes: sum over the ranks
            tb = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
            dist.all_reduce(tb)
            h2d, d2h = int(tb[0].item()), int(tb[1].item())
        e2e = {"value": N_job * e2e_iters * n_e2e / float(te.item()), "unit": "cell-iters/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * float(te.item()) / n_e2e, "steps": n_e2e, "timer": "host wall clock around the blocking C-ABI call, max over ranks"}

    # ---- validity of what was timed (size-independent properties; rank 0's shard) ---------------------
    checks = None
    try:
        f = resident_fit
        checks = {"n_iter": int(f["n_iter"]), "z_in_1..K": bool(f["z"].min() >= 1 and f["z"].max() <= K),
                  "p_jk_rows_sum_to_1": bool(np.allclose(f["p_jk"].sum(axis=1), 1.0, rtol=0, atol=1e-12)),
                  "beta_k_col0_ascending": bool(np.all(np.diff(f["beta_k"][:, 0]) >= 0)),                # final relabel, :527-559
                  "sigma_k_spd": bool(all(np.all(np.linalg.eigvalsh(f["sigma_k"][:, :, k]) > 0) for k in range(K))),
                  "params_finite": bool(all(np.all(np.isfinite(f[k])) for k in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk")))}
        if e2e is not None:                                       # the resident run and the host-buffer run are the same fit, bit for bit
            checks["resident_equals_e2e_bitwise"] = bool(all(np.array_equal(f[k], fit[k]) for k in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk", "z")))
    except Exception as ex:                                       # never lose the measurement over a check
        checks = {"error": repr(ex)}

    if dist is not None:                                          # all ranks leave the group together
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    out = {"metric": METRIC, "value": value, "unit": "cell-iters/s", "n_gpus": world, "steps": a.steps, "warmup": n_warm,
           "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak" if (weak or world == 1) else "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": config, "em_iterations_per_step": n_iter, "clocks": clocks, "e2e": e2e,
           "gpu_launches": int(launches), "roofline": roofline, "checks": checks}
    if not a.no_cpu_baseline and world == 1:                      # the CPU leg is reported at N = 1 only
        out["cpu_baseline"], _ = cpu_reference(x, groups, init_z, prm)
        out["cpu_baseline"]["note"] = "one single-threaded fit (the reference has no threads inside a fit); `--impl reference` runs one such fit per host core"
    emit(out)


if __name__ == "__main__":
    main()
