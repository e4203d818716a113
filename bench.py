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
cells (global parameters agreed through the in-kernel NVLink exchange).  Default `--scaling strong`: the BASELINE
workload (10 M cells) split over the GPUs, the configuration the metric names; `--scaling weak`: every GPU holds one
workload-sized shard (N x 10 M cells x N x 1000 samples in total).  The default run measures the other mode too and
reports it under `extra.other_scaling_mode`, next to short legs for the MLE variant, C4 (D = 6, K = 6; N = 1) and C5
(50 M cells, K = 5, robust + MLE; N = 8).  `checks` carries parity against the CPU oracle on the benchmarked data
(N = 1: the slice the `cpu_baseline` leg fits; N > 1: a small sharded fit before the timed runs).
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
    out = {"value": n * fit["n_iter"] / dt, "unit": METRIC, "cores": 1, "kind": "port",
           "sample": f"first {j1} samples / {n} cells of the workload, {fit['n_iter']} EM iterations, {dt:.1f} s, g++ -O2 single thread "
                     "(the reference's src/ has no OpenMP pragma); the per-component det / inverse of the density are hoisted "
                     "(RcppDist::dmvnorm recomputes them per call: see `faithful`)"}
    # the reference-faithful cost structure: det(S) and S^-1 recomputed inside every density call,
    # as RcppDist::dmvnorm does; same arithmetic, a small sample
    try:
        jf = max(int(np.searchsorted(bounds, min(150_000, groups.shape[0]))), 2)
        jf = min(jf, J); nf = int(bounds[jf])
        O.set_cost_faithful(True)
        t1 = time.perf_counter()
        ff = O.fit_sampleqc_robust_cpp(np.asfortranarray(x[:nf]), init_z[:nf], groups[:nf], prm["D"], jf, prm["K"], nf, 3, MCD_ALPHA, MCD_ITERS, SEED, rng_mode=0)
        dtf = time.perf_counter() - t1
        out["faithful"] = {"value": nf * ff["n_iter"] / dtf, "unit": METRIC, "sample": f"first {jf} samples / {nf} cells, {ff['n_iter']} EM iterations, {dtf:.1f} s"}
    except Exception as ex:
        out["faithful"] = {"error": repr(ex)}
    finally:
        O.set_cost_faithful(False)
    return out, dt, (fit, xs, gs, zs, j1, n, em_iters)


def parity_vs_oracle(ref_pack, prm, device):
    """the CUDA path on the very slice the CPU leg just fitted: parameters, labels, iteration count, likelihoods,
    R's RNG state (north_star bars: 1e-8 relative, same iterations, identical labels)"""
    from sampleqc_b200 import api
    ref, xs, gs, zs, j1, n, em_iters = ref_pack
    got = api.fit_sampleqc_robust_cpp(xs, zs, gs, prm["D"], j1, prm["K"], n, em_iters, MCD_ALPHA, MCD_ITERS, SEED, rng_mode=0, device=device)
    rel = {}
    for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk"):
        rel[key] = float(np.max(np.abs(got[key] - ref[key])) / max(np.max(np.abs(ref[key])), 1e-300))
    it = int(ref["n_iter"])
    l_rel = 0.0
    for key in ("like_1", "like_2"):
        a, b = got[key].ravel()[:it], ref[key].ravel()[:it]
        fin = np.isfinite(b)
        if not np.array_equal(fin, np.isfinite(a)):
            l_rel = float("inf")
        elif fin.any():
            l_rel = max(l_rel, float(np.max(np.abs(a[fin] - b[fin]) / np.abs(b[fin]))))
    return {"cells": int(n), "samples": int(j1), "em_iters": int(em_iters), "max_rel_param": max(rel.values()), "rel_by_param": rel,
            "max_rel_like": l_rel, "z_mismatches": int(np.sum(got["z"] != ref["z"])), "n_iter_equal": bool(got["n_iter"] == ref["n_iter"]),
            "rng_state_equal": bool(np.array_equal(got["rng_state"], ref["rng_state"])),
            "pass": bool(max(rel.values()) < 1e-8 and l_rel < 1e-8 and np.array_equal(got["z"], ref["z"]) and got["n_iter"] == ref["n_iter"])}


def multi_gpu_parity(dist, local, rank, world):
    """N > 1: one small sharded fit (R-compatible keys generated in per-rank shares) against the CPU oracle, before the
    timed runs: parameters, labels, iterations, R's final RNG state.  Rank 0 returns the verdict, the others None."""
    from sampleqc_b200 import multigpu
    from sampleqc_b200.simulate import init_clusts, simulate_qcs
    N, J, Kt, K, D, em = 200_000, 32, 4, 3, 3, 8
    sim = simulate_qcs(N, J, Kt, D, seed=7)
    iz, _ = init_clusts(sim.x, sim.groups, J, K, seed=7)
    j0, j1, i0, i1 = multigpu.shard_by_sample(sim.groups, J, world)[rank]
    xs = np.asfortranarray(sim.x[i0:i1]); gs = (sim.groups[i0:i1] - j0).astype(np.int32)
    loc = multigpu.fit_robust_sharded(dist, local, rank, world, xs, iz[i0:i1], gs, D, j1 - j0, K, i1 - i0, i0, N, em, MCD_ALPHA, MCD_ITERS, SEED, rng_mode=0)
    full = multigpu.assemble(dist, rank, world, loc, (j0, j1, i0, i1), J, N)
    if rank != 0:
        return None
    from oracle import oracle as O
    ref = O.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, em, MCD_ALPHA, MCD_ITERS, SEED, rng_mode=0)
    rel = max(float(np.max(np.abs(full[k] - ref[k])) / max(np.max(np.abs(ref[k])), 1e-300)) for k in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk"))
    zm = int(np.sum(full["z"] != ref["z"]))
    return {"case": f"{N} cells x {J} samples, K={K}, em_iters={em}, sharded over {world} GPUs vs the CPU oracle", "max_rel_param": rel, "z_mismatches": zm,
            "n_iter_equal": bool(full["n_iter"] == ref["n_iter"]), "rng_state_equal": bool(np.array_equal(full["rng_state"], ref["rng_state"])),
            "pass": bool(rel < 1e-8 and zm == 0 and full["n_iter"] == ref["n_iter"] and np.array_equal(full["rng_state"], ref["rng_state"]))}



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
    failed = 0
    try:
        fit = O.fit_sampleqc_robust_cpp(xs, zs, gs, prm["D"], per, prm["K"], i1 - i0, em_iters, MCD_ALPHA, MCD_ITERS, SEED, rng_mode=0)
        work = (i1 - i0) * fit["n_iter"]
    except Exception:                                             # (a slice whose initial labelling misses a cluster: counts as no work, and is reported)
        work, failed = 0, 1
    return work, time.perf_counter() - t0, failed


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
    return {"value": work / dt, "unit": METRIC, "cores": cores, "kind": "port", "failed_workers": int(sum(r[2] for r in res)),
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


def _pinned_alloc(torch):
    def pinned(shape, dtype=np.float64, order="C"):
        shape = (shape,) if np.isscalar(shape) else tuple(shape)
        n = int(np.prod(shape)) if shape else 1
        t = torch.empty(max(n, 1), dtype={np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8}[np.dtype(dtype)], pin_memory=True)
        arr = t.numpy()[:n]
        arr[...] = 0
        pinned.keep.append(t)
        return arr.reshape(shape, order=order)
    pinned.keep = []
    return pinned


# fp64 instruction model of the two per-cell passes: per (cell, component) the affine whitening
# D(D+1)/2 + D FMAs, D FMAs for the norm, 8 for exp + accumulation, 1 clamp (+ 2 for the label: log-density, compare);
# per cell D subtractions of alpha_j and ~1/4 of a 32-instruction log
def fp64_ops_per_cell(D, K, estep):
    per_k = D * (D + 1) // 2 + 2 * D + 9 + (2 if estep else 0)
    return K * per_k + D + 8


def measure_robust(ctx, x, groups, init_z, prm, weak, steps, n_warm, rng_name, do_e2e, want_fit=True):
    """time `steps` resident fits (device time, max over ranks) and, optionally, the same fit end to end through the
    reference-facing call with pinned host buffers"""
    torch, api, dist, rank, world, local = ctx["torch"], ctx["api"], ctx["dist"], ctx["rank"], ctx["world"], ctx["local"]
    D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
    N_job, J_job = (world * N, world * J) if weak else (N, J)
    rng_mode = api.SQC_RNG_R_COMPAT if rng_name == "r_compat" else api.SQC_RNG_COUNTER
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

    for _ in range(n_warm):
        sess.run()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); time.sleep(0.25)
    barrier()
    t0 = time.time()
    loop_ms, init_ms, fam_ms, fam_n, launches, n_iter = 0.0, 0.0, {}, {}, 0, 0
    for _ in range(steps):
        sess.run()
        st = sess.stats()
        loop_ms += st["loop_ms"] + st["init_ms"]; init_ms += st["init_ms"]
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
    value = N_job * n_iter * steps / (dev_ms / 1e3)

    # ---- dominant kernel: roofline ------------------------------------------------------------------
    hbm, peak_src = peaks()
    top = max(("mcd", "estep", "partition"), key=lambda k: fam_ms.get(k, 0.0))
    n_top = max(fam_n.get(top, 1), 1)
    top_ms = fam_ms[top] / n_top
    top_bytes = fam_bytes[top] / max(n_top // steps, 1)       # family_bytes is per run
    achieved = top_bytes / (top_ms / 1e3) / 1e9
    fams = {k: {"ms_per_step": fam_ms[k] / steps, "launches_per_step": fam_n[k] // steps, "avg_launch_ms": fam_ms[k] / fam_n[k],
                "GBps": (fam_bytes[k] / (fam_ms[k] / steps / 1e3) / 1e9) if fam_ms[k] > 0 and fam_bytes.get(k, 0) else None}
            for k in fam_ms if fam_n[k]}
    for k in fams:
        if fams[k]["GBps"]:
            fams[k]["hbm_frac"] = fams[k]["GBps"] / hbm
    roofline = {"bound": "hbm", "kernel": {"mcd": "k_mcd (persistent FastMCD M-step)", "estep": "k_estep", "partition": "k_partition"}[top],
                "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None, "peak_source": peak_src,
                "avg_launch_ms": top_ms, "algorithmic_bytes_per_launch": top_bytes,
                "share_of_step": fam_ms[top] / (dev_ms if dev_ms else 1.0),
                "families": fams, "mcd_rounds_per_step": mcd_rounds, "mcd_csteps_per_step": mcd_csteps,
                "init_ms_per_step": init_ms / steps,
                "note": "value counts the device time of the initial M-step and the EM loop of every fit; the one-off centring passes "
                        "(column means, centre_x, init_alpha_j: 3 passes at session creation) are outside it and inside e2e"}
    fp64 = ctx.get("fp64_dfma_per_s")
    if fp64:
        n_loc = i1 - i0
        co = {"peak_dfma_per_s": fp64, "peak_tflops": 2 * fp64 / 1e12, "peak_source": "measured in this run (8 independent DFMA chains per thread)",
              "model": "fp64 instructions per cell: K (D(D+1)/2 + 2D + 9 [+2 E-step]) + D + 8"}
        for fam, es in (("estep", True), ("partition", False)):
            if fam in fams:
                ops = fp64_ops_per_cell(D, K, es) * n_loc
                co[fam] = {"fp64_ops_per_launch": ops, "achieved_dfma_per_s": ops / (fams[fam]["avg_launch_ms"] / 1e3), "frac": ops / (fams[fam]["avg_launch_ms"] / 1e3) / fp64}
        roofline["fp64"] = co
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get(top)
        except Exception:
            pass
    resident_fit = sess.fetch() if want_fit else None            # outside the timed region: for the validity checks
    sess.close()

    # ---- e2e: the reference-facing call with pinned host buffers -------------------------------------
    e2e, fit = None, None
    if do_e2e:
        pinned = _pinned_alloc(torch)
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
                f = multigpu.fit_robust_sharded(dist, local, rank, world, xp, zp, gp, D, j1 - j0, K, i1 - i0, i0, N_job, EM_ITERS, MCD_ALPHA, MCD_ITERS, SEED, rng_mode, out=obufs)
            else:
                f = api.fit_sampleqc_robust_cpp(xp, zp, gp, D, J, K, N, EM_ITERS, MCD_ALPHA, MCD_ITERS, SEED, rng_mode=rng_mode, device=local, out=obufs)
            e2e_iters = f["n_iter"]
            return f
        for _ in range(2):
            fit = one()
        barrier()
        ts = time.perf_counter()
        n_e2e = max(min(steps, 5), 1)
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
    return {"value": value, "dev_ms": dev_ms, "ms_per_step": dev_ms / steps, "n_iter": n_iter, "launches": int(launches), "roofline": roofline, "clocks": clocks,
            "e2e": e2e, "resident_fit": resident_fit, "e2e_fit": fit, "N_job": N_job, "J_job": J_job, "cells_per_gpu": i1 - i0}


def measure_mle(ctx, x, groups, init_z, prm, weak, steps, n_warm, n_iter=20):
    """the MLE variant (fixed-count soft EM, fit_sampleqc_mle_cpp) on the same cells; init_gamma = softened one-hot of the
    initial labels"""
    torch, api, dist, rank, world, local = ctx["torch"], ctx["api"], ctx["dist"], ctx["rank"], ctx["world"], ctx["local"]
    D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
    N_job = world * N if weak else N
    if weak:
        j0, j1, i0, i1 = rank * J, (rank + 1) * J, rank * N, (rank + 1) * N
        xs = np.asfortranarray(x if rank == 0 else x + 1e-3 * np.random.default_rng(4000 + rank).standard_normal(x.shape))
        gs, zs = groups.astype(np.int32), init_z
    else:
        j0, j1, i0, i1 = shard_by_sample(groups, J, world)[rank]
        xs, gs, zs = np.asfortranarray(x[i0:i1]), (groups[i0:i1] - j0).astype(np.int32), init_z[i0:i1]
    gam = np.full((i1 - i0, K), 0.1 / K, order="F"); gam[np.arange(i1 - i0), zs] += 0.9
    kw = dict(variant=api.SQC_VARIANT_MLE, init_gamma=gam, seed=SEED, device=local)
    if world > 1:
        from sampleqc_b200 import multigpu
        kw.update(multigpu.exchange_kwargs(dist, local, rank, world, i0, N_job))
    sess = api.Session(xs, gs, D, j1 - j0, K, i1 - i0, n_iter, **kw)
    del gam
    sess.profile(True)
    for _ in range(n_warm):
        sess.run()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms, fam_ms, fam_n, fam_bytes = 0.0, 0.0, 0, 0.0
    for _ in range(steps):
        sess.run()
        st = sess.stats()
        ms += st["loop_ms"] + st["init_ms"]
        fam_ms += st["family_ms"]["mle_pass"]; fam_n += st["family_launches"]["mle_pass"]; fam_bytes = st["family_bytes"]["mle_pass"]
    sess.close()
    tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    hbm, _ = peaks()
    gbps = fam_bytes / (fam_ms / steps / 1e3) / 1e9 if fam_ms > 0 else None
    return {"value": N_job * n_iter * steps / (float(tt.item()) / 1e3), "unit": "cell-iters/s", "ms_per_step": float(tt.item()) / steps, "n_iter": n_iter, "steps": steps,
            "k_mle_pass": {"avg_launch_ms": fam_ms / max(fam_n, 1), "launches_per_step": fam_n // steps, "GBps": gbps, "hbm_frac": gbps / hbm if gbps else None}}


def pipeline_leg(ctx, x, groups, prm, rng_name):
    """.init_clusts -> fit -> re-centring -> outliers of .fit_one_sampleqc (R/fit_sampleqc.R:222-258)
    on ONE upload of the cells (sqc_resident_*), beside the same steps through the three independent entry points (x crosses
    PCIe three times).  Host buffers; the mixture on the 1e4-cell subsample (mclust in R, sklearn here) is the caller's and
    is timed separately."""
    from sklearn.mixture import GaussianMixture
    api, torch, local = ctx["api"], ctx["torch"], ctx["local"]
    D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
    rng_mode = api.SQC_RNG_R_COMPAT if rng_name == "r_compat" else api.SQC_RNG_COUNTER
    pinned = _pinned_alloc(torch)
    xp = pinned((N, D), order="F"); xp[...] = x
    gp = pinned(N, dtype=np.int32); gp[...] = groups
    maha = pinned((N, K), order="F"); flags = pinned(N, dtype=np.uint8)
    idx = np.sort(np.random.default_rng(11).choice(N, size=min(N, 10_000), replace=False))
    out, T = {}, {}
    for rep in range(2):                                          # first repetition warms the arena / page tables
        t = time.perf_counter()
        R = api.Resident(xp, gp, J, device=local); T["upload"] = time.perf_counter() - t
        t = time.perf_counter(); R.medians(); xs = R.subsample(idx); T["medians+subsample"] = time.perf_counter() - t
        t = time.perf_counter()
        for seed in range(5):
            gm = GaussianMixture(K, covariance_type="full", random_state=seed).fit(xs)
            sigma = np.ascontiguousarray(np.transpose(gm.covariances_, (1, 2, 0)))
            if np.unique(gm.predict(xs)).size == K:
                break
        T["host_mixture_fit"] = time.perf_counter() - t
        t = time.perf_counter(); R.classify(gm.weights_, gm.means_, sigma); T["classify"] = time.perf_counter() - t
        t = time.perf_counter(); fit = R.fit_robust(K, EM_ITERS, MCD_ALPHA, MCD_ITERS, SEED, rng_mode=rng_mode); T["fit+recentre"] = time.perf_counter() - t
        t = time.perf_counter(); _, fl, cut = R.outliers(0.01, out_maha=maha, out_flags=flags); T["outliers"] = time.perf_counter() - t
        R.close()
    dev_total = sum(v for k, v in T.items() if k != "host_mixture_fit")
    out["resident"] = {"ms": {k: 1e3 * v for k, v in T.items()}, "ms_total_without_host_mixture_fit": 1e3 * dev_total,
                       "h2d_bytes": int(xp.nbytes + gp.nbytes), "d2h_bytes": int(maha.nbytes + flags.nbytes + fit["z"].nbytes),
                       "em_iterations": int(fit["n_iter"]), "outliers": int(np.count_nonzero(flags)), "cells_per_s": N / dev_total}
    S = {}
    for rep in range(2):
        t = time.perf_counter(); med = api.sample_medians(xp, gp, J, device=local); S["medians"] = time.perf_counter() - t
        t = time.perf_counter(); _, _, z = api.gmm_classify(xp, gp, J, gm.weights_, gm.means_, sigma, want_gamma=False, device=local); S["classify"] = time.perf_counter() - t
        t = time.perf_counter(); f2 = api.fit_sampleqc_robust_cpp(xp, z, gp, D, J, K, N, EM_ITERS, MCD_ALPHA, MCD_ITERS, SEED, rng_mode=rng_mode, device=local); S["fit"] = time.perf_counter() - t
        t = time.perf_counter(); f2 = api.recentre(f2); m2, o2, _ = api.calc_outliers(xp, gp, f2["mu_0"], f2["alpha_j"], f2["beta_k"], f2["sigma_k"], 0.01, device=local); S["recentre+outliers"] = time.perf_counter() - t
    out["separate_calls"] = {"ms": {k: 1e3 * v for k, v in S.items()}, "ms_total": 1e3 * sum(S.values()), "h2d_bytes": int(3 * xp.nbytes + 3 * gp.nbytes + z.nbytes)}
    # (the re-centring sums run in C++ on one side and in numpy on the other: last-bit differences in the means)
    out["same_result"] = bool(np.array_equal(fit["z"], f2["z"]) and int(np.count_nonzero(np.asarray(flags, dtype=bool) != o2)) <= 2
                              and np.allclose(fit["beta_k"], f2["beta_k"], rtol=1e-9, atol=1e-12) and np.allclose(fit["alpha_j"], f2["alpha_j"], rtol=1e-9, atol=1e-12))
    out["outlier_flag_mismatches"] = int(np.count_nonzero(np.asarray(flags, dtype=bool) != o2))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--rng", default="r_compat", choices=["r_compat", "counter"])
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong (default) = the BASELINE workload split over the GPUs; weak = every GPU holds one workload-sized shard "
                         "(N x the cells and samples).  The default run reports the other mode as an extra key")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (other scaling mode, MLE, C4, C5)")
    a = ap.parse_args()
    guard_stdout()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(a.warmup, 0)

    x, groups, init_z, prm = workload(a.workload, a.scale)
    D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
    weak = world > 1 and a.scaling == "weak"

    def describe(wk):
        N_job, J_job = (world * N, world * J) if wk else (N, J)
        return {"workload": f"{a.workload}: {N} cells x {J} samples, D={D}, K={K}, robust EM (hard-assignment EM + FastMCD), "
                            f"em_iters={EM_ITERS} mcd_alpha={MCD_ALPHA} mcd_iters={MCD_ITERS} seed={SEED}"
                            + (f"; weak scaling: one such shard per GPU, the fit is ONE mixture over {N_job} cells x {J_job} samples" if wk else
                               (f"; strong scaling: the one workload sharded by sample over {world} GPUs" if world > 1 else "")),
                "rng_mode": a.rng, "sharding": f"by sample over {world} GPU(s)",
                "l2": "inputs larger than L2 (x + residual columns = %.0f MB per GPU)" % (2 * N * D * 8 / 1e6 / (1 if wk or world == 1 else world))}
    config = describe(weak)

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
        cb["failed_workers"] = int(sum(c.get("failed_workers", 0) for c, _ in vals))
        emit(({"impl": "reference", "metric": METRIC, "value": v, "unit": "cell-iters/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": W,
                          "ms_per_step": 1e3 * float(np.mean([d for _, d in vals])), "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": describe(False), "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "cell-iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    from sampleqc_b200 import api
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = {"torch": torch, "api": api, "dist": dist, "rank": rank, "world": world, "local": local}
    try:
        ctx["fp64_dfma_per_s"] = api.fp64_peak(local)
    except Exception:
        ctx["fp64_dfma_per_s"] = None

    # ---- parity where the driver runs: a sharded fit against the CPU oracle before anything is timed ----
    mg_parity = None
    if world > 1:
        try:
            mg_parity = multi_gpu_parity(dist, local, rank, world)
        except Exception as ex:
            mg_parity = {"error": repr(ex), "pass": False}

    n_warm = W if os.environ.get("SQC_BENCH_PROFILING") else max(W, 3)     # profiling runs may warm up less
    m = measure_robust(ctx, x, groups, init_z, prm, weak, a.steps, n_warm, a.rng, not a.no_e2e)

    # ---- validity of what was timed (size-independent properties; rank 0's shard) ---------------------
    checks = None
    try:
        f = m["resident_fit"]
        checks = {"n_iter": int(f["n_iter"]), "z_in_1..K": bool(f["z"].min() >= 1 and f["z"].max() <= K),
                  "p_jk_rows_sum_to_1": bool(np.allclose(f["p_jk"].sum(axis=1), 1.0, rtol=0, atol=1e-12)),
                  "beta_k_col0_ascending": bool(np.all(np.diff(f["beta_k"][:, 0]) >= 0)),                # final relabel, :527-559
                  "sigma_k_spd": bool(all(np.all(np.linalg.eigvalsh(f["sigma_k"][:, :, k]) > 0) for k in range(K))),
                  "params_finite": bool(all(np.all(np.isfinite(f[k])) for k in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk")))}
        if m["e2e"] is not None:                                  # the resident run and the host-buffer run are the same fit, bit for bit
            checks["resident_equals_e2e_bitwise"] = bool(all(np.array_equal(f[k], m["e2e_fit"][k]) for k in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk", "z")))
    except Exception as ex:                                       # never lose the measurement over a check
        checks = {"error": repr(ex)}
    if mg_parity is not None and checks is not None:
        checks["multi_gpu_parity"] = mg_parity

    # ---- extra legs (same JSON line, extra keys): the other scaling mode, the MLE variant, C4, C5 ------
    extra = {}
    if not a.no_extra:
        def leg(name, fn):
            try:
                extra[name] = fn()
            except Exception as ex:
                extra[name] = {"error": repr(ex)}
        if world > 1:
            other = not weak
            def other_mode():
                r = measure_robust(ctx, x, groups, init_z, prm, other, min(a.steps, 3), 3, a.rng, not a.no_e2e, want_fit=False)
                return {"scaling": "weak" if other else "strong", "value": r["value"], "unit": "cell-iters/s", "ms_per_step": r["ms_per_step"], "cells_in_job": r["N_job"],
                        "e2e": r["e2e"], "config": describe(other), "k_mcd_avg_launch_ms": r["roofline"]["families"].get("mcd", {}).get("avg_launch_ms")}
            leg("other_scaling_mode", other_mode)
        leg("mle_" + a.workload, lambda: dict(measure_mle(ctx, x, groups, init_z, prm, weak, 2, 2), config=f"MLE variant (fit_sampleqc_mle_cpp), n_iter=20, {a.workload} cells" + (", weak" if weak else "")))
        if a.workload == "C3" and a.scale == 1.0:
            if world == 1:
                def c4():
                    x4, g4, z4, p4 = workload("C4", 1.0)
                    r = measure_robust(ctx, x4, g4, z4, p4, False, 2, 3, a.rng, False, want_fit=False)
                    fam = r["roofline"]["families"]
                    return {"value": r["value"], "unit": "cell-iters/s", "ms_per_step": r["ms_per_step"], "em_iterations_per_step": r["n_iter"],
                            "config": f"C4: {p4['N']} cells x {p4['J']} samples, D={p4['D']}, K={p4['K']}, robust", "families": fam, "fp64": r["roofline"].get("fp64")}
                leg("robust_C4", c4)
                leg("pipeline_" + a.workload, lambda: pipeline_leg(ctx, x, groups, prm, a.rng))
            if world == 8:
                def c5():
                    x5, g5, z5, p5 = workload("C5", 1.0 / world)
                    r = measure_robust(ctx, x5, g5, z5, p5, True, 2, 3, a.rng, False, want_fit=False)
                    mm = measure_mle(ctx, x5, g5, z5, p5, True, 2, 2)
                    return {"config": f"C5: {world} x {p5['N']} cells x {world} x {p5['J']} samples, D={p5['D']}, K={p5['K']}, one mixture over all GPUs",
                            "robust": {"value": r["value"], "unit": "cell-iters/s", "ms_per_step": r["ms_per_step"], "em_iterations_per_step": r["n_iter"], "families": r["roofline"]["families"]},
                            "mle": mm}
                leg("C5_8gpu", c5)

    if dist is not None:                                          # all ranks leave the group together
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    out = {"metric": METRIC, "value": m["value"], "unit": "cell-iters/s", "n_gpus": world, "steps": a.steps, "warmup": n_warm,
           "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak" if weak else ("strong" if world > 1 else a.scaling), "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": config, "em_iterations_per_step": m["n_iter"], "clocks": m["clocks"], "e2e": m["e2e"],
           "gpu_launches": m["launches"], "roofline": m["roofline"], "checks": checks}
    if extra:
        out["extra"] = extra
    if not a.no_cpu_baseline and world == 1:                      # the CPU leg is reported at N = 1 only
        out["cpu_baseline"], _, pack = cpu_reference(x, groups, init_z, prm)
        out["cpu_baseline"]["note"] = "one single-threaded fit (the reference has no threads inside a fit); `--impl reference` runs one such fit per host core"
        try:                                                      # the same slice through the CUDA path: parity on the benchmarked data
            out["checks"]["parity_vs_oracle"] = parity_vs_oracle(pack, prm, local)
        except Exception as ex:
            out["checks"]["parity_vs_oracle"] = {"error": repr(ex), "pass": False}
    emit(out)


if __name__ == "__main__":
    main()
