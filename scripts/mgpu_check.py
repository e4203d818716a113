"""torchrun --nproc-per-node W scripts/mgpu_check.py : sharded robust + MLE fits against the CPU oracle"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sampleqc_b200 import api, multigpu
from sampleqc_b200.simulate import simulate_qcs, init_clusts

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
CASES = [(40000, 12, 4, 4, 3, 12, 0), (200000, 31, 4, 3, 3, 10, 1), (1_000_000, 100, 4, 4, 3, 8, 0), (3_000_001, 150, 4, 4, 3, 4, 0)]
if os.environ.get("MGPU_QUICK"):
    CASES = CASES[:2]
CASES.append((60_000, 10, 4, 4, 3, 6, 0, "absent"))      # one cluster has no cell at all on the last rank in the first M-step
for case in CASES:
    (N, J, Kt, K, D, em, rng), tag = case[:7], (case[7] if len(case) > 7 else "")
    sim = simulate_qcs(N, J, Kt, D, seed=7)
    iz, ig = init_clusts(sim.x, sim.groups, J, K, seed=7)
    j0, j1, i0, i1 = multigpu.shard_by_sample(sim.groups, J, world)[rank]
    if tag == "absent":
        lo = multigpu.shard_by_sample(sim.groups, J, world)[world - 1][2]
        iz = iz.copy(); iz[lo:][iz[lo:] == K - 1] = K - 2
    xs = np.asfortranarray(sim.x[i0:i1]); gs = (sim.groups[i0:i1] - j0).astype(np.int32)
    t0 = time.time()
    loc = multigpu.fit_robust_sharded(dist, local, rank, world, xs, iz[i0:i1], gs, D, j1 - j0, K, i1 - i0, i0, N, em, 0.5, 50, 0, rng_mode=rng)
    full = multigpu.assemble(dist, rank, world, loc, (j0, j1, i0, i1), J, N)
    t1 = time.time()
    if rank == 0:
        from oracle import oracle as O
        ref = O.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, em, 0.5, 50, 0, rng_mode=rng)
        one = api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, em, 0.5, 50, 0, rng_mode=rng, device=local)
        print(f"robust N={N} world={world} rng={rng} {tag}: {t1-t0:.3f}s n_iter {full['n_iter']} / {ref['n_iter']} / {one['n_iter']}")
        for key in ("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk", "like_1", "like_2"):
            fin = np.isfinite(ref[key])                  # plain densities may underflow: log 0 = -inf on both sides
            assert np.array_equal(fin, np.isfinite(full[key]))
            if not fin.any():
                print(f"   {key:8s} all -inf on both sides"); continue
            sc = max(np.max(np.abs(ref[key][fin])), 1e-300)
            d1 = np.max(np.abs(full[key][fin] - ref[key][fin])) / sc
            d2 = np.max(np.abs(full[key][fin] - one[key][fin])) / sc
            print(f"   {key:8s} vs oracle {d1:.2e}   vs 1 GPU {d2:.2e}")
            ok &= d1 < 1e-8
        zm = int(np.sum(full["z"] != ref["z"])); print("   z mismatches", zm); ok &= zm == 0 and full["n_iter"] == ref["n_iter"]
        if rng == 0:                                   # R's .Random.seed after the fit (the key stream is generated in shares)
            same = bool(np.array_equal(full["rng_state"], ref["rng_state"])) and full["rng_draws"] == ref["rng_draws"]
            print("   rng_state identical:", same); ok &= same
    # MLE
    loc = multigpu.fit_mle_sharded(dist, local, rank, world, xs, np.asfortranarray(ig[i0:i1]), gs, D, j1 - j0, K, i1 - i0, i0, N, 4, 0)
    full = multigpu.assemble(dist, rank, world, loc, (j0, j1, i0, i1), J, N)
    if rank == 0:
        ref = O.fit_sampleqc_mle_cpp(sim.x, ig, sim.groups, D, J, K, N, 4, 0)
        for key in ("alpha_j", "beta_k", "sigma_k", "p_k", "p_jk", "like_1", "like_2", "gamma_i"):
            d1 = np.max(np.abs(full[key] - ref[key])) / max(np.max(np.abs(ref[key])), 1e-300)
            print(f"   mle {key:8s} vs oracle {d1:.2e}"); ok &= d1 < 1e-7
    dist.barrier()
if rank == 0:
    print("MGPU CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
