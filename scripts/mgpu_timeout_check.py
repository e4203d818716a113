"""torchrun --nproc-per-node 2 scripts/mgpu_timeout_check.py : a rank whose peer never shows up gives up with SQC_ERR_COMM
after the wait budget (SQC_WAIT_BUDGET_MS), without a trap — its CUDA context survives and fits again."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SQC_WAIT_BUDGET_MS"] = "400"
import numpy as np, torch, torch.distributed as dist
from sampleqc_b200 import api, multigpu
from sampleqc_b200._abi import SampleQCError
from sampleqc_b200.simulate import simulate_qcs, init_clusts

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, J, K, D = 60_000, 8, 3, 3
sim = simulate_qcs(N, J, 3, D, seed=4)
iz, _ = init_clusts(sim.x, sim.groups, J, K, seed=4)
j0, j1, i0, i1 = multigpu.shard_by_sample(sim.groups, J, world)[rank]
xs = np.asfortranarray(sim.x[i0:i1]); gs = (sim.groups[i0:i1] - j0).astype(np.int32)
ok = True
if rank == 0:
    # rank 1 exports its window and joins the handle exchange (inside fit_robust_sharded on this side), then never fits
    t0 = time.time()
    try:
        multigpu.fit_robust_sharded(dist, local, rank, world, xs, iz[i0:i1], gs, D, j1 - j0, K, i1 - i0, i0, N, 6, 0.5, 50, 0)
        print("rank 0: the fit returned although its peer never ran"); ok = False
    except SampleQCError as e:
        dt = time.time() - t0
        print(f"rank 0: status {e.status} after {dt:.1f} s: {e.message}")
        ok &= e.status == -2 and dt < 30
    # the context is alive: an ordinary fit on this GPU still works and matches the oracle
    from oracle import oracle as O
    got = api.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, 6, 0.5, 50, 0, device=local)
    ref = O.fit_sampleqc_robust_cpp(sim.x, iz, sim.groups, D, J, K, N, 6, 0.5, 50, 0)
    ok &= bool(np.array_equal(got["z"], ref["z"]))
    print("MGPU TIMEOUT CHECK", "OK" if ok else "FAILED")
else:
    multigpu.exchange_kwargs(dist, local, rank, world, i0, N)      # the bootstrap only
    time.sleep(3.0)
dist.barrier()
dist.destroy_process_group()
