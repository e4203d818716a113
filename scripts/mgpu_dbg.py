import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sampleqc_b200 import api, multigpu
from sampleqc_b200.simulate import simulate_qcs, init_clusts
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, J, Kt, K, D, em = 40000, 12, 4, 4, 3, 5
sim = simulate_qcs(N, J, Kt, D, seed=7)
iz, ig = init_clusts(sim.x, sim.groups, J, K, seed=7)
j0, j1, i0, i1 = multigpu.shard_by_sample(sim.groups, J, world)[rank]
xs = np.asfortranarray(sim.x[i0:i1]); gs = (sim.groups[i0:i1] - j0).astype(np.int32)
for rng in (0, 1, 0):
    loc = multigpu.fit_robust_sharded(dist, local, rank, world, xs, iz[i0:i1], gs, D, j1 - j0, K, i1 - i0, i0, N, em, 0.5, 50, 0, rng_mode=rng)
    print(rank, "rng", rng, "like_1", loc["like_1"].ravel(), "like_2", loc["like_2"].ravel(), flush=True)
    dist.barrier()
dist.destroy_process_group()
