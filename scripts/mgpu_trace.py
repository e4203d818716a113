"""torchrun --nproc-per-node W scripts/mgpu_trace.py : per-round timeline of the last FastMCD kernel, weak-scaled C3 (debug)"""
import os, sys, ctypes as C
os.environ["SQC_MCD_TRACE"] = os.environ.get("TRACE_LEADER", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bench import workload
from sampleqc_b200 import api, multigpu
from sampleqc_b200._lib import lib
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, g, iz, prm = workload("C3", 1.0)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
if rank:
    x = np.asfortranarray(x + 1e-3 * np.random.default_rng(4000 + rank).standard_normal(x.shape))
em = int(sys.argv[1]) if len(sys.argv) > 1 else 6
kw = dict(variant=api.SQC_VARIANT_ROBUST, init_z=iz, mcd_alpha=0.5, mcd_iters=50, seed=0, rng_mode=int(os.environ.get("RNG", "0")), device=local)
kw.update(multigpu.exchange_kwargs(dist, local, rank, world, rank * N, world * N))
s = api.Session(x, g.astype(np.int32), D, J, K, N, em, **kw)
s.profile(True)
s.run(); s.run()
st = s.stats()
if rank == 0:
    print({k: v for k, v in st.items() if k != "family_bytes"})
    buf = np.zeros(16 * 512, dtype=np.int64)
    L = lib(); L.sqc_session_mcd_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.sqc_session_mcd_trace(s._h, buf.ctypes.data_as(C.c_void_p), buf.size)
    R = int(buf[0]); t = buf.reshape(-1, 16)
    names = "KH KC DH DC DONE".split()
    print("leader timeline: epoch  wait_for_workers_us  post_us  phase  ncand")
    for r in range(1, R + 1):
        t0, t1, t2, t3, ph, ch = t[r][:6]
        nc = ch >> 32
        u = lambda a, b: (b - a) / 1e3 if (a > 0 and b > 0) else 0.0
        t6, t7, t8, t9, t10, t11 = t[r][6:12]
        print(f"{r:4d} {(t1-t0)/1e3:8.1f} {(t3-t1)/1e3:8.1f}  {names[ph & 7]:3s} {nc:6d} | load {u(t1,t8):.1f} fhist-x+scan {u(t8,t6):.1f} pass2 {u(t6,t9):.1f} tie-x+rank {u(t9,t7):.1f} sum {u(t7,t10):.1f} linalg {u(t10,t3):.1f}")
    print("total us", (t[R][3] - t[1][0]) / 1e3)
dist.barrier()
s.close()
dist.destroy_process_group()
