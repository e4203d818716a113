"""per-round timeline of the last FastMCD kernel of a C3 fit (debug)"""
import os, sys, ctypes as C
os.environ["SQC_MCD_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import workload
from sampleqc_b200 import api
from sampleqc_b200._lib import lib
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
x, g, iz, prm = workload("C3", scale)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
em = int(sys.argv[2]) if len(sys.argv) > 2 else 6
s = api.Session(x, g, D, J, K, N, em, init_z=iz, rng_mode=1)
s.profile(True)
s.run(); s.run()
st = s.stats()
print({k: v for k, v in st.items() if k != "family_bytes"})
buf = np.zeros(8 * 512, dtype=np.int64)
L = lib(); L.sqc_session_mcd_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
L.sqc_session_mcd_trace(s._h, buf.ctypes.data_as(C.c_void_p), buf.size)
R = int(buf[0]); t = buf.reshape(-1, 8)
names = "KH KC DH DC DONE".split()
print("round  stream_us  sync1_us  post_us  sync2+map_us  chunks  ncand presort_us sort_us phases")
for r in range(1, R + 1):
    t0, t1, t2, t3, ph, ch, t6, t7 = t[r][:8]
    nc = ch >> 32; ch &= 0xffffffff
    nxt = t[r + 1][0] if r < R else t3
    phs = " ".join(names[(ph >> (3 * k)) & 7] for k in range(K))
    print(f"{r:4d} {(t1-t0)/1e3:9.1f} {(t2-t1)/1e3:9.1f} {(t3-t2)/1e3:8.1f} {(nxt-t3)/1e3:9.1f} {ch:8d} {nc:6d} {(t6-t2)/1e3 if nc or t6>t2 else 0:8.1f} {(t7-t6)/1e3 if t7>t6 else 0:8.1f}  {phs}")
