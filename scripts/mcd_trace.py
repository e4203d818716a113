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
buf = np.zeros(16 * 512, dtype=np.int64)
L = lib(); L.sqc_session_mcd_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
L.sqc_session_mcd_trace(s._h, buf.ctypes.data_as(C.c_void_p), buf.size)
R = int(buf[0]); t = buf.reshape(-1, 16)
names = "KH KC DH DC DONE".split()
print("round stream sync1  post  sync2 | finehist scan pass2 tie partials reduce+linalg | chunks ncand phases")
for r in range(1, R + 1):
    t0, t1, t2, t3, ph, ch, t6, t7, t8, t9, t10, t11 = t[r][:12]
    nc = ch >> 32; ch &= 0xffffffff
    nxt = t[r + 1][0] if r < R else t3
    phs = " ".join(names[(ph >> (3 * k)) & 7] for k in range(K))
    u = lambda a, b: (b - a) / 1e3 if (a > 0 and b > 0) else 0.0
    print(f"{r:4d} {u(t0,t1):6.1f} {u(t1,t2):5.1f} {u(t2,t3):5.1f} {u(t3,nxt):5.1f} | {u(t2,t8):6.1f} {u(t8,t6):5.1f} {u(t6,t9):5.1f} {u(t9,t7):5.1f} {u(t7,t10):6.1f} {u(t10,t3):6.1f} | {ch:6d} {nc:6d} {phs}  T={np.array([t[r][12]]).view(np.float64)[0]:.4f}")
