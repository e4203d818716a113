"""per-round timelines of ALL cluster leaders of the last FastMCD launch of a C3 fit (debug / design input)

    python scripts/mcd_trace_all.py [scale] [em_iters] [rng_mode]
"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import workload
from sampleqc_b200 import api
from sampleqc_b200._lib import lib

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
em = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rng = int(sys.argv[3]) if len(sys.argv) > 3 else 1
x, g, iz, prm = workload("C3", scale)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
L = lib(); L.sqc_session_mcd_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
names = "KH KC DH DC DONE DI ? ?".split()
rows = []
for leader in range(K):
    os.environ["SQC_MCD_TRACE"] = str(leader)
    s = api.Session(x, g, D, J, K, N, em, init_z=iz, rng_mode=rng)
    s.profile(True)
    s.run(); s.run()
    st = s.stats()
    if leader == 0:
        print({k: v for k, v in st.items() if k != "family_bytes"})
    buf = np.zeros(16 * 512, dtype=np.int64)
    L.sqc_session_mcd_trace(s._h, buf.ctypes.data_as(C.c_void_p), buf.size)
    s.close()
    R = int(buf[0]); t = buf.reshape(-1, 16)
    base = t[1][0]
    print(f"leader {leader}: rounds {R}  span {(t[R][3] - t[1][0]) / 1e3:.1f} us   kernel entry -> first publish {(t[0][2] - t[0][1]) / 1e3:.1f} us, entry -> end of last post {(t[R][3] - t[0][1]) / 1e3:.1f} us")
    print("  ep  start_us  wait_us  post_us  ph  dir  ncand   T")
    for r in range(1, R + 1):
        t0, t1, t2, t3, ph, ch = t[r][:6]
        nc = ch >> 32
        T = np.array([t[r][12]]).view(np.float64)[0]
        u = lambda a, b: (b - a) / 1e3 if (a > 0 and b > 0) else 0.0
        t6, t7, t8, t9, t10, t11 = t[r][6:12]
        print(f"  {r:3d} {(t0 - base) / 1e3:8.1f} {(t1 - t0) / 1e3:8.1f} {(t3 - t1) / 1e3:8.1f}  {names[ph & 7]:3s} {int(t[r][13]):3d} {nc:6d}  T={T:.6f} | fhist {u(t1, t8):.1f} scan {u(t8, t6):.1f} pass2 {u(t6, t9):.1f} tie {u(t9, t7):.1f} partials {u(t7, t10):.1f} reduce {u(t10, t11):.1f} linalg {u(t11, t3):.1f}")
