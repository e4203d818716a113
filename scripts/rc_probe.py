import os, sys
sys.path.insert(0, ".")
from bench import workload
from sampleqc_b200 import api
x, g, iz, prm = workload("C3", 1.0)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
for rng in (0, 1):
    s = api.Session(x, g, D, J, K, N, 12, init_z=iz, rng_mode=rng)
    s.profile(True); s.run(); s.run()
    st = s.stats()
    print(sys.argv[1:], "rng_mode", rng, "loop %.2f ms (%.0f us/iter)" % (st["loop_ms"], st["loop_ms"] / 12 * 1e3), {k: round(v, 2) for k, v in st["family_ms"].items() if v})
    s.close()
