import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from sampleqc_b200 import api
x, g, iz, prm = workload("C3", 1.0)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
for nb in (58, 49, 42, 37, 58, 49, 42, 37):
    os.environ["SQC_MT_BLOCKS"] = str(nb)
    s = api.Session(x, g, D, J, K, N, 50, init_z=iz, rng_mode=0)
    s.profile(False); s.run(); s.run()
    st = s.stats()
    print(f"blocks={nb}: loop {st['loop_ms']:.2f} ms ({st['loop_ms']/st['n_iter']*1e3:.0f} us/iter)")
    s.close()
