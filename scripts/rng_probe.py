import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from sampleqc_b200 import api
x, g, iz, prm = workload("C3", 1.0)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
for nb in (146, 73, 49, 37, 146):
    os.environ["SQC_MT_BLOCKS"] = str(nb)
    s = api.Session(x, g, D, J, K, N, 12, init_z=iz, rng_mode=0)
    s.profile(True); s.run(); s.run()
    st = s.stats()
    print(f"blocks={nb}: loop {st['loop_ms']:.2f} ms ({st['loop_ms']/12*1e3:.0f} us/iter), rng {st['family_ms']['rng']:.2f} ms, estep {st['family_ms']['estep']:.2f} mcd {st['family_ms']['mcd']:.2f} part {st['family_ms']['partition']:.2f}")
    s.close()
