import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from sampleqc_b200 import api
x, g, iz, prm = workload("C3", 1.0)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
for inline in (0,):
    for nb in (8, 12, 20, 40, 146):
        os.environ["SQC_RNG_INLINE"] = str(inline); os.environ["SQC_MT_BLOCKS"] = str(nb)
        s = api.Session(x, g, D, J, K, N, 6, init_z=iz, rng_mode=0)
        s.profile(True); s.run(); s.run()
        st = s.stats()
        print(f"inline={inline} blocks={nb}: loop {st['loop_ms']:.2f} ms, rng {st['family_ms']['rng']:.2f} ms over {st['family_launches']['rng']} launches, estep {st['family_ms']['estep']:.2f} mcd {st['family_ms']['mcd']:.2f} part {st['family_ms']['partition']:.2f}")
        s.close()
