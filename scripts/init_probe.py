"""timing of the device .init_clusts pieces at C3 size (10 M cells x 1000 samples), against numpy on the host"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import workload
from sampleqc_b200 import api
x, g, iz, prm = workload("C3", float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
D, J, K, N = prm["D"], prm["J"], prm["K"], prm["N"]
from sklearn.mixture import GaussianMixture
api.sample_medians(x[:1000], g[:1000], J)                   # context + module load
for rep in range(2):
    t0 = time.perf_counter(); med = api.sample_medians(x, g, J); t1 = time.perf_counter()
    print(f"device medians (incl. H2D of x): {1e3 * (t1 - t0):.1f} ms")
idx = np.random.default_rng(0).choice(N, 10000, replace=False)
gm = GaussianMixture(K, covariance_type="full", random_state=0).fit(x[idx] - med[g[idx]])
sigma = np.asfortranarray(np.transpose(gm.covariances_, (1, 2, 0)))
for rep in range(2):
    t0 = time.perf_counter(); m2, gamma, z = api.gmm_classify(x, g, J, gm.weights_, gm.means_, sigma); t1 = time.perf_counter()
    print(f"device medians + classify all cells (incl. H2D of x, D2H of gamma N x K and z): {1e3 * (t1 - t0):.1f} ms")
t0 = time.perf_counter()
bounds = np.searchsorted(g, np.arange(J + 1))
med_h = np.stack([np.median(x[bounds[j]:bounds[j + 1]], axis=0) for j in range(J)])
t1 = time.perf_counter()
ga = gm.predict_proba(x - med_h[g]); t2 = time.perf_counter()
print(f"host numpy medians {1e3 * (t1 - t0):.0f} ms, sklearn predict_proba {1e3 * (t2 - t1):.0f} ms")
print("medians identical:", np.array_equal(med, med_h), " max |gamma - sklearn|:", float(np.abs(gamma - ga).max()))
