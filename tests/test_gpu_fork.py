"""fit_sampleqc(K_list = ..., n_cores > 1) forks BiocParallel workers, each of which calls the export
(reference R/fit_sampleqc.R:77-92).  The library creates its CUDA context lazily in the calling process, so a parent
that has not touched CUDA can fork children that fit.  Run in a fresh interpreter (the pytest process has a context)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import os, sys, pickle
sys.path.insert(0, sys.argv[1])
import numpy as np
from sampleqc_b200 import api                      # loads the library; no CUDA call yet
from sampleqc_b200.simulate import simulate_qcs, init_clusts
from oracle import oracle as O
cases = []
for seed, K in ((1, 2), (2, 3)):                    # two "sample groups", as K_list has one entry per group
    N, J, D = 30_000, 6, 3
    sim = simulate_qcs(N, J, K, D, seed=seed)
    iz, _ = init_clusts(sim.x, sim.groups, J, K, seed=seed)
    cases.append((sim.x, iz, sim.groups, D, J, K, N))
pids = []
for i, c in enumerate(cases):
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:                                    # the forked worker: first CUDA use of this process
        os.close(r)
        try:
            fit = api.fit_sampleqc_robust_cpp(*c, 10, 0.5, 50, 0)
            payload = pickle.dumps({k: fit[k] for k in ("beta_k", "sigma_k", "z", "n_iter")})
        except Exception as e:
            payload = pickle.dumps({"error": repr(e)})
        os.write(w, len(payload).to_bytes(8, "little") + payload)
        os._exit(0)
    os.close(w)
    pids.append((pid, r))
ok = True
for (pid, r), c in zip(pids, cases):
    n = int.from_bytes(os.read(r, 8), "little")
    buf = b""
    while len(buf) < n:
        buf += os.read(r, n - len(buf))
    os.waitpid(pid, 0)
    got = pickle.loads(buf)
    assert "error" not in got, got
    ref = O.fit_sampleqc_robust_cpp(*c, 10, 0.5, 50, 0)
    ok &= got["n_iter"] == ref["n_iter"] and np.array_equal(got["z"], ref["z"]) and np.allclose(got["beta_k"], ref["beta_k"], rtol=1e-8, atol=1e-10)
# the parent can still fit afterwards (its own context is created now)
fit = api.fit_sampleqc_robust_cpp(*cases[0], 10, 0.5, 50, 0)
ref = O.fit_sampleqc_robust_cpp(*cases[0], 10, 0.5, 50, 0)
ok &= np.array_equal(fit["z"], ref["z"])
print("FORK FIT", "OK" if ok else "FAILED")
'''


def test_forked_workers_fit(tmp_path):
    script = tmp_path / "fork_fit.py"
    script.write_text(_SCRIPT)
    out = subprocess.run([sys.executable, str(script), ROOT], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "FORK FIT OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
