"""CPU-side checks: the C-ABI library loads and exports every symbol include/sampleqc_cuda.h declares, the ctypes
structures match the header, compute entry points fail loudly without a GPU (no CPU fallback), and the
multi-GPU host logic (sharding, handle exchange, result assembly) works under gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from sampleqc_b200 import build
    build.build()
    from sampleqc_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol():
    L = _lib()
    hdr = open(os.path.join(ROOT, "include", "sampleqc_cuda.h")).read()
    declared = set(re.findall(r"SQC_API\s+(?:const\s+)?\w+\*?\s+\*?(sqc_\w+)\s*\(", hdr))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    lib = L.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.sqc_abi_version() == 2
    assert lib.sqc_status_name(2) == b"SQC_ERR_EMPTY_CLUSTER"


def test_ctypes_structs_match_header_sizes(tmp_path):
    from sampleqc_b200 import _abi as A
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "sampleqc_cuda.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(sqc_fit_args), sizeof(sqc_fit_out), sizeof(sqc_maha_args), sizeof(sqc_run_stats), sizeof(sqc_init_args), sizeof(sqc_mmd_args));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(A.SqcFitArgs), C.sizeof(A.SqcFitOut), C.sizeof(A.SqcMahaArgs), C.sizeof(A.SqcRunStats), C.sizeof(A.SqcInitArgs), C.sizeof(A.SqcMmdArgs)]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from sampleqc_b200 import api
    from sampleqc_b200._abi import SampleQCError
    _lib()
    x = np.asfortranarray(np.random.default_rng(0).normal(size=(50, 3)))
    g = np.repeat(np.arange(5, dtype=np.int32), 10)
    with pytest.raises(SampleQCError) as e:
        api.fit_sampleqc_robust_cpp(x, np.zeros(50, np.int32), g, 3, 5, 1, 50, 3, 0.5, 10, 0)
    assert e.value.status == -1                      # SQC_ERR_CUDA
    with pytest.raises(SampleQCError):
        api.fit_sampleqc_mle_cpp(x, np.ones((50, 1)), g, 3, 5, 1, 50, 2, 0)
    with pytest.raises(SampleQCError):
        api.calc_outliers(x, g, np.zeros(3), np.zeros((5, 3)), np.zeros((1, 3)), np.eye(3).reshape(3, 3, 1), 0.01)


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sampleqc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "sqc_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_shard_by_sample_balanced_and_contiguous():
    from sampleqc_b200.multigpu import shard_by_sample
    rng = np.random.default_rng(1)
    n_j = rng.integers(5, 500, size=37)
    groups = np.repeat(np.arange(37), n_j)
    for world in (1, 2, 3, 4, 8):
        sh = shard_by_sample(groups, 37, world)
        assert sh[0][0] == 0 and sh[-1][1] == 37 and sh[0][2] == 0 and sh[-1][3] == groups.size
        for a, b in zip(sh[:-1], sh[1:]):
            assert a[1] == b[0] and a[3] == b[2]
        for (j0, j1, i0, i1) in sh:
            assert i1 - i0 == n_j[j0:j1].sum()
        assert max(i1 - i0 for (_, _, i0, i1) in sh) <= groups.size / world + n_j.max()
    with pytest.raises(ValueError):
        shard_by_sample(groups[::-1], 37, 2)
    with pytest.raises(ValueError):                               # a rank without cells: refused on every rank alike
        shard_by_sample(np.repeat(np.arange(3), [1000, 5, 5]), 3, 4)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
from sampleqc_b200 import multigpu
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
# handle all-gather: rank r contributes 64 bytes of value r + 1
h = multigpu.gather_handles(dist, bytes([rank + 1]) * 64, 2)
assert h == bytes([1]) * 64 + bytes([2]) * 64, h
# result assembly: global parameters identical, per-sample rows / labels concatenated in rank order
J, K, D, N = 5, 2, 3, 40
groups = np.repeat(np.arange(J), 8)
j0, j1, i0, i1 = multigpu.shard_by_sample(groups, J, 2)[rank]
loc = {"alpha_j": np.full((j1 - j0, D), rank + 1.0), "p_jk": np.full((j1 - j0, K), 0.5), "z": np.full((i1 - i0, 1), rank + 1.0),
       "beta_k": np.arange(K * D, dtype=float).reshape(K, D), "n_zero_like": rank + 3, "n_iter": 7}
full = multigpu.assemble(dist, rank, 2, loc, (j0, j1, i0, i1), J, N)
if rank == 0:
    assert full["alpha_j"].shape == (J, D) and full["z"].shape == (N, 1) and full["p_jk"].shape == (J, K)
    assert full["alpha_j"][0, 0] == 1.0 and full["alpha_j"][-1, 0] == 2.0 and full["z"][0, 0] == 1.0 and full["z"][-1, 0] == 2.0
    assert full["n_zero_like"] == 7 and full["n_iter"] == 7 and full["J"] == J and full["N"] == N
else:
    assert full is None
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


def test_multigpu_host_logic_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in o, o


def test_library_shard_plan_for_n_gpus():
    """sqc_fit_args.n_gpus: the library's own sample-aligned sharding (host-only hook): contiguous sample runs in any
    code order are cut at run boundaries with balanced cell counts; fewer samples than GPUs -> fewer shards;
    interleaved samples -> 0 (single GPU)."""
    import ctypes as C
    from sampleqc_b200 import _abi as A
    from sampleqc_b200._lib import lib
    L = lib()
    rng = np.random.default_rng(3)
    n_j = rng.integers(50, 4000, size=37)
    code = rng.permutation(37).astype(np.int32)                   # what factor(sample_ids) produces: runs in non-ascending code order
    g = np.repeat(code, n_j).astype(np.int32)
    N, J = int(g.size), 40                                        # codes 37..39 have no cells
    for n in (2, 3, 8):
        cells = (C.c_int64 * (n + 1))()
        owner = np.full(J, -7, dtype=np.int32)
        got = L.sqc_plan_shards(A.ptr_i(g), N, J, n, cells, A.ptr_i(owner))
        assert got == n
        b = list(cells)
        assert b[0] == 0 and b[-1] == N and all(b[i] < b[i + 1] for i in range(n))
        starts = np.concatenate([[0], np.cumsum(n_j)])
        assert all(x in starts for x in b)                        # cuts fall on sample boundaries
        sizes = np.diff(b)
        assert sizes.max() - sizes.min() <= 2 * n_j.max()         # balanced up to one sample
        for j in range(37):                                       # every sample lives in exactly the shard that holds its cells
            i0 = starts[np.where(code == j)[0][0]]
            assert b[owner[j]] <= i0 < b[owner[j] + 1]
        assert list(owner[37:]) == [-1, -1, -1]
    cells = (C.c_int64 * 9)()
    assert L.sqc_plan_shards(A.ptr_i(g[:n_j[0] + n_j[1]]), int(n_j[0] + n_j[1]), J, 8, cells, None) == 2      # 2 samples: 2 shards
    inter = np.tile(np.array([0, 1], dtype=np.int32), 500)
    assert L.sqc_plan_shards(A.ptr_i(inter), 1000, 2, 2, cells, None) == 0
    bad = g.copy(); bad[5] = 99
    assert L.sqc_plan_shards(A.ptr_i(bad), N, J, 2, cells, None) < 0


def test_mt_jump_polynomials_move_the_state(tmp_path):
    """the distributed key stream moves R's Mersenne-Twister state e words ahead with x^(e-1) mod phi derived on the
    host: check the derivation against an independent big-integer computation and against a directly generated stream
    (window[j] = XOR over the polynomial's terms i of x[i + j + 1])"""
    L = _lib()
    lib = L.lib()
    inc = open(os.path.join(ROOT, "sampleqc_b200", "csrc", "sqc_mt_jump.inc")).read()
    terms = [int(t) for t in re.search(r"sqc_mt_phi_terms\[\d+\] = \{([^}]*)\}", inc).group(1).split(",")]
    assert terms[-1] == 19937 and len(terms) == 135
    phi = sum(1 << t for t in terms)

    def mulmod(a, b):
        r = 0
        while a:
            low = a & -a
            r ^= b << (low.bit_length() - 1)
            a ^= low
        while r.bit_length() > 19937:
            r ^= phi << (r.bit_length() - 1 - 19937)
        return r

    def xpow(e):
        result, base = 1, 2
        while e:
            if e & 1:
                result = mulmod(result, base)
            base = mulmod(base, base)
            e >>= 1
        return result

    def lib_poly(e):
        out = np.zeros(624, dtype=np.uint32)
        assert lib.sqc_mt_jump_poly(e, out.ctypes.data_as(C.c_void_p)) == 0
        return sum(int(w) << (32 * i) for i, w in enumerate(out))

    for e in (1, 623, 19936, 19937, 624 * 1602 - 1, 3_000_001, 80_000_000 - 1, (1 << 40) + 12345):
        assert lib_poly(e) == xpow(e), e

    # against a direct stream: x[0..623] = seeded array, x[n + 624] = x[n + 397] ^ twist(x[n], x[n + 1])
    def stream(seed, n_words):
        s = seed & 0xFFFFFFFF
        for _ in range(50):
            s = (69069 * s + 1) & 0xFFFFFFFF
        s = (69069 * s + 1) & 0xFFFFFFFF
        x = []
        for _ in range(624):
            s = (69069 * s + 1) & 0xFFFFFFFF
            x.append(s)
        for n in range(n_words - 624):
            y = (x[n] & 0x80000000) | (x[n + 1] & 0x7FFFFFFF)
            x.append(x[n + 397] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0))
        return np.array(x, dtype=np.uint32)

    e = 624 * 40                                         # 40 chunks ahead
    x = stream(3, e + 624 + 19937 + 8)
    g = lib_poly(e - 1)
    idx = np.array([i for i in range(19937) if (g >> i) & 1], dtype=np.int64)
    js = np.arange(624)
    win = np.bitwise_xor.reduce(x[idx[:, None] + js[None, :] + 1], axis=0)
    np.testing.assert_array_equal(win, x[e:e + 624])


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with the
    contract's keys; here on the small C1 workload"""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "EM cell-iters/s" and d["unit"] == "cell-iters/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
