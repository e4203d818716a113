"""static SASS opcode histograms of the hot kernels of libsampleqc_cuda.so (cuobjdump -sass), plus the data-movement
mnemonics the profiling recipe asks about (UBLKCP / UTMALDG / LDGSTS / LDG.E.128 ...).

    python scripts/sass_histogram.py > profiles/r2_sass_histograms.md
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "sampleqc_b200", "libsampleqc_cuda.so")
KERNELS = [("k_mcd<3, false>", "_ZN3sqc5k_mcdILi3ELb0EEEvNS_3DevENS_7McdBufsE"),
           ("k_estep<3>", "_ZN3sqc7k_estepILi3EEEvNS_3DevE"),
           ("k_partition<3, 1>", "_ZN3sqc11k_partitionILi3ELi1EEEvNS_3DevE"),
           ("k_mle_pass<3, MLE_B>", None),
           ("k_mt_jump", "_ZN3sqc9k_mt_jumpEPKjS1_S1_PjxiPKNS_7ControlE"),
           ("k_mt_gen", "_ZN3sqc8k_mt_genEPKjPjS1_S1_PixiPKNS_7ControlE")]
print("# SASS opcode histograms (static), sm_100a, `cuobjdump -sass` of sampleqc_b200/libsampleqc_cuda.so\n")
print("Static counts: loops execute their bodies many times (the per-launch dynamic mix of the E-step / partition kernels is in "
      "`r2z_ncu_summary.md`).  Data movement is through `LDG` (128-bit where the layout allows: `LDG.E.NA.128.CONSTANT` = `ld.global.nc.L1::no_allocate.v2.f64` in the streaming loop of `k_mcd`) and `STG`; "
      "no `UBLKCP` / `UTMALDG` — a TMA-staged variant of the two per-cell passes was built and measured slower (DESIGN.md section 8).\n")
names = subprocess.run(["cuobjdump", "-elf", SO], capture_output=True, text=True).stdout
for title, sym in KERNELS:
    if sym is None:
        m = re.search(r"(_ZN3sqc10k_mle_passILi3ELi2EE\w+)", names)
        sym = m.group(1) if m else None
        if sym is None:
            continue
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", sym, SO], capture_output=True, text=True).stdout
    ops, wide = collections.Counter(), collections.Counter()
    for l in txt.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", l)
        if not m:
            continue
        t = re.sub(r"^@!?U?P\d+\s+", "", m.group(1))
        full = t.split()[0]
        ops[full.split(".")[0]] += 1
        if full.startswith(("LDG", "STG", "LDS", "STS", "LDL", "STL", "UBLKCP", "UTMALDG", "LDGSTS", "ATOM", "RED", "MUFU")):
            wide[full] += 1
    tot = sum(ops.values())
    if not tot:
        continue
    print(f"## {title}  ({tot} instructions)\n")
    print("| opcode | count | share |\n|---|---|---|")
    for o, n in ops.most_common(22):
        print(f"| {o} | {n} | {100 * n / tot:.1f} % |")
    print("\nmemory / special mnemonics: " + ", ".join(f"`{o}` {n}" for o, n in wide.most_common(14)) + "\n")
