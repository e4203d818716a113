"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.
usage: python scripts/summarize_profiles.py <tag> [launches.csv] [prof.ncu-rep]"""
import collections, csv, json, os, subprocess, sys
tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches.csv"
rep = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/prof_r1.ncu-rep"
out = []
if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try: v = float(r[vi].replace(",", ""))
        except ValueError: continue
        n = r[ki].split("(")[0]; agg[n][0] += 1; agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    out.append(f"# ncu launch list ({tag}): gpu__time_duration.sum per kernel, --clock-control none (cold-cache, serialised: compare shares)\n\n")
    out.append("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]/1e6:.3f} | {v[1]/v[0]/1e3:.1f} | {v[1]/tot:.3f} |\n")
traffic = {}
if os.path.exists(rep):
    # a .ncu-rep, or the CSV `ncu -i rep --page raw --csv` wrote on the GPU box (the reports are too large to travel back)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
            "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
            "smsp__average_warp_latency_per_inst_issued.ratio"]
    out.append(f"\n# ncu --set full ({tag}), one block per captured launch\n\n")
    def tobytes(v, u):
        v = float(v.replace(",", "")); return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        out.append(f"## {name}\n")
        for w in want:
            if w in idx: out.append(f"- {w}: {r[idx[w]]} {units[idx[w]]}\n")
        st = sorted(((float(r[idx[h]]) if r[idx[h]] not in ("", "n/a") else 0.0, h) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")), reverse=True)[:6]
        if st: out.append("- warps stalled per issue slot: " + ", ".join(f"{h.split('issue_stalled_')[1].split('_per')[0]} {v:.2f}" for v, h in st) + "\n")
        t = tobytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + tobytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        fam = "mcd" if "k_mcd" in name else "estep" if "k_estep" in name else "partition" if "k_partition" in name else name
        traffic.setdefault(fam, []).append(t)
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open("profiles/traffic.json", "w"), indent=1)
import re
for kn in ("k_estep", "k_partition", "k_mcd"):
    src = f"gpurun_out/prof_{tag}_src_{kn}.csv"
    if not os.path.exists(src):
        continue
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); h = rows[hi]; ix = {n: i for i, n in enumerate(h)}
    data = [r for r in rows[hi + 1:] if r and r[0].startswith("0x")]
    if len(data) > 1 and data[0][1] == data[len(data) // 2][1]: data = data[:len(data) // 2]      # (the page lists the kernel twice)
    op = collections.Counter()
    for r in data:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
        o = m.group(2).split(".")[0] if m else "?"
        op[o] += int(r[ix["Instructions Executed"]] or 0)
    T = sum(op.values())
    out.append(f"\n## dynamic instruction mix of {kn} (one captured launch, {T} warp instructions)\n\n| opcode | warp instructions | share |\n|---|---|---|\n")
    for o, n in op.most_common(16): out.append(f"| {o} | {n} | {100 * n / T:.1f} % |\n")
open(f"profiles/{tag}_ncu_summary.md", "w").writelines(out)
