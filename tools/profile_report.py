"""Dev tool: turn one gpurun_out/<tag>/ directory (tools/gpu_round.sh) into the tracked summaries under profiles/.

    python tools/profile_report.py gpurun_out/r1a r1a

Writes profiles/<tag>_launches_c2.md (ncu launch list: per-kernel count / mean / share), profiles/<tag>_ncu_<wl>.txt
(headline ncu --set full metrics per kernel), profiles/<tag>_bench.jsonl (the bench lines of that run) and
refreshes profiles/roofline_traffic.json (dram bytes per launch, read by bench.py).
"""
import collections
import csv
import json
import os
import subprocess
import sys

src, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
SHORT = {"wso_pass1_kernel": "K1", "wso_heights_kernel": "K2h", "wso_pass2_kernel": "K2"}


def short(name):
    for k, v in SHORT.items():
        if k in name:
            return v
    return name[:40]


# ---- launch list -------------------------------------------------------------------------------------
for f in sorted(os.listdir(src)):
    if f.startswith("launches_") and f.endswith(".csv"):
        rows = list(csv.reader(open(os.path.join(src, f))))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        hdr = rows[hi]
        ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
        agg = collections.OrderedDict()
        for r in rows[hi + 1:]:
            if len(r) <= vi:
                continue
            agg.setdefault((r[ki], r[gi], r[bi]), []).append(float(r[vi].replace(",", "")))
        tot = sum(sum(v) for v in agg.values())
        with open(os.path.join(out, f"{tag}_{f[:-4]}.md"), "w") as o:
            o.write(f"# ncu launch list ({f}, {tag})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over "
                    "`bench.py --steps 2 --warmup 3`; cold-cache, serialised: compare SHARES, not absolutes.\n\n")
            o.write("| kernel | grid | block | launches | mean us | share |\n|---|---|---|---|---|---|\n")
            for (k, g, b), v in agg.items():
                o.write(f"| {k} | {g} | {b} | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {sum(v) / tot:.3f} |\n")

# ---- full captures -----------------------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__waves_per_multiprocessor"]
tpath = os.path.join(out, "roofline_traffic.json")
traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
for f in sorted(os.listdir(src)):
    if f.startswith("prof_") and f.endswith(".ncu-rep"):
        wl = f[5:-8]
        res = subprocess.run(["ncu", "-i", os.path.join(src, f), "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(res.stdout.splitlines()))
        hdr, units = rows[0], rows[1]
        stall = [h for h in hdr if "average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
        seen = {}
        with open(os.path.join(out, f"{tag}_ncu_{wl}.txt"), "w") as o:
            o.write(f"ncu --set full --clock-control none --import-source on -k regex:wso_ ... python bench.py --workload {wl} "
                    f"--steps 2 --warmup 3   ({tag}; one row set per distinct kernel, first captured launch)\n"
                    "Durations are under the profiler (cold caches, serialised) - not bench values.\n")
            for vals in rows[2:]:
                name = vals[hdr.index("Kernel Name")]
                if name in seen:
                    continue
                seen[name] = 1
                o.write(f"\n---- {name}\n")
                for w in WANT + stall:
                    if w not in hdr:
                        continue
                    i = hdr.index(w)
                    try:
                        if "stalled" in w and float(vals[i].replace(",", "")) < 0.2:
                            continue
                    except ValueError:
                        pass
                    o.write(f"  {w.replace('smsp__average_warps_issue_stalled_', 'stall_'):68s} {vals[i]:>18s} {units[i]}\n")
                rd = float(vals[hdr.index("dram__bytes_read.sum")].replace(",", ""))
                wr = float(vals[hdr.index("dram__bytes_write.sum")].replace(",", ""))
                ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
                mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                traffic.setdefault(wl, {})[short(name)] = rd * mul.get(ur, 1) + wr * mul.get(uw, 1)
json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)

# ---- bench lines -------------------------------------------------------------------------------------
with open(os.path.join(out, f"{tag}_bench.jsonl"), "w") as o:
    for f in sorted(os.listdir(src)):
        if f.startswith("bench_") and f.endswith(".json"):
            for line in open(os.path.join(src, f)):
                if line.startswith("{"):
                    d = json.loads(line)
                    d["_file"] = f
                    o.write(json.dumps(d) + "\n")
for f in ("smi.txt", "host.txt"):
    p = os.path.join(src, f)
    if os.path.exists(p):
        open(os.path.join(out, f"{tag}_{f}"), "w").write(open(p).read())
print("profiles written for", tag)
