"""Dev tool: per-SASS-instruction memory work of one kernel from an .ncu-rep (ncu source page, no GPU needed).
    python tools/ncu_mem_instr.py REP KERNEL_REGEX
Lists every instruction that touched shared or global memory in the first matching launch: warp-level executions,
shared-memory wavefronts (actual / ideal / excess = bank conflicts), L1 tag requests (= 128-byte lines touched by global
accesses) and the L2 sectors it asked for; identical rows are grouped.  Totals per warp at the end."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                      f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
seen, groups, order = set(), {}, []
tot = {"wf": 0, "ex": 0, "tags": 0, "sect": 0}
warps = 0
for r in rows:
    if len(r) < len(hdr) or r[0] in seen or not r[0].startswith("0x"):
        continue
    seen.add(r[0])
    src = r[ix["Source"]].strip()
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    try:
        inst = int(r[ix["Instructions Executed"]])
        wf, ideal, ex = (int(r[ix[k]]) for k in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Wavefronts Shared Excessive"))
        tags, sect = int(r[ix["L1 Tag Requests Global"]]), int(r[ix["L2 Theoretical Sectors Global"]])
    except ValueError:
        continue
    warps = warps or inst  # the first instruction of the kernel: executed once by every warp
    if not (wf or tags):
        continue
    key = (op, inst, wf, ideal, ex, tags, sect)
    if key not in groups:
        order.append(key)
    groups[key] = groups.get(key, 0) + 1
    tot["wf"] += wf; tot["ex"] += ex; tot["tags"] += tags; tot["sect"] += sect
print(f"{'n':>3s} {'op':14s} {'warp execs':>10s} {'smem wf':>9s} {'ideal':>9s} {'excess':>8s} {'L1 tags':>9s} {'L2 sectors':>10s}  per exec")
for k in order:
    op, inst, wf, ideal, ex, tags, sect = k
    per = f"{wf / inst:.2f} wf" if wf else f"{tags / inst:.2f} lines, {sect / inst:.1f} sectors"
    print(f"{groups[k]:3d} {op:14s} {inst:10d} {wf:9d} {ideal:9d} {ex:8d} {tags:9d} {sect:10d}  {per}")
print(f"warps: {warps}; per warp: shared wavefronts {tot['wf'] / warps:.1f} (excess {tot['ex'] / warps:.1f}), "
      f"global lines {tot['tags'] / warps:.1f}, L2 sectors {tot['sect'] / warps:.1f}")
