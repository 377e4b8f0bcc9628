"""Dev tool: aggregate the ncu source page (per-SASS rows with -lineinfo) by CUDA source file:line.
    python tools/ncu_source.py REP KERNEL_REGEX [top]
For the first matching kernel launch: executed warp instructions and stall samples per source line, plus totals
per stall reason."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda",
                      "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = []      # (file, line, src, samples, inst)
fname, hdr, seen_fn = None, None, {}
stall_tot = {}
sass_ops = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        key = fname
        seen_fn[key] = seen_fn.get(key, 0) + 1
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or seen_fn.get(fname, 0) > 1:   # only the first launch's copy of each file
        continue
    i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    if r[0] != "":
        try:
            lines.append((fname, int(r[0]), r[1].strip()[:90], int(r[i_s]), int(r[i_i])))
        except ValueError:
            pass
        for j, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    stall_tot[h] = stall_tot.get(h, 0) + int(r[j])
                except ValueError:
                    pass
    elif len(r) > 3 and r[2].startswith("0x"):
        op = r[3].split()[0] if r[3].split() else "?"
        if op.startswith("@"):
            op = r[3].split()[1]
        op = op.split(".")[0]
        try:
            sass_ops[op] = sass_ops.get(op, 0) + int(r[i_i])
        except ValueError:
            pass
tot_s = sum(l[3] for l in lines) or 1
tot_i = sum(l[4] for l in lines) or 1
print(f"total samples {tot_s}, warp instructions {tot_i}")
print("stall totals:", {k: round(v / tot_s, 3) for k, v in sorted(stall_tot.items(), key=lambda kv: -kv[1]) if v / tot_s > 0.01})
print("top SASS opcodes by executed warp instructions (collapsed rows not included):")
for op, c in sorted(sass_ops.items(), key=lambda kv: -kv[1])[:18]:
    print(f"   {op:10s} {c:10d}")
print(f"\n{'file:line':32s} {'samples%':>8s} {'inst%':>7s}  source")
for f, ln, src, s, i in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{f + ':' + str(ln):32s} {100 * s / tot_s:8.2f} {100 * i / tot_i:7.2f}  {src}")
