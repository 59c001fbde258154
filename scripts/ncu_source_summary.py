"""Summarise an ncu source page (SASS view): opcode mix, stall samples and hot regions of the first kernel."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
ki = int(sys.argv[3]) if len(sys.argv) > 3 else 0        # which kernel of the report
hi = starts[ki]; end = starts[ki + 1] - 1 if len(starts) > ki + 1 else len(rows)
hdr = rows[hi]; data = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
ia = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); ith = hdr.index("Thread Instructions Executed")
tot_s = sum(int(r[isamp]) for r in data); tot_e = sum(int(r[iex]) for r in data); tot_t = sum(int(r[ith]) for r in data)
print("sass rows", len(data), "stall samples", tot_s, "warp inst executed", tot_e, "thread inst", tot_t)
op = collections.Counter(); ops = collections.Counter(); opt = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
    if not m:
        continue
    o = m.group(2).split(".")[0]
    op[o] += int(r[iex]); ops[o] += int(r[isamp]); opt[o] += int(r[ith])
print("opcode      warp-inst  share   stall-samples  thread-inst")
for o, c in op.most_common(30):
    print(f"{o:10s} {c:9d} {100 * c / tot_e:5.1f}%   {100 * ops[o] / max(tot_s, 1):5.1f}%   {opt[o]}")
# packed instructions (sm_100a FFMA2 / FMUL2 / FADD2) do two lanes' worth of work per thread instruction
fl = 2 * opt["FFMA"] + opt["FADD"] + opt["FMUL"] + opt["MUFU"] + opt["FMNMX"] + 4 * opt["FFMA2"] + 2 * opt["FMUL2"] + 2 * opt["FADD2"]
print("FP32 flop (2*FFMA + FADD + FMUL + MUFU + FMNMX + 4*FFMA2 + 2*FMUL2 + 2*FADD2), thread level:", fl)
stall_cols = [h for h in hdr if h.startswith("stall_")]
if stall_cols:
    tot = {h: sum(int(r[hdr.index(h)] or 0) for r in data) for h in stall_cols}
    print("stall reasons:", {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]})
n = len(data); B = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for b in range(B):
    seg = data[b * n // B:(b + 1) * n // B]
    e = sum(int(r[iex]) for r in seg); s = sum(int(r[isamp]) for r in seg)
    print(f"seg {b:2d} sass[{b * n // B:5d}:{(b + 1) * n // B:5d}] exec {100 * e / tot_e:5.1f}%  samples {100 * s / max(tot_s, 1):5.1f}%")
