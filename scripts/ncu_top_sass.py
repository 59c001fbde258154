"""Top SASS instructions of the first kernel in an ncu report by stall samples (with the dominant stall reason)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = starts[0]; end = starts[1] - 1 if len(starts) > 1 else len(rows)
hdr = rows[hi]; data = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
isrc = hdr.index("Source"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "rows", len(data))
agg = {}
for i, h in stall:
    agg[h] = sum(int(r[i] or 0) for r in data)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for n, r in sorted(enumerate(data), key=lambda kv: -int(kv[1][isamp]))[:top]:
    why = sorted(((int(r[i] or 0), h) for i, h in stall), reverse=True)[:2]
    print(f"{100 * int(r[isamp]) / tot:5.1f}%  #{n:5d} exec {r[iex]:>7s}  {r[isrc].strip()[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
