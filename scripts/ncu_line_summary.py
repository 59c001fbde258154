"""Per-CUDA-source-line instruction/stall summary from an ncu report (cuda,sass view)."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
# layout: blocks per file: "File Name", path ; header row starting with "Line No" ; per line rows, followed by sass rows?
cur_file = None; hdr = None
lines = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Name": cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r[0] in ("Line No", "#"): hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if hdr[0] == "Line No" and r[0].isdigit():
        try:
            ex = int(r[hdr.index("Instructions Executed")] or 0); sm = int(r[hdr.index("# Samples")] or 0)
        except (ValueError, IndexError):
            continue
        if ex or sm:
            lines[(cur_file, int(r[0]))] = (ex, sm, r[1].strip()[:110])
tot_e = sum(v[0] for v in lines.values()); tot_s = sum(v[1] for v in lines.values())
print("total warp inst", tot_e, "samples", tot_s, "lines", len(lines))
for (f, ln), (ex, sm, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*ex/tot_e:5.1f}% exec {100*sm/max(tot_s,1):5.1f}% stall  {f}:{ln:4d}  {src}")
