"""Static SASS footprint of the first kernel in an ncu report, per CUDA source line and per executed-count class
(rows executed >= `hot` times per launch form the hot loop)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
# find SASS table (header starts with "Address") and CUDA tables ("Line No")
cur = None; hdr = None; per_line = collections.OrderedDict(); sass = []
for r in rows:
    if not r: continue
    if r[0] == "File Name": cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] in ("Line No", "#", "Address"): hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if hdr[0] == "Address":
        sass.append(dict(zip(hdr, r)))
print("sass rows", len(sass))
ex = [int(x.get("Instructions Executed") or 0) for x in sass]
mx = max(ex)
hist = collections.Counter()
for e in ex:
    hist[round(e / mx * 8)] += 1
print("rows by executed count (in eighths of the max):", dict(sorted(hist.items())))
hot = [x for x, e in zip(sass, ex) if e >= 0.5 * mx]
print("hot rows (>= half the max executed count):", len(hot), "=", len(hot) * 16 / 1024, "KB")
warm = [x for x, e in zip(sass, ex) if 0.05 * mx <= e < 0.5 * mx]
print("warm rows (5%..50%):", len(warm), "=", len(warm) * 16 / 1024, "KB")
