#!/bin/bash
# launch list (gpu__time_duration) of the PPO update at 8192 envs: two updates of 2 epochs each (scripts/prof_update.py), learner kernels and library kernels
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:mma_kernel|tc_kernel|head_loss|rows_kernel|cutlass|gemv|sgemm|elementwise|reduce_kernel|scale_unless' --csv \
  --log-file gpurun_out/r02d_launches_update8192.csv python scripts/prof_update.py 8192 > gpurun_out/r02d_launches_update.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02d_launches_update8192.csv")) if len(r)>5]
hdr=[r for r in rows if "Kernel Name" in r][0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
t=collections.Counter(); n=collections.Counter()
for r in rows:
    if r is hdr or len(r)<=vi: continue
    try: v=float(r[vi].replace(",",""))
    except ValueError: continue
    k=r[ki].split("(")[0][:56]; t[k]+=v; n[k]+=1
tot=sum(t.values())
out=[f"{k:58s} {n[k]:5d} launches {v/1e6:9.2f} ms {100*v/tot:5.1f}%" for k,v in t.most_common(14)]
open("gpurun_out/r02d_launches_update8192_summary.txt","w").write("\n".join(out)+f"\ntotal {tot/1e6:.2f} ms over {sum(n.values())} launches (4 epochs)\n")
print("\n".join(out))
PY
