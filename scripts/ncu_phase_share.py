"""Share of the executed warp instructions of env_step_kernel per phase of a substep, from an ncu report with source import.
Every SASS row is assigned to the phase of the nearest preceding row (in address order) whose CUDA line lies inside a phase range of
env_device.cuh / env_kernels.cu; rows of inlined helpers (cross, operators, ...) thereby inherit the phase of their call site."""
import bisect, collections, csv, re, subprocess, sys
rep = sys.argv[1]
src = open(sys.argv[2]).read().splitlines() if len(sys.argv) > 2 else None
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[2]; iex = hdr.index("Instructions Executed")
end = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][1]
sass = []   # (address, line, text, executed)
cur = None
for r in rows[3:end]:
    if len(r) != len(hdr): continue
    if r[0].isdigit(): cur = (int(r[0]), r[1]); continue
    if r[2].startswith("0x") and r[iex].isdigit(): sass.append((int(r[2], 16), cur[0], cur[1], int(r[iex])))
sass.sort()
# phase anchors: markers in the source text of the line
PH = [("FK", r"SINCOS|leg_fk|quat_cols|k\.e\d|k\.j\d|k\.toe"), ("RNEA bias forces", r"aj\d|al\d|F\d =|N\d =|n\d =|f\d_|d\.hl|d\.hb|fb|nb"),
      ("composite inertias, B / D blocks", r"hC|IC|mC|P\d =|L\d =|d\.B\[|D\.|add_point_mass|s\dl"), ("Schur complement, Cholesky, solves", r"d\.Y\[|S\[|A0\[|d\.L\[|Dinv|fwd6|bwd6|rsqrtf\(dj\)|tri\("),
      ("contact detection + set-up", r"xf|xc\d|hit\d|allhits|contact_setup|ct\.Q|ct\.G|ct\.T|E\[|Jl\d|cf\.c|cb\.c|vpre|qw_|box_reach|terrain_sample"),
      ("Jacobi sweeps", r"sweep|frozen|solve_one_contact|maxd|maxl|cf\.lam|dl\b|gs_visit|lnz|lt2|vn2|den\b"), ("velocity + configuration update", r"ytot|b\.v =|b\.w =|qd = qd|b\.p =|ow|ox|oy|oz|kk|q = axpy"),
      ("PD + clamp", r"torque_last|motor_|kp0|kd0|tt\[k\]|ilow_|rr_")]
anchors = []
for a, ln, text, e in sass:
    for name, pat in PH:
        if re.search(pat, text):
            anchors.append((a, name)); break
addr = [a for a, _ in anchors]
base = sass[0][3]                        # warps per launch (the first instruction runs once per warp)
tot = collections.Counter(); hot = 0
for a, ln, text, e in sass:
    if e < 7 * base: continue              # keep the substep loop (executed loop_count = 8 times per warp), drop prologue / epilogue
    i = bisect.bisect_right(addr, a) - 1
    name = anchors[i][1] if i >= 0 else "other"
    tot[name] += e; hot += e
for name, e in tot.most_common():
    print(f"{100 * e / hot:5.1f} %  {name}")
