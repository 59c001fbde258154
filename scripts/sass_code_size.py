"""Static code size of the step kernel's substep loop, attributed to the calling line (not the inlined leaf).
usage: python scripts/sass_code_size.py [kernel-name-substring] ; needs the in-tree build (build/env_kernels.o)."""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "high_speed_quadrupedal_locomotion_by_irrl_b200", "build", "env_kernels.o")
name = sys.argv[1] if len(sys.argv) > 1 else "env_step_kernelILi128ELb1ELb0"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", OBJ], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if l.startswith(".text.") and name in l)
end = next((i for i in range(start + 1, len(txt)) if txt[i].startswith("//-----")), len(txt))
body = txt[start:end]
ins = []   # (addr, opcode, [frames innermost..outermost])
frames, pending = [], []
for l in body:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        pending.append(m); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?(\S+)', l)
    if m:
        if pending:     # a chain: "A inlined at B", "B inlined at C", ..., "C" (the last line repeats the outermost frame)
            frames = [(os.path.basename(pending[0].group(1)), int(pending[0].group(2)))]
            for q in pending:
                if q.group(3): frames.append((os.path.basename(q.group(3)), int(q.group(4))))
            pending = []
        ins.append((int(m.group(1), 16), m.group(2).split(".")[0], list(frames)))
# loop = largest backward branch span
best = (0, 0, 0); loops = []
# nvdisasm -c prints labels; find loop by label positions
labels = {}
addr = 0
for l in body:
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/', l)
    if m: addr = int(m.group(1), 16)
    m2 = re.match(r'(\.L_x_\d+):', l)
    if m2: labels[m2.group(1)] = None
cur = None
for l in body:
    m2 = re.match(r'(\.L_x_\d+):', l)
    if m2: cur = m2.group(1); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/', l)
    if m and cur: labels[cur] = int(m.group(1), 16); cur = None
for l in body:
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+.*BRA.*`\((\.L_x_\d+)\)', l)
    if m and labels.get(m.group(2)) is not None:
        a, t = int(m.group(1), 16), labels[m.group(2)]
        if t < a and a - t > 16384: loops.append((a - t, t, a))
loops.sort(key=lambda x: x[1])
print("loops > 16 KB:", ", ".join("0x%x..0x%x (%d B)" % (t, a, sp) for sp, t, a in loops))
span, lo, hi = loops[int(os.environ.get("LOOP", "0"))]      # LOOP=1: the second (complete) copy of the substep loop
print("kernel %s: %d instructions; substep loop 0x%x..0x%x = %d B (%d instructions)" % (name, len(ins), lo, hi, span, span // 16))
_src = open(os.path.join(ROOT, "high_speed_quadrupedal_locomotion_by_irrl_b200", "csrc", "env_device.cuh")).read().split("\n")
IS_LO = next(i for i, l in enumerate(_src) if "void integrate_substep(" in l or "bool integrate_substep(" in l) + 1
IS_HI = next(i for i in range(IS_LO, len(_src)) if _src[i].startswith("}")) + 1
def phase(fr):
    # outermost frame inside env_device.cuh functions of interest, else the kernel line
    for f, ln in reversed(fr):
        if f == "env_device.cuh" and IS_LO <= ln <= IS_HI: return "integrate_substep:%d" % ln
    for f, ln in reversed(fr):
        if f == "env_kernels.cu": return "kernel:%d" % ln
    return ("%s:%d" % fr[-1]) if fr else "?:0"
def mid(fr):
    # second-level attribution: the frame just inside integrate_substep (dynamics / contact_setup / leg_fk line)
    idx = None
    for i, (f, ln) in enumerate(fr):
        if f == "env_device.cuh" and IS_LO <= ln <= IS_HI: idx = i
    if idx is None or idx == 0: return None
    f, ln = fr[idx - 1]
    return "%s:%d" % (f, ln)
by_phase = collections.Counter(); by_mid = collections.Counter()
src = open(os.path.join(ROOT, "high_speed_quadrupedal_locomotion_by_irrl_b200", "csrc", "env_device.cuh")).read().split("\n")
for a, op, fr in ins:
    if lo <= a <= hi:
        by_phase[phase(fr)] += 1
        m_ = mid(fr)
        if m_: by_mid[(phase(fr), m_)] += 1
print("--- by line of integrate_substep")
for k, v in sorted(by_phase.items(), key=lambda kv: -kv[1])[:45]:
    s = src[int(k.split(":")[1]) - 1].strip()[:90] if k.startswith("integrate") else ""
    print("%5d  %-24s %s" % (v, k, s))
print("--- inside dynamics()/contact_setup()/leg_fk (second level)")
for (p, m_), v in sorted(by_mid.items(), key=lambda kv: -kv[1])[:60]:
    ln = int(m_.split(":")[1]); s = src[ln - 1].strip()[:90] if m_.startswith("env_device") else ""
    print("%5d  %-22s %-20s %s" % (v, p, m_, s))
