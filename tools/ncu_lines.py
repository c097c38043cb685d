"""Map ncu SASS-level stall samples back to CUDA source lines (profiling helper).
usage: tools_ncu_lines.py <report.ncu-rep> <lib.so> [kernel-substring]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import csv, re, subprocess, sys, os, tempfile, collections
rep, lib = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "sb_run_kernel"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = []
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):      # one per translation unit: take the one with the kernel
    d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    if any(l.startswith("\t.section\t.text.") and kern in l for l in d):
        dis = d
        break
# instructions of the kernel with their line info (innermost line + inline chain)
ins = []; on = False; cur = None
for l in dis:
    if l.startswith("\t.section\t.text."):
        on = kern in l; continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        chain = [(os.path.basename(m.group(1)), int(m.group(2)))]
        for f, n in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3)):
            chain.append((os.path.basename(f), int(n)))
        cur = chain; continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        ins.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(src))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; body = rows[hi + 1:]
si = hdr.index("# Samples"); ei = hdr.index("Instructions Executed")
print("sass rows", len(body), "disasm instructions", len(ins))
inner = collections.Counter(); outer = collections.Counter(); execd = collections.Counter()
tot = 0
lines = {}
for f in ("sb_core.cuh", "sb_platform.cuh", "sb_counts.cuh", "sb_eval.cuh"):
    p = os.path.join(os.path.dirname(os.path.abspath(lib)), f)
    if os.path.exists(p): lines[f] = open(p).read().splitlines()
for r, ch in zip(body, ins):
    s = int(r[si] or 0); tot += s
    if not ch: continue
    inner[ch[0]] += s
    core = [c for c in ch if c[0] == "sb_core.cuh"]
    if core: outer[core[-1]] += s; 
    for c in core: execd[c] += 0
def show(cnt, title, n=40):
    print("==", title)
    for (f, ln), s in cnt.most_common(n):
        txt = lines.get(f, [""] * (ln + 1))[ln - 1].strip()[:100] if f in lines and ln - 1 < len(lines[f]) else ""
        print(f"{s:7d} {100*s/max(tot,1):5.1f}%  {f}:{ln:<5d} {txt}")
show(inner, "innermost line")
show(outer, "outermost sb_core.cuh line (call site in the step function)")
