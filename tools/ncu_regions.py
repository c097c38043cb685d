"""Executed warp instructions and stall samples of an ncu capture, aggregated by device function
(the function whose body contains the outermost sb_core.cuh line of each SASS instruction).
usage: tools_ncu_regions.py <report.ncu-rep> <lib.so> [kernel-substring]"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))   # repo root (bench.py, sbayes_b200/, oracle/)
import csv, re, subprocess, sys, os, tempfile, collections
rep, lib = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "sb_run_kernel"
src_path = os.path.join(os.path.dirname(os.path.abspath(lib)), "sb_core.cuh")
src = open(src_path).read().splitlines()
# function start lines
starts = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template <[^>]*>\s*)?SB_(?:DEV|KERNEL_BIG|KERNEL|HD)\s+[\w:<> \*&]*?\b(\w+)\(", l)
    if m: starts.append((i, m.group(1)))
def func_of(line):
    name = "?"
    for s, n in starts:
        if s <= line: name = n
        else: break
    return name
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
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
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; body = rows[hi + 1:]
si = hdr.index("# Samples"); ei = hdr.index("Instructions Executed")
ex = collections.Counter(); sm = collections.Counter(); exl = collections.Counter()
for r, ch in zip(body, ins):
    if not ch: continue
    core = [c for c in ch if c[0] == "sb_core.cuh"]
    if not core: continue
    # innermost sb_core function = first core entry
    fn = func_of(core[0][1])
    e = int(r[ei] or 0); s = int(r[si] or 0)
    ex[fn] += e; sm[fn] += s
tot_e = sum(ex.values()); tot_s = sum(sm.values())
print(f"total warp instructions {tot_e}, samples {tot_s}")
for fn, e in ex.most_common(40):
    print(f"{fn:28s} instr {e:12d} {100*e/tot_e:5.1f}%   samples {sm[fn]:8d} {100*sm[fn]/max(tot_s,1):5.1f}%")
