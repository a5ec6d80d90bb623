#!/usr/bin/env python
"""Dev helper (CPU only): generate the poset-specialised forward kernel of a bench config, compile it with nvcc for
sm_100a and print registers + the static SASS opcode histogram of the sample loop.
usage: python tools_jitsass.py cfg3"""
import collections, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mccbn_b200 as mc
from bench import make_problem

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
cfg, p, N, L, sampling, desc = make_problem(name)
r = mc.jit_compile_forward(cfg["poset"], False)
print([l for l in r["log"].splitlines() if "registers" in l or "spill" in l])
open("/tmp/spec.cu", "w").write(r["source"])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17",
                       "-lineinfo", "-I", os.path.join(root, "mccbn_b200/csrc"), "-o", "/tmp/spec.cubin", "/tmp/spec.cu"])
txt = subprocess.run(["cuobjdump", "-sass", "/tmp/spec.cubin"], capture_output=True, text=True).stdout
open("/tmp/spec.sass", "w").write(txt)
lines = [l for l in txt.splitlines() if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
ops = []
for l in lines:
    m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", l)
    if m:
        ops.append((int(m.group(1), 16), m.group(2)))
# the sample loop = the longest backward-branch span
best = (0, 0)
for l in lines:
    m = re.search(r"/\*([0-9a-f]{4,6})\*/.*BRA\s+(?:U?R?\w+,\s*)?0x([0-9a-f]+)", l)
    if m:
        a, t = int(m.group(1), 16), int(m.group(2), 16)
        if t < a and a - t > best[0] - best[1] and (a - t) < 0x4000:
            pass
        if t < a and (a - t) > (best[0] - best[1]):
            best = (a, t)
print("total static instructions", len(ops), "; largest loop %#x..%#x" % (best[1], best[0]))
def hist(lo, hi):
    c = collections.Counter(o.split(".")[0] + ("." + o.split(".")[1] if o.startswith(("IMAD", "MUFU", "LDS", "STS", "SHF")) and "." in o else "") for a, o in ops if lo <= a <= hi)
    return c
# inner sample loop: the smallest loop containing MUFU.LG2
loops = []
for l in lines:
    m = re.search(r"/\*([0-9a-f]{4,6})\*/.*BRA\s+(?:U?R?\w+,\s*)?0x([0-9a-f]+)", l)
    if m:
        a, t = int(m.group(1), 16), int(m.group(2), 16)
        if t < a:
            n_mufu = sum(1 for x, o in ops if t <= x <= a and o.startswith("MUFU.LG2"))
            if n_mufu:
                loops.append((a - t, t, a, n_mufu))
loops.sort()
if loops:
    _, t, a, n = loops[0]
    c = hist(t, a)
    print("sample loop %#x..%#x: %d instructions, %d MUFU.LG2" % (t, a, sum(c.values()), n))
    print(c.most_common(40))
