#!/usr/bin/env python
"""Dev helper: per-kernel register/spill table from an nvcc -Xptxas -v log, and SASS opcode histograms."""
import collections, re, subprocess, sys

def regs(log):
    blocks = re.split(r"ptxas info    : Compiling entry function '", open(log).read())[1:]
    for b in blocks:
        name = b.split("'")[0]
        dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
        r = re.search(r"Used (\d+) registers", b)
        sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", b)
        sm = re.search(r"(\d+) bytes smem", b)
        print(f"{r.group(1):>4} regs stack {sp.group(1):>5} spill {sp.group(2):>5} smem {sm.group(1) if sm else 0:>6}  {dem.replace('mccbn::','')[:100]}")

def sass(so, pattern):
    txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s+Function : ", txt)[1:]:
        name = f.split("\n")[0]
        if pattern in name:
            ops = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)[\.\s;]", f)
            c = collections.Counter(ops)
            print(name, len(ops))
            print(c.most_common(60))
            open('/tmp/kernel.sass', 'w').write(f)

if __name__ == '__main__':
    if sys.argv[1] == 'regs':
        regs(sys.argv[2])
    else:
        sass(sys.argv[2], sys.argv[3])
