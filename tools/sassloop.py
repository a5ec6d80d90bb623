#!/usr/bin/env python
"""Dev helper (CPU only): static SASS opcode histogram of the sample loop of one kernel in a cuobjdump -sass listing.
usage: python tools/sassloop.py /tmp/spec.sass mccbn_fwd_spec [MUFU.LG2]
The sample loop = the smallest backward-branch span that contains the most `marker` instructions."""
import collections, re, sys

def functions(txt):
    out, cur, name = {}, None, None
    for l in txt.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            name = m.group(1)
            out[name] = []
            continue
        if name and re.search(r"/\*[0-9a-f]{4,6}\*/\s", l) and not l.strip().startswith("/* 0x"):
            out[name].append(l)
    return out

def analyse(lines, marker="MUFU.LG2"):
    ops = []
    for l in lines:
        m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", l)
        if m:
            ops.append((int(m.group(1), 16), m.group(2), l))
    loops = []
    for a, o, l in ops:
        if o.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)\s*;", l)
            if m:
                t = int(m.group(1), 16)
                if t < a:
                    n = sum(1 for x, oo, _ in ops if t <= x <= a and oo.startswith(marker))
                    loops.append((n, -(a - t), t, a))
    if not loops:
        return None
    nmax = max(x[0] for x in loops)
    n, _, t, a = max(x for x in loops if x[0] == nmax)
    def cls(o):
        parts = o.split(".")
        if parts[0] in ("IMAD", "MUFU", "LDS", "STS", "SHF", "LEA", "FFMA", "LOP3", "IADD3", "FMNMX", "FMNMX3", "PRMT"):
            return ".".join(parts[:2]) if len(parts) > 1 and parts[1] in ("WIDE", "HI", "LG2", "EX2", "128", "R", "L", "SAT", "MOV", "X", "U32") else parts[0]
        return parts[0]
    c = collections.Counter(cls(o) for x, o, _ in ops if t <= x <= a)
    return t, a, n, c, len(ops)

if __name__ == "__main__":
    txt = open(sys.argv[1]).read()
    fn = functions(txt)
    key = sys.argv[2] if len(sys.argv) > 2 else "mccbn_fwd_spec"
    marker = sys.argv[3] if len(sys.argv) > 3 else "MUFU.LG2"
    for name, lines in fn.items():
        if key in name:
            r = analyse(lines, marker)
            if r:
                t, a, n, c, tot = r
                print(f"{name}: {tot} static instr; sample loop {t:#x}..{a:#x}: {sum(c.values())} instr, {n} {marker}")
                print("  ", c.most_common(60))
