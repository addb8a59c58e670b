#!/usr/bin/env python
"""Lists the loops of a kernel in `cuobjdump -sass` output with their instruction mix, and
optionally prints one loop body.  Used to produce the SASS evidence under profiles/.

    cuobjdump -sass lordfast_b200/liblfgpu.so > /tmp/all.sass
    python tools/sass_loops.py /tmp/all.sass k_myers_smallILi4ELb0 [--body N]
"""
import collections
import re
import sys


def parse(path):
    cur, fn = None, {}
    for line in open(path):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); fn[cur] = []; continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if cur and m:
            fn[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return fn


def loops(ins):
    out = []
    for addr, txt in ins:
        if "BRA" in txt:
            t = re.search(r"0x([0-9a-f]+)", txt)
            if t and int(t.group(1), 16) < addr:
                lo = int(t.group(1), 16)
                out.append((lo, addr, [(a, x) for a, x in ins if lo <= a <= addr]))
    return out


def mix(body):
    return collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in body)


if __name__ == "__main__":
    fn = parse(sys.argv[1])
    pat = sys.argv[2]
    body_idx = int(sys.argv[sys.argv.index("--body") + 1]) if "--body" in sys.argv else None
    for name in [k for k in fn if pat in k]:
        print(f"{name}: {len(fn[name])} instructions")
        for i, (lo, hi, body) in enumerate(loops(fn[name])):
            print(f"  loop {i}: {lo:#x}-{hi:#x} {len(body)} instr  {dict(mix(body).most_common(12))}")
            if body_idx == i:
                for a, x in body:
                    print(f"      {a:#06x}  {x}")
