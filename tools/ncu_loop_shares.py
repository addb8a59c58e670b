#!/usr/bin/env python
"""Per-loop share of executed thread instructions and stall samples from an ncu source page
(`ncu -i X.ncu-rep --page source --csv > src.csv`; kernels compiled with -lineinfo).

    python tools/ncu_loop_shares.py src.csv "k_myers_small<(int)6, (bool)0>"
"""
import csv
import re
import sys


def sections(path):
    cur, out = None, {}
    for r in csv.reader(open(path)):
        if r and r[0] == "Kernel Name":
            cur = r[1]
            out.setdefault(cur, [])
            out[cur].append([])
        elif cur is not None:
            out[cur][-1].append(r)
    return out


def main():
    secs = sections(sys.argv[1])
    for name, parts in secs.items():
        if len(sys.argv) > 2 and sys.argv[2] not in name:
            continue
        rows = parts[0]  # first part = SASS view
        hdr = rows[0]
        idx = {h: i for i, h in enumerate(hdr)}
        ins = []
        for r in rows[1:]:
            try:
                ins.append((int(r[idx["Address"]], 16), r[idx["Source"]].strip(), float(r[idx["Thread Instructions Executed"]] or 0),
                            float(r[idx["Warp Stall Sampling (All Samples)"]] or 0)))
            except Exception:
                pass
        if not ins:
            continue
        base = ins[0][0]
        ins = [(a - base, s, t, w) for a, s, t, w in ins]
        tot = sum(t for _, _, t, _ in ins) or 1
        totw = sum(w for *_, w in ins) or 1
        loops = set()
        for a, s, t, w in ins:
            if "BRA" in s:
                m = re.search(r"0x([0-9a-f]+)", s)
                if m and int(m.group(1), 16) - base < a and int(m.group(1), 16) >= base:
                    loops.add((int(m.group(1), 16) - base, a))
        loops = sorted(loops, key=lambda x: x[1] - x[0])
        acc = {l: [0.0, 0.0] for l in loops}
        other = [0.0, 0.0]
        for a, s, t, w in ins:
            for l in loops:
                if l[0] <= a <= l[1]:
                    acc[l][0] += t; acc[l][1] += w
                    break
            else:
                other[0] += t; other[1] += w
        print(f"{name}: {tot / 1e9:.3f} G thread instructions")
        for l in sorted(acc, key=lambda x: x[0]):
            print(f"  loop {l[0]:#07x}-{l[1]:#07x} ({(l[1] - l[0]) // 16 + 1:4d} instr): {100 * acc[l][0] / tot:5.1f}% of thread instr, {100 * acc[l][1] / totw:5.1f}% of stall samples")
        print(f"  outside loops: {100 * other[0] / tot:5.1f}% of thread instr, {100 * other[1] / totw:5.1f}% of stall samples")


if __name__ == "__main__":
    main()
