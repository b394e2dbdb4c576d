#!/usr/bin/env python
"""Count the instructions a row costs on the hot path of a fused-engine kernel, offline (no GPU).

    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DE3_DEV_ONLY=256 \\
         -c simpleds_b200/csrc/sds_engine_f32.cu -o /tmp/e.o          # one instantiation: seconds
    cuobjdump -sass /tmp/e.o -fun '_ZN3sds17engine_f32_kernelILi256ELi7ELb0EEEvNS_12EngineParamsE' > /tmp/e.sass
    python scripts/sass_hot_path.py /tmp/e.sass                      # lists the branches of the step loop
    LOOP=2340:7010 python scripts/sass_hot_path.py /tmp/e.sass 26d0 2df0 5360   # ... with these forward branches TAKEN

Walks the SASS from the head of the step loop (the largest backward branch, or LOOP=head:tail in hex)
to its back edge, falling through conditional branches unless their address is listed, and prints the
number of instructions on that path and their opcode histogram.  The counts quoted in DESIGN.md
(675 -> 571 -> 586 instructions per 256-sample row) come from this walk; ncu's
smsp__inst_executed / rows agrees with it to a few per cent (the close of a time and the item prologue
are not on the walked path)."""
import collections
import os
import re
import sys


def main():
    lines = open(sys.argv[1]).read().split("\n")
    ins = []
    for ln in lines:
        m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tg = int(m.group(1), 16)
            if tg < a and (best is None or a - tg > best[1] - best[0]) and (a - tg) < 0x7000:
                best = (tg, a)
    head, tail = best
    if os.environ.get("LOOP"):
        head, tail = [int(x, 16) for x in os.environ["LOOP"].split(":")]
    taken = set(int(x, 16) for x in sys.argv[2:])

    def op(t):
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        return t.split()[0].split(".")[0]

    hist, n, i, brs = collections.Counter(), 0, addr[head], []
    while True:
        a, t = ins[i]
        hist[op(t)] += 1
        n += 1
        if a == tail:
            break
        m = re.search(r"BRA(?:\.U|\.DIV)?\s+(?:!?U?P\d+,\s*)*(?:UR\d+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tg = int(m.group(1), 16)
            uncond = not t.startswith("@") and not re.search(r"BRA(\.U)?\s+!?U?P\d", t) and "BRA.DIV" not in t
            brs.append((a, t[:40], tg, (tg - a) // 16 - 1, uncond))
            if (uncond or a in taken) and tg > a:
                i = addr[tg]
                continue
        i += 1
    print("loop %x..%x  path instructions: %d" % (head, tail, n))
    for b in brs:
        print("  %x %-40s -> %x skip %d %s" % (b[0], b[1], b[2], b[3], "UNCOND" if b[4] else ("TAKEN" if b[0] in taken else "")))
    print(" ".join(f"{k}:{v}" for k, v in hist.most_common(40)))


if __name__ == "__main__":
    main()
