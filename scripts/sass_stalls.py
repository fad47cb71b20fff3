#!/usr/bin/env python
"""Static issue model of a SASS function (no GPU needed): per basic-block-ish region between labels, the number of
instructions, the sum of the stall fields (bits 105..108 of each 128-bit instruction = cycles before the same warp may
issue again: the single-warp lower bound), and the ALU / FMA / LSU pipe mix.

    cuobjdump -sass build/viterbi.o > v.sass
    python scripts/sass_stalls.py v.sass <function-name-substring> [--min 50]
"""
import re
import sys
from collections import Counter

ALU = ("IADD3", "LOP3", "SHF", "PRMT", "VIADDMNMX", "VIMNMX", "IMNMX", "ISETP", "SEL", "LEA", "IABS", "VIADD", "FMNMX",
       "FSEL", "FSETP", "POPC", "FLO", "BMSK", "SGXT", "IADD", "MOV", "R2P", "P2R", "PLOP3", "LOP", "SHL", "SHR", "I2I", "F2I", "I2F", "F2F")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2")
LSU = ("LDS", "STS", "LDG", "STG", "LD", "ST", "ATOMS", "ATOM", "RED", "LDL", "STL", "LDC", "ULDC")


def pipe(op):
    base = op.split(".")[0]
    if base in FMA:
        return "fma"
    if base in LSU:
        return "lsu"
    if base in ("I2F", "F2I", "F2F", "MUFU", "I2I", "POPC", "FLO"):
        return "xu"
    if base in ALU:
        return "alu"
    return "other"


def parse(path, fname):
    lines = open(path).read().splitlines()
    start = None
    for i, l in enumerate(lines):
        if "Function :" in l and fname in l:
            start = i
            break
    if start is None:
        raise SystemExit("function not found")
    out = []
    i = start + 1
    ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/")
    hi_re = re.compile(r"^\s+/\* 0x([0-9a-f]{16}) \*/")
    lab_re = re.compile(r"^\s+(\.L_x_\d+|\.L_\d+):")
    while i < len(lines) and "Function :" not in lines[i]:
        m = ins_re.match(lines[i])
        if m:
            addr = int(m.group(1), 16)
            text = m.group(2).strip()
            hi = 0
            if i + 1 < len(lines):
                h = hi_re.match(lines[i + 1])
                if h:
                    hi = int(h.group(1), 16)
            stall = (hi >> 41) & 0xF
            yld = (hi >> 45) & 1
            toks = text.split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            out.append((addr, op, stall, text, yld))
        else:
            lm = lab_re.match(lines[i])
            if lm:
                out.append((None, "LABEL", 0, lm.group(1), 0))
        i += 1
    return out


def main():
    path, fname = sys.argv[1], sys.argv[2]
    minlen = 50
    if "--min" in sys.argv:
        minlen = int(sys.argv[sys.argv.index("--min") + 1])
    ins = parse(path, fname)
    # regions: between labels / branches
    regions = []
    cur = []
    name = "entry"
    for a, op, st, text, y in ins:
        if op == "LABEL":
            if cur:
                regions.append((name, cur))
            cur = []
            name = text
            continue
        cur.append((a, op, st, text))
        if op.startswith("BRA") or op.startswith("EXIT") or op.startswith("RET") or op.startswith("CALL"):
            regions.append((name, cur))
            cur = []
            name = "after@%x" % a
    if cur:
        regions.append((name, cur))
    tot_i = tot_s = 0
    for name, r in regions:
        n = len(r)
        s = sum(x[2] for x in r)
        tot_i += n
        tot_s += s
        if n < minlen:
            continue
        pc = Counter(pipe(x[1]) for x in r)
        oc = Counter(x[1].split(".")[0] for x in r)
        print("%-14s @%05x  instr %5d  stall-sum %5d  alu %4d fma %4d lsu %3d xu %3d other %3d | %s" % (
            name, r[0][0], n, s, pc["alu"], pc["fma"], pc["lsu"], pc["xu"], pc["other"],
            " ".join("%s:%d" % kv for kv in oc.most_common(9))))
    print("total instr %d, stall-sum %d" % (tot_i, tot_s))


if __name__ == "__main__":
    main()
