"""Opcode / pipe histogram of the loops of one kernel from `cuobjdump -sass` output.

Usage: cuobjdump -sass lib.so | python tools/sass_mix.py <substring of the mangled kernel name> [min_loop_len]
Finds every backward branch, treats [target, branch] as a loop body, and prints the opcode mix of the longest bodies
(static counts: predicated-off paths are included).
"""
import collections
import re
import sys

ALU = ("LOP3", "IADD3", "IADD", "SHF", "PRMT", "SEL", "ISETP", "LEA", "IMNMX", "VIMNMX", "FMNMX", "POPC", "FLO", "BREV", "SGXT", "BMSK", "PLOP3", "LOP", "IABS", "VOTE", "MOV", "CS2R", "VIADD", "FSETP", "FSEL", "P2R", "R2P")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2")
LSU = ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDC", "ATOM", "RED", "LDSM", "LDL", "STL", "ATOMS", "ATOMG")


def pipe(op):
    base = op.split(".")[0]
    if base.startswith("U"):
        return "uniform"
    if base in FMA:
        return "fma"
    if base in ALU:
        return "alu"
    if base in LSU:
        return "lsu"
    if base in ("SHFL", "BAR", "S2R", "MUFU", "I2F", "F2I", "I2FP", "F2F", "DADD", "DMUL", "DFMA", "MEMBAR", "ERRBAR", "FENCE", "SYNCS", "UBLKCP", "S2UR", "R2UR", "CCTL", "NANOSLEEP", "DSETP"):
        return "other(" + base + ")"
    if base in ("BRA", "EXIT", "BSSY", "BSYNC", "CALL", "RET", "WARPSYNC", "NOP", "BREAK", "YIELD", "BPT"):
        return "ctrl"
    return "?" + base


def main():
    name = sys.argv[1]
    min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    ins = []  # (addr, opcode)
    active = False
    pat = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)(.*?);")
    for line in sys.stdin:
        if "Function :" in line:
            active = name in line
            if active:
                print("kernel:", line.strip())
            continue
        if not active:
            continue
        m = pat.search(line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    print("instructions:", len(ins))
    addr_index = {a: i for i, (a, _, _) in enumerate(ins)}
    loops = []
    for i, (a, op, rest) in enumerate(ins):
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t:
                ta = int(t.group(1), 16)
                if ta <= a and ta in addr_index:
                    loops.append((addr_index[ta], i))
    loops = [l for l in loops if l[1] - l[0] >= min_len]
    for lo, hi in loops:
        body = ins[lo : hi + 1]
        ops = collections.Counter(op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "IMAD")) and "." in op else "") for _, op, _ in body)
        pipes = collections.Counter(pipe(op) for _, op, _ in body)
        print(f"\nloop 0x{ins[lo][0]:x}..0x{ins[hi][0]:x}: {len(body)} instr")
        print("  pipes:", dict(pipes.most_common()))
        print("  ops:  ", dict(ops.most_common()))


if __name__ == "__main__":
    main()
