"""Instruction budget of a kernel's loops, from the SASS of the built library (no GPU needed):

    python tools/sass_budget.py 'psa_pack_fill_kernelILi8ELi19ELi1ELi1E' [--top 4 | --inner 2 --has VIADDMNMX] [--lib path/to/libpsa.so]

For every kernel whose mangled name matches the regex: registers are not shown (see *.ptxas.log); every backward
branch closes a loop, and for the largest loop bodies the tool prints the instruction count by issue pipe -- the
numbers DESIGN.md quotes per lane-step (ALU-pipe integer instructions such as VIADDMNMX / VIMNMX3 / PRMT / IADD3 /
LOP3 / ISETP, FMA-pipe IMAD, shuffles, shared / global memory, control).  A body count is the STATIC length of the
loop; divide by the cells one trip updates for instructions per cell."""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PIPES = [
    ("alu", r"^(VIADDMNMX|VIMNMX3?|VIADD|PRMT|IADD3?|LOP3|ISETP|SEL|SHF|LEA|IABS|IMNMX|PLOP3|BMSK|SGXT|FLO|POPC|VABSDIFF4?|IDP|I2I|VOTE|P2R|R2P|CS2R|MOV)\b"),
    ("fma (IMAD)", r"^(IMAD|FFMA|FMUL|FADD|HFMA2|HADD2|HMUL2)\b"),
    ("shuffle", r"^(SHFL|REDUX)\b"),
    ("shared mem", r"^(LDS|STS|LDSM|ATOMS)\b"),
    ("global / const mem", r"^(LDG|STG|LD|ST|LDC|LDCU|ULDC|ATOMG|ATOM|RED|LDL|STL)\b"),
    ("uniform datapath", r"^(U[A-Z0-9]+|R2UR|S2UR)\b"),
    ("control", r"^(BRA|BSSY|BSYNC|EXIT|WARPSYNC|BAR|NANOSLEEP|YIELD|CALL|RET|BREAK|NOP|DEPBAR|MEMBAR|ERRBAR|CCTL|FENCE|S2R|BPT|JMP|BRX)\b"),
]


def classify(op):
    for name, pat in PIPES:
        if re.match(pat, op):
            return name
    return "other"


def kernels(sass):
    name, body = None, []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            text = m.group(2).strip()
            text = re.sub(r"^@!?U?P\d+\s+", "", text)               # drop the predicate guard
            body.append((int(m.group(1), 16), text))
    if name:
        yield name, body


def report(name, body, top, inner=0, min_len=100, has=None):
    addr = [a for a, _ in body]
    index = {a: k for k, a in enumerate(addr)}
    loops = []
    for k, (a, text) in enumerate(body):
        m = re.match(r"BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?`?\(?0x([0-9a-f]+)\)?", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in index:
                loops.append((k - index[tgt] + 1, index[tgt], k))
    print(f"== {name}\n   {len(body)} instructions, {len(loops)} loops")
    chosen = sorted(loops, reverse=True)[:top]
    if inner:           # the smallest loops that are still whole steps (>= min_len instructions): the steady-state bodies
        def keeps(x):
            return x[0] >= min_len and (has is None or any(t.startswith(has) for _, t in body[x[1]:x[2] + 1]))
        chosen = sorted(x for x in loops if keeps(x))[:inner]
    for length, lo, hi in chosen:
        ops = collections.Counter()
        pipes = collections.Counter()
        for _, text in body[lo:hi + 1]:
            op = text.split()[0].split(".")[0] if text else "?"
            ops[text.split()[0]] += 1
            pipes[classify(op)] += 1
        print(f"   loop 0x{addr[lo]:04x}..0x{addr[hi]:04x}: {length} instructions")
        print("      by pipe: " + ", ".join(f"{k} {v}" for k, v in pipes.most_common()))
        print("      top opcodes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pattern")
    ap.add_argument("--top", type=int, default=3, help="report the N largest loop bodies")
    ap.add_argument("--inner", type=int, default=0, help="instead: the N smallest loop bodies of at least --min-len instructions")
    ap.add_argument("--min-len", type=int, default=100)
    ap.add_argument("--has", default=None, help="with --inner: only loops that contain this opcode (prefix), e.g. VIADDMNMX")
    ap.add_argument("--lib", default=os.path.join(ROOT, "cse305_parallel_sequence_alignment_b200", "libpsa.so"))
    args = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
    hit = False
    for name, body in kernels(sass):
        if re.search(args.pattern, name):
            report(name, body, args.top, args.inner, args.min_len, args.has)
            hit = True
    if not hit:
        sys.exit(f"no kernel matches {args.pattern!r}")


if __name__ == "__main__":
    main()
