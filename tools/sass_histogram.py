#!/usr/bin/env python3
"""SASS opcode histogram of a kernel and of its loops, grouped by issue pipe (evidence for "what bounds this kernel").

  python tools/sass_histogram.py <object or .so> <kernel name substring> [--loops N] [--md out.md]

The kernel is disassembled with `cuobjdump -sass`; a loop is a backward branch (target address < branch address) and is
reported innermost-first as the address range [target, branch].  Pipes (Blackwell SM, per sub-partition and clock):
  ALU   integer/logic/compare/select/shift/min-max/convert-free ops: 16 lanes/clk  -> a warp instruction every 2 clocks
  FMA   FFMA/FMUL/FADD and IMAD (integer multiply-add shares the FMA pipe): 32 lanes/clk (heavy + lite halves)
  XU    MUFU (rcp, sqrt, ex2 ...), I2F/F2I, POPC, FLO, BREV: 4..16 lanes/clk
  LSU   shared / global / local loads and stores, atomics, LDC (constant loads issue through the same front end)
  CTRL  branches, convergence barriers, votes, shuffles (SHFL goes to the LSU crossbar; listed apart because it syncs)
  UNI   uniform-datapath instructions (U*), one per warp, not per lane
"""
import argparse
import collections
import re
import subprocess

PIPES = {
    "ALU": {"IADD3", "IADD", "LOP3", "LOP", "SHF", "SHL", "SHR", "ISETP", "FSETP", "SEL", "FSEL", "IMNMX", "VIMNMX", "VIMNMX3", "FMNMX", "VIADD",
            "PRMT", "LEA", "IABS", "BMSK", "SGXT", "PLOP3", "P2R", "R2P", "MOV", "CS2R", "ICMP", "FCHK", "VABSDIFF", "VABSDIFF4", "IDP",
            "HADD2", "HMUL2", "HFMA2", "FSET", "ISET", "CSEL", "DSETP"},
    "FMA": {"FFMA", "FMUL", "FADD", "IMAD", "FFMA32I", "FMUL32I", "FADD32I", "IMUL", "IMAD32I"},
    "XU": {"MUFU", "I2F", "F2I", "I2FP", "F2F", "POPC", "FLO", "BREV", "F2IP", "FRND", "I2I"},
    "LSU": {"LDS", "STS", "LDG", "STG", "LDL", "STL", "LD", "ST", "ATOMS", "ATOMG", "ATOM", "RED", "LDC", "LDSM", "LDGSTS", "LDGDEPBAR",
            "DEPBAR", "CCTL", "MEMBAR", "ERRBAR", "LDCU", "UBLKCP", "SYNCS", "FENCE", "UTMALDG", "UTMASTG", "UBLKRED"},
    "CTRL": {"BRA", "BRX", "JMP", "BSSY", "BSYNC", "BREAK", "WARPSYNC", "VOTE", "VOTEU", "SHFL", "BAR", "EXIT", "RET", "CALL", "NOP", "NANOSLEEP",
             "YIELD", "BMOV", "S2R", "REDUX", "MATCH", "ELECT", "ENDCOLLECTIVE", "ACQBULK", "S2UR", "KILL", "BPT", "RPCMOV"},
}


def pipe_of(op):
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED"):
        return "UNI"
    for p, ops in PIPES.items():
        if base in ops:
            return p
    return "OTHER"


def disassemble(obj, name):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    picked = [f for f in funcs[1:] if name in f.split("\n", 1)[0]]
    if not picked:
        raise SystemExit(f"no kernel matching {name!r}; have: " + ", ".join(f.split(chr(10), 1)[0][:60] for f in funcs[1:]))
    f = picked[0]
    insts = []
    for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", f):
        addr, op, rest = int(m.group(1), 16), m.group(2), m.group(3)
        tgt = None
        if op.split(".")[0] == "BRA":
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t:
                tgt = int(t.group(1), 16)
        insts.append((addr, op, tgt))
    return f.split("\n", 1)[0].strip(), insts


def hist(insts):
    by_pipe, by_op = collections.Counter(), collections.Counter()
    for _, op, _ in insts:
        by_pipe[pipe_of(op)] += 1
        by_op[op.split(".")[0]] += 1
    return by_pipe, by_op


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--loops", type=int, default=6)
    ap.add_argument("--md")
    a = ap.parse_args()
    fname, insts = disassemble(a.obj, a.kernel)
    out = [f"### `{fname}`", "", f"{len(insts)} SASS instructions in the kernel.", ""]
    bp, bo = hist(insts)
    out.append("| scope | instr | ALU | FMA | XU | LSU | CTRL | UNI | other | top opcodes |")
    out.append("|---|---:|---:|---:|---:|---:|---:|---:|---:|---|")

    def row(label, sub):
        p, o = hist(sub)
        top = ", ".join(f"{k} {v}" for k, v in o.most_common(9))
        return f"| {label} | {len(sub)} | {p['ALU']} | {p['FMA']} | {p['XU']} | {p['LSU']} | {p['CTRL']} | {p['UNI']} | {p['OTHER']} | {top} |"
    out.append(row("whole kernel", insts))
    loops = sorted({(t, addr) for addr, op, t in insts if t is not None and t < addr}, key=lambda r: r[1] - r[0], reverse=True)
    for t, addr in loops[:a.loops]:
        sub = [i for i in insts if t <= i[0] <= addr]
        p, _ = hist(sub)
        # issue-bound estimate per sub-partition: the ALU and the FMA pipe each take a warp instruction every 2 clocks, the
        # scheduler issues one instruction per clock
        n = len(sub)
        cyc = max(2 * p["ALU"], 2 * p["FMA"], n)
        out.append(row(f"loop 0x{t:04x}..0x{addr:04x} (issue floor {cyc} clk/iter = {n / cyc:.2f} IPC per sub-partition)", sub))
    text = "\n".join(out) + "\n"
    print(text)
    if a.md:
        with open(a.md, "a") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
