"""
SASS instruction-class summary per kernel of libsk_b200.so (runs without a GPU):

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt

For every kernel in the sm_100a cubin: number of SASS instructions and the count per class (FP64 arithmetic, FP32/int
ALU, shared / global / local memory, 256-bit global accesses, atomics, barriers, shuffles / votes / redux, conversions,
branches), plus the Blackwell/Hopper-only mnemonics the profiling recipe asks about (UTMALDG/UTMASTG/UBLKCP: TMA,
LDGSTS: cp.async, SYNCS: mbarrier, UTC*/tcgen05).  The K(r) path is FP64 spread / FFT / interpolate: no tensor-core
instruction is expected; what the summary shows instead is where each kernel's instructions go.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spectralkernels.jl_b200", "libsk_b200.so")

CLASSES = [
    ("fp64", re.compile(r"^(DFMA|DADD|DMUL|DSETP|DMNMX|DSEL)")),
    ("mufu/f2f", re.compile(r"^(MUFU|F2F|F2I|I2F|FRND|I2I)")),
    ("shared ld/st", re.compile(r"^(LDS|STS|LDSM)")),
    ("shared atom", re.compile(r"^ATOMS")),
    ("global ld/st", re.compile(r"^(LDG(?!STS)|STG|LD\b|ST\b)")),
    ("global atom/red", re.compile(r"^(ATOMG|ATOM\b|RED)")),
    ("local ld/st", re.compile(r"^(LDL|STL)")),
    ("const ld", re.compile(r"^(LDC|ULDC)")),
    ("barrier", re.compile(r"^(BAR|WARPSYNC|NANOSLEEP|MEMBAR|ERRBAR|CCTL)")),
    ("shfl/vote/redux", re.compile(r"^(SHFL|VOTE|REDUX|MATCH|VOTEU)")),
    ("branch", re.compile(r"^(BRA|BRX|JMP|CALL|RET|EXIT|BSSY|BSYNC|BREAK|YIELD)")),
    ("tma/cp.async/mbarrier/tcgen05", re.compile(r"^(UTMALDG|UTMASTG|UBLKCP|LDGSTS|SYNCS|UTC|UTMA|UBLK)")),
]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    wide = collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            kernels[cur]["total"] += 1
            hit = False
            for name, rx in CLASSES:
                if rx.match(op):
                    kernels[cur][name] += 1
                    hit = True
                    break
            if not hit:
                kernels[cur]["int/fp32/other alu"] += 1
            if re.match(r"^(LDG|STG).*\.256", op) or ".ENL2.256" in op:
                wide[cur] += 1
    names = [c[0] for c in CLASSES] + ["int/fp32/other alu"]
    print(f"libsk_b200.so, sm_100a SASS: {len(kernels)} kernels, {sum(k['total'] for k in kernels.values())} instructions")
    print("(k_interp_session, k_hankel_interp, k_hankel_interp2 and the width variants 4..14 are the A/B references "
          "selected by sk_ctx_set_interp_mode / sk_ctx_set_nufft_eps)\n")
    for k, c in kernels.items():
        if k.startswith("void cub::") or "cub::" in k:
            k = k[:90]
        parts = ", ".join(f"{n} {c[n]}" for n in names if c[n])
        extra = f", 256-bit global accesses {wide[k]}" if wide[k] else ""
        print(f"{k}\n    {c['total']} instructions: {parts}{extra}")


if __name__ == "__main__":
    main()
