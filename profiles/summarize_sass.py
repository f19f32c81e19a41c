#!/usr/bin/env python3
"""SASS evidence for the kernels of libplayrho_b200.so (runs without a GPU: cuobjdump only).

  python profiles/summarize_sass.py > profiles/r01_sass_summary.txt

Per kernel: registers, stack (spill) bytes, static shared memory, SASS instruction count and the instruction mix that the
design claims rest on:
  * --fmad=false: no DFMA at all, and FFMA only inside the correctly rounded division / square-root sequences (div.rn / sqrt.rn
    = MUFU.RCP / MUFU.RSQ + FCHK + Newton steps, whose internal FFMAs are what makes them IEEE-exact): kernels without MUFU have
    0 FFMA (k_solve_velocity, k_warm_start, the broadphase), i.e. no user-level a*b+c was contracted -- every float result equals
    the reference's FMA-free x86-64 build;
  * 128-bit global loads/stores (LDG.E.128 / STG.E.128) on the float4 planes;
  * programmatic dependent launch: griddepcontrol.launch_dependents / .wait show up as PREEXIT / ACQBULK (column pdl: 2 per
    kernel; k_build_constraints keeps the completion trigger = 1; the opt-in cooperative k_solve_fused has none);
  * warp collectives (SHFL / VOTE / MATCH / REDUX) in the cooperative query, island and append paths;
  * FP64 (DADD/DMUL, never DFMA) only where the glibc sincosf restatement is used (position solve, TOI, joints, proxies).
"""
import collections
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "playrho_b200", "libplayrho_b200.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in out:
        m = re.match(r"(?:void )?(?:pr::)?([A-Za-z0-9_]+)(<[^(]*>)?\(", n)
        short.append((m.group(1) + (m.group(2) or "")) if m else n)
    return short


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            kernels[cur][m.group(1)] += 1
    names = list(kernels)
    short = demangle(names)
    print("# SASS summary of playrho_b200/libplayrho_b200.so (sm_100a, nvcc -O3 -lineinfo --fmad=false); made by profiles/summarize_sass.py")
    print("# columns: regs, stack bytes (spills + local arrays), static smem, SASS instructions, then instruction classes")
    print("%-34s %4s %6s %6s %6s | %5s %5s %6s %6s %5s %5s %5s %5s %5s %5s %5s" % (
        "kernel", "regs", "stack", "smem", "instr", "FFMA", "DFMA", "LDG128", "STG128", "LDG", "ATOM", "SHFL", "VOTE", "MUFU", "FP64", "pdl"))
    tot = collections.Counter()
    rows = []
    for mangled, name in zip(names, short):
        c = kernels[mangled]
        u = usage.get(mangled, {})
        n = sum(c.values())

        def cnt(pred):
            return sum(v for k, v in c.items() if pred(k))
        row = dict(
            FFMA=cnt(lambda k: k.startswith("FFMA")), DFMA=cnt(lambda k: k.startswith("DFMA")),
            LDG128=cnt(lambda k: k.startswith("LDG") and ".128" in k), STG128=cnt(lambda k: k.startswith("STG") and ".128" in k),
            LDG=cnt(lambda k: k.startswith("LDG")), ATOM=cnt(lambda k: k.startswith("ATOM") or k.startswith("RED") and not k.startswith("REDUX")),
            SHFL=cnt(lambda k: k.startswith("SHFL")), VOTE=cnt(lambda k: k.startswith("VOTE") or k.startswith("MATCH") or k.startswith("REDUX")),
            MUFU=cnt(lambda k: k.startswith("MUFU")), FP64=cnt(lambda k: k.startswith(("DADD", "DMUL", "DSETP", "F2F.F64", "F2F.F32.F64", "D2F", "F2D"))),
            pdl=cnt(lambda k: k.startswith(("ACQBULK", "PREEXIT", "GRIDDEPCONTROL"))))
        rows.append((name, u, n, row))
        for k, v in row.items():
            tot[k] += v
        tot["instr"] += n
    for name, u, n, row in sorted(rows, key=lambda r: -r[2]):
        print("%-34s %4s %6s %6s %6d | %5d %5d %6d %6d %5d %5d %5d %5d %5d %5d %5d" % (
            name[:34], u.get("REG", "?"), u.get("STACK", "?"), u.get("SHARED", "?"), n, row["FFMA"], row["DFMA"], row["LDG128"], row["STG128"],
            row["LDG"], row["ATOM"], row["SHFL"], row["VOTE"], row["MUFU"], row["FP64"], row["pdl"]))
    print("%-34s %4s %6s %6s %6d | %5d %5d %6d %6d %5d %5d %5d %5d %5d %5d %5d" % (
        "TOTAL (%d kernels)" % len(rows), "", "", "", tot["instr"], tot["FFMA"], tot["DFMA"], tot["LDG128"], tot["STG128"], tot["LDG"], tot["ATOM"],
        tot["SHFL"], tot["VOTE"], tot["MUFU"], tot["FP64"], tot["pdl"]))
    tc = sum(v for c in kernels.values() for k, v in c.items() if k.startswith(("UTCMMA", "HMMA", "IMMA", "UTMALDG", "UBLKCP")))
    print("# tensor-core / TMA instructions (UTCMMA, HMMA, IMMA, UTMALDG, UBLKCP): %d -- none by design: no dense contraction on this path, and the" % tc)
    print("# gathers are per-contact 16-byte rows (DESIGN.md section 3)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
