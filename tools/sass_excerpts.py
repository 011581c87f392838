#!/usr/bin/env python
"""tools/sass_excerpts.py -- per-kernel SASS evidence from the built libb200slam.so (cuobjdump, no GPU needed):
opcode histogram, the mnemonics DESIGN.md's claims rest on with a few sample lines each, and the check that the
bit-exact kernels contain no fused multiply-add.  Writes profiles/r02_sass.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hardware-acceleration-of-lidar-slam_b200", "libb200slam.so")
WANT = {   # kernel name fragment -> (label, mnemonics to show)
    "edt_tma_kernelILi9ELi3ELi3ELb0": ("edt_tma_kernel<R=9, NW=3, NST=3> (max_dist 10)",
                                       ["UTMALDG", "SYNCS", "VIADDMNMX", "VIMNMX3", "VOTE", "BREV", "FLO", "ACQBULK", "BRA.U"]),
    "lattice_kernelILi16ELi2ELi4ELi1ELb0ELi2E": ("lattice_kernel<16, 2, 4, 1, false, Q=2> (config 3, row reuse)",
                                                 ["LDG", "FADD", "FSEL", "IMAD.WIDE", "VIADDMNMX", "ACQBULK", "FFMA", "FADD2"]),
    "lattice_kernelILi4ELi1ELi8ELi1ELb0ELi0E": ("lattice_kernel<4, 1, 8, 1, false, 0> (config 1, throughput policy)",
                                                ["LDG", "FADD", "VIADDMNMX", "FFMA"]),
    "fastmatch_kernel": ("fastmatch_kernel (FastMatch-sized lattices, one CTA)", ["LDG", "FADD", "F2I", "SHFL", "VOTE", "FFMA", "ACQBULK"]),
    "poses_kernel": ("poses_kernel (pose / particle lists)", ["FMUL2", "FADD2", "FADD", "LDG", "F2I", "FFMA"]),
    "csv_parse_kernel": ("csv_parse_kernel", ["DMUL", "DFMA", "MUFU.RCP64H", "F2F", "ATOMG"]),
    "resample_push_kernel": ("resample_push_kernel (sharded particle filter: offspring to the peers + barrier)", ["STG", "MEMBAR", "ATOMG", "CCTL"]),
}


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            ins = re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip().rstrip(";").strip()
            funcs[cur].append(ins)
    lines = ["SASS evidence extracted from libb200slam.so by tools/sass_excerpts.py (cuobjdump -sass; sm_100a)", ""]
    for frag, (label, mnems) in WANT.items():
        names = [n for n in funcs if frag in n]
        if not names:
            lines.append(f"## {label}: NOT FOUND ({frag})")
            continue
        name = sorted(names, key=len)[0]
        body = funcs[name]
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for i in body if i)
        lines.append(f"## {label}")
        lines.append(f"   symbol {name}")
        lines.append(f"   {len(body)} instructions; top opcodes: " + ", ".join(f"{k} x{v}" for k, v in ops.most_common(14)))
        for mn in mnems:
            hits = [i for i in body if re.search(r"(^|\s)" + re.escape(mn) + r"(\.|\s|$)", i)]
            lines.append(f"   {mn:12s} x{len(hits):4d}" + ("".join(f"\n        {h}" for h in hits[:3]) if hits else "   (absent)"))
        lines.append("")
    # no FFMA / FFMA2 in any bit-exact scoring kernel
    for frag in ("lattice_kernel", "poses_kernel", "fastmatch_kernel", "scan_transform_kernel", "scan_read_kernel"):
        n = sum(1 for name, body in funcs.items() if frag in name for i in body if re.search(r"\bFFMA2?\b", i))
        lines.append(f"FFMA / FFMA2 instructions in *{frag}*: {n}")
    # nothing of tensor cores (nothing here is a dense contraction)
    n = sum(1 for body in funcs.values() for i in body if re.search(r"\b(HMMA|UTCHMMA|UTCMMA|IMMA|QMMA)\b", i))
    lines.append(f"tensor-core instructions in the whole library: {n}")
    path = os.path.join(ROOT, "profiles", "r02_sass.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
