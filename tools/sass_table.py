"""SASS opcode table of the built libsrk.so (cuobjdump -sass): per kernel the instruction count and the Blackwell-specific opcodes.
Usage: python tools/sass_table.py > profiles/<round>_sass_opcode_table.txt   (runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tpu_superresolution_b200", "lib", "libsrk.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UBLKRED", "UTMALDG", "UTMASTG", "UTMAREDG", "SYNCS", "HMMA", "STL", "LDL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"^void ", "", cur).replace("srk::", "")
            cur = re.sub(r"\(.*$", "", cur)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur]["instr"] += 1
            kernels[cur][m.group(1)] += 1
    print("# SASS opcode table of tpu_superresolution_b200/lib/libsrk.so (cuobjdump -sass, sm_100a; tools/sass_table.py): per kernel, the")
    print("# instruction count and the Blackwell-specific opcodes (UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UBLKCP =")
    print("# cp.async.bulk g->s / s->g, UBLKRED = cp.reduce.async.bulk, UTMALDG / UTMASTG / UTMAREDG = tensor-map TMA load / store / reduce,")
    print("# SYNCS = mbarrier ops); HMMA (legacy mma.sync) must be 0; STL / LDL = local-memory (spill) accesses.\n")
    print(f"{'kernel':66s}" + "".join(f"{o:>9s}" for o in ["instr"] + OPS))
    for k, c in kernels.items():
        print(f"{k[:64]:66s}" + "".join(f"{c[o]:9d}" for o in ["instr"] + OPS))


if __name__ == "__main__":
    sys.exit(main())
