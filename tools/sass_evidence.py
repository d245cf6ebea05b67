#!/usr/bin/env python
"""(CPU) Count the Blackwell tensor-core / TMA / tensor-memory instructions in every kernel of libkdb200.so:

    python tools/sass_evidence.py > profiles/rN_sass_evidence.txt

UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,
UTCATOMSWS = tcgen05.alloc / dealloc, SYNCS = mbarrier operations, ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents,
MUFU.* = special-function unit (EX2 / TANH / RSQ ...), FFMA2 = packed fp32 FMA."""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "k-diffusion_b200" / "k_diffusion" / "_lib" / "libkdb200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "ACQBULK", "PREEXIT", "FFMA2", "HFMA2",
        "MUFU.EX2", "MUFU.TANH", "MUFU.RSQ", "MUFU.RCP", "REDUX", "LDGSTS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], check=True, capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, order, cur, it = collections.defaultdict(collections.Counter), [], None, iter(names)
    for line in sass.splitlines():
        if "Function :" in line:
            cur = next(it)
            cur = re.sub(r"\((?:anonymous namespace|int|bool|kdb::\w+)\)", "", cur)      # casts / namespaces inside template arguments
            cur = re.sub(r"\(.*", "", cur).replace("::::", "::")                        # drop the parameter list
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k.startswith("MUFU") and op.startswith(k)):
                    counts[cur][k] += 1
    print(f"SASS evidence: cuobjdump -sass {LIB.name} (sm_100a).  Instruction sites per kernel; kernels without tensor-core / TMA / tensor-memory instructions omitted.")
    print("\n".join(__doc__.strip().split("\n")[4:]))
    print()
    total = collections.Counter()
    for name in order:
        c = counts[name]
        total.update(c)
        if any(c[k] for k in ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "ACQBULK")):
            print(f"{name[:100]:<100} " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
    print()
    print("library-wide: " + "  ".join(f"{k}={total[k]}" for k in KEYS if total[k]) + f"  ({len(order)} kernels)")


if __name__ == "__main__":
    sys.exit(main())
