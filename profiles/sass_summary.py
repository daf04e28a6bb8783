#!/usr/bin/env python
"""Per-kernel SASS evidence: counts of the Blackwell tensor-core / TMEM / TMA mnemonics in the built library.

    python profiles/sass_summary.py [path/to/libdeco_b200.so] > profiles/sass_rN.txt

Runs `cuobjdump -sass` (no GPU needed) and, per kernel, counts
    UTCHMMA  tcgen05.mma (kind::f16)            LDTM / STTM   tcgen05.ld / tcgen05.st (tensor memory)
    UTMALDG  cp.async.bulk.tensor load (TMA)    UTMASTG       cp.async.bulk.tensor store (TMA)
    HMMA     legacy mma.sync                    UTCBAR        tcgen05.commit
so that "which kernels are genuine tcgen05/TMEM/TMA kernels" can be read off without disassembling the .so again.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "deco_b200", "_C", "libdeco_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "MUFU", "LDG", "STG"]


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def strip_params(name):
    """Demangled signature -> kernel name with its template arguments, without the parameter list."""
    depth = 0
    for i, ch in enumerate(name):
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return name[:i]
    return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op == k or (k in ("LDG", "STG") and op.startswith(k)) or (k == "HMMA" and op.startswith("HMMA")):
                    counts[cur][k] += 1
            counts[cur]["_all"] += 1
    names = demangle(order)
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; profiles/sass_summary.py)")
    print("# " + " ".join(f"{k:>8}" for k in KEYS) + "    instr  kernel")
    rows = []
    for fn in order:
        c = counts[fn]
        short = strip_params(names[fn]).replace("void ", "")
        rows.append((short, c))
    for short, c in sorted(rows, key=lambda r: r[0]):
        print("  " + " ".join(f"{c[k]:>8}" for k in KEYS) + f"  {c['_all']:>7}  {short}")
    tot = collections.Counter()
    for _, c in rows:
        tot.update(c)
    print("# total: " + ", ".join(f"{k} {tot[k]}" for k in KEYS))
    tc = sorted({re.sub(r"<.*", "", s) for s, c in rows if c["UTCHMMA"]})
    legacy = sorted({s for s, c in rows if c["HMMA"] and not c["UTCHMMA"]})
    print(f"# kernels with tcgen05.mma (UTCHMMA): {len(tc)}")
    print(f"# kernels on legacy mma.sync only (HMMA, no UTCHMMA): {len(legacy)}")
    for s in legacy:
        print(f"#     {s}")


if __name__ == "__main__":
    main()
