#!/usr/bin/env python
"""Compares the SASS of two builds of one object file kernel by kernel (cuobjdump -sass): IDENTICAL / same modulo
parameter offsets / DIFFERENT. Used to show that a source change did not touch the generated code of kernels whose
measurements are on record (e.g. adding the CONV instantiation next to the Linear GEMM kernels).

    python tools/sass_diff.py old.o new.o [--map 's/EEEv14/ELb0EEEv14/']     (optional sed-like rename old -> new)
"""
import re
import subprocess
import sys


def parse(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(.*?)\s*/\*", line)
        if cur and m:
            funcs[cur].append(m.group(1))
    return funcs


def main():
    old, new = parse(sys.argv[1]), parse(sys.argv[2])
    rename = None
    if "--map" in sys.argv:
        _, a, b, _ = sys.argv[sys.argv.index("--map") + 1].split("/")
        rename = (a, b)
    norm = lambda L: [re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[0][X]", x) for x in L]
    worst = 0
    for f, code in old.items():
        g = f if f in new else (f.replace(*rename) if rename else f)
        if g not in new:
            print("missing   ", f)
            worst = max(worst, 1)
            continue
        if code == new[g]:
            verdict = "IDENTICAL "
        elif norm(code) == norm(new[g]):
            verdict = "same-modulo-param-offsets"
        else:
            verdict, worst = "DIFFERENT ", 2
        print(verdict, len(code), len(new[g]), f)
    for g in new:
        if g not in old and not (rename and g.replace(rename[1], rename[0]) in old):
            print("new       ", len(new[g]), g)
    sys.exit(0 if worst < 2 else 1)


if __name__ == "__main__":
    main()
