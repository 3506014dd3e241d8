"""Compares the SASS of every kernel in two builds (`cuobjdump -sass lib.so`, encodings stripped): prints the kernels
whose instructions differ.  Used to prove that adding gated experimental code leaves the validated kernels
byte-identical:  python tools/sass_same.py old_sass.txt new_sass.txt"""
import re
import sys


def split(path):
    d, cur = {}, None
    for ln in open(path):
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif re.match(r"\s*Fatbin (elf|ptx) code", ln):
            cur = None                      # next object of the fat binary: not part of the previous function
        elif cur and "identifier" not in ln and not re.match(r"\s*/\* 0x", ln):
            d[cur].append(ln)
    return d


if __name__ == "__main__":
    a, b = split(sys.argv[1]), split(sys.argv[2])
    diff = [k for k in a if k in b and a[k] != b[k]]
    gone = [k for k in a if k not in b]
    new = [k for k in b if k not in a]
    print(f"{len(a)} kernels before, {len(b)} after: {len(a) - len(diff) - len(gone)} identical, {len(diff)} changed, {len(gone)} removed, {len(new)} new")
    for k in diff:
        print("  changed:", k)
    for k in gone:
        print("  removed:", k)
    sys.exit(1 if diff or gone else 0)
