"""List the loops (backward branches) of every kernel in a cubin with their SASS instruction counts and opcode mix.

Offline tuning aid: `python tools/sass_loops.py file.cubin [name-filter]`.  A loop is reported as the instruction
range between a backward branch and its target label; nested ranges show up as separate lines.
"""
import collections
import re
import subprocess
import sys


def functions(cubin):
    txt = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name is not None:
            body.append(ln)
    if name:
        yield name, body


def loops(body):
    labels, ins = {}, []
    for ln in body:
        m = re.match(r"\s*(\.L_x_\d+):", ln)
        if m:
            labels[m.group(1)] = len(ins)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
        if m:
            ins.append(m.group(1).strip())
    out = []
    for i, s in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\S+,\s*)?`\((\.L_x_\d+)\)", s)
        if m and m.group(1) in labels and labels[m.group(1)] <= i:
            lo = labels[m.group(1)]
            ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for x in ins[lo:i + 1])
            out.append((lo, i, ops))
    return len(ins), out


if __name__ == "__main__":
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    for name, body in functions(sys.argv[1]):
        if flt not in name:
            continue
        n, ls = loops(body)
        print(f"{name}: {n} instructions")
        for lo, hi, ops in ls:
            top = " ".join(f"{k}:{v}" for k, v in ops.most_common(14))
            print(f"   loop [{lo:5d},{hi:5d}] {hi - lo + 1:5d} instr  {top}")
