"""Builds tests/emu/_build/libdada2b_emu.so: the CUDA sources of dada2_b200/csrc compiled for the host SIMT emulator.

TEST INFRASTRUCTURE ONLY (see cuda_emu.h).  The .cu files are not modified: a textual pass turns
    kern<<<grid, block, smem, stream>>>(args);   ->  cuemu::launch("kern", [=]() { kern(args); }, grid, block, smem, stream);
    extern __shared__ T name[];                  ->  T *name = (T *)cuemu::dyn_smem();
and points the driver's dlopen of libnccl at the emulator's in-process stand-in (ranks = threads), into
tests/emu/_build/*.cpp, which g++ compiles against cuda_emu.h (-ffp-contract=off mirrors nvcc's -fmad=false).
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "dada2_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libdada2b_emu.so")

_LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;()]*>)?)\s*<<<(.+?)>>>\s*\((.*)\)\s*;")
_DYN = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w:]+(?:\s+\w+)*?)\s+(\w+)\s*\[\s*\]\s*;")


def transform(text):
    out = []
    for line in text.splitlines():
        if "<<<" in line:
            new = _LAUNCH.sub(lambda m: f"cuemu::launch(\"{m.group(1)}\", [=]() {{ {m.group(1)}({m.group(3)}); }}, {m.group(2)});", line)
            if new == line:
                raise RuntimeError("build_emu: cannot rewrite launch: " + line.strip())
            line = new
        elif "extern __shared__" in line:
            new = _DYN.sub(lambda m: f"{m.group(1)} *{m.group(2)} = ({m.group(1)} *)cuemu::dyn_smem();", line)
            if new == line:
                raise RuntimeError("build_emu: cannot rewrite dynamic shared memory: " + line.strip())
            line = new
        if "dlopen(name" in line:          # NCCL is resolved with dlopen: point it at the emulator's in-process stand-in
            line = line.replace("dlopen(name", "dlopen(cuemu_self_path()")
        out.append(line)
    return "\n".join(out) + "\n"


def sources():
    sys.path.insert(0, ROOT)
    from dada2_b200.build import SOURCES
    return list(SOURCES)


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp", "build_emu.py")]
    deps += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))] + [os.path.join(ROOT, "dada2_b200", "build.py")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, extra_sources=(), lib=LIB, opt="-O1", asan=False):
    """asan=True builds libdada2b_emu_asan.so with AddressSanitizer (load it with LD_PRELOAD=`g++ -print-file-name=libasan.so`):
    every out-of-bounds access of a kernel to a "device" buffer or beyond its dynamic shared memory is reported."""
    if asan:
        lib = LIB.replace(".so", "_asan.so")
    if not force and not needs_build(lib) and not extra_sources:
        return lib
    os.makedirs(OUT, exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    asan_dir = []
    if asan:                                 # a compiler wrapper earlier on PATH may not know where libasan lives
        for probe in (cxx, "/usr/bin/g++"):
            try:
                f = subprocess.run([probe, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
            except OSError:
                continue
            if os.path.isabs(f):
                asan_dir = ["-L", os.path.dirname(f)]
                break
    flags = [opt, "-g1", "-std=c++17", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-strict-aliasing", "-w", "-DDADA2B_EMU=1",
             "-I", os.path.join(HERE, "stub"), "-I", CSRC, "-include", os.path.join(HERE, "cuda_emu.h")]
    san = ["-fsanitize=address", "-fno-omit-frame-pointer", "--param", "asan-stack=0"] if asan else []     # heap + shared memory checks; fiber stacks stay uninstrumented
    flags += san
    tag = ".asan" if asan else ""
    procs, objs = [], []
    for src in list(sources()) + list(extra_sources):
        path = src if os.path.isabs(src) else os.path.join(CSRC, src)
        cpp = os.path.join(OUT, os.path.basename(src).replace(".cu", tag + ".emu.cpp"))
        with open(path) as f:
            body = transform(f.read())
        # keep relative includes of the original directory working
        with open(cpp, "w") as f:
            f.write(f'#line 1 "{path}"\n' + body)
        obj = cpp.replace(".cpp", ".o")
        cmd = [cxx] + flags + ["-I", os.path.dirname(path), "-c", cpp, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    obj = os.path.join(OUT, "cuda_emu" + tag + ".o")
    cmd = [cxx, "-O2", "-g1", "-std=c++17", "-fPIC"] + san + ["-c", os.path.join(HERE, "cuda_emu.cpp"), "-o", obj]
    procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs.append(obj)
    for cmd, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode:
            sys.stderr.write(out[-6000:])
            raise RuntimeError("emulator build failed: " + " ".join(cmd))
    subprocess.check_call([cxx, "-shared", "-o", lib] + san + asan_dir + objs + ["-ldl", "-pthread"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, asan="--asan" in sys.argv))
