"""Test tooling: host-side cost of do_upload (validate + SIMD packing pipelined with the H2D copies) at full size without a GPU.
The CUDA sources are built for the host emulator with -O2 into /tmp/emu_<TAG> (device copies become memcpy, so "H2D" here is a
single-threaded memcpy issued by the calling thread under the packing); DADA2B_VERBOSE prints the library's own phase timers.
  N=1000000 python tools/upload_bench.py            (CSRC=<dir> TAG=<name> builds another copy of csrc/ for A/B)
Build container, 8 cores, 1e6 uniques: 160 ms per upload before the rewrite (validate 9.5, pack 122, copy 28), 75-80 ms after."""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import numpy as np
import build_emu
tag = os.environ.get("TAG", "o2")
build_emu.OUT = '/tmp/emu_' + tag
os.makedirs(build_emu.OUT, exist_ok=True)
if os.environ.get("CSRC"): build_emu.CSRC = os.environ["CSRC"]
lib = build_emu.build(lib=build_emu.OUT + '/libdada2b_emu.so', opt='-O2', force='--force' in sys.argv)
from dada2_b200 import _abi
from tests import cases
n = int(os.environ.get("N", "1000000"))
cache = '/tmp/syn%d' % n
if not os.path.exists(cache + '_q.npy'):
    from tools import synth
    seqs, ab, q, truth = synth.illumina(n, seed=12345)
    np.save(cache + '_q.npy', q); np.save(cache + '_ab.npy', ab); open(cache + '_s.txt', 'w').write("\n".join(seqs))
q = np.load(cache + '_q.npy'); ab = np.load(cache + '_ab.npy'); seqs = open(cache + '_s.txt').read().split("\n")
pin = _abi.PackedIn(seqs, ab, None, cases.tperr1(), q)
h = C.CDLL(lib)
h.dada2b_upload.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.c_char_p]
h.dada2b_reupload.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
ctx = C.c_void_p(); eb = C.create_string_buffer(256)
os.environ["DADA2B_VERBOSE"] = "1"
rc = h.dada2b_upload(C.byref(pin.struct), 0, C.byref(ctx), eb); print("upload rc", rc, eb.value, flush=True)
for i in range(6):
    t = time.time(); rc = h.dada2b_reupload(ctx, C.byref(pin.struct), eb); print("reupload rc", rc, "ms", round((time.time() - t) * 1e3, 1), flush=True)
