"""Randomised differential campaign of the SURVEY 8(f) kernels on the host SIMT emulator against the CPU oracle (test tooling, not part
of the suite): bimera pairs and tables (both alignment kernels, random scores / bands / one-off), mergePairs pairings, dereplication.
  python tools/fuzz_emu.py <seed> <seconds>      (round 1: seeds 1-3, 20 min each: 5 398 iterations, 431 840 bimera pairs, all identical)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import numpy as np
import build_emu
lib = build_emu.build()
import dada2_b200.api as api
api._LIBPATH = lib
from dada2_b200 import bimera, merge, derep
from oracle import port, derep as OD
from tools import synth
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
def rs(L): return "".join("ACGT"[i] for i in rng.integers(0,4,L))
def mut(s,k):
    s=list(s)
    for p in rng.choice(len(s),min(k,len(s)),replace=False): s[p]="ACGT"[("ACGT".index(s[p])+1+rng.integers(0,3))%4]
    return "".join(s)
def indel(s):
    i=int(rng.integers(0,len(s)))
    return s[:i]+rs(int(rng.integers(1,5)))+s[i:] if rng.random()<0.5 else s[:i]+s[i+int(rng.integers(1,5)):]
t_end = time.time() + float(sys.argv[2]) if len(sys.argv) > 2 else time.time() + 600
it = 0; npairs = 0
while time.time() < t_end:
    it += 1
    # ---- bimera pairs, both kernels, random scores / band ----
    base = rs(int(rng.integers(30, 160)))
    seqs = []
    for _ in range(24):
        m = int(rng.integers(0, 6)); s = base
        if m == 1: s = mut(s, int(rng.integers(1, 6)))
        elif m == 2: s = indel(s)
        elif m == 3: o = rs(len(base)); bp = int(rng.integers(5, len(base) - 5)); s = base[:bp] + o[bp:]
        elif m == 4: s = s[int(rng.integers(0, 10)):len(s) - int(rng.integers(0, 10))]
        elif m == 5: s = rs(int(rng.integers(30, 160)))
        if len(s) >= 8: seqs.append(s)
    q = rng.integers(0, len(seqs), 40); p = rng.integers(0, len(seqs), 40)
    o = dict(allow_one_off=bool(rng.integers(0, 2)), max_shift=int(rng.choice([0, 1, 5, 16, 16, 30, 64])), match=int(rng.integers(1, 7)),
             mismatch=-int(rng.integers(1, 9)), gap_p=-int(rng.integers(1, 12)))
    if it % 2 == 0:                      # every other round: runs of one query against equally long parents (what dd_bimfwd16.cu pairs up)
        same = [x for x in range(len(seqs)) if len(seqs[x]) == len(base)]
        if len(same) >= 2:
            q = np.repeat(rng.integers(0, len(seqs), 5), 8); p = rng.choice(same, 40)
    for fwd in (None, "1", "2"):
        if fwd: os.environ["DADA2B_BIMFWD"] = fwd
        else: os.environ.pop("DADA2B_BIMFWD", None)
        got = bimera.test_bimera_pairs(seqs, q, p, **o)
        for n, (a, b) in enumerate(zip(q, p)):
            r = port.bimera_pair(seqs[a], seqs[b], **o)
            want = [r["left"], r["right"], r["left_oo"] if o["allow_one_off"] else got[n][2], r["right_oo"] if o["allow_one_off"] else got[n][3], r["ham"]]
            assert list(got[n]) == want, ("bimera", fwd, o, seqs[a], seqs[b], list(got[n]), want)
        npairs += len(q)
    os.environ.pop("DADA2B_BIMFWD", None)
    # ---- small table, random options ----
    sq, mat = synth.bimera_table(int(rng.integers(12, 40)), int(rng.integers(1, 5)), seed=int(rng.integers(0, 1 << 30)), L=int(rng.integers(40, 100)), lenvar=int(rng.integers(0, 8)))
    to = dict(allow_one_off=bool(rng.integers(0, 2)), min_fold=float(rng.choice([0.5, 1.0, 1.5, 2.0])), min_abund=int(rng.integers(1, 9)),
              max_shift=int(rng.choice([0, 4, 16, 32])), min_one_off_par_dist=int(rng.integers(1, 6)))
    want = port.table_bimera(mat, sq, **to)
    for fwd in (None, "1", "2"):
        if fwd: os.environ["DADA2B_BIMFWD"] = fwd
        else: os.environ.pop("DADA2B_BIMFWD", None)
        g = bimera.C_table_bimera2(mat, sq, **to)
        assert np.array_equal(g["nflag"], want[0]) and np.array_equal(g["nsam"], want[1]), ("table", fwd, to)
    os.environ.pop("DADA2B_BIMFWD", None)
    # ---- merge ----
    ms, a, b = [], [], []
    for _ in range(12):
        amp = rs(int(rng.integers(40, 200))); lf, lr = int(rng.integers(20, 120)), int(rng.integers(20, 120))
        f, r = amp[:lf], amp[max(0, len(amp) - lr):]
        k = int(rng.integers(0, 4))
        if k == 1: r = mut(r, 2)
        elif k == 2: r = indel(r)
        elif k == 3: f = rs(4) + f
        if len(f) < 8 or len(r) < 8: continue
        ms += [f, r]; a.append(len(ms) - 2); b.append(len(ms) - 1)
    pref = rng.integers(1, 3, len(a)).astype(np.int32)
    mo = dict(mismatch=-int(rng.choice([1, 8, 64])), gap_p=-int(rng.choice([1, 8, 64])), trim_overhang=bool(rng.integers(0, 2)), band=int(rng.choice([-1, -1, 8, 40])))
    if rng.random() < 0.3: mo["homo_gap_p"] = -1
    g = merge.merge_align(ms, a, b, pref, **mo)
    for x, (i, j) in enumerate(zip(a, b)):
        w = port.merge_pair(ms[i], ms[j], prefer=int(pref[x]), **mo)
        assert (int(g["nmatch"][x]), int(g["nmismatch"][x]), int(g["nindel"][x]), g["sequence"][x]) == (w["nmatch"], w["nmismatch"], w["nindel"], w["sequence"]), ("merge", mo, ms[i], ms[j])
    # ---- derep ----
    from tests import derep_cases as D
    s, qq = D.synthetic(int(rng.integers(50, 900)), seed=int(rng.integers(0, 1 << 30)), L=int(rng.integers(14, 90)), nvar=int(rng.integers(2, 30)), zero_len=int(rng.integers(0, 3)))
    n = int(rng.choice([1000000, 7, 64, 300]))
    D.assert_same(derep.derep_reads(s, qq, n=n), OD.derep_reads(s, qq, n=n), "derep fuzz")
    if it % 8 == 0:                      # derep left resident on the device -> dada() == the two-call form (bit-identical)
        import dada2_b200
        from tests import cases
        keep = [i for i in range(len(s)) if len(s[i]) > 10 or len(s[i]) == 0]
        s3, q3 = [s[i] for i in keep], [qq[i] for i in keep]
        d3, r3 = derep.derep_reads(s3, q3, n=n, resident=True)
        e3 = np.full((16, 45), 0.01); e3[[0, 5, 10, 15]] = 0.97
        cases.assert_same(r3.run(e3), dada2_b200.dada_uniques(d3["uniques"], d3["abundances"], None, e3, d3["quals"]), rtol=0, label="derep resident fuzz")
        r3.close()
print("FUZZ OK iterations", it, "bimera pairs", npairs)
