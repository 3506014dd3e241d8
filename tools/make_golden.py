"""Generate tests/golden/* from the REFERENCE ITSELF (oracle/_ref/libdada2ref.so =
/root/reference/src compiled unmodified + R-nmath ppois restatement) and from mpmath.

Run in the build container (needs /root/reference for the fixtures and _ref build):
    python tools/make_golden.py
Outputs (all committed):
    tperr1.npy                      16x41 error matrix of data/tperr1.rda
    config1_sam1F_input.npz         derep of inst/extdata/sam1F.fastq.gz (BASELINE config 1)
    e2e_<case>.npz                  reference outputs per tests/cases.py case (+ config1)
    pairs.npz                       random / adversarial pair corpus with reference
                                    alignments (all aligners), Subs and lambdas
    ppois_grid.json                 Poisson upper tails by exact summation (mpmath)
"""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import refdata, synth  # noqa: E402
from tools.poisson_exact import upper_tail  # noqa: E402
from oracle import ref  # noqa: E402
from tests import cases  # noqa: E402

G = cases.GOLDEN
REF = "/root/reference"


def save_res(path, res):
    flat = cases.flatten(res)
    np.savez_compressed(path, **{k.replace(".", "__"): v for k, v in flat.items()})


def gen_pairs(rng, err):
    """Random + adversarial (centre, raw) pairs."""
    out = []
    def rseq(n): return "".join("ACGT"[c] for c in rng.integers(0, 4, size=n))
    def mutate(s, nsub, nindel, homop=False):
        s = list(s)
        for _ in range(nsub):
            p = int(rng.integers(0, len(s))); s[p] = "ACGT"[("ACGT".index(s[p]) + int(rng.integers(1, 4))) % 4]
        for _ in range(nindel):
            p = int(rng.integers(1, len(s) - 1)); n = int(rng.integers(1, 5))
            if rng.random() < 0.5: del s[p:p + n]
            else: s[p:p] = [s[p]] * n if homop else list(rseq(n))
        return "".join(s)
    for L in (30, 63, 64, 65, 100, 250, 251):
        for _ in range(14):
            a = rseq(L)
            r = rng.random()
            if r < 0.3: b = mutate(a, int(rng.integers(0, 6)), 0)
            elif r < 0.6: b = mutate(a, int(rng.integers(0, 20)), int(rng.integers(1, 4)))
            elif r < 0.8: b = mutate(a, int(rng.integers(0, 8)), int(rng.integers(1, 3)), homop=True)
            else: b = rseq(L + int(rng.integers(-5, 6)))
            if rng.random() < 0.3:  # length tweak
                d = int(rng.integers(1, 9))
                b = b[:-d] if rng.random() < 0.5 and len(b) > d + 8 else b + rseq(d)
            out.append((a, b))
    # low complexity / homopolymer-rich
    for _ in range(20):
        L = int(rng.integers(40, 200))
        unit = rseq(int(rng.integers(1, 4)))
        a = (unit * L)[:L]
        a = mutate(a, int(rng.integers(0, 5)), 0)
        b = mutate(a, int(rng.integers(0, 4)), int(rng.integers(0, 3)), homop=True)
        out.append((a, b))
    return out


def main():
    os.makedirs(G, exist_ok=True)
    err = refdata.read_err_rda(os.path.join(REF, "data", "tperr1.rda"))
    np.save(os.path.join(G, "tperr1.npy"), err)
    # ---- config 1
    u, ab, q, _ = refdata.derep_fastq(os.path.join(REF, "inst", "extdata", "sam1F.fastq.gz"))
    np.savez_compressed(os.path.join(G, "config1_sam1F_input.npz"), seqs=np.array(u), abund=ab, quals=q)
    res = ref.dada_uniques(u, ab, None, err, q)
    assert len(res["clustering"]["sequence"]) == 10
    save_res(os.path.join(G, "e2e_config1.npz"), res)
    # ---- synthetic e2e cases
    for name in cases.E2E_CASES:
        seqs, ab, pri, e, qq, opts = cases.build_case(name)
        o = dict(opts)
        o.setdefault("homo_gap", o.get("gap", -8))
        res = ref.dada_uniques(seqs, ab, pri, e, qq, **o)
        save_res(os.path.join(G, "e2e_%s.npz" % name), res)
        print(name, "nclust", len(res["clustering"]["sequence"]), "nbs", len(res["birth_subs"]["pos"]))
    # ---- pair corpus
    rng = np.random.default_rng(2024)
    pairs = gen_pairs(rng, err)
    rec = dict(a=[], b=[], qa=[], qb=[])
    modes = [("vec16", dict(band_size=16)), ("vec5", dict(band_size=5)), ("vec33", dict(band_size=33)),
             ("vec_unb", dict(band_size=-1)),
             ("sc16", dict(band_size=16, vectorized_alignment=False)),
             ("homo16", dict(band_size=16, vectorized_alignment=False, homo_gap=-1)),
             ("homo7", dict(band_size=7, vectorized_alignment=False, homo_gap=-2, gap=-9))]
    cols = {m: dict(al0=[], al1=[], lam=[], nsubs=[], kind=[]) for m, _ in modes}
    for a, b in pairs:
        qa = rng.integers(2, 41, size=len(a)).astype(np.uint8)
        qb = rng.integers(2, 41, size=len(b)).astype(np.uint8)
        rec["a"].append(a); rec["b"].append(b); rec["qa"].append(qa.tobytes()); rec["qb"].append(qb.tobytes())
        for m, o in modes:
            oo = dict(o); oo.setdefault("homo_gap", oo.get("gap", -8))
            r = ref.pair(a, qa, b, qb, err, use_kmers=True, kdist_cutoff=0.42, **oo)
            c = cols[m]
            c["al0"].append(r["al0"]); c["al1"].append(r["al1"]); c["lam"].append(r["lam"]); c["nsubs"].append(r["nsubs"])
            c["kind"].append(0 if r["shrouded"] else (1 if "-" not in (r["al0"] + r["al1"])[:0] and r["kodist"] == r["kdist"] else 2))
    flat = {"a": np.array(rec["a"]), "b": np.array(rec["b"]), "qa": np.array(rec["qa"], dtype=object), "qb": np.array(rec["qb"], dtype=object)}
    save = {"a": flat["a"], "b": flat["b"],
            "qa": np.frombuffer(b"".join(rec["qa"]), dtype=np.uint8), "qb": np.frombuffer(b"".join(rec["qb"]), dtype=np.uint8)}
    for m, _ in modes:
        for k, v in cols[m].items():
            save["%s__%s" % (m, k)] = np.array(v)
    np.savez_compressed(os.path.join(G, "pairs.npz"), **save)
    with open(os.path.join(G, "pairs_modes.json"), "w") as f:
        json.dump({m: o for m, o in modes}, f, indent=1)
    print("pairs", len(pairs))
    # ---- ppois grid (exact summation)
    grid = []
    prng = np.random.default_rng(5)
    pts = [(r, e) for r in (1, 2, 3, 5, 10, 17, 50, 137, 1000, 20000) for e in (1e-300, 1e-30, 1e-8, 1e-3, 0.5, 1.0, 3, 50, 137, 999, 1001, 1e5)]
    for _ in range(400):
        r = int(10 ** prng.uniform(0, 5.5))
        e = r * 10 ** prng.uniform(-6, 0.5) if prng.random() < 0.7 else 10 ** prng.uniform(-20, 6)
        pts.append((r, float(e)))
    import mpmath as mp
    for r, e in pts:
        ex = upper_tail(r, mp.mpf(e))
        grid.append({"reads": r, "E": e, "p": mp.nstr(ex, 25) if ex > mp.mpf("1e-4000") else "0"})
    with open(os.path.join(G, "ppois_grid.json"), "w") as f:
        json.dump(grid, f)
    print("ppois grid", len(grid))


if __name__ == "__main__":
    main()
