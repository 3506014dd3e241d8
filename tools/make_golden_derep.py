"""Generates tests/golden/sam1F_reads.npz: the reads of the reference's fixture inst/extdata/sam1F.fastq.gz (1 500 x 250 nt, Phred+33)
as arrays -- sequences and numeric qualities -- so that the dereplication tests can run where /root/reference does not exist.
The committed config-1 input of dada() (tests/golden/config1_sam1F_input.npz, tools/make_golden.py) is what dereplicating these
reads must reproduce.  Build container only.   python tools/make_golden_derep.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import refdata            # noqa: E402
from oracle import derep             # noqa: E402


def main():
    seqs, quals = refdata.read_fastq("/root/reference/inst/extdata/sam1F.fastq.gz")
    r = derep.derep_reads(seqs, quals)
    z = np.load(os.path.join(ROOT, "tests", "golden", "config1_sam1F_input.npz"))
    assert r["uniques"] == z["seqs"].tolist() and np.array_equal(r["abundances"], z["abund"]) and np.array_equal(r["quals"], z["quals"], equal_nan=True)
    path = os.path.join(ROOT, "tests", "golden", "sam1F_reads.npz")
    np.savez_compressed(path, seqs=np.array(seqs), quals=np.stack([q.astype(np.uint8) for q in quals]))
    print(path, os.path.getsize(path), "bytes;", len(seqs), "reads ->", len(r["uniques"]), "uniques")


if __name__ == "__main__":
    main()
