#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference (run in the build container only).

Inputs: the reference's own fixtures under /root/reference/test-data (copied, they are data) and
the UNMODIFIED reference binary oracle/_ref/bam-readcount (oracle/build_ref.sh).  Outputs:
  expected_*                       the reference's four golden files, verbatim
  test_bam.npz, test_bam_bad_rg.npz  test.bam decoded to the compact batch + the reference
                                   window of contig 21 that its reads touch (ref.fa is 10.5 MB)
  ref_<case>_<flags>.txt.gz        reference-binary STDOUT on the deterministic synthetic cases
                                   of tests/cases.py (deletions, insertions, -q/-b, -i, -p, -d)
  edge_*.txt.gz                    reference-binary STDOUT on the hand-built edge-case reads
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import edge_cases  # noqa: E402
from bam_readcount_b200 import synth  # noqa: E402
from bam_readcount_b200.bamio import Fasta, read_bam  # noqa: E402
from oracle.oracle import REF_BIN, REF_SAMTOOLS, run_reference_binary  # noqa: E402

TD = "/root/reference/test-data"


def decode_fixture(bam, out):
    hdr, b = read_bam(os.path.join(TD, bam))
    fa = Fasta(os.path.join(TD, "ref.fa"))
    wb, we = 10402000, 10406000
    win = np.frombuffer(fa.fetch("21", wb, we), dtype=np.uint8)
    assert b.pos.min() >= wb and b.ref_end().max() <= we
    np.savez_compressed(os.path.join(HERE, out), tid=b.tid, pos=b.pos, flag=b.flag, mapq=b.mapq, lib=b.lib, l_qseq=b.l_qseq,
                        nm=b.nm, sm=b.sm, cigar_off=b.cigar_off, cigar=b.cigar, seq_off=b.seq_off, seq=b.seq,
                        qual_off=b.qual_off, qual=b.qual, ref_win=win, ref_win_beg=np.int64(wb),
                        chrom_len=np.int64(fa.length("21")), lib_names=np.array("\t".join(hdr.lib_names)))


def write_case_files(case, d):
    name, L, seq, wb = case["contigs"][0]
    assert wb == 0
    synth.write_fasta(os.path.join(d, "ref.fa"), name, np.frombuffer(seq, dtype=np.uint8))
    synth.write_sam(os.path.join(d, "s.sam"), case["batch"], [(name, L)], n_libs=len(case["lib_names"]),
                    read_group=case.get("read_group", True))
    if "sam_header_extra" in case:
        pass
    subprocess.check_call([REF_SAMTOOLS, "view", "-b", "-o", os.path.join(d, "s.bam"), os.path.join(d, "s.sam")])
    subprocess.check_call([REF_SAMTOOLS, "index", os.path.join(d, "s.bam")])


def reference_stdout(case, flags, d, site_list):
    name = case["contigs"][0][0]
    argv = ["-w", "0", "-f", os.path.join(d, "ref.fa")] + cases.flags_to_argv(flags)
    if site_list:
        with open(os.path.join(d, "sites"), "w") as fh:
            for (_, b1, e1) in case["regions"]:
                fh.write(f"{name}\t{b1}\t{e1}\n")
        argv += ["-l", os.path.join(d, "sites"), os.path.join(d, "s.bam")]
    else:
        argv += [os.path.join(d, "s.bam")] + [f"{name}:{b1}-{e1}" for (_, b1, e1) in case["regions"]]
    out, err, rc = run_reference_binary(argv)
    assert rc == 0, err[-2000:]
    return out


def main():
    assert os.path.exists(REF_BIN), "run oracle/build_ref.sh first"
    for f in ("expected_all_lib", "expected_per_lib", "expected_insertion_centric_all_lib",
              "expected_insertion_centric_per_lib"):
        shutil.copy(os.path.join(TD, f), os.path.join(HERE, f))
    decode_fixture("test.bam", "test_bam.npz")
    decode_fixture("test_bad_rg.bam", "test_bam_bad_rg.npz")

    # config 2b: the CRAM fixture through the reference binary (the reference ships no expected file for it)
    for flags, name in ((["-p"], "ref_cram_twolib_perlib.txt"), ([], "ref_cram_twolib_alllib.txt")):
        out, err, rc = run_reference_binary(flags + ["-l", "twolib_site_list.txt", "-f", "rand1k.fa", "twolib.sorted.cram"], cwd=TD)
        assert rc == 0
        with gzip.open(os.path.join(HERE, name + ".gz"), "wb", compresslevel=9) as fh:
            fh.write(out.encode("latin-1"))
        print(name, len(out.splitlines()), "lines")

    jobs = []
    syn = cases.synthetic_case(L=12000, depth=30, seed=11, regions=((0, 1000, 4000),), site_list=False)
    for fname, fl in cases.FLAG_SETS.items():
        jobs.append((syn, fname, fl, False, f"ref_syn_{fname}.txt"))
    deep = cases.deep_case(n_sites=3, depth=5000, seed=5)
    jobs.append((deep, "perlib_deep", dict(per_lib=True, max_cnt=100000000), True, "ref_deep_perlib.txt"))
    jobs.append((deep, "alllib_deep", dict(max_cnt=100000000), True, "ref_deep_alllib.txt"))
    for ec in edge_cases.all_cases():
        for fname, fl in ec["flag_sets"].items():
            jobs.append((ec, fname, fl, ec["site_list"], f"edge_{ec['name']}_{fname}.txt"))
    done = {}
    for case, fname, fl, sl, outname in jobs:
        key = case["name"]
        if key not in done:
            d = tempfile.mkdtemp()
            write_case_files(case, d)
            done[key] = d
        txt = reference_stdout(case, fl, done[key], sl)
        with gzip.open(os.path.join(HERE, outname + ".gz"), "wb", compresslevel=9) as fh:
            fh.write(txt.encode("latin-1"))
        print(outname, len(txt.splitlines()), "lines")


if __name__ == "__main__":
    main()
