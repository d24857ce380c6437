"""Micro known-answers from SURVEY.md Appendix E (generated with the reference binary during the survey):
hand-built reads on the 50 bp contig `c` = ACGT x12 + AC, region c:10-10, and the `-d` max-count table.
CPU: the oracle; GPU: the CUDA engine through the C ABI."""
import numpy as np
import pytest

import cases
from bam_readcount_b200.batch import BatchBuilder

REF = ("ACGT" * 12 + "AC").encode()


def _case(reads, regions=((0, 10, 10),)):
    bb = BatchBuilder()
    for r in reads:
        bb.add_sam(tid=0, flag=0, mapq=60, lib=0, sm=None, qname="r", **r)
    return dict(name="ka", contigs=[("c", 50, REF, 0)], batch=bb.build(), regions=list(regions), site_list=False, lib_names=["lib0"])


def _mut(seq, i, b):
    return seq[:i] + b + seq[i + 1:]


R20 = REF[:20].decode()
KNOWN = [
    (dict(pos=0, cigar="20M", seq=_mut(R20, 5, "T"), qual="I" * 20, nm=1), "C:1:60.00:40.00:60.00:1:0:0.90:0.05:40.00:1:0.45:20.00:0.45"),
    (dict(pos=0, cigar="5=1X14=", seq=_mut(R20, 5, "T"), qual="I" * 20, nm=1), "C:1:60.00:40.00:60.00:1:0:0.90:0.05:0.00:1:0.45:20.00:0.45"),
    (dict(pos=0, cigar="3S17M", seq="TTT" + REF[:17].decode(), qual="I" * 20, nm=0), "C:1:60.00:40.00:60.00:1:0:0.94:0.00:0.00:1:0.30:17.00:0.30"),
    (dict(pos=0, cigar="5H3S17M", seq="TTT" + REF[:17].decode(), qual="I" * 20, nm=0), "C:1:60.00:40.00:60.00:1:0:0.59:0.00:0.00:1:0.30:17.00:0.25"),
]
D_TABLE = {100: [20, 20, 20, 40, 40], 10: [10, 10, 10, 11, 11], 5: [5, 5, 5, 6, 6], 1: [1, 1, 1, 2, 2], 0: [1, 1, 1, 2, 2]}


def _block(text, allele="C"):
    line = text.strip().split("\n")[0].split("\t")
    assert line[:2] == ["c", "10"] and line[3] == "1"
    return [f for f in line[4:] if f.startswith(allele + ":")][0]


def _d_case():
    reads = [dict(pos=4, cigar="20M", seq=REF[4:24].decode(), qual="I" * 20, nm=0) for _ in range(20)] + \
            [dict(pos=7, cigar="20M", seq=REF[7:27].decode(), qual="I" * 20, nm=0) for _ in range(20)]
    return _case(reads, regions=((0, 5, 9),))


@pytest.mark.parametrize("k", range(len(KNOWN)))
def test_oracle_known_answers(k):
    read, want = KNOWN[k]
    text, _, _ = cases.run_oracle(_case([read]), dict())
    assert _block(text) == want


@pytest.mark.parametrize("d", sorted(D_TABLE))
def test_oracle_max_count_table(d):
    text, _, _ = cases.run_oracle(_d_case(), dict(max_cnt=d))
    assert [int(l.split("\t")[3]) for l in text.strip().split("\n")] == D_TABLE[d]


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(KNOWN)))
def test_engine_known_answers(k):
    read, want = KNOWN[k]
    text, _, _, _ = cases.run_engine(_case([read]), dict(), want_dump=False)
    assert _block(text) == want


@pytest.mark.gpu
@pytest.mark.parametrize("d", sorted(D_TABLE))
def test_engine_max_count_table(d):
    text, _, _, _ = cases.run_engine(_d_case(), dict(max_cnt=d), want_dump=False)
    assert [int(l.split("\t")[3]) for l in text.strip().split("\n")] == D_TABLE[d]
