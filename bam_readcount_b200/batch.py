"""Compact decoded-read batch: the host-side container streamed to the device.

One ``ReadBatch`` is the struct-of-arrays equivalent of the ``bam1_t`` records the reference
hands to ``fetch_func`` (R:src/exe/bam-readcount/bamreadcount.cpp:114) — only the fields the
pileup hot path consumes: core fields, CIGAR, 4-bit sequence, base
qualities, the NM/SM integer tags and a dense library id replacing ``bam_get_library``'s
string (V:bam.c:77-101).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

TAG_ABSENT = np.int32(-2**31)   # NM / SM tag missing
LIB_NONE = np.uint16(0xFFFF)    # read has no RG or its @RG has no LB

BAM_FUNMAP = 4
BAM_FREVERSE = 16
BAM_FPROPER_PAIR = 2

CIGAR_OPS = "MIDNSHP=XB"


@dataclass
class ReadBatch:
    tid: np.ndarray        # int32 [n]
    pos: np.ndarray        # int32 [n] 0-based leftmost
    flag: np.ndarray       # uint16 [n]
    mapq: np.ndarray       # uint8 [n]
    lib: np.ndarray        # uint16 [n]
    l_qseq: np.ndarray     # int32 [n]
    nm: np.ndarray         # int32 [n]
    sm: np.ndarray         # int32 [n]
    cigar_off: np.ndarray  # uint64 [n+1]
    cigar: np.ndarray      # uint32 [sum n_cigar]
    seq_off: np.ndarray    # uint64 [n+1]
    seq: np.ndarray        # uint8, BAM 4-bit packing
    qual_off: np.ndarray   # uint64 [n+1]
    qual: np.ndarray       # uint8
    qname: Optional[List[str]] = None

    @property
    def n_reads(self) -> int:
        return int(self.pos.shape[0])

    def ref_end(self) -> np.ndarray:
        """bam_endpos per read (V:htslib-1.10/sam.c:507-513)."""
        n = self.n_reads
        ops = self.cigar & 0xF
        lens = (self.cigar >> 4).astype(np.int64)
        consumes = np.isin(ops, (0, 2, 3, 7, 8))
        contrib = np.where(consumes, lens, 0)
        csum = np.concatenate([[0], np.cumsum(contrib)])
        rlen = csum[self.cigar_off[1:].astype(np.int64)] - csum[self.cigar_off[:-1].astype(np.int64)]
        n_cig = (self.cigar_off[1:] - self.cigar_off[:-1]).astype(np.int64)
        unm = (self.flag & BAM_FUNMAP) != 0
        end = self.pos.astype(np.int64) + np.where(unm | (n_cig == 0), 1, rlen)
        assert end.shape[0] == n
        return end

    def select(self, idx: np.ndarray) -> "ReadBatch":
        """Gather a subset of reads (file order preserved by the caller)."""
        idx = np.asarray(idx, dtype=np.int64)

        def gather(off, pool):
            lo = off[:-1][idx].astype(np.int64)
            hi = off[1:][idx].astype(np.int64)
            ln = hi - lo
            new_off = np.zeros(idx.shape[0] + 1, dtype=np.uint64)
            new_off[1:] = np.cumsum(ln)
            tot = int(new_off[-1])
            if tot == 0:
                return new_off, pool[:0].copy()
            # vectorised ragged gather
            starts = np.repeat(lo - new_off[:-1].astype(np.int64), ln)
            out = pool[starts + np.arange(tot, dtype=np.int64)]
            return new_off, out

        co, c = gather(self.cigar_off, self.cigar)
        so, s = gather(self.seq_off, self.seq)
        qo, q = gather(self.qual_off, self.qual)
        return ReadBatch(
            tid=self.tid[idx], pos=self.pos[idx], flag=self.flag[idx], mapq=self.mapq[idx], lib=self.lib[idx],
            l_qseq=self.l_qseq[idx], nm=self.nm[idx], sm=self.sm[idx], cigar_off=co, cigar=c, seq_off=so, seq=s,
            qual_off=qo, qual=q, qname=[self.qname[i] for i in idx] if self.qname is not None else None)

    @staticmethod
    def concat(batches: Sequence["ReadBatch"]) -> "ReadBatch":
        def cat_off(offs):
            out = [np.zeros(1, dtype=np.uint64)]
            base = np.uint64(0)
            for o in offs:
                out.append(o[1:] + base)
                base = base + o[-1]
            return np.concatenate(out)
        qn = None
        if all(b.qname is not None for b in batches):
            qn = [q for b in batches for q in b.qname]
        return ReadBatch(
            tid=np.concatenate([b.tid for b in batches]), pos=np.concatenate([b.pos for b in batches]),
            flag=np.concatenate([b.flag for b in batches]), mapq=np.concatenate([b.mapq for b in batches]),
            lib=np.concatenate([b.lib for b in batches]), l_qseq=np.concatenate([b.l_qseq for b in batches]),
            nm=np.concatenate([b.nm for b in batches]), sm=np.concatenate([b.sm for b in batches]),
            cigar_off=cat_off([b.cigar_off for b in batches]), cigar=np.concatenate([b.cigar for b in batches]),
            seq_off=cat_off([b.seq_off for b in batches]), seq=np.concatenate([b.seq for b in batches]),
            qual_off=cat_off([b.qual_off for b in batches]), qual=np.concatenate([b.qual for b in batches]), qname=qn)

    def fetch(self, tid: int, beg: int, end: int) -> np.ndarray:
        """Indices of the records the index iterator yields for [beg, end) on ``tid``
        (V:htslib-1.10/hts.c:3229-3236): ``endpos > max(beg,0)`` and ``pos < end``."""
        e = self.ref_end()
        m = (self.tid == tid) & (e > max(beg, 0)) & (self.pos.astype(np.int64) < end)
        return np.nonzero(m)[0]


def parse_cigar_string(s: str) -> np.ndarray:
    out = []
    num = 0
    for ch in s:
        if ch.isdigit():
            num = num * 10 + ord(ch) - 48
        else:
            out.append((num << 4) | CIGAR_OPS.index(ch))
            num = 0
    return np.array(out, dtype=np.uint32)


_NT16 = np.full(256, 15, dtype=np.uint8)
for _i, _c in enumerate("=ACMGRSVTWYHKDBN"):
    _NT16[ord(_c)] = _i
    _NT16[ord(_c.lower())] = _i


def pack_seq(ascii_seq: bytes) -> np.ndarray:
    """ASCII bases -> BAM 4-bit packing (high nibble first)."""
    codes = _NT16[np.frombuffer(ascii_seq, dtype=np.uint8)]
    if codes.shape[0] & 1:
        codes = np.concatenate([codes, np.zeros(1, dtype=np.uint8)])
    return ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)


@dataclass
class BatchBuilder:
    """Per-read append interface (the shape of ``fetch_func(b)`` + ``bam_plbuf_push(b)``)."""
    tid: list = field(default_factory=list)
    pos: list = field(default_factory=list)
    flag: list = field(default_factory=list)
    mapq: list = field(default_factory=list)
    lib: list = field(default_factory=list)
    l_qseq: list = field(default_factory=list)
    nm: list = field(default_factory=list)
    sm: list = field(default_factory=list)
    cigars: list = field(default_factory=list)
    seqs: list = field(default_factory=list)
    quals: list = field(default_factory=list)
    qname: list = field(default_factory=list)

    def add(self, *, tid, pos, flag, mapq, lib, cigar, seq4, qual, nm=None, sm=None, l_qseq=None, qname="r"):
        cigar = np.asarray(cigar, dtype=np.uint32)
        qual = np.asarray(qual, dtype=np.uint8)
        l = int(qual.shape[0]) if l_qseq is None else int(l_qseq)
        self.tid.append(tid); self.pos.append(pos); self.flag.append(flag); self.mapq.append(mapq)
        self.lib.append(int(LIB_NONE) if lib is None else lib); self.l_qseq.append(l)
        self.nm.append(int(TAG_ABSENT) if nm is None else nm); self.sm.append(int(TAG_ABSENT) if sm is None else sm)
        self.cigars.append(cigar); self.seqs.append(np.asarray(seq4, dtype=np.uint8)); self.quals.append(qual)
        self.qname.append(qname)

    def add_sam(self, *, tid, pos, flag, mapq, lib, cigar: str, seq: str, qual, nm=None, sm=None, qname="r"):
        if isinstance(qual, str):
            qual = np.frombuffer(qual.encode(), dtype=np.uint8) - 33
        self.add(tid=tid, pos=pos, flag=flag, mapq=mapq, lib=lib, cigar=parse_cigar_string(cigar),
                 seq4=pack_seq(seq.encode()), qual=qual, nm=nm, sm=sm, l_qseq=len(seq), qname=qname)

    def build(self) -> ReadBatch:
        n = len(self.pos)

        def offs(parts):
            o = np.zeros(n + 1, dtype=np.uint64)
            if n:
                o[1:] = np.cumsum([p.shape[0] for p in parts])
            return o

        def cat(parts, dt):
            return np.concatenate(parts).astype(dt) if n else np.zeros(0, dtype=dt)
        return ReadBatch(
            tid=np.array(self.tid, dtype=np.int32), pos=np.array(self.pos, dtype=np.int32),
            flag=np.array(self.flag, dtype=np.uint16), mapq=np.array(self.mapq, dtype=np.uint8),
            lib=np.array(self.lib, dtype=np.uint16), l_qseq=np.array(self.l_qseq, dtype=np.int32),
            nm=np.array(self.nm, dtype=np.int32), sm=np.array(self.sm, dtype=np.int32),
            cigar_off=offs(self.cigars), cigar=cat(self.cigars, np.uint32),
            seq_off=offs(self.seqs), seq=cat(self.seqs, np.uint8),
            qual_off=offs(self.quals), qual=cat(self.quals, np.uint8), qname=list(self.qname))
