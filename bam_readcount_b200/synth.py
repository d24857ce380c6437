"""Synthetic workloads of SURVEY.md §8(d): a uniform-ACGT contig and sorted 150 bp reads.

Distributions (fixed by the survey so CPU and GPU numbers stay comparable):
CIGAR mix 90 % ``150M``, 3 % ``70M2I78M``, 3 % ``60M3D90M``, 4 % ``10S140M``; per-base
substitution 0.5 %; base qualities iid from {37,37,37,30,25,12,2}; 20 % of forward reads get a
trailing Q2 run of 1–19; strand 50/50 (flag 0/16, unpaired so SE-mapq = mapq); MAPQ iid from
{60,60,60,40,20,0}; ``NM`` = substitutions + indel bases; library = read index mod n_libs.
Everything is vectorised numpy (PCG64, seeded) so a 10 Mb x 30x batch builds in seconds and
any window regenerates bit-identically on the GPU box.
"""
from __future__ import annotations

import numpy as np

from .batch import ReadBatch, TAG_ABSENT

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_NIB = np.array([1, 2, 4, 8], dtype=np.uint8)
_QUALS = np.array([37, 37, 37, 30, 25, 12, 2], dtype=np.uint8)
_MAPQS = np.array([60, 60, 60, 40, 20, 0], dtype=np.uint8)

READ_LEN = 150
# (cigar ops, reference span)
_CIGARS = [
    (np.array([(150 << 4) | 0], dtype=np.uint32), 150),
    (np.array([(70 << 4) | 0, (2 << 4) | 1, (78 << 4) | 0], dtype=np.uint32), 148),
    (np.array([(60 << 4) | 0, (3 << 4) | 2, (90 << 4) | 0], dtype=np.uint32), 153),
    (np.array([(10 << 4) | 4, (140 << 4) | 0], dtype=np.uint32), 140),
]
_MAX_SPAN = 153


def synth_reference(length: int, seed: int = 1234) -> np.ndarray:
    """Uniform ACGT contig as ASCII uint8."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return _ASCII[rng.integers(0, 4, size=length, dtype=np.uint8)]


def _ref_codes(ref_ascii: np.ndarray) -> np.ndarray:
    lut = np.zeros(256, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
    return lut[ref_ascii]


def synth_reads(ref_ascii: np.ndarray, depth: float, seed: int = 1234, n_libs: int = 8, tid: int = 0,
                start_lo: int = 0, start_hi: int | None = None, chunk: int = 262144,
                n_reads: int | None = None) -> ReadBatch:
    """Reads starting uniformly in [start_lo, start_hi) at ``depth`` x coverage, sorted by start."""
    L = int(ref_ascii.shape[0])
    hi = (L - _MAX_SPAN) if start_hi is None else min(start_hi, L - _MAX_SPAN)
    lo = start_lo
    assert hi > lo
    n = int(round(depth * (hi - lo) / READ_LEN)) if n_reads is None else int(n_reads)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    starts = np.sort(rng.integers(lo, hi, size=n, dtype=np.int64)).astype(np.int32)
    kind = rng.choice(4, size=n, p=[0.90, 0.03, 0.03, 0.04]).astype(np.int8)
    reverse = rng.random(n) < 0.5
    mapq = _MAPQS[rng.integers(0, _MAPQS.shape[0], size=n)]
    codes = _ref_codes(ref_ascii)

    seq = np.empty((n, (READ_LEN + 1) // 2), dtype=np.uint8)
    qual = np.empty((n, READ_LEN), dtype=np.uint8)
    nm = np.empty(n, dtype=np.int32)
    # query position -> reference offset from start (or -1 for non-reference bases), per kind
    q = np.arange(READ_LEN)
    idx_kind = np.stack([
        q,
        np.where(q < 70, q, np.where(q < 72, -1, q - 2)),
        np.where(q < 60, q, q + 3),
        np.where(q < 10, -1, q - 10),
    ]).astype(np.int64)
    indel_bases = np.array([0, 2, 3, 0], dtype=np.int32)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        m = c1 - c0
        k = kind[c0:c1]
        off = idx_kind[k]                                    # [m,150]
        from_ref = off >= 0
        gidx = starts[c0:c1, None].astype(np.int64) + np.where(from_ref, off, 0)
        base = codes[gidx]
        rnd = rng.integers(0, 4, size=(m, READ_LEN), dtype=np.uint8)
        base = np.where(from_ref, base, rnd)
        sub = (rng.random((m, READ_LEN)) < 0.005) & from_ref
        shift = rng.integers(1, 4, size=(m, READ_LEN), dtype=np.uint8)
        base = np.where(sub, (base + shift) & 3, base).astype(np.uint8)
        nm[c0:c1] = sub.sum(axis=1).astype(np.int32) + indel_bases[k]
        nib = _NIB[base]
        seq[c0:c1] = (nib[:, 0::2] << 4) | nib[:, 1::2]
        ql = _QUALS[rng.integers(0, _QUALS.shape[0], size=(m, READ_LEN))]
        tail = np.where((~reverse[c0:c1]) & (rng.random(m) < 0.2), rng.integers(1, 20, size=m), 0)
        ql = np.where(q[None, :] >= (READ_LEN - tail)[:, None], np.uint8(2), ql)
        qual[c0:c1] = ql

    n_cig = np.array([c[0].shape[0] for c in _CIGARS], dtype=np.int64)[kind]
    cigar_off = np.zeros(n + 1, dtype=np.uint64)
    cigar_off[1:] = np.cumsum(n_cig)
    cigar = np.empty(int(cigar_off[-1]), dtype=np.uint32)
    co = cigar_off[:-1].astype(np.int64)
    for kk, (ops, _) in enumerate(_CIGARS):
        sel = co[kind == kk]
        for j in range(ops.shape[0]):
            cigar[sel + j] = ops[j]
    step_s = (READ_LEN + 1) // 2
    return ReadBatch(
        tid=np.full(n, tid, dtype=np.int32), pos=starts, flag=np.where(reverse, 16, 0).astype(np.uint16), mapq=mapq,
        lib=(np.arange(n) % n_libs).astype(np.uint16), l_qseq=np.full(n, READ_LEN, dtype=np.int32), nm=nm,
        sm=np.full(n, TAG_ABSENT, dtype=np.int32), cigar_off=cigar_off, cigar=cigar,
        seq_off=(np.arange(n + 1, dtype=np.uint64) * np.uint64(step_s)), seq=seq.reshape(-1),
        qual_off=(np.arange(n + 1, dtype=np.uint64) * np.uint64(READ_LEN)), qual=qual.reshape(-1), qname=None)


def synth_deep_panel(ref_ascii: np.ndarray, sites: np.ndarray, depth: int, seed: int = 1234, n_libs: int = 8,
                     tid: int = 0):
    """Config-5 shape: for each 0-based ``site``, ``depth`` reads whose span covers it, starts uniform.
    Returns (batch, read_lo[n_sites+1]) with reads grouped per site, each group sorted by start."""
    parts, bounds = [], [0]
    for i, s in enumerate(np.asarray(sites, dtype=np.int64)):
        lo = max(0, int(s) - 139)
        hi = int(s) + 1
        b = synth_reads(ref_ascii, 0, seed=seed + 7919 * (i + 1), n_libs=n_libs, tid=tid, start_lo=lo, start_hi=hi,
                        n_reads=depth)
        parts.append(b)
        bounds.append(bounds[-1] + b.n_reads)
    return ReadBatch.concat(parts), np.array(bounds, dtype=np.int64)


_DEC = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)


def write_sam(path: str, batch: ReadBatch, contigs, n_libs: int = 8, read_group: bool = True) -> None:
    """Write the batch as SAM text (for the reference binary via samtools view -b)."""
    ops = "MIDNSHP=XB"
    with open(path, "w") as fh:
        fh.write("@HD\tVN:1.6\tSO:coordinate\n")
        for name, ln in contigs:
            fh.write(f"@SQ\tSN:{name}\tLN:{ln}\n")
        if read_group:
            for i in range(n_libs):
                fh.write(f"@RG\tID:rg{i}\tSM:s\tLB:lib{i}\n")
        names = [c[0] for c in contigs]
        co = batch.cigar_off.astype(np.int64)
        so = batch.seq_off.astype(np.int64)
        qo = batch.qual_off.astype(np.int64)
        for i in range(batch.n_reads):
            cig = batch.cigar[co[i]:co[i + 1]]
            cs = "".join(f"{int(c) >> 4}{ops[int(c) & 15]}" for c in cig) or "*"
            l = int(batch.l_qseq[i])
            pk = batch.seq[so[i]:so[i + 1]]
            nibs = np.empty(pk.shape[0] * 2, dtype=np.uint8)
            nibs[0::2] = pk >> 4
            nibs[1::2] = pk & 15
            s = _DEC[nibs[:l]].tobytes().decode()
            ql = (batch.qual[qo[i]:qo[i + 1]] + 33).tobytes().decode()
            tags = []
            if int(batch.nm[i]) != int(TAG_ABSENT):
                tags.append(f"NM:i:{int(batch.nm[i])}")
            if int(batch.sm[i]) != int(TAG_ABSENT):
                tags.append(f"SM:i:{int(batch.sm[i])}")
            if read_group and int(batch.lib[i]) != 0xFFFF:
                tags.append(f"RG:Z:rg{int(batch.lib[i])}")
            qn = batch.qname[i] if batch.qname is not None else f"r{i}"
            fh.write("\t".join([qn, str(int(batch.flag[i])), names[int(batch.tid[i])], str(int(batch.pos[i]) + 1),
                                str(int(batch.mapq[i])), cs, "*", "0", "0", s, ql] + tags) + "\n")


def write_fasta(path: str, name: str, ref_ascii: np.ndarray, width: int = 60) -> None:
    L = int(ref_ascii.shape[0])
    with open(path, "wb") as fh:
        fh.write(f">{name}\n".encode())
        full = (L // width) * width
        if full:
            body = np.empty((L // width, width + 1), dtype=np.uint8)
            body[:, :width] = ref_ascii[:full].reshape(-1, width)
            body[:, width] = 10
            fh.write(body.tobytes())
        if L > full:
            fh.write(ref_ascii[full:].tobytes() + b"\n")
    n_lines_full = L // width
    with open(path + ".fai", "w") as fh:
        fh.write(f"{name}\t{L}\t{len(name) + 2}\t{width}\t{width + 1}\n")
    _ = n_lines_full
