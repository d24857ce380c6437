"""Host-side decode: BGZF/BAM records, SAM header @RG->LB map, FASTA (+.fai).

The north star leaves file decode on the host ("BAM/CRAM decode and BAI/CRAI region
iteration left on the host"); this module is the Python host's decoder feeding
``ReadBatch``.  It implements the published BAM container layout (SAM spec §4) directly —
no htslib.  Region selection mirrors the index iterator's overlap rule
(V:htslib-1.10/hts.c:3229-3236) with a linear scan; a BAI reader is not needed for parity.
CRAM is out of scope for this decoder.
"""
from __future__ import annotations

import gzip
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

from .batch import LIB_NONE, TAG_ABSENT, ReadBatch


class BamHeader:
    def __init__(self, text: str, names: List[str], lengths: List[int]):
        self.text = text
        self.target_names = names
        self.target_lengths = lengths
        self.tid_of = {n: i for i, n in enumerate(names)}
        # @RG ID -> LB, like sam_hdr_find_tag_id(h,"RG","ID",rg,"LB") (V:bam.c:88)
        self.rg_lb: Dict[str, Optional[str]] = {}
        for line in text.split("\n"):
            if line.startswith("@RG"):
                tags = dict(f.split(":", 1) for f in line.split("\t")[1:] if ":" in f)
                if "ID" in tags and tags["ID"] not in self.rg_lb:
                    self.rg_lb[tags["ID"]] = tags.get("LB")
        # std::set<std::string> order == byte-lexicographic (R:bamreadcount.cpp:92-111)
        self.lib_names: List[str] = sorted({lb for lb in self.rg_lb.values() if lb is not None},
                                           key=lambda s: s.encode())
        self.lib_id = {lb: i for i, lb in enumerate(self.lib_names)}

    def lib_of_rg(self, rg: Optional[str]) -> int:
        if rg is None:
            return int(LIB_NONE)
        lb = self.rg_lb.get(rg)
        if lb is None:
            return int(LIB_NONE)
        return self.lib_id[lb]


_AUX_SIZE = {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4, "A": 1}
_AUX_FMT = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}


def _scan_aux(aux: bytes) -> Tuple[Optional[int], Optional[int], Optional[str]]:
    """Return (NM, SM, RG): first occurrence of each, like bam_aux_get's linear scan."""
    nm = sm = rg = None
    i, n = 0, len(aux)
    while i + 3 <= n:
        tag = aux[i:i + 2]
        typ = chr(aux[i + 2])
        i += 3
        if typ in _AUX_SIZE:
            sz = _AUX_SIZE[typ]
            if typ in _AUX_FMT and tag in (b"NM", b"SM"):
                v = struct.unpack_from(_AUX_FMT[typ], aux, i)[0]
                if tag == b"NM" and nm is None:
                    nm = v
                elif tag == b"SM" and sm is None:
                    sm = v
            i += sz
        elif typ in "ZH":
            j = aux.index(b"\0", i)
            if tag == b"RG" and rg is None and typ == "Z":
                rg = aux[i:j].decode()
            i = j + 1
        elif typ == "B":
            sub = chr(aux[i])
            cnt = struct.unpack_from("<I", aux, i + 1)[0]
            i += 5 + cnt * _AUX_SIZE[sub]
        else:
            break
    return nm, sm, rg


def read_bam(path: str) -> Tuple[BamHeader, ReadBatch]:
    """Decode a whole BAM file (BGZF members are gzip members) into one ReadBatch, file order."""
    with open(path, "rb") as fh:
        raw = gzip.decompress(fh.read())
    assert raw[:4] == b"BAM\1", "not a BAM file"
    l_text = struct.unpack_from("<i", raw, 4)[0]
    text = raw[8:8 + l_text].split(b"\0")[0].decode()
    o = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, o)[0]; o += 4
    names, lens = [], []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", raw, o)[0]; o += 4
        names.append(raw[o:o + ln - 1].decode()); o += ln
        lens.append(struct.unpack_from("<i", raw, o)[0]); o += 4
    hdr = BamHeader(text, names, lens)

    tid, pos, flag, mapq, lib, lq, nm, sm = [], [], [], [], [], [], [], []
    cig_parts, seq_parts, qual_parts, qnames = [], [], [], []
    n_raw = len(raw)
    while o + 4 <= n_raw:
        bs = struct.unpack_from("<i", raw, o)[0]; o += 4
        (refid, p, l_rn, mq, _bin, n_cig, fl, l_seq, _nref, _npos, _tlen) = struct.unpack_from("<iiBBHHHiiii", raw, o)
        q = o + 32
        qnames.append(raw[q:q + l_rn - 1].decode()); q += l_rn
        cig = np.frombuffer(raw, dtype="<u4", count=n_cig, offset=q); q += 4 * n_cig
        sq = np.frombuffer(raw, dtype=np.uint8, count=(l_seq + 1) // 2, offset=q); q += (l_seq + 1) // 2
        ql = np.frombuffer(raw, dtype=np.uint8, count=l_seq, offset=q); q += l_seq
        a_nm, a_sm, a_rg = _scan_aux(raw[q:o + bs])
        o += bs
        tid.append(refid); pos.append(p); flag.append(fl); mapq.append(mq); lq.append(l_seq)
        lib.append(hdr.lib_of_rg(a_rg))
        nm.append(int(TAG_ABSENT) if a_nm is None else a_nm)
        sm.append(int(TAG_ABSENT) if a_sm is None else a_sm)
        cig_parts.append(cig); seq_parts.append(sq); qual_parts.append(ql)
    n = len(pos)

    def offs(parts):
        off = np.zeros(n + 1, dtype=np.uint64)
        if n:
            off[1:] = np.cumsum([x.shape[0] for x in parts])
        return off

    def cat(parts, dt):
        return np.concatenate(parts).astype(dt) if n else np.zeros(0, dtype=dt)
    batch = ReadBatch(
        tid=np.array(tid, dtype=np.int32), pos=np.array(pos, dtype=np.int32), flag=np.array(flag, dtype=np.uint16),
        mapq=np.array(mapq, dtype=np.uint8), lib=np.array(lib, dtype=np.uint16), l_qseq=np.array(lq, dtype=np.int32),
        nm=np.array(nm, dtype=np.int32), sm=np.array(sm, dtype=np.int32),
        cigar_off=offs(cig_parts), cigar=cat(cig_parts, np.uint32), seq_off=offs(seq_parts), seq=cat(seq_parts, np.uint8),
        qual_off=offs(qual_parts), qual=cat(qual_parts, np.uint8), qname=qnames)
    return hdr, batch


class Fasta:
    """Indexed FASTA (.fai) reader; ``fetch`` returns raw characters, case preserved (fai_fetch)."""

    def __init__(self, path: str):
        self.path = path
        self.index: Dict[str, Tuple[int, int, int, int]] = {}
        with open(path + ".fai") as fh:
            for line in fh:
                f = line.rstrip("\n").split("\t")
                self.index[f[0]] = (int(f[1]), int(f[2]), int(f[3]), int(f[4]))

    def length(self, name: str) -> int:
        return self.index[name][0]

    def fetch(self, name: str, beg: int = 0, end: Optional[int] = None) -> bytes:
        ln, off, lb, lw = self.index[name]
        end = ln if end is None else min(end, ln)
        beg = max(0, beg)
        if end <= beg:
            return b""
        fo = off + (beg // lb) * lw + beg % lb
        fe = off + ((end - 1) // lb) * lw + (end - 1) % lb + 1
        with open(self.path, "rb") as fh:
            fh.seek(fo)
            raw = fh.read(fe - fo)
        return raw.replace(b"\n", b"").replace(b"\r", b"")
