"""ctypes binding of libbrc_synth.so (include/brc_synth.h): the counter-based generator of the C3/C4/C5 workloads.

``Spec`` names a workload; ``window_host`` returns a window as a numpy ``ReadBatch`` (what the oracle and the reference
binary see), ``DeviceWindow`` is a reusable set of device buffers that ``fill`` regenerates in HBM — byte-identical to
the host copy, because both run the same integer code.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .batch import ReadBatch
from .engine import CReadBatch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbrc_synth.so")

BLOCK_BP = 1280
BLOCK_READS = 256
READ_LEN = 150
MAX_SPAN = 153
WGS, DEEP = 0, 1


class CSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("mode", C.c_int32), ("n_libs", C.c_int32), ("contig_len", C.c_int64),
                ("depth", C.c_int32), ("site_stride", C.c_int32)]


class COut(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("tid", C.c_void_p), ("pos", C.c_void_p), ("flag", C.c_void_p), ("mapq", C.c_void_p),
                ("lib", C.c_void_p), ("l_qseq", C.c_void_p), ("nm", C.c_void_p), ("sm", C.c_void_p), ("cigar_off", C.c_void_p),
                ("cigar", C.c_void_p), ("seq_off", C.c_void_p), ("seq", C.c_void_p), ("qual_off", C.c_void_p), ("qual", C.c_void_p),
                ("region_of_read", C.c_void_p)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: python -m bam_readcount_b200.build")
        lib = C.CDLL(LIB_PATH)
        lib.brc_synth_window_reads.restype = C.c_int64
        lib.brc_synth_window_reads.argtypes = [C.POINTER(CSpec), C.c_int64, C.c_int64]
        lib.brc_synth_ref_host.argtypes = [C.POINTER(CSpec), C.c_int32, C.c_int64, C.c_int64, C.c_void_p]
        lib.brc_synth_ref_device.argtypes = [C.POINTER(CSpec), C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        lib.brc_synth_fill_host.argtypes = [C.POINTER(CSpec), C.c_int32, C.c_int64, C.c_int64, C.POINTER(COut), C.c_int]
        lib.brc_synth_fill_device.argtypes = [C.POINTER(CSpec), C.c_int32, C.c_int64, C.c_int64, C.POINTER(COut), C.c_void_p, C.c_void_p]
        lib.brc_synth_write_sam.argtypes = [C.POINTER(CSpec), C.c_int32, C.c_int64, C.c_int64, C.c_char_p, C.c_char_p, C.c_int64, C.c_int]
        lib.brc_synth_checksum_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


class Spec:
    """A synthetic workload.  WGS: ``n_contigs`` contigs of ``contig_len`` bp at 30x; DEEP: panel sites at
    ``500 + k * site_stride`` of contig 0 under ``depth`` reads each."""

    def __init__(self, seed=1234, mode=WGS, n_libs=8, contig_len=10_000_000 // BLOCK_BP * BLOCK_BP, n_contigs=1, depth=50_000,
                 site_stride=1000, n_sites=0):
        assert contig_len % BLOCK_BP == 0
        self.c = CSpec(seed, mode, n_libs, contig_len, depth, site_stride)
        self.mode, self.n_libs, self.contig_len, self.n_contigs = mode, n_libs, contig_len, n_contigs
        self.depth, self.site_stride, self.n_sites = depth, site_stride, n_sites

    # ---- geometry -------------------------------------------------------------------------------
    def blocks_per_contig(self) -> int:
        return self.contig_len // BLOCK_BP

    def window_reads(self, lo: int, hi: int) -> int:
        return int(load().brc_synth_window_reads(C.byref(self.c), lo, hi))

    def site_pos(self, k: int) -> int:
        return 500 + k * self.site_stride

    def deep_contig_len(self) -> int:
        return 500 + self.n_sites * self.site_stride + 500

    # ---- host generation --------------------------------------------------------------------------
    def ref_host(self, contig: int, beg: int, length: int) -> bytes:
        buf = np.empty(length, dtype=np.uint8)
        rc = load().brc_synth_ref_host(C.byref(self.c), contig, beg, length, buf.ctypes.data)
        assert rc == 0, rc
        return buf.tobytes()

    def window_host(self, contig: int, lo: int, hi: int, threads: int = 0) -> tuple:
        """(ReadBatch, region_of_read) of blocks / sites [lo, hi) in file order."""
        n = self.window_reads(lo, hi)
        a = dict(tid=np.empty(n, np.int32), pos=np.empty(n, np.int32), flag=np.empty(n, np.uint16), mapq=np.empty(n, np.uint8),
                 lib=np.empty(n, np.uint16), l_qseq=np.empty(n, np.int32), nm=np.empty(n, np.int32), sm=np.empty(n, np.int32),
                 cigar_off=np.zeros(n + 1, np.uint64), cigar=np.zeros(3 * n + 16, np.uint32), seq_off=np.zeros(n + 1, np.uint64),
                 seq=np.zeros(75 * n + 64, np.uint8), qual_off=np.zeros(n + 1, np.uint64), qual=np.zeros(150 * n + 64, np.uint8))
        ror = np.zeros(n, np.int32)
        out = COut(n, *[a[k].ctypes.data for k in ("tid", "pos", "flag", "mapq", "lib", "l_qseq", "nm", "sm", "cigar_off", "cigar",
                                                    "seq_off", "seq", "qual_off", "qual")], ror.ctypes.data)
        rc = load().brc_synth_fill_host(C.byref(self.c), contig, lo, hi, C.byref(out), threads or min(32, os.cpu_count() or 1))
        assert rc == 0, rc
        nc = int(a["cigar_off"][n]) if n else 0
        a["cigar"] = a["cigar"][:nc]
        a["seq"] = a["seq"][:75 * n]
        a["qual"] = a["qual"][:150 * n]
        return ReadBatch(qname=None, **a), ror


def write_sample_bam(spec: "Spec", contig: int, lo: int, hi: int, workdir: str, samtools: str, contig_name: str = "chr1") -> dict:
    """ref.fa (+.fai) and s.bam (+.bai) of blocks / sites [lo, hi): the window as files, for the reference binary and the CLI.
    The FASTA covers the contig from 0 to the end of the window (+ 400 bp), which is also the @SQ length."""
    from . import synth
    end = (hi * BLOCK_BP if spec.mode == WGS else spec.site_pos(hi)) + 400
    if spec.mode == WGS:
        end = min(end, spec.contig_len)
    ref = np.frombuffer(spec.ref_host(contig, 0, end), dtype=np.uint8)
    fa = os.path.join(workdir, "ref.fa")
    synth.write_fasta(fa, contig_name, ref)
    sam = os.path.join(workdir, "s.sam")
    rc = load().brc_synth_write_sam(C.byref(spec.c), contig, lo, hi, sam.encode(), contig_name.encode(), end, min(16, os.cpu_count() or 1))
    assert rc == 0, rc
    import subprocess
    bam = os.path.join(workdir, "s.bam")
    subprocess.check_call([samtools, "view", "-@", "8", "-b", "-o", bam, sam])
    subprocess.check_call([samtools, "index", bam])
    os.remove(sam)
    return dict(fasta=fa, bam=bam, length=end, contig=contig_name)


class DeviceWindow:
    """Device buffers for windows of up to ``max_reads`` reads (torch owns the memory)."""

    def __init__(self, spec: Spec, max_reads: int, device, scratch=None, with_region: bool = True):
        import torch
        self.spec, self.cap = spec, int(max_reads)
        n = self.cap
        z = lambda m, dt: torch.empty(m, dtype=dt, device=device)   # noqa: E731
        self.t = dict(pos=z(n, torch.int32), flag=z(n, torch.int16), mapq=z(n, torch.uint8), lib=z(n, torch.int16), l_qseq=z(n, torch.int32),
                      nm=z(n, torch.int32), sm=z(n, torch.int32), cigar_off=z(n + 1, torch.int64), cigar=z(3 * n + 16, torch.int32),
                      seq_off=z(n + 1, torch.int64), seq=z(75 * n + 64, torch.uint8), qual_off=z(n + 1, torch.int64),
                      qual=z(150 * n + 64, torch.uint8), region=z(n if with_region else 1, torch.int32),
                      scratch=scratch if scratch is not None else z(n // BLOCK_READS + 4, torch.int64))
        self.with_region = with_region
        self.n_reads = 0

    @staticmethod
    def bytes_per_read() -> int:
        return 4 + 2 + 1 + 2 + 4 + 4 + 4 + 8 + 12 + 8 + 75 + 8 + 150

    def fill(self, contig: int, lo: int, hi: int, stream_ptr: int) -> int:
        """Enqueue the generator kernels for blocks / sites [lo, hi) on the stream; returns the read count."""
        n = self.spec.window_reads(lo, hi)
        assert n <= self.cap, (n, self.cap)
        t = self.t
        out = COut(n, None, *[t[k].data_ptr() for k in ("pos", "flag", "mapq", "lib", "l_qseq", "nm", "sm", "cigar_off", "cigar", "seq_off",
                                                         "seq", "qual_off", "qual")], t["region"].data_ptr() if self.with_region else None)
        rc = load().brc_synth_fill_device(C.byref(self.spec.c), contig, lo, hi, C.byref(out), t["scratch"].data_ptr(), stream_ptr)
        assert rc == 0, rc
        self.n_reads = n
        return n

    def c_batch(self) -> CReadBatch:
        t = self.t
        return CReadBatch(self.n_reads, None, t["pos"].data_ptr(), t["flag"].data_ptr(), t["mapq"].data_ptr(), t["lib"].data_ptr(),
                          t["l_qseq"].data_ptr(), t["nm"].data_ptr(), t["sm"].data_ptr(), t["cigar_off"].data_ptr(), t["cigar"].data_ptr(),
                          t["seq_off"].data_ptr(), t["seq"].data_ptr(), t["qual_off"].data_ptr(), t["qual"].data_ptr())

    def to_host(self) -> ReadBatch:
        """Copy the current window back (tests: device == host generator)."""
        n = self.n_reads
        t = {k: v.cpu().numpy() for k, v in self.t.items()}
        nc = int(t["cigar_off"][n]) if n else 0
        return ReadBatch(tid=np.zeros(n, np.int32), pos=t["pos"][:n], flag=t["flag"][:n].view(np.uint16), mapq=t["mapq"][:n],
                         lib=t["lib"][:n].view(np.uint16), l_qseq=t["l_qseq"][:n], nm=t["nm"][:n], sm=t["sm"][:n],
                         cigar_off=t["cigar_off"][:n + 1].view(np.uint64), cigar=t["cigar"][:nc].view(np.uint32),
                         seq_off=t["seq_off"][:n + 1].view(np.uint64), seq=t["seq"][:75 * n], qual_off=t["qual_off"][:n + 1].view(np.uint64),
                         qual=t["qual"][:150 * n], qname=None)


def checksum_device(tensor, acc, stream_ptr: int) -> None:
    """acc (int64 device tensor of 1 element) += checksum of ``tensor``'s bytes."""
    nbytes = tensor.numel() * tensor.element_size()
    rc = load().brc_synth_checksum_device(tensor.data_ptr(), nbytes, acc.data_ptr(), stream_ptr)
    assert rc == 0, rc
