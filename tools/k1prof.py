"""Debug: build with -DBRC_K1_PROFILE and print where K1's consumer / producer cycles go."""
import ctypes as C, os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["BRC_NVCC_EXTRA"] = "-DBRC_K1_PROFILE " + os.environ.get("BRC_NVCC_EXTRA", "")
from bam_readcount_b200 import build
build.build(force=True)
import numpy as np, torch
from bam_readcount_b200 import synth
from bam_readcount_b200.engine import Engine, CReadBatch, CRegion
L = 10_000_000
ref = synth.synth_reference(L, 1234); batch = synth.synth_reads(ref, 30, seed=1234); n = batch.n_reads
eng = Engine(min_mapq=20, min_bq=20); eng.set_reference(0, "chr1", L, ref.tobytes(), 0)
def dev(a):
    a = np.ascontiguousarray(a)
    v = {np.dtype(np.uint16): np.int16, np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64}.get(a.dtype)
    return torch.from_numpy(a.view(v) if v else a).cuda()
pad = np.zeros(64, np.uint8)
d = dict(pos=dev(batch.pos), flag=dev(batch.flag), mapq=dev(batch.mapq), lib=dev(batch.lib), l_qseq=dev(batch.l_qseq), nm=dev(batch.nm), sm=dev(batch.sm),
         cigar_off=dev(batch.cigar_off), cigar=dev(np.concatenate([batch.cigar, np.zeros(16, np.uint32)])), seq_off=dev(batch.seq_off),
         seq=dev(np.concatenate([batch.seq, pad])), qual_off=dev(batch.qual_off), qual=dev(np.concatenate([batch.qual, pad])))
cb = CReadBatch(n, None, *[d[k].data_ptr() for k in ("pos", "flag", "mapq", "lib", "l_qseq", "nm", "sm", "cigar_off", "cigar", "seq_off", "seq", "qual_off", "qual")])
reg = CRegion(0, 0, L, 0, 0, n, 0, 0, L)
eng._check(eng.lib.brc_plan_device(eng.h, C.byref(reg), 1, n, 0))
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = (C.c_ulonglong * 8)()
for it in range(3):
    eng._check(eng.lib.brc_run_device(eng.h, C.byref(cb), None, sp))
    eng.lib.brc_debug_k1prof(out, 1)
v = list(out)
warps = 444 * 8
print("k1 ms", eng.stage_ms(1))
print("per consumer warp: wait cycles %.0f  busy cycles %.0f  (wait share %.1f%%)" % (v[0] / warps, v[1] / warps, 100 * v[0] / max(1, v[0] + v[1])))
print("per producer: wait-empty cycles %.0f  fill cycles %.0f  items %d  fill cycles/item %.0f" % (v[2] / 444, v[3] / 444, v[4], v[3] / max(1, v[4])))
