"""Config-5 shape (ultra-deep panel): N single-base regions x DEPTH reads x 8 libraries, -p -d 100000000.
Times the push path and the kernels, and checks a few sites bit-for-bit against the CPU oracle.
usage: deep_panel.py [n_sites] [depth] [alllib]
  BRC_DEEP_MIN_READS=2147483647   pileup_kernel only (deep_site_kernel off)
  BRC_ENGINE_LIB=bam_readcount_b200/libbrc_engine_prof.so   (python -c "from bam_readcount_b200 import build; build.build_profile_variant()")
                                  prints deep_site_kernel's cycle profile per block of 256 reads"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from bam_readcount_b200 import synth
from bam_readcount_b200.batch import ReadBatch
from bam_readcount_b200.engine import Engine, admitted

n_sites = int(sys.argv[1]) if len(sys.argv) > 1 else 200
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
L = 400 * (n_sites + 2)
ref = synth.synth_reference(L, 99)
sites = (np.arange(n_sites) + 1) * 400
t0 = time.time()
batch, bounds = synth.synth_deep_panel(ref, sites, depth, seed=7, n_libs=8)
print(f"generated {batch.n_reads} reads in {time.time()-t0:.1f}s")
libs = [f"lib{i}" for i in range(8)]
flags = dict(per_lib=(len(sys.argv) <= 3 or sys.argv[3] != 'alllib'), max_cnt=100000000)
eng = Engine(lib_names=libs, **flags)
eng.set_reference(0, "chr1", L, ref.tobytes(), 0)
t0 = time.time()
subs = []
for i, s in enumerate(sites):
    sub = batch.select(np.arange(bounds[i], bounds[i + 1]))
    subs.append(sub)
t_sel = time.time() - t0
t0 = time.time()
for i, s in enumerate(sites):
    eng.begin_region(0, int(s), int(s) + 1, True)
    eng.push_reads(subs[i])
    eng.end_region()
t_push = time.time() - t0
t0 = time.time()
res = eng.compute()
t_comp = time.time() - t0
k0, k1 = eng.stage_ms(0), eng.stage_ms(1)
ev = int(res.ncover.sum())
print(f"push {t_push:.2f}s compute {t_comp:.2f}s  K0 {k0:.2f} ms  K1 {k1:.2f} ms  events {ev}  -> {ev/((k0+k1)/1e3):.3e} events/s (kernels), sites {n_sites}")
if hasattr(eng.lib, "brc_debug_deepprof"):
    import ctypes as C
    out = (C.c_ulonglong * 8)()
    eng.lib.brc_debug_deepprof(out, 1)
    v = list(out); nb = max(1, v[3])
    print("deep kernel, cycles per block of 256 reads (thread 0): phase1a+barrier %.0f  phase1b %.0f  partition+scatter %.0f  phase2(thread0) %.0f  wait-for-owners %.0f  blocks %d" % (v[0] / nb, v[1] / nb, v[2] / nb, v[4] / nb, v[5] / nb, v[3]))
text = eng.format_text()
# oracle on the first 2 and last site
import cases
from oracle.oracle import Oracle
chk = [0, 1, n_sites - 1][: min(3, n_sites)]
o = Oracle(lib_names=libs, **flags)
want = []
for i in chk:
    s = int(sites[i])
    o.region(subs[i], tid=0, beg=s, end=s + 1, contig="chr1", chrom_len=L, ref_seq=ref.tobytes(), site_list_mode=True)
lines = text.splitlines()
got = "\n".join(lines[i] for i in chk) + "\n"
print("oracle match on sites", chk, ":", got == o.text())
eng.close()
