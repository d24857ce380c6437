"""Wall-clock of the C++ host (brc-readcount) vs the reference binary on the same synthetic BAM (one process each)."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bam_readcount_b200 import synth, build
from oracle.oracle import REF_BIN, REF_SAMTOOLS
L = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
ref_sample = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000
d = tempfile.mkdtemp()
ref = synth.synth_reference(L, 1234); b = synth.synth_reads(ref, 30, seed=1234)
synth.write_fasta(d + "/ref.fa", "chr1", ref)
t0 = time.time(); synth.write_sam(d + "/s.sam", b, [("chr1", L)])
subprocess.check_call([REF_SAMTOOLS, "view", "-b", "-o", d + "/s.bam", d + "/s.sam"]); subprocess.check_call([REF_SAMTOOLS, "index", d + "/s.bam"])
print(f"made BAM: {b.n_reads} reads, {os.path.getsize(d+'/s.bam')/1e6:.1f} MB in {time.time()-t0:.1f}s")
exe = build.build_cli()
env = dict(os.environ, BRC_CLI_TIMING="1")
for out in ("/dev/null", d + "/out.txt"):
    t0 = time.time()
    with open(out, "wb") as fh:
        p = subprocess.run([exe, "-w", "0", "-q", "20", "-b", "20", "-f", d + "/ref.fa", d + "/s.bam", f"chr1:1-{L}"], stdout=fh, stderr=subprocess.PIPE, env=env)
    dt = time.time() - t0
    print(f"brc-readcount -> {out}: {dt:.2f}s  ({L/dt:.3e} positions/s) rc={p.returncode}", " | ".join(p.stderr.decode().strip().splitlines()[-3:]))
t0 = time.time()
with open(d + "/ref_out.txt", "wb") as fh:
    subprocess.run([REF_BIN, "-w", "0", "-q", "20", "-b", "20", "-f", d + "/ref.fa", d + "/s.bam", f"chr1:1-{ref_sample}"], stdout=fh, stderr=subprocess.DEVNULL)
dt = time.time() - t0
print(f"reference binary on the first {ref_sample} bp: {dt:.2f}s ({ref_sample/dt:.3e} positions/s)")
a = open(d + "/ref_out.txt", "rb").read()
bb = open(d + "/out.txt", "rb").read()
print("first", ref_sample, "positions byte-identical:", bb.startswith(a[: a.rfind(b"\n", 0, len(a) - 1) + 1]) if a else None)
