#!/bin/bash
# tools/gpu_cli_timing2.sh <tag> [blocks] — phase times of the C++ host on a generated BAM (no tests)
TAG=${1:-r02t2}; NB=${2:-8000}; O=gpurun_out/$TAG; mkdir -p $O /tmp/brc_txt
python - <<PY > $O/gen.log 2>&1
import sys, time, argparse
sys.path.insert(0, ".")
import bench
from bam_readcount_b200 import synth_cb
from oracle.oracle import REF_SAMTOOLS
spec = bench.make_spec("c4", argparse.Namespace(contigs=None, contig_blocks=None, c5_sites=None, c5_depth=None, c5_sites_per_window=None))
t = time.time(); info = synth_cb.write_sample_bam(spec, 0, 0, $NB, "/tmp/brc_txt", REF_SAMTOOLS); print(info, "written in %.1f s" % (time.time() - t))
PY
tail -1 $O/gen.log
N=$(( NB * 1280 ))
run() { name=$1; shift
  for rep in 1 2 3; do
    s=$(date +%s.%N)
    env "$@" BRC_CLI_TIMING=1 bam_readcount_b200/brc-readcount -w 0 -i -f /tmp/brc_txt/ref.fa /tmp/brc_txt/s.bam chr1:1-$N > /dev/null 2> $O/cli_${name}_$rep.err
    e=$(date +%s.%N); echo "$name rep $rep: wall $(python -c "print('%.3f' % ($e - $s))") s for $N bp"; grep "brc timing\] \(reference\|startup\)" $O/cli_${name}_$rep.err
  done
}
run par X=1
run clean BRC_CLI_CLEAN_EXIT=1
run lazy CUDA_MODULE_LOADING=LAZY
s=$(date +%s.%N); bam_readcount_b200/brc-readcount -h > /dev/null 2>&1; e=$(date +%s.%N); echo "-h (no CUDA): $(python -c "print('%.3f' % ($e - $s))") s"
ldd bam_readcount_b200/brc-readcount | head -20
