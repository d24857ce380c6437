#!/bin/bash
# tools/gpu_cli_timing.sh <tag> [blocks] — where the C++ host spends its time on a generated BAM (BRC_CLI_TIMING)
TAG=${1:-r02t}; NB=${2:-8000}; O=gpurun_out/$TAG; mkdir -p $O
python - <<PY > $O/gen.log 2>&1
import sys, time, argparse
sys.path.insert(0, ".")
import bench
from bam_readcount_b200 import synth_cb
from oracle.oracle import REF_SAMTOOLS
spec = bench.make_spec("c4", argparse.Namespace(contigs=None, contig_blocks=None, c5_sites=None, c5_depth=None, c5_sites_per_window=None))
t = time.time(); info = synth_cb.write_sample_bam(spec, 0, 0, $NB, "/tmp/brc_txt", REF_SAMTOOLS); print(info, "written in %.1f s" % (time.time() - t))
PY
cat $O/gen.log
N=$(( NB * 1280 ))
for rep in 1 2; do
  /usr/bin/env time -v true > /dev/null 2>&1
  s=$(date +%s.%N)
  BRC_CLI_TIMING=1 bam_readcount_b200/brc-readcount -w 0 -i -f /tmp/brc_txt/ref.fa /tmp/brc_txt/s.bam chr1:1-$N > /dev/null 2> $O/cli_$rep.err
  e=$(date +%s.%N); echo "rep $rep: wall $(python -c "print('%.3f' % ($e - $s))") s for $N bp"; grep "brc timing" $O/cli_$rep.err
done
s=$(date +%s.%N); bam_readcount_b200/brc-readcount -w 0 -i -f /tmp/brc_txt/ref.fa /tmp/brc_txt/s.bam chr1:1-$N 2>/dev/null | wc -c; e=$(date +%s.%N); echo "to wc: $(python -c "print('%.3f' % ($e - $s))") s"
