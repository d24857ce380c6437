"""Wide differential fuzz of the CPU oracle against the unmodified reference binary (needs oracle/_ref; CPU only).
usage: fuzz_oracle_vs_reference.py [first_seed] [n_seeds] [argv]
  site-list mode (default): round 1, seeds 200..419: 1320 runs, 0 mismatches
  argv: adjacent / overlapping command-line regions in arbitrary order (the never-cleared deletion queue): seeds 500..599,
        600 runs, 0 mismatches"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases, edge_cases
from oracle.oracle import REF_SAMTOOLS, run_reference_binary
from bam_readcount_b200 import synth

first = int(sys.argv[1]) if len(sys.argv) > 1 else 200
count = int(sys.argv[2]) if len(sys.argv) > 2 else 220
argv_mode = len(sys.argv) > 3 and sys.argv[3] == "argv"
bad = n = 0
t0 = time.time()
for seed in range(first, first + count):
    rng = np.random.default_rng(seed)
    safe = bool(seed % 2)
    case = edge_cases.fuzz_case(seed, L=int(rng.integers(300, 900)), n_reads=int(rng.integers(100, 500)), name=f"s{seed}", per_lib_safe=safe,
                                n_libs=int(rng.integers(1, 7)), overhang=bool(seed % 3), force_perlib=not safe)
    name, L, seq, _ = case["contigs"][0]
    d = tempfile.mkdtemp()
    synth.write_fasta(d + "/ref.fa", name, np.frombuffer(seq, dtype=np.uint8))
    synth.write_sam(d + "/s.sam", case["batch"], [(name, L)], n_libs=len(case["lib_names"]))
    subprocess.check_call([REF_SAMTOOLS, "view", "-b", "-o", d + "/s.bam", d + "/s.sam"])
    subprocess.check_call([REF_SAMTOOLS, "index", d + "/s.bam"])
    if argv_mode:
        a = int(rng.integers(20, L - 120)); b = int(rng.integers(1, L - 5))
        regs = [(0, a, a + 30), (0, a + 31, a + 31), (0, a + 32, a + 60), (0, a + 50, a + 70), (0, b, b + 3), (0, a + 10, a + 12)]
    else:
        regs = [(0, 1, L)] + [(0, int(a), int(a) + int(w)) for a, w in zip(rng.integers(1, L - 40, 6), rng.integers(0, 30, 6))]
    case = dict(case, regions=regs)
    with open(d + "/sites", "w") as fh:
        fh.write("".join(f"{name}\t{b}\t{e}\n" for _, b, e in regs))
    for fname, fl in case["flag_sets"].items():
        tail = [d + "/s.bam"] + [f"{name}:{b}-{e}" for _, b, e in regs] if argv_mode else ["-l", d + "/sites", d + "/s.bam"]
        out, err, rc = run_reference_binary(["-w", "0", "-f", d + "/ref.fa"] + cases.flags_to_argv(fl) + tail)
        want, _, _ = cases.run_oracle(case, fl, site_list=not argv_mode)
        n += 1
        if rc != 0 or want != out:
            bad += 1
            print("MISMATCH seed", seed, fname, "rc", rc)
    subprocess.call(["rm", "-rf", d])
print(f"runs {n}  mismatches {bad}  in {time.time() - t0:.1f}s")
sys.exit(1 if bad else 0)
