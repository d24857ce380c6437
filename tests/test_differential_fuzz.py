"""Differential fuzz on FRESH seeds (none of these cases has a committed golden):
  CPU: the oracle against the unmodified reference binary (oracle/_ref, when it has been built) — pins the restatement on
       inputs nobody looked at: random CIGARs (D/N/=/X/H/P, P-then-I, leading deletions), missing NM/SM tags, filtered flags,
       reads without a library, -q/-b/-i/-p/-d.
  GPU: the engine against the oracle on the same cases, raw accumulators bit-for-bit, with and without the deep-site kernel
       forced onto the small tiles."""
import os
import subprocess

import numpy as np
import pytest

import cases
import edge_cases

SEEDS = (101, 102, 103, 104, 105, 106)


def _case(seed):
    rng = np.random.default_rng(seed)
    safe = bool(seed % 2)
    return edge_cases.fuzz_case(seed, L=int(rng.integers(300, 700)), n_reads=int(rng.integers(150, 380)), name=f"fresh{seed}",
                                per_lib_safe=safe, n_libs=int(rng.integers(1, 6)), overhang=bool(seed % 3), force_perlib=not safe)


@pytest.mark.parametrize("seed", SEEDS)
def test_oracle_equals_reference_binary_on_fresh_fuzz(seed, tmp_path):
    from oracle.oracle import REF_SAMTOOLS, have_reference_binary, run_reference_binary
    if not have_reference_binary() or not os.path.exists(REF_SAMTOOLS):
        pytest.skip("oracle/_ref not built (oracle/build_ref.sh needs /root/reference)")
    from bam_readcount_b200 import synth
    case = _case(seed)
    name, L, seq, _ = case["contigs"][0]
    d = str(tmp_path)
    synth.write_fasta(os.path.join(d, "ref.fa"), name, np.frombuffer(seq, dtype=np.uint8))
    synth.write_sam(os.path.join(d, "s.sam"), case["batch"], [(name, L)], n_libs=len(case["lib_names"]))
    subprocess.check_call([REF_SAMTOOLS, "view", "-b", "-o", os.path.join(d, "s.bam"), os.path.join(d, "s.sam")])
    subprocess.check_call([REF_SAMTOOLS, "index", os.path.join(d, "s.bam")])
    with open(os.path.join(d, "sites"), "w") as fh:
        for (_, b1, e1) in case["regions"]:
            fh.write(f"{name}\t{b1}\t{e1}\n")
    for fname, fl in case["flag_sets"].items():
        out, err, rc = run_reference_binary(["-w", "0", "-f", os.path.join(d, "ref.fa")] + cases.flags_to_argv(fl) +
                                            ["-l", os.path.join(d, "sites"), os.path.join(d, "s.bam")])
        assert rc == 0, err[-1500:]
        want, _, _ = cases.run_oracle(case, fl, site_list=True)
        assert want == out, f"{fname}: oracle differs from the reference binary"


@pytest.mark.gpu
@pytest.mark.parametrize("deep", (False, True), ids=("pileup", "deep-forced"))
@pytest.mark.parametrize("seed", SEEDS)
def test_engine_equals_oracle_on_fresh_fuzz(seed, deep, monkeypatch):
    if deep:
        monkeypatch.setenv("BRC_DEEP_MIN_READS", "1")
    case = _case(seed)
    # single-base site-list lines as well as the whole contig: the small tiles are what the deep-site kernel takes
    L = case["contigs"][0][1]
    rng = np.random.default_rng(seed + 1000)
    case = dict(case, regions=[(0, 1, L)] + [(0, int(p), int(p)) for p in np.sort(rng.integers(2, L - 2, 25))])
    for fname, fl in case["flag_sets"].items():
        otext, odump, owarn = cases.run_oracle(case, fl, site_list=True)
        etext, edump, ewarn, _ = cases.run_engine(case, fl, site_list=True, want_dump=True)
        assert edump == odump, fname
        assert etext == otext, fname
        assert (ewarn[0], ewarn[1], ewarn[3]) == owarn, fname
