"""The C++ host `brc-readcount`: same command line and STDOUT as bam-readcount.
CPU: flag handling that needs no device.  GPU: the reference's six integration-test command lines
(R:integration-test/bam-readcount_test.py:29-116) against its golden files, through BGZF/BAM/BAI/FASTA
decode -> C ABI -> CUDA kernels -> text emitter."""
import os
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _cli():
    from bam_readcount_b200 import build
    build.build()
    return build.build_cli()


def test_cli_help_and_version_exit_code_1():
    exe = _cli()
    for flag in ("-h", "-v"):
        p = subprocess.run([exe, flag], capture_output=True)
        assert p.returncode == 1          # R:src/exe/bam-readcount/bamreadcount.cpp:467-475
    p = subprocess.run([exe], capture_output=True)
    assert p.returncode == 1 and b"Usage: bam-readcount" in p.stdout


def _write_ref(tmp):
    """ref.fa of contig 21: N everywhere except the window the fixture reads touch (stored in test_bam.npz)."""
    z = np.load(os.path.join(GOLDEN, "test_bam.npz"))
    L, wb = int(z["chrom_len"]), int(z["ref_win_beg"])
    seq = np.full(L, ord("N"), dtype=np.uint8)
    seq[wb:wb + z["ref_win"].shape[0]] = z["ref_win"]
    from bam_readcount_b200 import synth
    synth.write_fasta(os.path.join(tmp, "ref.fa"), "21", seq)
    return os.path.join(tmp, "ref.fa")


@pytest.mark.gpu
@pytest.mark.parametrize("args,bam,golden", [
    (["-w", "1", "-l", "site_list"], "test.bam", "expected_all_lib"),
    (["-w", "1", "-p", "-l", "site_list"], "test.bam", "expected_per_lib"),
    (["-w", "1", "-i", "-l", "site_list"], "test.bam", "expected_insertion_centric_all_lib"),
    (["-w", "1", "-i", "-p", "-l", "site_list"], "test.bam", "expected_insertion_centric_per_lib"),
    (["-w", "1", "REGIONS"], "test.bam", "expected_all_lib"),
    (["-w", "1", "REGIONS"], "test_bad_rg.bam", "expected_all_lib"),
])
def test_cli_reproduces_reference_goldens(tmp_path, args, bam, golden):
    exe = _cli()
    ref = _write_ref(str(tmp_path))
    argv = [exe, "-f", ref]
    regions = []
    for a in args:
        if a == "REGIONS":
            regions = ["21:10402985-10402985", "21:10405200-10405200"]
        elif a == "site_list":
            argv.append(os.path.join(GOLDEN, "site_list"))
        else:
            argv.append(a)
    argv.append(os.path.join(GOLDEN, bam))
    argv += regions
    p = subprocess.run(argv, capture_output=True)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    assert p.stdout.decode("latin-1") == cases.load_golden_text(golden)
    assert b"Minimum mapping quality is set to 0" in p.stderr


@pytest.mark.gpu
def test_cli_synthetic_bam_matches_oracle(tmp_path):
    """A coordinate-sorted synthetic BAM (written with the samtools the oracle build leaves in oracle/_ref) through the
    CLI, whole-contig region and a site list, against the CPU oracle."""
    from oracle.oracle import REF_SAMTOOLS
    if not os.path.exists(REF_SAMTOOLS):
        pytest.skip("oracle/_ref/samtools not built")
    from bam_readcount_b200 import synth
    exe = _cli()
    case = cases.synthetic_case(L=40000, depth=30, seed=31, regions=((0, 1001, 38000),), site_list=False)
    name, L, seq, _ = case["contigs"][0]
    d = str(tmp_path)
    synth.write_fasta(os.path.join(d, "ref.fa"), name, np.frombuffer(seq, dtype=np.uint8))
    synth.write_sam(os.path.join(d, "s.sam"), case["batch"], [(name, L)])
    subprocess.check_call([REF_SAMTOOLS, "view", "-b", "-o", os.path.join(d, "s.bam"), os.path.join(d, "s.sam")])
    subprocess.check_call([REF_SAMTOOLS, "index", os.path.join(d, "s.bam")])
    for fl, argv in ((dict(min_mapq=20, min_bq=20), ["-q", "20", "-b", "20"]), (dict(per_lib=True, insertion_centric=True), ["-p", "-i"])):
        want, _, _ = cases.run_oracle(case, fl, site_list=False)
        p = subprocess.run([exe, "-w", "0", "-f", os.path.join(d, "ref.fa")] + argv + [os.path.join(d, "s.bam"), "chr1:1001-38000"], capture_output=True)
        assert p.returncode == 0, p.stderr.decode()[-2000:]
        assert p.stdout.decode("latin-1") == want


@pytest.mark.gpu
def test_cli_windowed_long_region_equals_unsplit(tmp_path):
    """brc-readcount cuts long regions into windows (bounded memory); the concatenation must equal the unsplit output,
    deletions across window edges included."""
    from oracle.oracle import REF_SAMTOOLS
    if not os.path.exists(REF_SAMTOOLS):
        pytest.skip("oracle/_ref/samtools not built")
    from bam_readcount_b200 import synth
    exe = _cli()
    case = cases.synthetic_case(L=30000, depth=30, seed=41, regions=((0, 1, 30000),), site_list=False)
    name, L, seq, _ = case["contigs"][0]
    d = str(tmp_path)
    synth.write_fasta(os.path.join(d, "ref.fa"), name, np.frombuffer(seq, dtype=np.uint8))
    synth.write_sam(os.path.join(d, "s.sam"), case["batch"], [(name, L)])
    subprocess.check_call([REF_SAMTOOLS, "view", "-b", "-o", os.path.join(d, "s.bam"), os.path.join(d, "s.sam")])
    subprocess.check_call([REF_SAMTOOLS, "index", os.path.join(d, "s.bam")])
    outs = []
    for win in ("100000000", "3777"):
        p = subprocess.run([exe, "-w", "0", "-p", "-f", os.path.join(d, "ref.fa"), os.path.join(d, "s.bam"), "chr1:1-30000"],
                           capture_output=True, env=dict(os.environ, BRC_CLI_WINDOW=win))
        assert p.returncode == 0, p.stderr.decode()[-2000:]
        outs.append(p.stdout)
    assert outs[0] == outs[1] and outs[0].count(b"\n") > 29000


def test_cli_bgzf_bai_region_fetch_matches_python_decoder(tmp_path):
    """CPU: the C++ host's BGZF (multi-threaded read-ahead) / BAM / BAI region fetch, checked against the Python decoder's
    linear scan: same records per region (count, sum of positions, sum of qualities)."""
    exe = _cli()
    from bam_readcount_b200.bamio import read_bam
    bam = os.path.join(GOLDEN, "test.bam")
    hdr, b = read_bam(bam)
    regions = [("21", 10402985, 10402985), ("21", 10405200, 10405200), ("21", 10402700, 10405300), ("21", 1, 10402000), ("21", 10403000, 10403100)]
    sl = tmp_path / "sites"
    sl.write_text("".join(f"{c}\t{s}\t{e}\n" for c, s, e in regions))
    p = subprocess.run([exe, "-l", str(sl), bam], capture_output=True, env=dict(os.environ, BRC_CLI_DECODE_ONLY="1", BRC_CLI_WINDOW="2000000000"))
    assert p.returncode == 0, p.stderr.decode()
    got = [tuple(int(x) for x in line.split("\t")) for line in p.stdout.decode().strip().splitlines()]
    assert len(got) == len(regions)
    qo = b.qual_off.astype(np.int64)
    for (c, s, e), g in zip(regions, got):
        tid = hdr.tid_of[c]
        idx = b.fetch(tid, s - 2, e)          # samfetch(d.beg-1, d.end) with d.beg = s-1
        want = (tid, s - 1, e, len(idx), int(b.pos[idx].astype(np.int64).sum()), int(sum(int(b.qual[qo[i]:qo[i + 1]].astype(np.int64).sum()) for i in idx)))
        assert g == want


def _site_list_regions(L, rng):
    """Sorted dense single sites, overlapping / nested / repeated regions, a backwards jump and far jumps."""
    regs = [(int(p), int(p)) for p in range(2000, 5500, 7)]
    regs += [(6000, 6400), (6100, 6150), (6100, 6150), (6149, 6700), (6700, 6700)]
    regs += [(3000, 3010)]                                     # backwards
    regs += [(int(p), int(p) + int(w)) for p, w in zip(np.sort(rng.integers(7000, L - 500, 150)), rng.integers(0, 40, 150))]
    regs += [(L - 300, L), (1, 50)]
    return regs


def _make_bam(case, d):
    from oracle.oracle import REF_SAMTOOLS
    from bam_readcount_b200 import synth
    name, L, seq, _ = case["contigs"][0]
    synth.write_fasta(os.path.join(d, "ref.fa"), name, np.frombuffer(seq, dtype=np.uint8))
    synth.write_sam(os.path.join(d, "s.sam"), case["batch"], [(name, L)], n_libs=len(case["lib_names"]))
    subprocess.check_call([REF_SAMTOOLS, "view", "-b", "-o", os.path.join(d, "s.bam"), os.path.join(d, "s.sam")])
    subprocess.check_call([REF_SAMTOOLS, "index", os.path.join(d, "s.bam")])
    return os.path.join(d, "s.bam"), os.path.join(d, "ref.fa")


def test_cli_site_list_fetch_merging_yields_samfetch_records(tmp_path):
    """CPU: consecutive site-list lines share one forward pass over the BAM instead of one index seek
    each; every region must still receive exactly the records samfetch yields.  Checked against the per-region seek path
    (BRC_CLI_NO_MERGE) and against the Python decoder."""
    from oracle.oracle import REF_SAMTOOLS
    if not os.path.exists(REF_SAMTOOLS):
        pytest.skip("oracle/_ref/samtools not built")
    exe = _cli()
    case = cases.synthetic_case(L=60000, depth=30, seed=77, regions=((0, 1, 60000),), site_list=True)
    bam, _ = _make_bam(case, str(tmp_path))
    regs = _site_list_regions(60000, np.random.default_rng(5))
    sl = tmp_path / "sites"
    sl.write_text("".join(f"chr1\t{s}\t{e}\n" for s, e in regs))
    outs, stats = [], []
    for extra in ({}, {"BRC_CLI_NO_MERGE": "1"}):
        p = subprocess.run([exe, "-l", str(sl), bam], capture_output=True, env=dict(os.environ, BRC_CLI_DECODE_ONLY="1", BRC_CLI_TIMING="1", **extra))
        assert p.returncode == 0, p.stderr.decode()
        outs.append(p.stdout.decode())
        line = [ln for ln in p.stderr.decode().splitlines() if "index seeks" in ln][0].split()
        stats.append((int(line[4]), int(line[7])))
    assert outs[0] == outs[1]
    assert stats[0][0] < 20 and stats[1][0] == len(regs)          # a handful of seeks instead of one per line
    assert stats[0][1] * 20 < stats[1][1]                          # and far fewer records decoded
    b = case["batch"]
    qo = b.qual_off.astype(np.int64)
    got = [tuple(int(x) for x in line.split("\t")) for line in outs[0].strip().splitlines()]
    for (s, e), g in zip(regs, got):
        idx = b.fetch(0, max(s - 2, 0), e)
        want = (0, s - 1, e, len(idx), int(b.pos[idx].astype(np.int64).sum()), int(sum(int(b.qual[qo[i]:qo[i + 1]].astype(np.int64).sum()) for i in idx)))
        assert g == want, (s, e)


@pytest.mark.gpu
def test_cli_dense_site_list_matches_oracle(tmp_path):
    """A few hundred site-list lines (dense, overlapping, nested, repeated, out of order) through the merged fetch, the
    engine and the emitter, against the CPU oracle run region by region like the reference's -l loop."""
    from oracle.oracle import REF_SAMTOOLS
    if not os.path.exists(REF_SAMTOOLS):
        pytest.skip("oracle/_ref/samtools not built")
    exe = _cli()
    regs = _site_list_regions(60000, np.random.default_rng(5))
    case = cases.synthetic_case(L=60000, depth=30, seed=77, regions=tuple((0, s, e) for s, e in regs), site_list=True)
    bam, ref = _make_bam(case, str(tmp_path))
    sl = tmp_path / "sites"
    sl.write_text("".join(f"chr1\t{s}\t{e}\n" for s, e in regs))
    for fl, argv in ((dict(min_mapq=20, min_bq=20), ["-q", "20", "-b", "20"]), (dict(per_lib=True), ["-p"])):
        want, _, _ = cases.run_oracle(case, fl, site_list=True)
        outs = []
        for extra in ({}, {"BRC_CLI_NO_MERGE": "1"}):
            p = subprocess.run([exe, "-w", "0", "-f", ref] + argv + ["-l", str(sl), bam], capture_output=True, env=dict(os.environ, **extra))
            assert p.returncode == 0, p.stderr.decode()[-2000:]
            outs.append(p.stdout.decode("latin-1"))
        assert outs[0] == outs[1]
        assert outs[0] == want


def test_cli_fetch_merging_two_contigs_unsorted_and_past_the_end(tmp_path):
    """CPU: the merged fetch across contig changes, lines that go backwards, duplicates, a line past the contig's last read
    and an unknown contig — record sets identical to the one-seek-per-line path and to the Python decoder."""
    from oracle.oracle import REF_SAMTOOLS
    if not os.path.exists(REF_SAMTOOLS):
        pytest.skip("oracle/_ref/samtools not built")
    import dataclasses
    from bam_readcount_b200 import synth
    from bam_readcount_b200.batch import ReadBatch
    exe = _cli()
    a = cases.synthetic_case(L=30000, depth=20, seed=3, regions=((0, 1, 30000),), site_list=True)["batch"]
    b = cases.synthetic_case(L=20000, depth=25, seed=4, regions=((0, 1, 20000),), site_list=True)["batch"]
    b = dataclasses.replace(b, tid=np.ones_like(b.tid))
    both = ReadBatch.concat([a, b])
    d = str(tmp_path)
    synth.write_sam(os.path.join(d, "s.sam"), both, [("chrA", 30000), ("chrB", 20000)])
    subprocess.check_call([REF_SAMTOOLS, "view", "-b", "-o", os.path.join(d, "s.bam"), os.path.join(d, "s.sam")])
    subprocess.check_call([REF_SAMTOOLS, "index", os.path.join(d, "s.bam")])
    lines = [("chrA", 100, 100), ("chrA", 101, 130), ("chrB", 5000, 5000), ("chrB", 5001, 5001), ("chrA", 120, 125), ("chrA", 120, 125),
             ("chrB", 19990, 25000), ("chrA", 29999, 30000), ("chrZ", 5, 6), ("chrB", 1, 1), ("chrB", 2, 2), ("chrB", 3, 400)]
    sl = tmp_path / "sites"
    sl.write_text("".join(f"{c}\t{s}\t{e}\n" for c, s, e in lines))
    outs = []
    for extra in ({}, {"BRC_CLI_NO_MERGE": "1"}):
        p = subprocess.run([exe, "-l", str(sl), os.path.join(d, "s.bam")], capture_output=True, env=dict(os.environ, BRC_CLI_DECODE_ONLY="1", **extra))
        assert p.returncode == 0, p.stderr.decode()
        outs.append(p.stdout.decode())
        assert b"chrZ not found in bam file" in p.stderr
    assert outs[0] == outs[1]
    got = [tuple(int(x) for x in ln.split("\t")) for ln in outs[0].strip().splitlines()]
    known = [ln for ln in lines if ln[0] != "chrZ"]
    assert len(got) == len(known)
    qo = both.qual_off.astype(np.int64)
    for (c, s, e), g in zip(known, got):
        tid = 0 if c == "chrA" else 1
        idx = both.fetch(tid, max(s - 2, 0), e)
        want = (tid, s - 1, e, len(idx), int(both.pos[idx].astype(np.int64).sum()), int(sum(int(both.qual[qo[i]:qo[i + 1]].astype(np.int64).sum()) for i in idx)))
        assert g == want, (c, s, e)
