"""CPU: the oracle restatement (oracle/brc_oracle.c) must reproduce, byte for byte, every
golden the reference provides (its four expected_* files through the six integration-test
command lines) and every committed reference-binary output (tests/golden/*.txt.gz)."""
import pytest

import cases
import golden_jobs

JOBS = golden_jobs.jobs()


@pytest.mark.parametrize("job", JOBS, ids=[j[0] for j in JOBS])
def test_oracle_matches_reference_golden(job):
    _, getter, flags, site_list, golden = job
    case = getter()
    text, dump, warn = cases.run_oracle(case, flags, site_list=site_list)
    want = cases.load_golden_text(golden)
    assert text == want
    assert dump.count("\nS ") + dump.startswith("S ") >= len(want.splitlines())


def test_zero_count_print_form():
    """R:test/lib/bamrc/TestIndelQueueEntry.cpp:24-34 pins the all-zero block."""
    case = cases.testbam_case()
    text, _, _ = cases.run_oracle(case, dict(), site_list=True)
    assert "\tC:0:0.00:0.00:0.00:0:0:0.00:0.00:0.00:0:0.00:0.00:0.00\t" in text
