"""Enumerates every (case, flag set) that has a committed reference-binary golden under tests/golden/."""
from __future__ import annotations

import functools

import cases
import edge_cases


@functools.lru_cache(maxsize=None)
def _syn():
    return cases.synthetic_case(L=12000, depth=30, seed=11, regions=((0, 1000, 4000),), site_list=False)


@functools.lru_cache(maxsize=None)
def _deep():
    return cases.deep_case(n_sites=3, depth=5000, seed=5)


@functools.lru_cache(maxsize=None)
def _edge():
    return {c["name"]: c for c in edge_cases.all_cases()}


@functools.lru_cache(maxsize=None)
def _testbam(bad):
    return cases.testbam_case(bad_rg=bad)


def jobs():
    """[(id, case_getter, flags, site_list, golden_file)]"""
    out = []
    for fname, fl in cases.FLAG_SETS.items():
        out.append((f"syn-{fname}", _syn, fl, False, f"ref_syn_{fname}.txt"))
    out.append(("deep-perlib", _deep, dict(per_lib=True, max_cnt=100000000), True, "ref_deep_perlib.txt"))
    out.append(("deep-alllib", _deep, dict(max_cnt=100000000), True, "ref_deep_alllib.txt"))
    for name, c in _edge().items():
        for fname, fl in c["flag_sets"].items():
            out.append((f"edge-{name}-{fname}", (lambda n=name: _edge()[n]), fl, c["site_list"], f"edge_{name}_{fname}.txt"))
    # the reference's own integration tests (R:integration-test/bam-readcount_test.py:29-116)
    tb = lambda: _testbam(False)
    out.append(("testbam-all", tb, dict(), True, "expected_all_lib"))
    out.append(("testbam-perlib", tb, dict(per_lib=True), True, "expected_per_lib"))
    out.append(("testbam-ic", tb, dict(insertion_centric=True), True, "expected_insertion_centric_all_lib"))
    out.append(("testbam-ic-perlib", tb, dict(insertion_centric=True, per_lib=True), True, "expected_insertion_centric_per_lib"))
    out.append(("testbam-argv-regions", tb, dict(), False, "expected_all_lib"))
    out.append(("testbam-badrg-argv-regions", lambda: _testbam(True), dict(), False, "expected_all_lib"))
    return out
