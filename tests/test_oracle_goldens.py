"""Pins the CPU oracle against the reference's own golden PNGs.

Every case is a block of the reference's tests or an examples/*.nim program (see golden_cases.py);
the oracle in canonical (x86 row-kernel) semantics must reproduce the golden byte for byte, and
the scalar (-d:pixieNoSimd) semantics may differ by at most 1 LSB per channel."""
import pytest

import golden_cases as gc
from _oracle import OracleBackend


@pytest.mark.parametrize("name", sorted(gc.CASES))
def test_oracle_reproduces_golden(name):
    out = gc.CASES[name](OracleBackend(sem=0))
    if name in gc.SCORE_ONLY:  # masters the reference compares by score only (see golden_cases.SCORE_ONLY)
        score = gc.xray_score(out, gc.load_golden(name))
        assert score < gc.SCORE_ONLY[name], f"{name}: xray score {score}"
        return
    mism, mx = gc.compare(out, gc.load_golden(name))
    assert (mism, mx) == (0, 0), f"{name}: {mism} pixels differ, max |delta| {mx}"


@pytest.mark.parametrize("name", sorted(gc.CASES))
def test_scalar_semantics_within_one_lsb(name):
    if name in gc.SCORE_ONLY:
        pytest.skip("score-only master")
    out = gc.CASES[name](OracleBackend(sem=1))
    mism, mx = gc.compare(out, gc.load_golden(name))
    assert mx <= 1, f"{name}: scalar semantics differ by {mx} LSB"
    assert mism <= 0.05 * out.shape[0] * out.shape[1]
