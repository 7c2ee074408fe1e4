"""GPU parity on the reference's golden cases: CUDA path (through the C ABI) == oracle == golden."""
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(gc.CASES))
def test_gpu_matches_oracle_and_golden(name):
    from _gpu_backend import GpuBackend
    from _oracle import OracleBackend

    gb, ob = GpuBackend(), OracleBackend(0)
    got = gc.CASES[name](gb)
    want = gc.CASES[name](ob)
    mism, mx = gc.compare(got, want)
    assert (mism, mx) == (0, 0), f"{name}: GPU vs oracle {mism} px differ (max {mx})"
    if name in gc.SCORE_ONLY:  # masters the reference itself only scores (golden_cases.SCORE_ONLY)
        assert gc.xray_score(got, gc.load_golden(name)) < gc.SCORE_ONLY[name]
    else:
        assert gc.compare(got, gc.load_golden(name)) == (0, 0), f"{name}: GPU vs reference golden"
    assert gb.covered == ob.covered, f"{name}: covered-pixel count {gb.covered} != {ob.covered}"
