"""CPU: the control flow of k_theil_sen's bracket search (modem_b200/csrc/demod.cu: ts_slope), restated in
tools/ts_search_emulation.py, on the one window of bench.py's config3 batch in which a row went through the bisection fallback on
the GPU (tests/golden/ts_fallback_window.npz: tap PHASE of window 1397, written by tools/ts_dump_fallback_rows.py on a B200).
With the pilot stopped early, one lane's sub-queue of that row holds 65 pairs (cap 64) while the bracket as a whole holds 648:
the overflow zoom of the time "shrank" the bracket to 105 % of its width sixteen times over."""
import os
import re
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ts_search_emulation as E  # noqa: E402


@pytest.fixture(scope="module")
def window():
    return np.load(os.path.join(ROOT, "tests", "golden", "ts_fallback_window.npz"))


def test_emulation_uses_the_kernels_constants():
    src = open(os.path.join(ROOT, "modem_b200", "csrc", "demod.cu")).read()
    assert int(re.search(r"constexpr int kTsCap = (\d+);", src).group(1)) == E.K_CAP
    assert int(re.search(r"constexpr int kTsCandCap = (\d+);", src).group(1)) == E.CAND_CAP
    assert re.search(r"constexpr int kTsLaneCap = kTsCap / 32;", src) and E.LANE_CAP == E.K_CAP // 32
    # the zoom the emulation calls "new" is the one in the source
    assert "fminf((float)kTsCap / (3.f * (float)nq), (0.75f * (float)kTsLaneCap) / (float)nqmax)" in src


def test_device_slopes_of_the_window_are_the_exact_order_statistic(window):
    """what the B200 returned for the 50 rows (through the fallback for row 38) equals the brute-force upper median"""
    for r in (0, 17, 38, 49):
        y = window["phase"][r]
        i, j = np.triu_indices(y.size, 1)
        q = ((y[j] - y[i]).astype(np.float32) / (j - i).astype(np.float32)).astype(np.float32)
        assert np.partition(q, q.size // 2)[q.size // 2] == window["slope"][r]


def test_one_overfull_sub_queue_no_longer_stalls_the_search(window):
    y = window["phase"][38]
    sweeps_old, trace = E.search(y, steps=4, early_stop=0.4, new_zoom=False)
    assert sweeps_old == 116 and trace[0][4] == 65 and trace[0][3] < E.CAND_CAP   # 16 sweeps on the same bracket, then the bisection
    sweeps_new, _ = E.search(y, steps=4, early_stop=0.4, new_zoom=True)
    assert sweeps_new <= 3


@pytest.mark.parametrize("steps,early_stop", [(4, 0.0), (4, 0.4), (3, 0.0)])
def test_every_row_of_the_window_converges(window, steps, early_stop):
    total = 0
    for y in window["phase"]:
        sweeps, _ = E.search(y, steps=steps, early_stop=early_stop)
        assert sweeps <= 3
        total += sweeps
    assert total <= 70   # the GPU counted 179 for this window with the early stop, 64 are needed
