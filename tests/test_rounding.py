"""CPU: the two rounding flavours the reference mixes (builtin round on floats vs numpy's
rint(x*10^k)/10^k) -- oracle primitives against Python/numpy themselves."""
import numpy as np

import _util  # noqa: F401
import oracle


def test_pyround2_matches_builtin_round():
    L = oracle.lib()
    rng = np.random.default_rng(0)
    xs = list(rng.uniform(-4, 4, 20000)) + [k / 1000.0 for k in range(-3200, 3200)] + [0.005, 0.015, 0.025, 2.675, 1.005, 0.125, 0.375]
    for x in xs:
        assert L.uo_pyround2(float(x)) == round(float(x), 2), x
    assert np.isnan(L.uo_pyround2(float("nan")))


def test_npround_matches_numpy():
    L = oracle.lib()
    rng = np.random.default_rng(1)
    xs = list(rng.uniform(-60, 60, 20000)) + [k / 1000.0 + 0.0005 for k in range(-2000, 2000)]
    for x in xs:
        for nd in (2, 3):
            assert L.uo_npround(float(x), nd) == float(np.round(np.float64(x), nd)), (x, nd)
