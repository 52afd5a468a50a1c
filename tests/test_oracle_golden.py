"""CPU tests (no GPU): the oracle restatement against the committed golden vectors, i.e. against the
reference's own driver compiled over the same minipetsc (tests/golden/make_golden.py), bit for bit."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

from oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))
GOLD_U = np.load(os.path.join(HERE, "golden", "golden_u.npz"))


def _hex(v):
    return [float.fromhex(x) for x in v]


@pytest.mark.parametrize("name", sorted(GOLD.keys()))
def test_oracle_matches_golden_bitwise(name):
    g = GOLD[name]
    o = Oracle(g["options"])
    it, rn = o.solve()
    u, err = o.postprocess()
    o.close()
    assert it == g["num_iter"]
    want = np.array(_hex(g["rnorm_hex"]))
    # residual history: %.16e text round-trips doubles exactly, so equality is bitwise
    assert rn.shape == want.shape
    assert np.array_equal(rn, want) or all(
        (math.isnan(a) and math.isnan(b)) or a == b for a, b in zip(rn, want))
    assert np.array_equal(err, np.array(_hex(g["error_hex"])))
    assert list(u.shape) == g["u_shape"]
    assert hashlib.sha256(np.ascontiguousarray(u, dtype="<f8").tobytes()).hexdigest() == g["u_sha256"]
    if name in GOLD_U.files:
        assert np.array_equal(u, GOLD_U[name])


def test_known_answers_discretisation_error():
    """max|u - u_exact| ~ pi^2 h^2 / 12 for the 5-point scheme."""
    for name, npts in (("n17_l2_jacobi", 17), ("n129_l4_jacobi", 129), ("n1025_l10_jacobi", 1025)):
        err = _hex(GOLD[name]["error_hex"])[0]
        h = 1.0 / (npts - 1)
        assert err == pytest.approx(math.pi ** 2 * h * h / 12.0, rel=0.12)
        assert _hex(GOLD[name]["rnorm_hex"])[0] == 1.0


def test_survey_expected_iteration_counts():
    """BASELINE.md section 4 sanity table (independent numpy/scipy restatement made during the survey)."""
    expect = {"n17_l2_jacobi": 59, "n17_l2_jacobi23": 71, "n17_l2_sor": 13, "n129_l4_jacobi": 212,
              "n129_l4_sor": 43, "n129_l4_sor_forward": 85, "n129_l7_jacobi": 9,
              "n1025_l7_jacobi": 210, "n1025_l10_jacobi": 9}
    for k, v in expect.items():
        assert GOLD[k]["num_iter"] == v


def test_oracle_rejects_unsupported():
    with pytest.raises(ValueError):
        Oracle("-npts 17 -iter 10 -grids 3 -levels 2")        # several grids per level: out of scope
    with pytest.raises(ValueError):
        Oracle("-npts 17 -iter 10 -levels 2 -cycle 1")        # research cycles: out of scope
    with pytest.raises(ValueError):
        Oracle("-iter 10 -levels 2")                          # -npts missing


def test_red_black_numbering_is_a_permutation_of_natural():
    nat = Oracle("-npts 17 -iter 1 -levels 2 -map 2 -pc_type jacobi")
    rb = Oracle("-npts 17 -iter 1 -levels 2 -map 3 -pc_type jacobi")
    for l in range(2):
        g = rb.grid_to_global(l)
        ni, nj = rb.dims(l)
        assert sorted(g.tolist()) == list(range(ni * nj))
        ii, jj = np.divmod(np.arange(ni * nj), nj)
        nred = int((((ii + jj) % 2) == 0).sum())
        assert np.all(g[((ii + jj) % 2) == 0] < nred) and np.all(g[((ii + jj) % 2) == 1] >= nred)
        # same operator up to the permutation
        (_, _), ia, ja, va = nat.csr(0, l)
        (_, _), ib, jb, vb = rb.csr(0, l)
        x = np.random.default_rng(0).standard_normal(ni * nj)
        y_nat = nat.matmult(0, l, x)
        y_rb = rb.to_grid(l, rb.matmult(0, l, rb.from_grid(l, x.reshape(ni, nj)))).reshape(-1)
        assert np.allclose(y_nat, y_rb, rtol=1e-14, atol=1e-9)
    nat.close()
    rb.close()
