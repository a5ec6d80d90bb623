"""Golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py).

CPU part: the current oracle reproduces the frozen vectors (integers exactly, fp64 to 1e-12) and the std::mt19937
known answer of the C++ standard.  GPU part (`-m gpu`): every supplied-randomness seam of the C ABI reproduces them --
bit-exact for index work, <= 1e-10 relative for fp64 (BASELINE.json north_star)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
G = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
INT_KEYS = ("parent_mask", "compatible", "num_compatible", "num_incompatible_events", "neighbors_6_2",
            "neighbors_6_2_ring", "fwd_X", "fwd_dist", "cond_incompatible", "iw_backward_L_eff", "mt19937_kat")


def test_mt19937_known_answer(oracle):
    assert oracle.mt19937_kat() == 4123659995 == int(G["mt19937_kat"])


def test_oracle_reproduces_golden(oracle):
    import make_golden
    now = make_golden.build()
    assert set(now) == set(G.files)
    for k in G.files:
        a, b = np.asarray(now[k]), G[k]
        assert a.shape == b.shape, k
        if k in INT_KEYS or not np.issubdtype(b.dtype, np.floating):
            assert np.array_equal(a, b), k
        else:
            ok = np.isfinite(b)
            assert np.array_equal(np.isfinite(a), ok), k
            np.testing.assert_allclose(a[ok], b[ok], rtol=1e-12, atol=1e-300, err_msg=k)


@pytest.mark.gpu
def test_device_index_work_matches_golden():
    import mccbn_b200 as mc
    topo, mask = mc.flatten_poset(G["poset"])
    assert np.array_equal(mask, G["parent_mask"])
    r = mc.compatible_observations(G["obs"], G["poset"])
    assert np.array_equal(r["compatible"], G["compatible"])
    assert r["num_compatible"] == int(G["num_compatible"])
    assert r["num_incompatible_events"] == int(G["num_incompatible_events"])
    nb = mc.neighbors(6, 2, np.array([1, 0, 1, 1, 0, 0], np.int32))
    assert np.array_equal(nb[0], G["neighbors_6_2"]) and np.array_equal(nb[1], G["neighbors_6_2_ring"])
    lam, eps = mc.initialize_lambda(G["obs"], G["poset"], lambda_s=1.0)
    assert np.array_equal(lam, G["init_lambda"]) and eps == float(G["init_eps"])
    f = mc.forward_from_times(G["fwd_y"], G["poset"], 0.08, G["Z"], G["Ts"])
    for k in ("T_sum", "X", "dist", "w"):
        assert np.array_equal(f[k], G["fwd_" + k]), k


@pytest.mark.gpu
def test_device_fp64_seams_match_golden():
    import mccbn_b200 as mc
    e = mc.estep_forward_from_times(G["obs"], G["poset"], G["lambdas"], 0.08, G["Z"], G["Ts"], weights=G["weights"])
    for k in ("w", "w_sqrt", "dist", "Tdiff", "Tdiff_colsum", "lambda_new"):
        np.testing.assert_allclose(e[k], G["estep_" + k], rtol=1e-10, atol=1e-300, err_msg=k)
    for k in ("obs_llhood", "expected_dist", "N_eff", "eps_new"):
        assert abs(e[k] - float(G["estep_" + k])) <= 1e-10 * abs(float(G["estep_" + k])), k
    c = mc.conditional_from_uniforms(G["cond_X"], G["poset"], G["lambdas"], G["uniforms"], G["cond_Ts"])
    assert np.array_equal(c["incompatible"], G["cond_incompatible"])
    ok = G["cond_incompatible"] == 0
    np.testing.assert_allclose(c["Tdiff"][ok], G["cond_Tdiff"][ok], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(c["log_q"][ok], G["cond_log_q"][ok], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(c["log_pz"][ok], G["cond_log_pz"][ok], rtol=1e-10, atol=1e-12)
