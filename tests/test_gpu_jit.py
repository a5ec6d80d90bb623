"""NVRTC poset-specialised forward kernel (csrc/jit.cuh) against the generic kernel: same Philox stream, same records."""
import numpy as np
import pytest

import mccbn_b200 as mc
from mccbn_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p,times", [(5, False), (16, True), (23, False), (30, False), (32, True), (40, False), (64, True)])
def test_specialised_equals_generic(p, times):
    c = synth.make_config(p, 50, eps=0.07, seed=p, density=0.2 if p <= 32 else 0.06)  # p > 32: the wide (64-bit) kernels
    t = c["times"] if times else None
    try:
        mc.set_jit(0)
        g = mc.importance_weight(c["obs"], 700, c["poset"], c["lambdas"], 0.07, time=t, seed=11)
        mc.set_jit(1)
        j = mc.importance_weight(c["obs"], 700, c["poset"], c["lambdas"], 0.07, time=t, seed=11)
    finally:
        mc.set_jit(-1)
    # identical samples, identical arithmetic: agreement far below the fp32 accumulation error
    for k in ("w", "w_sqrt", "dist", "Tdiff"):
        np.testing.assert_allclose(j[k], g[k], rtol=1e-6, atol=1e-30, err_msg=k)


def test_specialised_mcem_matches_generic():
    c = synth.make_config(12, 200, eps=0.05, seed=3, density=0.25)
    kw = dict(L=256, eps=0.1, max_iter=40, update_step_size=20, seed=9)
    try:
        mc.set_jit(0)
        g = mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], **kw)
        mc.set_jit(1)
        j = mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], **kw)
    finally:
        mc.set_jit(-1)
    np.testing.assert_allclose(j["lambda_"], g["lambda_"], rtol=1e-4)
    assert abs(j["eps"] - g["eps"]) < 1e-5 and abs(j["llhood"] - g["llhood"]) < 1e-3 * abs(g["llhood"])


@pytest.mark.parametrize("sampling,p,kw", [("backward", 9, dict(neighborhood_dist=1)), ("backward", 25, dict(neighborhood_dist=1)),
                                           ("bernoulli", 14, {}), ("add-remove", 20, {})])
def test_specialised_conditional_equals_generic(sampling, p, kw):
    """NVRTC poset-specialised conditional-proposal kernel: same Philox stream and arithmetic as the generic kernel"""
    c = synth.make_config(p, 40, eps=0.0 if sampling == "bernoulli" else 0.06, seed=p, density=0.2)
    L = 26 * 20 if sampling == "backward" else 640
    try:
        mc.set_jit(0)
        g = mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.07, sampling=sampling, seed=11, **kw)
        mc.set_jit(1)
        j = mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.07, sampling=sampling, seed=11, **kw)
    finally:
        mc.set_jit(-1)
    for k in ("w", "w_sqrt", "dist", "Tdiff"):
        np.testing.assert_allclose(j[k], g[k], rtol=1e-5, atol=1e-30, err_msg=k)
    if "L_eff" in g:
        assert np.array_equal(j["L_eff"], g["L_eff"])


def test_specialised_otcbn_matches_generic():
    c = synth.make_config(10, 200, eps=0.0, seed=4, density=0.25)
    kw = dict(max_iter=30, nrOfSamples=8, seed=5)
    try:
        mc.set_jit(0)
        g = mc.estimate_mutation_rates(c["poset"], c["obs"], **kw)
        mc.set_jit(1)
        j = mc.estimate_mutation_rates(c["poset"], c["obs"], **kw)
    finally:
        mc.set_jit(-1)
    np.testing.assert_allclose(j["lambda_"], g["lambda_"], rtol=1e-4)
