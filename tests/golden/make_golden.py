#!/usr/bin/env python
"""Regenerates tests/golden/golden_v1.npz from the CPU oracle (oracle/mccbn_oracle.cpp).

What these vectors are -- and are not.  The reference (an R package whose C++ needs Rcpp, RcppEigen and Boost) cannot be
built or run in this image and ships no golden vectors / known-answer tests for this path, so
there is nothing of the reference's own to commit.  The vectors below are outputs of the oracle -- the line-by-line
restatement of the reference that the model-derived exact answers of tests/test_oracle_exact.py pin -- frozen at the
commit that introduced them.  They serve two purposes: (1) the oracle cannot drift silently (tests/test_golden.py
recomputes them with the current oracle, libm and libstdc++), (2) the `-m gpu` parity tests have fixed, reviewable
inputs and expected outputs for every seam of the C ABI.  The one externally specified value is the std::mt19937
known answer of the C++ standard (the reference's generator), checked by the same test.

usage: python tests/golden/make_golden.py      (needs only gcc; no GPU, no /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def build():
    import oracle as orc
    from mccbn_b200 import synth
    orc.build()
    g = {}
    # ---- inputs: the README-sized configuration (p = 10, seed 10 as in the reference's README example) ----
    c = synth.make_config(10, 60, eps=0.05, seed=10, density=0.15)
    p, N = 10, 60
    rng = np.random.default_rng(2024)
    L = 64
    g["poset"], g["obs"], g["lambdas"], g["lambda0"], g["times"] = c["poset"], c["obs"], c["lambdas"], c["lambda0"], c["times"]
    g["Z"] = rng.exponential(1.0, (L, p)) / c["lambdas"]
    g["Ts"] = rng.exponential(1.0, L)
    g["weights"] = rng.uniform(0.5, 2.0, N)
    g["uniforms"] = rng.random((L, p))
    g["eps"] = np.float64(0.08)
    # ---- indexing work (bit-exact tier) ----
    topo, mask, cyc = orc.flatten_poset(c["poset"])
    g["parent_mask"] = mask
    g["compatible"] = orc.compatible_observations(c["obs"], c["poset"])
    g["num_compatible"] = np.int64(orc.num_compatible_observations(c["obs"], c["poset"]))
    g["num_incompatible_events"] = np.int64(orc.num_incompatible_events(c["obs"], c["poset"]))
    nb = orc.neighbors(6, 2, np.array([1, 0, 1, 1, 0, 0], np.int32))
    g["neighbors_6_2"], g["neighbors_6_2_ring"] = nb
    lam_i, eps_i = orc.initialize_lambda(c["obs"], c["poset"], lambda_s=1.0)
    g["init_lambda"], g["init_eps"] = lam_i, np.float64(eps_i)
    y = c["obs"][7]
    f = orc.forward_from_times(y, c["poset"], 0.08, g["Z"], g["Ts"])
    g["fwd_y"] = y
    for k in ("T_sum", "X", "dist", "w"):
        g["fwd_" + k] = f[k]
    # ---- floating-point tier (<= 1e-10) from supplied randomness ----
    e = orc.estep_forward_from_times(c["obs"], c["poset"], c["lambdas"], 0.08, g["Z"], g["Ts"], weights=g["weights"])
    for k in ("w", "w_sqrt", "dist", "Tdiff", "Tdiff_colsum", "lambda_new"):
        g["estep_" + k] = e[k]
    for k in ("obs_llhood", "expected_dist", "N_eff", "eps_new"):
        g["estep_" + k] = np.float64(e[k])
    X = synth.sample_genotypes(L, c["poset"], 1.0, c["lambdas"], 0.0, np.random.default_rng(5))["hidden_genotypes"].copy()
    X[L // 2:] = np.random.default_rng(6).integers(0, 2, (L - L // 2, p))
    g["cond_X"] = X.astype(np.int32)
    g["cond_Ts"] = g["Ts"] + 0.01
    cu = orc.conditional_from_uniforms(g["cond_X"], c["poset"], c["lambdas"], g["uniforms"], g["cond_Ts"])
    for k in ("incompatible", "Tdiff", "log_q", "log_pz"):
        g["cond_" + k] = cu[k]
    g["scale_path"] = orc.scale_path_to_mutation(c["poset"], c["lambdas"])
    g["log_bernoulli"] = np.array([orc.log_bernoulli_process(d, 0.08, p) for d in range(p + 1)])
    # ---- the reference's own random stream (std::mt19937 through the StdRng seeding chain), one thread ----
    g["mt19937_kat"] = np.uint32(orc.mt19937_kat())
    g["rexp_seed42"] = orc.rexp(8, 1.5, 42)
    r = orc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], L=100, eps=0.05, max_iter=40, update_step_size=20, seed=10, thrds=1)
    g["mcem_lambda"], g["mcem_eps"], g["mcem_llhood"] = r["lambda_"], np.float64(r["eps"]), np.float64(r["llhood"])
    iw = orc.importance_weight(c["obs"][:12], 200, c["poset"], c["lambdas"], 0.08, sampling="backward", seed=3, thrds=1)
    g["iw_backward_w"], g["iw_backward_L_eff"], g["iw_backward_dist"] = iw["w"], iw["L_eff"], iw["dist"]
    return g


if __name__ == "__main__":
    out = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(out, **build())
    print("wrote", out, os.path.getsize(out), "bytes")
