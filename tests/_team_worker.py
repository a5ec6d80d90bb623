"""Worker of tests/test_gpu_multi.py::test_in_process_team (a fresh process: the team size is process state).

One process, G devices (mccbn_set_gpus): the reference-facing entries shard the observations over the devices
themselves.  Checked against the same calls on one device: same Philox samples (global observation indices), so only
the fp64 summation order of the statistic vector may differ."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    G = int(sys.argv[1])
    import mccbn_b200 as mc
    from mccbn_b200 import synth
    res = {}
    cases = (("forward", 512, 1, synth.make_config(14, 4001, eps=0.05, seed=5, density=0.2)),
             ("backward", 560, 2, synth.make_config(10, 2001, eps=0.02, seed=5, density=0.2)))
    for sampling, L, kdist, c in cases:
        kw = dict(L=L, eps=0.1, sampling=sampling, neighborhood_dist=kdist, seed=3, max_iter=6, update_step_size=3)
        out = {}
        for g in (1, G):
            mc.set_gpus(g)
            assert mc.get_gpus() == g, (mc.get_gpus(), g)
            fit = mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], **kw)
            ll = mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.05, L=L, sampling=sampling,
                                      neighborhood_dist=kdist, seed=9)
            iw = mc.importance_weight(c["obs"], L, c["poset"], c["lambdas"], 0.05, sampling=sampling,
                                      neighborhood_dist=kdist, seed=9)
            out[g] = (fit, ll, iw)
        (f1, l1, w1), (fg, lg, wg) = out[1], out[G]
        res[sampling] = {
            "lambda_rel": float(np.max(np.abs(fg["lambda_"] / f1["lambda_"] - 1))),
            "eps_abs": abs(fg["eps"] - f1["eps"]),
            "ll_rel": abs(fg["llhood"] / f1["llhood"] - 1),
            "obs_ll_rel": abs(lg / l1 - 1),
            "iw_w_rel": float(np.max(np.abs(wg["w"] / w1["w"] - 1))),
            "iw_dist_abs": float(np.max(np.abs(wg["dist"] - w1["dist"]))),
            "iw_Tdiff_rel": float(np.max(np.abs(wg["Tdiff"] / w1["Tdiff"] - 1))),
        }
    # weighted observations and given sampling times travel with their shard
    c = cases[0][3]
    wt = np.linspace(0.5, 1.5, c["obs"].shape[0])
    a = {}
    for g in (1, G):
        mc.set_gpus(g)
        a[g] = mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.05, weights=wt, times=c["times"], L=256, seed=4)
    res["weights_times_ll_rel"] = abs(a[G] / a[1] - 1)
    # the reference's abort travels through the team: eps = 0 and an observation no sample can match
    mc.set_gpus(G)
    bad = c["obs"].copy()
    u, v = [int(x[0]) for x in np.nonzero(c["poset"])]
    bad[:, :] = 0
    bad[-1, v] = 1  # child without its parent: incompatible, weight 0 for every sample at eps = 0
    try:
        mc.obs_loglikelihood(bad, c["poset"], c["lambdas"], 0.0, L=64, seed=1)
        res["zero_weight_error"] = "no error"
    except mc.MccbnError as e:
        res["zero_weight_error"] = [e.code, str(e)]
    # the team is still usable after an error
    res["after_error_ll_rel"] = abs(mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.05, weights=wt,
                                                         times=c["times"], L=256, seed=4) / a[1] - 1)
    # candidate-parallel annealing over the team: the single-device trajectory, bit for bit
    ca = synth.make_config(8, 400, eps=0.05, seed=5, density=0.3)
    start = np.zeros((8, 8), np.int32)
    kw = dict(L=64, max_iter=40, update_step_size=20, max_iter_asa=30, seed=7)
    os.environ["MCCBN_ASA_BATCH"] = "8"
    r = {}
    for g in (1, G):
        mc.set_gpus(g)
        d = tempfile.mkdtemp() + "/"
        r[g] = mc.adaptive_simulated_annealing(start, ca["obs"], outdir=d, **kw)
        r[g]["files"] = os.path.exists(d + "params.txt")
    res["asa_identical"] = bool(np.array_equal(r[1]["poset"], r[G]["poset"]) and
                                np.array_equal(r[1]["lambda_"], r[G]["lambda_"]) and r[1]["eps"] == r[G]["eps"] and
                                r[1]["llhood"] == r[G]["llhood"] and r[G]["files"])
    print("TEAM_RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
