"""adaptive.simulated.annealing on the device (BASELINE config 4 path): speculative batched scoring must not change the
result, outputs follow the reference's file formats, and the search behaves like the reference's (score never
decreases for the tracked ML poset; the returned poset is a transitively reduced DAG)."""
import os

import numpy as np
import pytest

import mccbn_b200 as mc
from mccbn_b200 import synth
from stat_criteria import joint_check

pytestmark = pytest.mark.gpu


def _run(tmp, batch, seed=7, **kw):
    if batch is None:  # automatic batch size (work-based cap, adapted to the acceptance pattern)
        os.environ.pop("MCCBN_ASA_BATCH", None)
    else:
        os.environ["MCCBN_ASA_BATCH"] = str(batch)
    out = os.path.join(tmp, f"b{batch}_")
    c = synth.make_config(8, 400, eps=0.05, seed=5, density=0.3)
    start = np.zeros((8, 8), np.int32)  # empty poset as the initial guess
    r = mc.adaptive_simulated_annealing(start, c["obs"], L=64, max_iter=40, update_step_size=20, max_iter_asa=30,
                                        outdir=out, seed=seed, **kw)
    return c, r, out


def test_asa_batched_equals_sequential(tmp_path):
    c, r1, o1 = _run(str(tmp_path), 1)
    for batch in (8, None):
        _, r8, o8 = _run(str(tmp_path), batch)
        assert np.array_equal(r1["poset"], r8["poset"])
        assert np.array_equal(r1["lambda_"], r8["lambda_"]) and r1["eps"] == r8["eps"] and r1["llhood"] == r8["llhood"]
        for name in ("asa.txt", "params.txt", "poset.txt"):
            assert open(o1 + name).read() == open(o8 + name).read(), name


def test_asa_outputs_and_search(tmp_path):
    c, r, out = _run(str(tmp_path), 8, seed=11)
    P = r["poset"]
    assert P.shape == (8, 8) and set(np.unique(P)) <= {0, 1}
    synth.topological_sort(P)  # raises if cyclic
    assert np.array_equal(synth.transitive_reduction(P), P)
    assert np.all(r["lambda_"] > 0) and 0 < r["eps"] < 0.5 and np.isfinite(r["llhood"])
    par = open(out + "params.txt").read().strip().split("\n")
    assert par[0] == "step\t llhood\t epsilon\t lambdas"
    assert len(par) == 1 + 30  # step 0 .. max_iter_asa - 1 (src/asa.cpp:527-552)
    assert [int(l.split("\t")[0]) for l in par[1:]] == list(range(30))
    asa = open(out + "asa.txt").read().strip().split("\n")
    assert asa[0] == "step\t llhood\t temperature\t acceptance rate"
    assert [int(l.split("\t")[0]) for l in asa[1:]] == [3, 6, 9, 12, 15, 18, 21, 24, 27]  # every step.size = 30 / 10
    lls = np.array([float(l.split("\t")[1]) for l in par[1:]])
    # learning from the empty poset: the best score seen improves on the start
    assert lls.max() > lls[0]
    # poset.txt holds the cover relations of the best poset seen ("u\tv" per line, event ids)
    edges = [tuple(map(int, l.split("\t"))) for l in open(out + "poset.txt").read().strip().split("\n") if l]
    assert all(0 <= u < 8 and 0 <= v < 8 for u, v in edges)
    # the final fit scores the ML poset: it is at least as good as the empty poset's fit (up to MC noise)
    ll0 = mc.MCEM_hcbn(*mc.initialize_lambda(c["obs"], np.zeros((8, 8), np.int32))[:1], np.zeros((8, 8), np.int32),
                       c["obs"], L=64, eps=0.05, max_iter=40, seed=3)["llhood"]
    assert r["llhood"] > ll0 - 5.0


def test_asa_not_acyclic_start():
    start = np.zeros((4, 4), np.int32)
    start[0, 1] = start[1, 0] = 1
    with pytest.raises(mc.MccbnError) as e:
        mc.adaptive_simulated_annealing(start, np.zeros((5, 4), np.int32), L=16, max_iter_asa=3, seed=1)
    assert e.value.code == 1


# ---- against the oracle's restatement of src/asa.cpp (one mt19937 for everything, like the reference) --------------
def _asa_summary(r, out):
    par = [l.split("\t") for l in open(out + "params.txt").read().strip().split("\n")[1:]]
    ll = np.array([float(x[1]) for x in par])
    return np.array([r["llhood"], ll.max(), ll[-1], float(np.count_nonzero(np.diff(ll))), float(r["poset"].sum())])


def _two_sample_z(a, b):
    a, b = np.asarray(a), np.asarray(b)
    se = np.sqrt(a.var(axis=0, ddof=1) / len(a) + b.var(axis=0, ddof=1) / len(b))
    return (a.mean(axis=0) - b.mean(axis=0)) / np.maximum(se, 1e-9)


@pytest.mark.parametrize("mode,n_seeds,kw", [
    ("pseudo", 48, dict(L=8, max_iter=2, update_step_size=1, max_iter_asa=60)),
    ("mcem", 16, dict(L=40, max_iter=20, update_step_size=10, max_iter_asa=40)),
])
def test_asa_matches_oracle_in_distribution(oracle, tmp_path, mode, n_seeds, kw):
    """The annealer (move generator, compatibility heuristic, memo tables, Metropolis rule, temperature schedule, ML
    tracking) against the oracle's line-by-line restatement of src/asa.cpp.  The random streams differ by design (the
    reference threads ONE mt19937 through proposals, acceptance and scoring; here proposals / acceptance / Philox
    scores are separate streams so that candidates can be scored speculatively), so the comparison is distributional:
    final score, best score, last score, number of score changes (= accepted moves) and number of cover relations of
    the result, over replicate seeds.  `pseudo` uses a deterministic score on both sides (no Monte-Carlo noise in the
    score: isolates the host logic), `mcem` the real MCEM scores."""
    c = synth.make_config(6, 150, eps=0.05, seed=21, density=0.35)
    start = np.zeros((6, 6), np.int32)
    os.environ["MCCBN_ASA_BATCH"] = "8"
    try:
        g, o = [], []
        for s in range(n_seeds):
            og = os.path.join(str(tmp_path), f"g{s}_")
            oo = os.path.join(str(tmp_path), f"o{s}_")
            rg = mc.adaptive_simulated_annealing(start, c["obs"], outdir=og, seed=1 + s, T0=5.0,
                                                 _test_structural_score=(mode == "pseudo"), **kw)
            ro = oracle.adaptive_simulated_annealing(start, c["obs"], outdir=oo, seed=1001 + s, T0=5.0, thrds=2,
                                                     score_mode=1 if mode == "pseudo" else 0,
                                                     **{("max_iter" if k == "max_iter" else k): v for k, v in kw.items()})
            g.append(_asa_summary(rg, og))
            o.append(_asa_summary(ro, oo))
            for r in (rg, ro):  # both return transitively reduced DAGs
                synth.topological_sort(r["poset"])
                assert np.array_equal(synth.transitive_reduction(r["poset"]), r["poset"])
        z = _two_sample_z(g, o)
        joint_check(z, 2 * n_seeds - 2, f"ASA {mode}")
    finally:
        os.environ.pop("MCCBN_ASA_BATCH", None)


def test_short_annealing_run_exits_cleanly():
    """A short run leaves background compilations of look-ahead kernels in flight when the process ends (the default
    context is never destroyed): the library waits for them at exit, in front of libnvrtc's own teardown -- the process must
    end with status 0, not with a SIGSEGV after it has printed its result."""
    import subprocess
    import sys
    code = (
        "import sys, tempfile; sys.path.insert(0, %r)\n"
        "import numpy as np, mccbn_b200 as mc\n"
        "from mccbn_b200 import synth\n"
        "c = synth.make_config(20, 10000, eps=0.05, seed=10)\n"
        "r = mc.adaptive_simulated_annealing(c['poset'], c['obs'], L=1000, max_iter=100, update_step_size=20, max_iter_asa=8,\n"
        "                                    outdir=tempfile.mkdtemp() + '/', seed=10)\n"
        "print('llhood', r['llhood'])\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, MCCBN_ASA_PREFETCH_GATE="0", MCCBN_CACHE_DIR=str(__import__("tempfile").mkdtemp()))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert "llhood" in r.stdout, r.stderr[-2000:]
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
