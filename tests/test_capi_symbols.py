"""CPU-only: the C-ABI library loads, exports every symbol include/mccbn_b200.h declares, its host-only entry points
agree with the oracle, and compute entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import os
import re

import numpy as np
import pytest

import mccbn_b200 as mc
from mccbn_b200 import _capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "mccbn_b200.h")).read()
    declared = set(re.findall(r"\b(mccbn_[A-Za-z_0-9]+)\s*\(", hdr))
    declared -= {"mccbn_em_options"}
    assert declared, "no declarations parsed"
    lib = _capi.lib()
    for s in sorted(declared):
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert declared == set(_capi.SYMBOLS)


def test_product_does_not_reference_oracle():
    # the product path must never import / link the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mccbn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "libmccbn_oracle" not in txt, f


def test_flatten_poset_matches_oracle(oracle):
    rng = np.random.default_rng(0)
    for p in (1, 2, 7, 16, 30, 48, 64):
        poset = synth.random_poset(p, 0.2, rng=rng)
        perm = rng.permutation(p)  # relabel so that the matrix is not upper-triangular
        poset = np.asfortranarray(poset[np.ix_(perm, perm)])
        topo, mask = mc.flatten_poset(poset)
        topo_o, mask_o, cyc = oracle.flatten_poset(poset)
        assert not cyc
        assert np.array_equal(mask, mask_o)  # parent sets: bit-exact
        pos = np.empty(p, int)
        pos[topo] = np.arange(p)
        assert sorted(topo) == list(range(p))
        for u, v in zip(*np.nonzero(poset)):
            assert pos[u] < pos[v]  # any valid topological order is equivalent


def test_cycle_and_self_loop_rejected(oracle):
    poset = np.zeros((5, 5), np.int32)
    poset[0, 1] = poset[1, 2] = poset[2, 3] = poset[3, 1] = 1
    with pytest.raises(mc.MccbnError) as e:
        mc.flatten_poset(poset)
    assert e.value.code == 1 and str(e.value) == "Poset does not correspond to an acyclic graph"
    assert oracle.flatten_poset(poset)[2]
    loop = np.zeros((3, 3), np.int32)
    loop[1, 1] = 1
    with pytest.raises(mc.MccbnError):
        mc.flatten_poset(loop)


def test_neighbors_matches_oracle(oracle):
    rng = np.random.default_rng(1)
    for p, k in ((1, 0), (1, 1), (5, 1), (6, 2), (7, 3), (4, 4), (25, 1), (12, 2)):
        y = rng.integers(0, 2, p)
        rows, ring = mc.neighbors(p, k, y)
        rows_o, ring_o = oracle.neighbors(p, k, y)
        assert np.array_equal(rows, rows_o) and np.array_equal(ring, ring_o)  # same enumeration order


def test_no_cpu_fallback():
    if mc.device_count() > 0:
        pytest.skip("CUDA device present")
    c = synth.make_config(5, 10, seed=1)
    with pytest.raises(mc.MccbnError) as e:
        mc.MCEM_hcbn(c["lambda0"], c["poset"], c["obs"], L=32, eps=0.1, seed=1)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)
    with pytest.raises(mc.MccbnError):
        mc.compatible_observations(c["obs"], c["poset"])


def test_device_team_arguments():
    """mccbn_set_gpus: argument range; without devices a team request degrades to "one device" and the compute entries
    still fail loudly (no CPU fallback behind the team path either)."""
    with pytest.raises(mc.MccbnError) as e:
        mc.set_gpus(-1)
    assert e.value.code == 4
    with pytest.raises(mc.MccbnError):
        mc.set_gpus(17)
    if mc.device_count() > 0:
        pytest.skip("CUDA device present")
    mc.set_gpus(8)
    try:
        assert mc.get_gpus() == 1
        c = synth.make_config(5, 10, seed=1)
        with pytest.raises(mc.MccbnError) as e:
            mc.obs_loglikelihood(c["obs"], c["poset"], c["lambdas"], 0.1, L=8, seed=1)
        assert e.value.code == 3 and "no CPU fallback" in str(e.value)
    finally:
        mc.set_gpus(1)


def test_r_wrapper_argument_rules():
    with pytest.raises(ValueError):
        mc.obs_loglikelihood(np.zeros((2, 3), np.int32), np.zeros((3, 3), np.int32), np.ones(3), 0.1, L=4,
                             sampling="naive")
    with pytest.raises(TypeError):
        mc.MCEM_hcbn(np.ones(3), np.zeros((3, 3), np.int32), np.zeros((2, 3), np.int32))


def test_jit_source_generation_and_nvrtc_compile():
    """The NVRTC specialisation is a pure compiler step: it can be checked without a GPU."""
    c = synth.make_config(9, 5, seed=2, density=0.3)
    try:
        r = mc.jit_compile_forward(c["poset"])
    except mc.MccbnError as e:
        if "could not be loaded" in str(e):
            pytest.skip("libnvrtc not available")
        raise
    assert r["cubin_bytes"] > 10000
    assert "mccbn_fwd_spec" in r["source"] and "estep_forward_body<float, 12, false>" in r["source"]
    topo, _ = mc.flatten_poset(c["poset"])
    assert r["source"].count("fast_lg2(") == 9 + 1  # one exponential per event + the sampling time
    assert r["source"].count("(uint64_t)0xD2511F53u") == 10 * 2  # 10 uniforms of 24 bits = 2 Philox blocks x 10 rounds
    assert r["source"].count("fmaxf(") == sum(max(0, int(d) - 1) for d in c["poset"].sum(axis=0))  # |E| - #non-roots
    assert "registers" in r["log"]


def test_jit_conditional_source_generation_and_nvrtc_compile():
    """the conditional-proposal specialisation (backward / bernoulli / OT-CBN / add-remove) is a pure compiler step too"""
    c = synth.make_config(9, 5, seed=2, density=0.3)
    for mode in (0, 3):
        try:
            r = mc.jit_compile_conditional(c["poset"], mode=mode)
        except mc.MccbnError as e:
            if "could not be loaded" in str(e):
                pytest.skip("libnvrtc not available")
            raise
        assert r["cubin_bytes"] > 10000
        assert "mccbn_cond_spec" in r["source"] and f"estep_conditional_body<float, 12, false, {mode}>" in r["source"]
        assert r["source"].count("CondMath<float>::z_of(") == 9 + 1  # one draw per event + the sampling time
        assert "spill stores" in r["log"]
