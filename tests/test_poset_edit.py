"""CPU-only: the annealer's poset edit operations (host C++, poset.hpp) against brute-force closure semantics.
Reference: Model::add_edge / transitive_reduction_dag / add_relation / remove_relation (src/model_impl.cpp:145-310)."""
import numpy as np

import mccbn_b200 as mc
from mccbn_b200 import synth


def _closure(A):
    return synth.transitive_closure(A)


def test_add_edge_flags_and_result():
    rng = np.random.default_rng(0)
    for p in (4, 7, 12, 20):
        for _ in range(10):
            A = synth.random_poset(p, 0.25, rng=rng)
            perm = rng.permutation(p)
            A = np.asfortranarray(A[np.ix_(perm, perm)])
            C = _closure(A)
            for _ in range(15):
                u, v = rng.integers(0, p, 2)
                if u == v:
                    continue
                r = mc.poset_edit(A, "add_edge", u, v)
                if A[u, v]:
                    assert not r["added"] and np.array_equal(r["poset"], A)  # existing edge: caller toggles it off
                    continue
                assert r["added"]
                assert r["cycle"] == bool(C[v, u])  # cycle <=> u already reachable from v
                if r["cycle"]:
                    continue
                B = A.copy()
                B[u, v] = 1
                red = synth.transitive_reduction(B)
                assert r["reduction_flag"] == bool(np.array_equal(red, B))
                assert np.array_equal(r["poset"], red)  # redundant edges are dropped


def test_add_and_remove_relation_preserve_order_relations():
    rng = np.random.default_rng(1)
    for p in (5, 9, 14):
        for _ in range(12):
            A = synth.random_poset(p, 0.3, rng=rng)
            C = _closure(A)
            # add a new cover relation between incomparable (or forward-comparable) nodes
            for _ in range(10):
                u, v = rng.integers(0, p, 2)
                if u == v or A[u, v]:
                    continue
                r = mc.poset_edit(A, "add_relation", u, v)
                assert r["added"]
                if C[u, v]:
                    assert not r["reduction_flag"]  # v already reachable from u
                    continue
                if C[v, u]:
                    assert r["cycle"]
                    continue
                B = A.copy()
                B[u, v] = 1
                assert np.array_equal(_closure(r["poset"]), _closure(B))  # same partial order as A + (u,v)
                assert np.array_equal(r["poset"], synth.transitive_reduction(B))  # and transitively reduced
                assert r["reduction_flag"]
            # remove every cover relation in turn: only the pair (u, v) leaves the order
            for u, v in zip(*np.nonzero(A)):
                r = mc.poset_edit(A, "remove_relation", u, v)
                want = C.copy()
                want[u, v] = 0
                assert np.array_equal(_closure(r["poset"]), want), (u, v)


def test_transitive_reduction_and_swap():
    rng = np.random.default_rng(2)
    for p in (6, 15):
        A = synth.random_poset(p, 0.4, trans_reduced=False, rng=rng)
        r = mc.poset_edit(A, "transitive_reduction")
        red = synth.transitive_reduction(A)
        assert np.array_equal(r["poset"], red)
        assert r["reduction_flag"] == bool(np.array_equal(red, A))
        s = mc.poset_edit(red, "swap_node", 1, 3)["poset"]  # adjacency over EVENT ids after swapping labels 1 <-> 3
        perm = np.arange(p)
        perm[[1, 3]] = perm[[3, 1]]
        assert np.array_equal(s, red[np.ix_(perm, perm)])
