"""CPU-only: numpy emulations of two device-side algorithms whose correctness does not need a GPU to be stated.

1. warp_transpose_sum (csrc/kernels_forward.cuh): the butterfly that reduces 32 per-lane partial-sum vectors of 32
   components so that lane l ends with the total of component l, in 31 exchanges.
2. the order-statistics resampling of estep_pool_resample_kernel (csrc/kernels_pool.cuh): L iid indices from Discrete(q)
   are distributed like the cells that the order statistics of L iid uniforms fall into along the cumulative weights,
   and those order statistics can be generated one by one in increasing order (the smallest of m iid U(a, 1) is
   1 - (1 - a) V^(1/m)) -- the reference draws the L indices directly (rdiscrete_std, src/mcem_hcbn.cpp:44-54)."""
import numpy as np
from scipy import stats


def _warp_transpose_sum(v):
    """v[lane, component] -> per-lane result, following the exchange pattern of the device function literally"""
    v = v.copy()
    lanes = np.arange(32)
    n_exchanges = 0
    o = 16
    while o >= 1:
        up = (lanes & o) != 0
        for i in range(o):
            lo, hi = v[:, i].copy(), v[:, i + o].copy()
            send = np.where(up, lo, hi)
            keep = np.where(up, hi, lo)
            v[:, i] = keep + send[lanes ^ o]  # __shfl_xor_sync(send, o)
            n_exchanges += 1
        o >>= 1
    return v[:, 0], n_exchanges


def test_warp_transpose_sum_emulation():
    rng = np.random.default_rng(0)
    v = rng.standard_normal((32, 32))
    got, n = _warp_transpose_sum(v)
    assert n == 31
    np.testing.assert_allclose(got, v.sum(axis=0), rtol=1e-13, atol=1e-13)  # lane l holds the total of component l
    # components beyond the live ones are zero on every lane: the live totals are unaffected
    v[:, 25:] = 0.0
    got, _ = _warp_transpose_sum(v)
    np.testing.assert_allclose(got[:25], v[:, :25].sum(axis=0), rtol=1e-13, atol=1e-13)
    assert np.all(got[25:] == 0.0)


def _resample_counts_order_statistics(q, L, rng):
    """counts per entry, generated like the device sweep: one pass over the entries with the running cumulative weight,
    the next order statistic drawn only when the previous one has been placed"""
    S = q.sum()
    cnt = np.zeros(q.size, int)
    m, rem = L, 1.0
    thr = None

    def next_stat():
        nonlocal rem, m
        rem *= np.exp(np.log(rng.random()) / m)  # 1 - U_(next) = (1 - U_(prev)) V^(1/m)
        return S * (1.0 - rem)

    thr = next_stat()
    C = 0.0
    for k in range(q.size):
        C += q[k]
        while m > 0 and thr < C:
            cnt[k] += 1
            m -= 1
            if m > 0:
                thr = next_stat()
    cnt[np.nonzero(q)[0][-1]] += m  # leftovers of a rounding of S go to the last entry with positive weight
    return cnt


def test_order_statistics_resampling_is_multinomial():
    rng = np.random.default_rng(1)
    q = rng.uniform(0.1, 1.0, 12) ** 2  # no near-empty cells: their Pearson terms are too skewed for a 4000-run mean
    q[[2, 7]] = 0.0  # entries with weight zero are never drawn
    L, reps = 40, 4000
    counts = np.array([_resample_counts_order_statistics(q, L, rng) for _ in range(reps)])
    assert np.all(counts.sum(axis=1) == L) and np.all(counts[:, [2, 7]] == 0)
    prob = q / q.sum()
    live = prob > 0
    # marginal means and variances of a Multinomial(L, prob)
    z = (counts.mean(axis=0)[live] - L * prob[live]) / np.sqrt(L * prob[live] * (1 - prob[live]) / reps)
    assert np.abs(z).max() < stats.norm.isf(1e-3 / (2 * live.sum()))
    # pairwise covariance -L p_a p_b (the draws compete for the same L slots), largest pair
    a, b = np.argsort(prob)[-2:]
    cov = np.cov(counts[:, a], counts[:, b])[0, 1]
    assert abs(cov / (-L * prob[a] * prob[b]) - 1.0) < 0.15
    # the pooled chi-square of the per-run counts against the multinomial: sum over runs of Pearson statistics
    pearson = (((counts[:, live] - L * prob[live]) ** 2) / (L * prob[live])).sum(axis=1)
    df = live.sum() - 1
    assert abs(pearson.mean() - df) < 4 * np.sqrt(2 * df / reps)
