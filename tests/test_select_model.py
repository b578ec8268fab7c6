"""CPU model of the score-linear selection the CUDA ranking kernels rest on (csrc/topk.cu
fused_select_sort / counting_rank_sort, csrc/bucketsort.cu second level), in numpy float32 arithmetic
with the kernels' own operations: key = order-preserving u32 image of the score, sub-bin =
trunc((hi - score) * scale) saturated, NaN keys in the last sub-bin.  Checks the two properties the
exactness argument needs -- the sub-bin is non-decreasing in the key, equal keys share one -- and that
selection + ordering through the sub-bins reproduces a plain stable sort with pandas' keep="all" tie rule
(measuring_bias.py:29,57 -> selectn.py:106-190) on random, tie-heavy and special-value inputs.
The CUDA kernels themselves are checked against the oracle in the -m gpu tests; this file pins the
algorithm."""
import numpy as np
import pytest

BINS = 2048
NAN_KEY = np.uint32(0xFFFFFFFF)
FLT_MAX = np.float32(3.4028234664e38)


def score_to_key(s: np.ndarray) -> np.ndarray:
    """Ascending u32 key = descending score; -0 == +0; NaN -> 0xFFFFFFFF (dvl_common.cuh score_to_key)."""
    s = s.astype(np.float32).copy()
    s[s == 0] = 0.0
    u = s.view(np.uint32)
    key = np.where(u & np.uint32(0x80000000), u, u ^ np.uint32(0x7FFFFFFF)).astype(np.uint32)
    key[np.isnan(s)] = NAN_KEY
    return key


def key_score(k: np.ndarray) -> np.ndarray:
    k = k.astype(np.uint32)
    sign = (k.view(np.int32) >> 31).view(np.uint32)
    return (k ^ (~sign & np.uint32(0x7FFFFFFF))).view(np.float32)


def subbin(k: np.ndarray, hi: np.float32, scale: np.float32) -> np.ndarray:
    with np.errstate(over="ignore", invalid="ignore"):
        f = (np.float32(hi) - key_score(k)) * np.float32(scale)        # float32 throughout
    f = np.where(np.isnan(f) | (f < 0), np.float32(0), f)              # __float2uint_rz saturates: negative, NaN -> 0
    sb = np.minimum(np.floor(np.minimum(f, np.float32(4e9))).astype(np.int64), BINS - 1)
    return np.where(k == NAN_KEY, BINS - 1, sb)


def plan(keys: np.ndarray):
    real = keys[keys != NAN_KEY]
    hi, scale = np.float32(0), np.float32(0)
    if real.size and real.max() > real.min():
        hi = np.minimum(key_score(real.min()[None])[0], FLT_MAX)
        lo = np.maximum(key_score(real.max()[None])[0], -FLT_MAX)
        if hi > lo:
            with np.errstate(over="ignore"):
                scale = np.float32(BINS - 2) / np.float32(hi - lo)
    if not scale <= FLT_MAX:
        scale = np.float32(0)
    return hi, scale


def fused_select(keys: np.ndarray, idx: np.ndarray, k: int):
    """(kept entries ordered by (key, idx), kth key) the way fused_select_sort finds them."""
    n = keys.size
    hi, scale = plan(keys)
    sb = subbin(keys, hi, scale)
    counts = np.bincount(sb, minlength=BINS)
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    kk = min(k, n)
    bstar = int(np.nonzero((start < kk) & (kk <= start + counts))[0][0])
    take = sb <= bstar
    kprime = int(take.sum())
    assert kprime == start[bstar] + counts[bstar]
    # scatter by sub-bin (arbitrary order inside), then rank inside the sub-bin by (key, idx)
    order = np.lexsort((idx[take], keys[take], sb[take]))
    okeys, oidx = keys[take][order], idx[take][order]
    kth = int(okeys[k - 1]) if n > k else int(NAN_KEY)
    end = kprime
    if n > k:
        later = np.nonzero(okeys[k:] > kth)[0]
        if later.size:
            end = k + int(later[0])
    return okeys[:end], oidx[:end], kth


def reference(keys, idx, k):
    order = np.lexsort((idx, keys))
    okeys, oidx = keys[order], idx[order]
    if keys.size > k:
        kth = okeys[k - 1]
        keep = okeys <= kth
        return okeys[keep], oidx[keep], int(kth)
    return okeys, oidx, int(NAN_KEY)


def cases():
    rng = np.random.default_rng(7)
    yield "gauss", rng.standard_normal(3000).astype(np.float32), 1000
    yield "tail", np.sort(rng.standard_normal(200000).astype(np.float32))[-2500:], 1000
    yield "coarse grid", (rng.integers(-40, 40, 4000) / 64).astype(np.float32), 1000
    yield "two values", rng.choice(np.array([0.25, 0.5], np.float32), 3000), 100
    yield "constant", np.full(500, 0.125, np.float32), 10
    yield "signed zero", rng.choice(np.array([0.0, -0.0, 1.0, -1.0], np.float32), 2000), 700
    special = rng.standard_normal(3000).astype(np.float32)
    special[::97] = np.nan
    special[5] = np.inf
    special[9] = -np.inf
    special[11:14] = np.float32(1e-42)       # denormals
    yield "nan inf denormal", special, 1000
    yield "huge range", np.concatenate([rng.standard_normal(1500), [3e38, -3e38]]).astype(np.float32), 500
    yield "k >= n", rng.standard_normal(300).astype(np.float32), 1000
    padded = rng.standard_normal(1000).astype(np.float32)
    yield "exactly k real + padding", np.concatenate([padded, np.full(224, np.nan, np.float32)]), 1000


@pytest.mark.parametrize("name,scores,k", list(cases()), ids=[c[0] for c in cases()])
def test_subbin_is_monotone_in_the_key(name, scores, k):
    keys = score_to_key(scores)
    hi, scale = plan(keys)
    order = np.argsort(keys, kind="stable")
    sb = subbin(keys[order], hi, scale)
    assert np.all(np.diff(sb) >= 0), "a better key landed in a later sub-bin"
    same = np.diff(keys[order].astype(np.int64)) == 0
    assert np.all(np.diff(sb)[same] == 0), "equal keys in different sub-bins"
    real = keys[order] != NAN_KEY
    if scale > 0:
        assert sb[real].max() <= BINS - 2 or np.isinf(scores).any() or np.abs(scores[~np.isnan(scores)]).max() > 1e38


@pytest.mark.parametrize("name,scores,k", list(cases()), ids=[c[0] for c in cases()])
def test_fused_selection_equals_sort_with_keep_all(name, scores, k):
    keys = score_to_key(scores)
    idx = np.arange(keys.size, dtype=np.int64)
    got = fused_select(keys, idx, k)
    want = reference(keys, idx, k)
    assert got[2] == want[2]
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1], want[1])
