"""The synthetic operator of BASELINE config 5: numpy twin of the device generator (host side only)."""
import numpy as np

import arpack_ng_b200 as ab


def test_randsparse_rule():
    m, k, per = 300, 64, 16
    A = ab.randsparse_numpy(0, m, k, per)
    assert A.shape == (m, k)
    # exactly `per` draws per row (duplicate columns are summed, so stored entries <= per)
    assert (np.diff(A.indptr) <= per).all() and np.diff(A.indptr).min() >= 1
    # row blocks are slices of the same matrix (what lets every rank generate its own shard)
    B = ab.randsparse_numpy(120, 100, k, per)
    assert (A[120:220] != B).nnz == 0
    # the rule itself, entry (row 7, draw 3): col = splitmix64(seed + per*7 + 3) mod k, value from the next hash
    def mix(x):
        x = (x + 0x9E3779B97F4A7C15) & (2**64 - 1)
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        return x ^ (x >> 31)
    row = np.zeros(k)
    for d in range(per):
        h1 = mix(0x5EED + per * 7 + d)
        row[h1 % k] += 2.0 * ((mix(h1) >> 11) / 9007199254740992.0) - 1.0
    assert np.array_equal(A[7].toarray().ravel(), row)
    assert np.abs(A.data).max() <= 16.0 and A.data.min() < 0 < A.data.max()
