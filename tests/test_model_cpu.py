"""CPU checks of the model layer's host logic (src/model.jl:15-19, :49-51, :79-90): pair ordering and the lookup."""
import numpy as np
import pytest


def test_dense_index_pairs_is_the_column_major_product_order():
    import spectralkernels_jl_b200 as sk
    # Julia: vec(collect(Iterators.product(1:3, 1:3))) filtered by x[1] <= x[2] gives
    # (1,1) (1,2) (2,2) (1,3) (2,3) (3,3); 0-based here
    assert sk.dense_index_pairs(np.zeros(3)).tolist() == [[0, 0], [0, 1], [1, 1], [0, 2], [1, 2], [2, 2]]
    pr = sk.dense_index_pairs(np.zeros((57, 2)))
    assert pr.shape == (57 * 58 // 2, 2) and np.all(pr[:, 0] <= pr[:, 1])
    assert np.all(np.diff(pr[:, 1]) >= 0)                     # second index slowest


def test_spectral_kernel_lookup_and_matrix():
    import spectralkernels_jl_b200 as sk
    pts = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 2.0]])
    pairs = sk.dense_index_pairs(pts)
    vals = np.arange(pairs.shape[0], dtype=float)
    k = sk.SpectralKernel(pts, pairs, vals)
    assert k(pts[0], pts[1]) == k(pts[1], pts[0]) == 1.0 and k(pts[2], pts[2], "params") == 5.0
    with pytest.raises(KeyError):
        k(np.array([9.0, 9.0]), pts[0])
    M = k.matrix()
    assert np.array_equal(M, M.T) and M[1, 2] == 4.0
    assert len(k.store) == pairs.shape[0]
