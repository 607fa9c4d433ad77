"""Host-side sampler (tlod_anchor_subsample_host / tlod_numpy_permutation) against numpy's own
calls: same permutations, same labels, and the same position of numpy's global stream afterwards
(lib/model/rpn/anchor_target_layer.py:118-145 consumes np.random.permutation)."""
import ctypes

import numpy as np
import pytest

from model.rpn import anchor_target_layer as atl
from tlod_b200._lib import check, lib


def _native_permutation(n):
    addr = atl._numpy_mt19937_address()
    out = np.empty(n, np.int64)
    check(lib.tlod_numpy_permutation(addr, ctypes.cast(addr + 4 * 624, ctypes.POINTER(ctypes.c_int)), n,
                                     out.ctypes.data), "tlod_numpy_permutation")
    return out


def test_numpy_state_is_reachable():
    assert atl._numpy_mt19937_address() is not None
    assert atl._native_sampler_matches_numpy()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 8, 9, 17, 128, 129, 1000, 17434, 65536, 65537, 70000])
@pytest.mark.parametrize("seed", [0, 3, 12345])
def test_permutation_and_stream_position_equal_numpy(n, seed):
    np.random.seed(seed)
    a, a2, tail_a = np.random.permutation(n), np.random.permutation(max(n // 2, 1)), np.random.rand(3)
    sa = np.random.get_state()
    np.random.seed(seed)
    b, b2, tail_b = _native_permutation(n), _native_permutation(max(n // 2, 1)), np.random.rand(3)
    sb = np.random.get_state()
    assert np.array_equal(a, b) and np.array_equal(a2, b2) and np.array_equal(tail_a, tail_b)
    assert sa[2] == sb[2] and np.array_equal(sa[1], sb[1])


@pytest.mark.parametrize("B,n,p", [(1, 17434, (0.02, 0.97, 0.01)), (2, 17434, (0.3, 0.69, 0.01)),
                                   (3, 500, (0.1, 0.2, 0.7)), (2, 300, (0.9, 0.08, 0.02)), (1, 1, (0.0, 1.0, 0.0)),
                                   (4, 4000, (0.0, 0.5, 0.5))])
def test_subsample_equals_reference_transcription(B, n, p):
    """Too many fg and bg, too few of either, none at all: labels, the last image's example count
    and the stream position must equal the numpy transcription of the reference loop."""
    rs = np.random.RandomState(B * 1000 + n)
    lab = rs.choice(np.array([-1.0, 0.0, 1.0], np.float32), size=(B, n), p=list(p))
    for seed in (3, 11):
        a, b = lab.copy(), lab.copy()
        np.random.seed(seed)
        ra = atl.subsample_labels_numpy(a, 128, 256)
        xa = np.random.rand()
        np.random.seed(seed)
        rb = atl.subsample_labels(b, 128, 256)
        xb = np.random.rand()
        assert ra == rb and np.array_equal(a, b) and xa == xb
        assert (b == 1).sum(1).max() <= 128 and ((b == 1).sum(1) + (b == 0).sum(1)).max() <= 256
