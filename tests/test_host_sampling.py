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


def test_numpy_fallback_paths_stay_usable(monkeypatch):
    """The native samplers write into numpy's private MT19937 state (a contract with numpy internals);
    when the one-time self check fails they fall back to the numpy transcriptions.  Force that fallback
    and require the same labels / samples / stream position as the native path."""
    from model.rpn import proposal_target_layer_cascade as ptl
    from model.utils.config import cfg
    rs = np.random.RandomState(7)
    lab = rs.choice(np.array([-1.0, 0.0, 1.0], np.float32), size=(2, 5000), p=[0.1, 0.8, 0.1])
    mo = (rs.rand(3, 900) ** 2).astype(np.float32)
    st = {"copied": None, "host": __import__("torch").from_numpy(mo)}
    layer = ptl._ProposalTargetLayer(9)
    old_bs = cfg.TRAIN.BATCH_SIZE
    cfg.TRAIN.BATCH_SIZE = 64
    try:
        results = []
        for native in (True, False):
            monkeypatch.setattr(atl, "_native_ok", native)
            monkeypatch.setattr(ptl, "_native_ok", native)
            a = lab.copy()
            np.random.seed(21)
            n_ex = atl.subsample_labels(a, 128, 256)
            keep, fg = layer.sample(st)
            results.append((a, n_ex, keep.copy(), fg.copy(), np.random.get_state()))
        (a0, n0, k0, f0, s0), (a1, n1, k1, f1, s1) = results
        assert np.array_equal(a0, a1) and n0 == n1 and np.array_equal(k0, k1) and np.array_equal(f0, f1)
        assert s0[2] == s1[2] and np.array_equal(s0[1], s1[1])
    finally:
        cfg.TRAIN.BATCH_SIZE = old_bs


@pytest.mark.parametrize("seed", range(6))
def test_proposal_sampling_native_equals_numpy(seed):
    """tlod_proposal_sample_host vs the numpy transcription of proposal_target_layer_cascade.py:140-181:
    fg + bg, fg only, bg only images; permutation + rand draws; same stream position afterwards."""
    from model.rpn import proposal_target_layer_cascade as ptl
    r = np.random.RandomState(seed)
    mo = (r.rand(4, 600) ** (1 + seed % 3)).astype(np.float32)
    mo[1] *= 0.3   # no foreground
    mo[2] = 0.9    # no background
    for rpi, fgr in ((256, 64), (16, 4), (1, 1)):
        np.random.seed(seed)
        a = ptl._sample_numpy(mo, rpi, fgr)
        sa = np.random.get_state()
        np.random.seed(seed)
        b = ptl._sample_native(mo, rpi, fgr)
        sb = np.random.get_state()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert sa[2] == sb[2] and np.array_equal(sa[1], sb[1])
    with pytest.raises(ValueError):
        ptl._sample_native(np.full((1, 10), 0.05, np.float32) * 0 - 1, 4, 1)  # neither fg nor bg


@pytest.mark.parametrize("blocks", [0, 1, 5, 40, 400])
def test_pregenerated_blocks_give_the_same_labels_and_stream(blocks):
    """tlod_mt_pregen + tlod_anchor_subsample_host_ahead: key blocks generated before the labels
    arrive.  Too few blocks (the stream continues in numpy's own key), exactly enough, far too many:
    labels and numpy's state afterwards always equal numpy's own calls."""
    from model.rpn import anchor_target_layer as atl
    if not atl._native_sampler_matches_numpy():
        pytest.skip("numpy without the ctypes MT19937 state interface")
    rs = np.random.RandomState(5)
    lab = rs.choice(np.array([-1.0, 0.0, 1.0], np.float32), size=(3, 9000), p=[0.2, 0.7, 0.1])
    saved = np.random.get_state()
    try:
        a, b = lab.copy(), lab.copy()
        np.random.seed(31)
        np.random.rand(100)  # start in the middle of a key block
        ra = atl.subsample_labels_numpy(a, 128, 256)
        sa = np.random.get_state()
        np.random.seed(31)
        np.random.rand(100)
        ahead = atl.pregenerate_stream(blocks * 624) if blocks else None
        before = np.random.get_state()
        if ahead is not None:
            assert ahead.size == (blocks + 1) * 624 and np.array_equal(ahead[:624], before[1])
        rb = atl.subsample_labels(b, 128, 256, ahead)
        sb = np.random.get_state()
        assert ra == rb and np.array_equal(a, b)
        assert sa[2] == sb[2] and np.array_equal(sa[1], sb[1])
        assert np.random.randint(0, 1 << 30) >= 0 and np.array_equal(np.random.get_state()[1], np.random.get_state()[1])
    finally:
        np.random.set_state(saved)


def test_stale_pregenerated_blocks_are_ignored():
    """Somebody draws from numpy's stream between pregenerate_stream and the subsampling: the blocks
    no longer start at numpy's key and must not be used."""
    from model.rpn import anchor_target_layer as atl
    if not atl._native_sampler_matches_numpy():
        pytest.skip("numpy without the ctypes MT19937 state interface")
    rs = np.random.RandomState(6)
    lab = rs.choice(np.array([-1.0, 0.0, 1.0], np.float32), size=(2, 5000), p=[0.2, 0.7, 0.1])
    saved = np.random.get_state()
    try:
        a, b = lab.copy(), lab.copy()
        np.random.seed(32)
        np.random.rand(700)
        ra = atl.subsample_labels_numpy(a, 128, 256)
        sa = np.random.get_state()
        np.random.seed(32)
        ahead = atl.pregenerate_stream(30 * 624)
        np.random.rand(700)  # crosses a refill: numpy's key changes
        rb = atl.subsample_labels(b, 128, 256, ahead)
        sb = np.random.get_state()
        assert ra == rb and np.array_equal(a, b) and sa[2] == sb[2] and np.array_equal(sa[1], sb[1])
    finally:
        np.random.set_state(saved)


def test_stream_count_blocks_of_sixteen_against_numpy_for_many_sizes():
    """The vectorised 'count only' part of the keep-last permutation (sixteen draws at a time, blocks
    with a position-dependent draw resolved one by one): sizes around the power-of-two epochs."""
    from model.rpn import anchor_target_layer as atl
    if not atl._native_sampler_matches_numpy():
        pytest.skip("numpy without the ctypes MT19937 state interface")
    saved = np.random.get_state()
    try:
        for n in (130, 257, 300, 1023, 1024, 1025, 4097, 16383, 16385, 17434, 33300):
            lab = np.zeros((1, n), np.float32)  # all background: one permutation of n, keep 256
            lab[0, :5] = 1.0
            a, b = lab.copy(), lab.copy()
            np.random.seed(n)
            ra = atl.subsample_labels_numpy(a, 128, 256)
            sa = np.random.get_state()
            np.random.seed(n)
            rb = atl.subsample_labels(b, 128, 256)
            sb = np.random.get_state()
            assert ra == rb and np.array_equal(a, b), n
            assert sa[2] == sb[2] and np.array_equal(sa[1], sb[1]), n
    finally:
        np.random.set_state(saved)
