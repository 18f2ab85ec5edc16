"""Host-side checks (CPU tensors) of the decoder-callback helpers, mirror of /root/reference/src/sionna/phy/fec/ldpc/utils.py:
the ragged message container and the three callbacks operate on plain torch tensors, so their arithmetic can be pinned
without a GPU (the GPU tests run them inside the unfused decoder)."""
import numpy as np
import torch

from sionna_b200.phy.fec.ldpc.utils import RaggedMessages, EXITCallback, DecoderStatisticsCallback, WeightedBPCallback
from sionna_b200.phy.fec.utils import llr2mi


def _ragged(rng, lengths, batch):
    splits = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    vals = rng.normal(size=(splits[-1], batch)).astype(np.float32)
    return RaggedMessages(torch.from_numpy(vals), torch.from_numpy(splits)), vals, splits


def test_ragged_messages_reductions_match_numpy():
    rng = np.random.default_rng(0)
    lengths = [3, 1, 5, 2, 7, 1]
    msg, vals, splits = _ragged(rng, lengths, 4)
    assert msg.shape == (6, None, 4) and msg.nrows() == 6
    assert msg.row_lengths().tolist() == lengths
    assert msg.value_rowids().tolist() == sum(([i] * n for i, n in enumerate(lengths)), [])
    rows = [vals[splits[i]:splits[i + 1]] for i in range(6)]
    np.testing.assert_allclose(msg.reduce_sum().numpy(), np.stack([r.sum(0) for r in rows]), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(msg.reduce_prod().numpy(), np.stack([r.prod(0) for r in rows]), rtol=1e-5, atol=1e-7)
    assert np.array_equal(msg.reduce_min().numpy(), np.stack([r.min(0) for r in rows]))
    assert np.array_equal(msg.reduce_max().numpy(), np.stack([r.max(0) for r in rows]))
    per_node = torch.arange(24, dtype=torch.float32).reshape(6, 4)
    assert np.array_equal(msg.gather_rows(per_node).numpy(), np.repeat(per_node.numpy(), lengths, axis=0))
    doubled = msg.map_flat_values(lambda v, k: v * k, 2.0)
    assert torch.equal(doubled.flat_values, msg.flat_values * 2) and doubled.row_splits is msg.row_splits
    assert torch.equal((3.0 * msg).flat_values, msg.flat_values * 3)


def test_decoder_statistics_callback_counts_satisfied_codewords():
    """utils.py:56-153: a codeword counts as decoded in iteration `it` when every check node sees an even number of
    negative messages (sign(0) := +)."""
    lengths = [3, 2, 4]
    splits = torch.tensor(np.concatenate([[0], np.cumsum(lengths)]))
    #            codeword:  0     1     2
    flat = torch.tensor([[1.0, -1.0, 1.0],     # check 0
                         [-2.0, -1.0, 0.0],
                         [-3.0, 4.0, 2.0],
                         [1.0, 1.0, -1.0],     # check 1
                         [5.0, -1.0, -2.0],
                         [-1.0, 1.0, 1.0],     # check 2
                         [-1.0, 1.0, -0.0],
                         [2.0, 1.0, 3.0],
                         [2.0, 1.0, 1.0]])
    cb = DecoderStatisticsCallback(num_iter=3)
    out = cb(RaggedMessages(flat, splits), 1)
    assert torch.equal(out.flat_values, flat)
    # codeword 0: checks have 2, 0, 2 negatives -> satisfied; codeword 1: 2, 1, 0 -> not; codeword 2: 0, 2, 0 -> satisfied
    assert cb.num_samples.tolist() == [0, 3, 0] and cb.num_decoded_cws.tolist() == [0, 2, 0]
    cb(RaggedMessages(flat, splits), 0)
    cb(RaggedMessages(flat.abs(), splits), 2)
    assert cb.num_decoded_cws.tolist() == [2, 2, 3]
    np.testing.assert_allclose(cb.success_rate.numpy(), [2 / 3, 2 / 3, 1.0])
    np.testing.assert_allclose(float(cb.avg_number_iterations), (1 + 1 + 0) / 3)
    cb.reset_stats()
    assert cb.num_samples.sum() == 0


def test_exit_callback_tracks_mutual_information():
    """utils.py:12-54: mi[it] is the running mean of llr2mi(-msg) over the calls of iteration `it`."""
    rng = np.random.default_rng(1)
    msg, vals, _ = _ragged(rng, [4, 4, 4], 64)
    cb = EXITCallback(num_iter=2)
    cb(msg, 0)
    cb(msg * 4.0, 0)
    cb(msg * 4.0, 2)
    want0 = 0.5 * (float(llr2mi(torch.from_numpy(-vals))) + float(llr2mi(torch.from_numpy(-4.0 * vals))))
    np.testing.assert_allclose(float(cb.mi[0]), want0, rtol=1e-6)
    np.testing.assert_allclose(float(cb.mi[2]), float(llr2mi(torch.from_numpy(-4.0 * vals))), rtol=1e-6)
    assert np.isnan(float(cb.mi[1]))                                        # no sample for iteration 1: 0 / 0 as in the reference


def test_weighted_bp_callback_scales_edges():
    rng = np.random.default_rng(2)
    msg, vals, _ = _ragged(rng, [2, 3], 5)
    cb = WeightedBPCallback(num_edges=5)
    assert torch.equal(cb(msg).flat_values, msg.flat_values)                # weights start at one
    cb.weights[:] = torch.tensor([0.5, 1.0, 2.0, 0.0, -1.0])
    np.testing.assert_allclose(cb(msg).flat_values.numpy(), vals * np.array([0.5, 1.0, 2.0, 0.0, -1.0], np.float32)[:, None])


def test_j_function_and_inverse():
    """fec/utils.py:184-267 (Brannstrom approximation): J(0+) = 0, J(inf) = 1, monotone, J^-1(J(mu)) = mu; the inverse is
    clipped to 20 and both clip their arguments like the reference (1e-10 .. 1000, 1e-10 .. 1)."""
    from sionna_b200.phy.fec.utils import j_fun, j_fun_inv
    mu = np.array([1e-3, 0.1, 0.5, 1.0, 2.0, 4.0, 8.0, 15.0])
    mi = j_fun(mu)
    assert np.all(np.diff(mi) > 0) and mi[0] < 1e-3 and mi[-1] > 0.98 and j_fun(100.0) > 0.999999
    np.testing.assert_allclose(j_fun_inv(mi), mu, rtol=1e-6)
    h1, h2, h3 = 0.3073, 0.8935, 1.1064
    np.testing.assert_allclose(j_fun(1.0), (1 - 2 ** (-h1 * 2.0 ** h2)) ** h3, rtol=1e-12)
    assert j_fun(-5.0) == j_fun(1e-10) and j_fun(1e9) == j_fun(1000.0)
    assert j_fun_inv(1.0) == 20.0 and j_fun_inv(2.0) == 20.0 and j_fun_inv(-1.0) == j_fun_inv(1e-10)


def test_llr2mi_of_consistent_gaussian_llrs_follows_the_j_function():
    """Logits (log p(1)/p(0)) of an all-zero codeword, ~ N(-mu, 2 mu): llr2mi estimates the mutual information that the
    J-function approximates (fec/utils.py:116-182; the EXIT callback negates the decoder's internal messages, which use
    the opposite sign, before the call)."""
    from sionna_b200.phy.fec.utils import j_fun
    rng = np.random.default_rng(5)
    for mu in (0.5, 2.0, 6.0):
        logits = torch.from_numpy(rng.normal(-mu, np.sqrt(2 * mu), size=400000).astype(np.float32))
        mi = float(llr2mi(logits))
        assert abs(mi - float(j_fun(mu))) < 0.01
        signs = torch.from_numpy(rng.choice([-1.0, 1.0], size=logits.shape).astype(np.float32))
        np.testing.assert_allclose(float(llr2mi(signs * logits, s=signs)), mi, rtol=1e-6)       # sign-adjusted variant
    per_row = llr2mi(torch.zeros(3, 8), reduce_dims=False)
    assert per_row.shape == (3,) and torch.allclose(per_row, torch.zeros(3))                    # llr = 0: no information
    try:
        llr2mi(torch.zeros(4, dtype=torch.int32))
        assert False
    except TypeError:
        pass
