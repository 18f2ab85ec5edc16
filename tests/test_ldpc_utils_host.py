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
