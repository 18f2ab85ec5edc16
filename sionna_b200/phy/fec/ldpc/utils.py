"""Decoder callbacks and the ragged message container they operate on (mirror of
/root/reference/src/sionna/phy/fec/ldpc/utils.py:12-260).

The reference hands callbacks a ``tf.RaggedTensor`` of shape ``[num_nodes, None, batch_size]``. `RaggedMessages` is the
torch counterpart used by the unfused decoder path (``csrc/ldpc_bp_flat.cu``): the same flat ``[num_edges, batch]``
value tensor plus the row partition, with the handful of operations the reference's callbacks and typical user node
updates need (``flat_values``, ``with_flat_values``, ``value_rowids``, ``row_lengths``, segment reductions over the ragged
axis). Callbacks are Python-level analysis hooks: they run torch operations between the decoder's kernel launches.
"""
import numpy as np
import torch

from ...block import Object
from ..utils import llr2mi


class RaggedMessages:
    """Messages of all nodes: ``flat_values [num_edges, batch]`` (row-major over nodes, then the node's edges in list
    order) and ``row_splits [num_nodes + 1]``; ``shape == (num_nodes, None, batch)``."""

    def __init__(self, flat_values, row_splits):
        self.flat_values = flat_values
        self.row_splits = row_splits                                   # int64 tensor on the values' device
        self._lengths = None
        self._rowids = None

    @property
    def shape(self):
        return (int(self.row_splits.shape[0]) - 1, None, int(self.flat_values.shape[-1]))

    @property
    def dtype(self):
        return self.flat_values.dtype

    def nrows(self):
        return int(self.row_splits.shape[0]) - 1

    def row_lengths(self):
        if self._lengths is None:
            self._lengths = self.row_splits[1:] - self.row_splits[:-1]
        return self._lengths

    def value_rowids(self):
        if self._rowids is None:
            self._rowids = torch.repeat_interleave(torch.arange(self.nrows(), device=self.flat_values.device),
                                                   self.row_lengths())
        return self._rowids

    def with_flat_values(self, flat_values):
        r = RaggedMessages(flat_values, self.row_splits)
        r._lengths, r._rowids = self._lengths, self._rowids
        return r

    def map_flat_values(self, fn, *args, **kwargs):
        """``tf.ragged.map_flat_values(fn, self)``."""
        return self.with_flat_values(fn(self.flat_values, *args, **kwargs))

    def _reduce(self, how):
        return torch.segment_reduce(self.flat_values, how, lengths=self.row_lengths(), axis=0, unsafe=True)

    def reduce_sum(self):
        """Sum over the ragged axis -> ``[num_nodes, batch]`` (``tf.reduce_sum(x, axis=1)``)."""
        out = torch.zeros((self.nrows(),) + tuple(self.flat_values.shape[1:]), dtype=self.flat_values.dtype,
                          device=self.flat_values.device)
        return out.index_add_(0, self.value_rowids(), self.flat_values)

    def reduce_prod(self):
        return self._reduce("prod")

    def reduce_min(self):
        return self._reduce("min")

    def reduce_max(self):
        return self._reduce("max")

    def gather_rows(self, node_values):
        """Broadcast one value per node ``[num_nodes, batch]`` back to the edges ``[num_edges, batch]``."""
        return node_values.index_select(0, self.value_rowids())

    def __mul__(self, other):
        return self.with_flat_values(self.flat_values * other)

    __rmul__ = __mul__


class EXITCallback:
    """EXITCallback(num_iter): tracks the mutual information of the messages after every iteration (utils.py:12-54);
    requires all-zero codeword simulations. Register in ``v2c_callbacks`` and / or ``c2v_callbacks``."""

    def __init__(self, num_iter):
        self._mi = torch.zeros(num_iter + 1, dtype=torch.float64)
        self._num_samples = torch.zeros(num_iter + 1, dtype=torch.float64)

    @property
    def mi(self):
        """Mutual information after each iteration"""
        return (self._mi / self._num_samples).to(torch.float32)

    def __call__(self, msg, it, *args, **kwargs):
        self._mi[it] += float(llr2mi(-1.0 * msg.flat_values))
        self._num_samples[it] += 1.0
        return msg


class DecoderStatisticsCallback:
    """DecoderStatisticsCallback(num_iter): counts, per iteration, the codewords whose check nodes are all satisfied
    (utils.py:56-153). Register in ``c2v_callbacks``."""

    def __init__(self, num_iter):
        self._num_iter = num_iter
        self.reset_stats()

    @property
    def num_samples(self):
        """Total number of processed codewords"""
        return self._num_samples

    @property
    def num_decoded_cws(self):
        """Number of decoded codewords after each iteration"""
        return self._decoded_samples

    @property
    def success_rate(self):
        """Success rate after each iteration"""
        return self._decoded_samples.to(torch.float64) / self._num_samples.to(torch.float64)

    @property
    def avg_number_iterations(self):
        """Average number of decoding iterations"""
        num_decoded = self._decoded_samples.to(torch.float64)
        num_samples = self._num_samples.to(torch.float64)
        return (num_samples - num_decoded).sum() / num_samples[0]

    def reset_stats(self):
        """Reset internal statistics"""
        self._num_samples = torch.zeros(self._num_iter, dtype=torch.int64)
        self._decoded_samples = torch.zeros(self._num_iter, dtype=torch.int64)

    def __call__(self, msg, it, *args, **kwargs):
        # sign of every message with sign(0) := +1, product over the node, all nodes of a codeword positive
        negative = (msg.flat_values < 0).to(torch.float32)
        odd = msg.with_flat_values(negative).reduce_sum().remainder(2.0)      # parity of the negative signs per node
        cw_success = (odd == 0).all(dim=0)
        self._num_samples[it] += int(msg.flat_values.shape[-1])
        self._decoded_samples[it] += int(cw_success.sum())
        return msg


class WeightedBPCallback(Object):
    """WeightedBPCallback(num_edges): multiplies every message by a per-edge weight (weighted BP [Nachmani],
    utils.py:155-260). The weights are a plain tensor (``weights``); the decoder kernels are not differentiable, so the
    weights can be set / loaded but not trained through this package."""

    def __init__(self, num_edges, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        self._edge_weights = torch.ones(num_edges, dtype=self.rdtype)

    @property
    def weights(self):
        return self._edge_weights

    def show_weights(self, size=7):
        raise NotImplementedError("plotting helpers are not part of this package")

    def __call__(self, msg, *args):
        w = self._edge_weights.to(device=msg.flat_values.device, dtype=msg.flat_values.dtype)
        return msg.with_flat_values(msg.flat_values * w[:, None])
