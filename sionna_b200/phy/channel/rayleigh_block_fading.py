"""RayleighBlockFading (mirror of /root/reference/src/sionna/phy/channel/rayleigh_block_fading.py:13-115)."""
import torch

from ..block import Block
from ..config import config
from ..utils.misc import complex_normal


class RayleighBlockFading(Block):
    """RayleighBlockFading(num_rx, num_rx_ant, num_tx, num_tx_ant, precision=None): single-path i.i.d. CN(0, 1) gains,
    constant over the block. ``__call__(batch_size, num_time_steps, sampling_frequency=None)`` -> ``a [batch, num_rx,
    num_rx_ant, num_tx, num_tx_ant, 1, num_time_steps]``, ``tau [batch, num_rx, num_tx, 1]`` (zeros)."""

    def __init__(self, num_rx, num_rx_ant, num_tx, num_tx_ant, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        self.num_rx, self.num_rx_ant, self.num_tx, self.num_tx_ant = num_rx, num_rx_ant, num_tx, num_tx_ant

    def __call__(self, batch_size, num_time_steps, sampling_frequency=None):
        return self._invoke(batch_size, num_time_steps, sampling_frequency)

    def call(self, batch_size, num_time_steps, sampling_frequency=None):
        h = complex_normal([batch_size, self.num_rx, self.num_rx_ant, self.num_tx, self.num_tx_ant, 1, 1])
        h = h.expand(-1, -1, -1, -1, -1, -1, int(num_time_steps)).contiguous()
        tau = torch.zeros((batch_size, self.num_rx, self.num_tx, 1), dtype=torch.float32, device=config.device)
        return h, tau
