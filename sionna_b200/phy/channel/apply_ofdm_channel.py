"""Frequency-domain channel application (mirror of /root/reference/src/sionna/phy/channel/apply_ofdm_channel.py:10-80)."""
import torch

from ..block import Block
from ..config import config
from ..mapping import _broadcast_inner
from ..._lib import lib, check, ptr, current_stream


class ApplyOFDMChannel(Block):
    """``y[b, rx, rx_ant, s, f] = sum_{tx, tx_ant} h[b, rx, rx_ant, tx, tx_ant, s, f] x[b, tx, tx_ant, s, f] + w`` with
    ``w ~ CN(0, no)``; ``call(x, h_freq, no=None)``, ``no`` broadcast from the left over ``[b, rx, rx_ant, s, f]``."""

    def __init__(self, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)

    def call(self, x, h_freq, no=None):
        if self.precision != "single":
            raise NotImplementedError("sb_apply_ofdm_channel is a complex64 kernel.")
        dev = self.device
        x = x.to(device=dev, dtype=torch.complex64).contiguous()
        h = h_freq.to(device=dev, dtype=torch.complex64).contiguous()
        b, rx, ra, tx, ta, s_, f_ = h.shape
        if tuple(x.shape) != (b, tx, ta, s_, f_):
            raise ValueError(f"x of shape {tuple(x.shape)} does not match h_freq of shape {tuple(h.shape)} "
                             "([batch, num_rx, num_rx_ant, num_tx, num_tx_ant, num_ofdm_symbols, fft_size])")
        y = torch.empty((b, rx, ra, s_, f_), dtype=torch.complex64, device=dev)
        no_t, inner, add = None, 1, 0
        seed, off = 0, 0
        if no is not None:
            no_t, inner = _broadcast_inner(no, y.shape, dev, torch.float32)
            add = 1
            seed, off = config.next_philox()
        check(lib().sb_apply_ofdm_channel(ptr(x), ptr(h), ptr(no_t), inner, ptr(y), b, rx * ra, tx * ta, s_ * f_, add,
                                          seed, off, current_stream()), "sb_apply_ofdm_channel")
        return y
