"""OFDM modulator (mirror of /root/reference/src/sionna/phy/ofdm/modulator.py:13-124)."""
import numpy as np
import torch

from ..block import Block
from ..._lib import lib, check, ptr, current_stream


class OFDMModulator(Block):
    """OFDMModulator(cyclic_prefix_length=0): ``[..., num_ofdm_symbols, fft_size]`` frequency-domain grid (DC in the
    centre) -> ``[..., num_ofdm_symbols*fft_size + sum(cyclic_prefix_length)]`` time samples. ``cyclic_prefix_length`` is an
    int or one value per OFDM symbol (modulator.py:42-124)."""

    def __init__(self, cyclic_prefix_length=0, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        self.cyclic_prefix_length = cyclic_prefix_length
        self._tabs = None

    @property
    def cyclic_prefix_length(self):
        return self._cyclic_prefix_length

    @cyclic_prefix_length.setter
    def cyclic_prefix_length(self, value):
        value = np.asarray(value)
        if not (np.issubdtype(value.dtype, np.integer) or np.all(value == value.astype(int))):
            raise ValueError("`cyclic_prefix_length` must be integer.")
        value = value.astype(np.int32)
        if np.any(value < 0):
            raise ValueError("`cyclic_prefix_length` must be nonnegative.")
        if value.ndim > 1:
            raise ValueError("`cyclic_prefix_length` must be of rank 0 or 1.")
        self._cyclic_prefix_length = value
        self._tabs = None

    def build(self, input_shape):
        nsym, fft_size = input_shape[-2], input_shape[-1]
        cp = self._cyclic_prefix_length
        if np.any(cp > fft_size):
            raise ValueError("shape(inputs)[-1] must not be smaller than `cylic_prefix_length`")
        if cp.ndim == 1 and cp.shape[0] != nsym:
            raise ValueError("`cyclic_prefix_length` must be of size [num_ofdm_symbols]")

    def call(self, inputs):
        if self.precision != "single":
            raise NotImplementedError("sb_ofdm_modulate is a complex64 kernel.")
        dev = self.device
        x = inputs.to(device=dev, dtype=torch.complex64).contiguous()
        nsym, n = x.shape[-2], x.shape[-1]
        cp = np.broadcast_to(self._cyclic_prefix_length, (nsym,)).astype(np.int32)
        off = np.concatenate([[0], np.cumsum(n + cp)[:-1]]).astype(np.int32)
        out_len = int(np.sum(n + cp))
        if self._tabs is None or self._tabs[2] != (nsym, n, dev):
            self._tabs = (torch.from_numpy(np.ascontiguousarray(cp)).to(dev), torch.from_numpy(off).to(dev), (nsym, n, dev))
        rows = x.numel() // (nsym * n)
        out = torch.empty(list(x.shape[:-2]) + [out_len], dtype=torch.complex64, device=dev)
        check(lib().sb_ofdm_modulate(ptr(x), ptr(out), rows, nsym, n, ptr(self._tabs[0]), ptr(self._tabs[1]), out_len, 1,
                                     current_stream()), "sb_ofdm_modulate")
        return out
