"""OFDM demodulator (mirror of /root/reference/src/sionna/phy/ofdm/demodulator.py:13-203)."""
import numpy as np
import torch

from ..block import Block
from ..._lib import lib, check, ptr, current_stream


class OFDMDemodulator(Block):
    """OFDMDemodulator(fft_size, l_min, cyclic_prefix_length=0): ``[..., num_samples]`` time samples ->
    ``[..., num_ofdm_symbols, fft_size]``; trailing samples that do not fill an OFDM symbol are dropped, the cyclic
    prefix is removed, the FFT output is multiplied by ``exp(-j 2 pi k l_min / fft_size)`` to undo the timing offset of a
    channel whose first tap has (negative) index ``l_min`` and finally fft-shifted (demodulator.py:86-203)."""

    def __init__(self, fft_size, l_min, cyclic_prefix_length=0, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        self._fft_size = int(fft_size)
        self._l_min = int(l_min)
        cp = np.asarray(cyclic_prefix_length)
        if np.any(cp < 0):
            raise ValueError("`cyclic_prefix_length` must be nonnegative.")
        if cp.ndim > 1:
            raise ValueError("`cyclic_prefix_length` must be of rank 0 or 1.")
        self._cyclic_prefix_length = cp.astype(np.int32)
        self._tabs = None

    @property
    def fft_size(self):
        return self._fft_size

    @property
    def l_min(self):
        return self._l_min

    @property
    def cyclic_prefix_length(self):
        return self._cyclic_prefix_length

    def _layout(self, num_samples):
        n, cp = self._fft_size, self._cyclic_prefix_length
        if cp.ndim == 0:
            nsym = num_samples // (n + int(cp))                          # remainder dropped (:138-144)
            cps = np.full(nsym, int(cp), np.int32)
        else:
            cps = cp.astype(np.int32)
            nsym = len(cps)
            if int(np.sum(n + cps)) > num_samples:
                raise ValueError("shape(inputs)[-1] must be larger or equal than the total symbol length.")
        off = np.concatenate([[0], np.cumsum(n + cps)[:-1]]).astype(np.int32) if nsym else np.zeros(0, np.int32)
        return nsym, cps, off

    def call(self, inputs):
        if self.precision != "single":
            raise NotImplementedError("sb_ofdm_demodulate is a complex64 kernel.")
        dev = self.device
        x = inputs.to(device=dev, dtype=torch.complex64).contiguous()
        ns = x.shape[-1]
        nsym, cps, off = self._layout(ns)
        if self._tabs is None or self._tabs[2] != (ns, dev):
            self._tabs = (torch.from_numpy(np.ascontiguousarray(cps)).to(dev), torch.from_numpy(off).to(dev), (ns, dev))
        rows = x.numel() // ns
        out = torch.empty(list(x.shape[:-1]) + [nsym, self._fft_size], dtype=torch.complex64, device=dev)
        check(lib().sb_ofdm_demodulate(ptr(x), ptr(out), rows, nsym, self._fft_size, ptr(self._tabs[0]),
                                       ptr(self._tabs[1]), ns, self._l_min, 1, current_stream()), "sb_ofdm_demodulate")
        return out
