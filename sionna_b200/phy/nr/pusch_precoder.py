"""PUSCHPrecoder: codebook precoding of layer grids onto antenna ports (reference: src/sionna/phy/nr/pusch_precoder.py)."""
import numpy as np
import torch
from ..block import Block
from ..._lib import lib, check, ptr, current_stream


class PUSCHPrecoder(Block):
    """PUSCHPrecoder(precoding_matrices, precision=None)

    ``precoding_matrices``: list (one per transmitter) of ``[num_antenna_ports, num_layers]`` matrices.
    ``call(x)``: ``[batch, num_tx, num_layers, num_symbols, fft_size]`` -> ``[batch, num_tx, num_antenna_ports,
    num_symbols, fft_size]`` (kernel ``sb_pusch_precode``)."""

    def __init__(self, precoding_matrices, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        w = np.stack([np.asarray(m) for m in precoding_matrices])
        assert w.ndim == 3, "Each precoding matrix must have the shape [num_antenna_ports, num_layers]"
        self._num_tx, self._num_ports, self._num_layers = w.shape
        self._w_np = w.astype(np.complex64)
        self._w = None

    @property
    def w(self):
        """[num_tx, num_antenna_ports, num_layers] complex64 device tensor"""
        if self._w is None or self._w.device != self.device:
            self._w = torch.from_numpy(self._w_np).to(self.device)
        return self._w

    @property
    def w_t(self):
        """[num_tx, num_layers, num_antenna_ports]: the transposed matrices (effective-channel computation)."""
        if getattr(self, "_w_t", None) is None or self._w_t.device != self.device:
            self._w_t = torch.from_numpy(np.ascontiguousarray(self._w_np.transpose(0, 2, 1))).to(self.device)
        return self._w_t

    def effective_channel(self, h):
        """h [batch, num_rx, num_rx_ant, num_tx, num_antenna_ports, num_symbols, fft_size] -> the channel seen by the
        layers, h_eff[..., t, l, s, f] = sum_p h[..., t, p, s, f] W[t, p, l] (pusch_receiver.py:243-252): the same
        per-resource-element contraction as precoding with W^T, so the same kernel."""
        b, r, ra, num_tx, ports, s, f = h.shape
        assert num_tx == self._num_tx and ports == self._num_ports
        x = h.to(torch.complex64).contiguous()
        y = torch.empty((b, r, ra, num_tx, self._num_layers, s, f), dtype=torch.complex64, device=x.device)
        check(lib().sb_pusch_precode(ptr(x), ptr(self.w_t), ptr(y), b * r * ra, num_tx, ports, self._num_layers, s * f,
                                     current_stream()), "sb_pusch_precode")
        return y

    def call(self, inputs):
        if self.precision != "single":
            raise NotImplementedError("PUSCHPrecoder runs complex64 kernels only.")
        b, num_tx, num_layers, s, f = inputs.shape
        assert num_tx == self._num_tx, \
            f"The input shape is for {num_tx} transmitters, but you have configured precoding matrices for {self._num_tx}."
        assert num_layers == self._num_layers, \
            f"You have configured precoding matrices for {self._num_layers} layers, but the input provides {num_layers} layers."
        x = inputs.to(torch.complex64).contiguous()
        y = torch.empty((b, num_tx, self._num_ports, s, f), dtype=torch.complex64, device=x.device)
        check(lib().sb_pusch_precode(ptr(x), ptr(self.w), ptr(y), b, num_tx, num_layers, self._num_ports, s * f,
                                     current_stream()), "sb_pusch_precode")
        return y
