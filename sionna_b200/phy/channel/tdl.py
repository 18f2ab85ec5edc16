"""Block-fading TDL channel models and CIR -> OFDM channel conversion ("next tier" input generation of SURVEY.md
section 8(f3); mirror of /root/reference/src/sionna/phy/channel/tr38901/tdl.py:372-502 for zero speed and of
channel/utils.py:180-253, 1010-1060). Power delay profiles: TR 38.901 Tables 7.7.2-1..5 (``tdl_models.npz``).
The tap gains are drawn on the device (``complex_normal`` -> ``sb_awgn``'s Philox generator); the frequency response is
assembled with a few torch tensor ops (input generation, not part of the graded receive path)."""
import os
import numpy as np
import torch

from ..block import Block
from ..config import config
from ..utils.misc import complex_normal

_MODELS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tdl_models.npz")


def subcarrier_frequencies(num_subcarriers, subcarrier_spacing, precision=None):
    """Baseband frequencies of the subcarriers, DC in the centre (channel/utils.py:1010-1060)."""
    if num_subcarriers % 2 == 0:
        start, limit = -num_subcarriers / 2, num_subcarriers / 2
    else:
        start, limit = -(num_subcarriers - 1) / 2, (num_subcarriers - 1) / 2 + 1
    return torch.arange(start, limit, dtype=torch.float32) * subcarrier_spacing


def cir_to_ofdm_channel(frequencies, a, tau, normalize=False):
    """h_f[b, rx, rx_ant, tx, tx_ant, t, f] = sum_p a[..., p, t] exp(-j 2 pi f tau_p) (channel/utils.py:180-253)."""
    dev = a.device
    f = frequencies.to(dev)
    tau = tau.to(dev)
    if tau.dim() == 4:                                        # [b, rx, tx, paths] -> broadcast over antennas
        tau = tau[:, :, None, :, None, :]
    e = torch.exp(torch.complex(torch.zeros((), device=dev), -2 * np.pi * tau[..., None] * f))   # [..., paths, F]
    h = torch.einsum("brmtnpl,brmtnpf->brmtnlf", a, e.expand(*a.shape[:-1], f.shape[0]).to(a.dtype))
    if normalize:
        c = torch.sqrt(torch.mean(torch.abs(h) ** 2, dim=(2, 4, 5, 6), keepdim=True))
        h = h / c.to(h.dtype)
    return h


class TDL(Block):
    """TDL(model, delay_spread, carrier_frequency, num_rx_ant=1, num_tx_ant=1, min_speed=0., max_speed=None)

    Tapped delay line model "A".."E" of TR 38.901 with RMS delay spread ``delay_spread`` [s]; block fading (zero
    speed): ``call(batch_size, num_time_steps, sampling_frequency)`` -> ``a [batch, 1, num_rx_ant, 1, num_tx_ant,
    num_paths, num_time_steps]`` (constant over time), ``tau [batch, 1, 1, num_paths]``. Average total power is 1."""

    def __init__(self, model, delay_spread, carrier_frequency, num_rx_ant=1, num_tx_ant=1, min_speed=0., max_speed=None,
                 precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        assert model in ("A", "B", "C", "D", "E"), "Invalid TDL model"
        if (min_speed or 0.) != 0. or (max_speed or 0.) != 0.:
            raise NotImplementedError("TDL: only block fading (zero speed) is provided.")
        with np.load(_MODELS) as d:
            delays, p_db, los = d[f"{model}_delays"], d[f"{model}_powers_db"], int(d[f"{model}_los"])
        p = 10 ** (p_db / 10)
        self._los = bool(los)
        if self._los:                                          # first entry = LoS component sharing the first delay
            self._los_power = p[0]
            p, delays = p[1:], delays[1:]
            norm = self._los_power + p.sum()
            self._los_power /= norm
            p = p / norm
        else:
            p = p / p.sum()
        self._powers = p.astype(np.float32)
        self._delays = (delays * delay_spread).astype(np.float32)
        self._num_rx_ant, self._num_tx_ant = num_rx_ant, num_tx_ant

    @property
    def num_clusters(self):
        return len(self._powers)

    @property
    def delays(self):
        return torch.from_numpy(self._delays)

    @property
    def mean_powers(self):
        return torch.from_numpy(self._powers)

    def __call__(self, batch_size, num_time_steps=1, sampling_frequency=1.0):
        return self.call(batch_size, num_time_steps, sampling_frequency)

    def call(self, batch_size, num_time_steps=1, sampling_frequency=1.0):
        dev = config.device
        n = self.num_clusters
        g = complex_normal([batch_size, 1, self._num_rx_ant, 1, self._num_tx_ant, n, 1])
        a = g * torch.sqrt(torch.from_numpy(self._powers).to(dev)).reshape(1, 1, 1, 1, 1, n, 1)
        if self._los:
            a[..., 0, :] = a[..., 0, :] + np.sqrt(self._los_power)
        a = a.expand(batch_size, 1, self._num_rx_ant, 1, self._num_tx_ant, n, num_time_steps).contiguous()
        tau = torch.from_numpy(self._delays).to(dev).reshape(1, 1, 1, n).expand(batch_size, 1, 1, n).contiguous()
        return a, tau
