"""Time-domain channel: discrete-time taps from a CIR, time-variant filtering, and the `OFDMChannel` / `TimeChannel`
convenience blocks (mirror of /root/reference/src/sionna/phy/channel/{utils.py:123-350, apply_time_channel.py,
generate_time_channel.py, time_channel.py, generate_ofdm_channel.py, ofdm_channel.py})."""
import numpy as np
import torch

from ..block import Block
from ..config import config
from ..mapping import _broadcast_inner
from ..._lib import lib, check, ptr, current_stream
from .apply_ofdm_channel import ApplyOFDMChannel
from .tdl import cir_to_ofdm_channel, subcarrier_frequencies


def time_lag_discrete_time_channel(bandwidth, maximum_delay_spread=3e-6):
    """(l_min, l_max) = (-6, ceil(maximum_delay_spread * bandwidth) + 6) (utils.py:123-178)."""
    return -6, int(np.ceil(maximum_delay_spread * bandwidth)) + 6


def cir_to_time_channel(bandwidth, a, tau, l_min, l_max, normalize=False):
    """Discrete-time taps hm[b, rx, rx_ant, tx, tx_ant, t, l] = sum_p a[..., p, t] sinc(l - tau_p W), l = l_min..l_max
    (utils.py:256-350); ``normalize``: unit average total tap energy per link (:341-348). Same three kernels as
    `cir_to_ofdm_channel` with the real sinc table in place of the phase table."""
    from .tdl import _cir_convert
    key = (int(l_min), int(l_max))
    lags = _LAGS.get(key)
    if lags is None:
        lags = _LAGS[key] = torch.arange(int(l_min), int(l_max) + 1, dtype=torch.float32)
    return _cir_convert(a, tau, lags, 1, float(bandwidth), normalize, 1.0)


_LAGS = {}


def time_to_ofdm_channel(h_t, rg, l_min):
    """Frequency response seen by each OFDM symbol from time-domain taps (utils.py:352-458): the taps at the first
    sample after every cyclic prefix, zero-padded to ``fft_size``, rolled by ``l_min``, un-normalised DFT + fftshift
    (``sb_ofdm_demodulate`` with the shift folded in). ``h_t [..., num_time_samples + l_tot - 1, l_tot]`` ->
    ``[..., num_ofdm_symbols, fft_size]``."""
    n = rg.fft_size
    ofdm_length = n + rg.cyclic_prefix_length
    h = h_t[..., rg.cyclic_prefix_length:rg.num_time_samples:ofdm_length, :].to(torch.complex64)
    h = torch.cat([h, torch.zeros(list(h.shape[:-1]) + [n - h.shape[-1]], dtype=h.dtype, device=h.device)], -1)
    h = torch.roll(h, int(l_min), dims=-1).contiguous()
    rows = h.numel() // n
    zero = torch.zeros(1, dtype=torch.int32, device=h.device)
    out = torch.empty_like(h)
    check(lib().sb_ofdm_demodulate(ptr(h), ptr(out), rows, 1, n, ptr(zero), ptr(zero), n, 0, 1, current_stream()),
          "sb_ofdm_demodulate")
    return out * float(np.sqrt(n))                            # the kernel applies 1 / sqrt(N); tf.signal.fft does not


class ApplyTimeChannel(Block):
    """ApplyTimeChannel(num_time_samples, l_tot, precision=None): ``call(x, h_time, no=None)`` filters
    ``x [batch, num_tx, num_tx_ant, num_time_samples]`` with the time-variant taps ``h_time [batch, num_rx, num_rx_ant,
    num_tx, num_tx_ant, num_time_samples + l_tot - 1, l_tot]`` and adds noise -> ``[batch, num_rx, num_rx_ant,
    num_time_samples + l_tot - 1]`` (kernel ``sb_apply_time_channel``)."""

    def __init__(self, num_time_samples, l_tot, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        self._n, self._l = int(num_time_samples), int(l_tot)

    def call(self, x, h_time, no=None):
        if self.precision != "single":
            raise NotImplementedError("sb_apply_time_channel is a complex64 kernel.")
        dev = self.device
        x = x.to(device=dev, dtype=torch.complex64).contiguous()
        h = h_time.to(device=dev, dtype=torch.complex64).contiguous()
        b, rx, ra, tx, ta, nt, l = h.shape
        assert x.shape[-1] == self._n and l == self._l and nt == self._n + self._l - 1, \
            "h_time must hold num_time_samples + l_tot - 1 time steps of l_tot taps"
        y = torch.empty((b, rx, ra, nt), dtype=torch.complex64, device=dev)
        no_t, inner, add, seed, off = None, 1, 0, 0, 0
        if no is not None:
            no_t, inner = _broadcast_inner(no, y.shape, dev, torch.float32)
            add = 1
            seed, off = config.next_philox()
        check(lib().sb_apply_time_channel(ptr(x), ptr(h), ptr(no_t), inner, ptr(y), b, rx * ra, tx * ta, self._n, self._l,
                                          add, seed, off, current_stream()), "sb_apply_time_channel")
        return y


class GenerateOFDMChannel(Block):
    """GenerateOFDMChannel(channel_model, resource_grid, normalize_channel=False): ``__call__(batch_size)`` samples a CIR
    per OFDM symbol and returns ``h_freq [batch, num_rx, num_rx_ant, num_tx, num_tx_ant, num_ofdm_symbols, fft_size]``."""

    def __init__(self, channel_model, resource_grid, normalize_channel=False, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        self._cir_sampler = channel_model
        self._num_ofdm_symbols = resource_grid.num_ofdm_symbols
        self._normalize_channel = normalize_channel
        self._sampling_frequency = 1. / resource_grid.ofdm_symbol_duration
        self._frequencies = subcarrier_frequencies(resource_grid.fft_size, resource_grid.subcarrier_spacing)

    def call(self, batch_size=None):
        h, tau = self._cir_sampler(batch_size, self._num_ofdm_symbols, self._sampling_frequency)
        return cir_to_ofdm_channel(self._frequencies, h, tau, self._normalize_channel)


class OFDMChannel(Block):
    """OFDMChannel(channel_model, resource_grid, normalize_channel=False, return_channel=False): ``call(x, no=None)``
    -> ``y`` (and ``h_freq``): `GenerateOFDMChannel` + `ApplyOFDMChannel`."""

    def __init__(self, channel_model, resource_grid, normalize_channel=False, return_channel=False, precision=None,
                 **kwargs):
        super().__init__(precision=precision, **kwargs)
        self._return_channel = return_channel
        self._generate_channel = GenerateOFDMChannel(channel_model, resource_grid, normalize_channel, self.precision)
        self._apply_channel = ApplyOFDMChannel(self.precision)

    def call(self, x, no=None):
        h_freq = self._generate_channel(x.shape[0])
        y = self._apply_channel(x, h_freq, no)
        return (y, h_freq) if self._return_channel else y


class GenerateTimeChannel(Block):
    """GenerateTimeChannel(channel_model, bandwidth, num_time_samples, l_min, l_max, normalize_channel=False):
    ``__call__(batch_size)`` -> ``h_time [batch, num_rx, num_rx_ant, num_tx, num_tx_ant, num_time_samples + l_max -
    l_min, l_max - l_min + 1]``."""

    def __init__(self, channel_model, bandwidth, num_time_samples, l_min, l_max, normalize_channel=False, precision=None,
                 **kwargs):
        super().__init__(precision=precision, **kwargs)
        self._cir_sampler = channel_model
        self._l_min, self._l_max = int(l_min), int(l_max)
        self._l_tot = self._l_max - self._l_min + 1
        self._bandwidth, self._num_time_steps = bandwidth, int(num_time_samples)
        self._normalize_channel = normalize_channel

    def call(self, batch_size=None):
        h, tau = self._cir_sampler(batch_size, self._num_time_steps + self._l_tot - 1, self._bandwidth)
        return cir_to_time_channel(self._bandwidth, h, tau, self._l_min, self._l_max, self._normalize_channel)


class TimeChannel(Block):
    """TimeChannel(channel_model, bandwidth, num_time_samples, maximum_delay_spread=3e-6, l_min=None, l_max=None,
    normalize_channel=False, return_channel=False): ``call(x, no=None)`` -> ``y [batch, num_rx, num_rx_ant,
    num_time_samples + l_max - l_min]`` (and ``h_time``)."""

    def __init__(self, channel_model, bandwidth, num_time_samples, maximum_delay_spread=3e-6, l_min=None, l_max=None,
                 normalize_channel=False, return_channel=False, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        l_min_d, l_max_d = time_lag_discrete_time_channel(bandwidth, maximum_delay_spread)
        l_min = l_min_d if l_min is None else l_min
        l_max = l_max_d if l_max is None else l_max
        self._return_channel = return_channel
        self._generate_channel = GenerateTimeChannel(channel_model, bandwidth, num_time_samples, l_min, l_max,
                                                     normalize_channel, precision=self.precision)
        self._apply_channel = ApplyTimeChannel(num_time_samples, l_max - l_min + 1, precision=self.precision)

    def call(self, x, no=None):
        h_time = self._generate_channel(x.shape[0])
        y = self._apply_channel(x, h_time, no)
        return (y, h_time) if self._return_channel else y
