"""OFDM MIMO detection (mirror of /root/reference/src/sionna/phy/ofdm/detection.py:20-317, 740-847): ``LinearDetector``
= fused LMMSE equalisation (``sb_ofdm_lmmse``) + demapping with the per-symbol effective noise variance (``sb_demap``)."""
from ..block import Block
from ..mapping import Constellation, Demapper
from .equalization import LMMSEEqualizer


class LinearDetector(Block):
    """LinearDetector(equalizer, output, demapping_method, resource_grid, stream_management, constellation_type=None, num_bits_per_symbol=None, constellation=None, hard_out=False, precision=None)

    ``call(y, h_hat, err_var, no)`` -> LLRs ``[batch, num_tx, num_streams, num_data_symbols*num_bits_per_symbol]``
    (``output="bit"``); ``equalizer="lmmse"`` (detection.py:740-847; PUSCH default ``("lmmse","bit","maxlog")``)."""

    def __init__(self, equalizer, output, demapping_method, resource_grid, stream_management, constellation_type=None,
                 num_bits_per_symbol=None, constellation=None, hard_out=False, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        assert equalizer == "lmmse", "Only the 'lmmse' equalizer is provided (zf / mf are out of scope)."
        assert output == "bit", "Only output='bit' is provided."
        self._constellation = Constellation.check_or_create(constellation_type=constellation_type,
                                                            num_bits_per_symbol=num_bits_per_symbol,
                                                            constellation=constellation, precision=precision)
        self._equalizer = LMMSEEqualizer(resource_grid, stream_management, precision=precision)
        self._demapper = Demapper(demapping_method, constellation=self._constellation, hard_out=hard_out,
                                  precision=precision)

    def call(self, y, h_hat, err_var, no):
        x_hat, no_eff = self._equalizer(y, h_hat, err_var, no)
        return self._demapper(x_hat, no_eff)
