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
        # same argument checks and error types as the reference (mimo/detection.py:103-115)
        assert not isinstance(equalizer, str) or equalizer in ["lmmse", "zf", "mf"], "Unknown equalizer."
        assert output in ("bit", "symbol"), "Unknown output"
        assert demapping_method in ("app", "maxlog"), "Unknown demapping method"
        if equalizer != "lmmse":
            raise NotImplementedError(f"equalizer={equalizer!r}: only 'lmmse' has a fused OFDM kernel here.")
        if output != "bit":
            raise NotImplementedError("output='symbol' (SymbolDemapper) is not provided; use output='bit'.")
        self._constellation = Constellation.check_or_create(constellation_type=constellation_type,
                                                            num_bits_per_symbol=num_bits_per_symbol,
                                                            constellation=constellation, precision=precision)
        self._equalizer = LMMSEEqualizer(resource_grid, stream_management, precision=precision)
        self._demapper = Demapper(demapping_method, constellation=self._constellation, hard_out=hard_out,
                                  precision=precision)

    def call(self, y, h_hat, err_var, no):
        x_hat, no_eff = self._equalizer(y, h_hat, err_var, no)
        return self._demapper(x_hat, no_eff)
