"""PUSCHReceiver (reference: src/sionna/phy/nr/pusch_receiver.py:17-270)."""
import numpy as np
import torch
from ..block import Block
from ..ofdm import OFDMDemodulator, LinearDetector
from ..mimo import StreamManagement
from .pusch_transmitter import PUSCHTransmitter
from .pusch_channel_estimation import PUSCHLSChannelEstimator
from .layer_mapping import LayerDemapper
from .tb_decoder import TBDecoder


class PUSCHReceiver(Block):
    """PUSCHReceiver(pusch_transmitter, channel_estimator=None, mimo_detector=None, tb_decoder=None, return_tb_crc_status=False, stream_management=None, input_domain="freq", l_min=None, precision=None)

    ``call(y, no, h=None)``: optional `OFDMDemodulator` (``input_domain="time"``) -> channel estimation
    (`PUSCHLSChannelEstimator` with linear interpolation by default; ``"perfect"`` uses the provided ``h`` -- time-domain
    taps are converted with `time_to_ofdm_channel` -- multiplied by the precoding matrices when the transmitter precodes) -> MIMO detection (default LMMSE `LinearDetector`, max-log
    bit LLRs) -> `LayerDemapper` -> `TBDecoder`. Returns ``b_hat [batch, num_tx, tb_size]`` (and ``tb_crc_status
    [batch, num_tx]`` if ``return_tb_crc_status``)."""

    def __init__(self, pusch_transmitter, channel_estimator=None, mimo_detector=None, tb_decoder=None,
                 return_tb_crc_status=False, stream_management=None, input_domain="freq", l_min=None, precision=None,
                 **kwargs):
        super().__init__(precision=precision, **kwargs)
        assert input_domain in ["time", "freq"], "input_domain must be 'time' or 'freq'"
        assert isinstance(pusch_transmitter, PUSCHTransmitter), "pusch_transmitter must be a PUSCHTransmitter"
        tx = pusch_transmitter
        self._input_domain = input_domain
        self._return_tb_crc_status = return_tb_crc_status
        self._resource_grid = tx.resource_grid
        if input_domain == "time":
            assert l_min is not None, "l_min must be provided for input_domain==time"
            self._l_min = l_min
            self._ofdm_demodulator = OFDMDemodulator(fft_size=tx._num_subcarriers, l_min=l_min,
                                                     cyclic_prefix_length=tx._cyclic_prefix_length,
                                                     precision=self.precision)
        self._perfect_csi = False
        self._w = None
        if channel_estimator is None:
            self._channel_estimator = PUSCHLSChannelEstimator(self._resource_grid, tx._dmrs_length,
                                                              tx._dmrs_additional_position,
                                                              tx._num_cdm_groups_without_data,
                                                              interpolation_type="lin", precision=self.precision)
        elif isinstance(channel_estimator, str) and channel_estimator == "perfect":
            self._perfect_csi = True
            if tx._precoding == "codebook":
                self._w = tx._precoder
        else:
            self._channel_estimator = channel_estimator
        if stream_management is None:
            stream_management = StreamManagement(np.ones([1, tx._num_tx], bool), tx._num_layers)
        self._stream_management = stream_management
        self._default_detector = mimo_detector is None
        if mimo_detector is None:
            mimo_detector = LinearDetector("lmmse", "bit", "maxlog", tx.resource_grid, stream_management, "qam",
                                           tx._num_bits_per_symbol, precision=self.precision)
        self._mimo_detector = mimo_detector
        # default estimator + default detector: one fused launch from the resource grid to the LLRs (ofdm/frontend.py)
        self._fused = None
        if channel_estimator is None and self._default_detector and not self._perfect_csi:
            from ..ofdm.frontend import FusedLSLinearDetector, fusable
            if fusable(tx.resource_grid, stream_management, self._channel_estimator, mimo_detector._constellation):
                self._fused = FusedLSLinearDetector(self._channel_estimator, tx.resource_grid, stream_management, "maxlog",
                                                    constellation=mimo_detector._constellation, precision=self.precision)
        self._layer_demapper = LayerDemapper(tx._layer_mapper, num_bits_per_symbol=tx._num_bits_per_symbol,
                                             precision=self.precision)
        self._tb_decoder = TBDecoder(tx._tb_encoder, precision=self.precision) if tb_decoder is None else tb_decoder

    resource_grid = property(lambda self: self._resource_grid)
    fuse_front_end = True      # set False to run estimator and detector as separate blocks (same numbers up to rounding)

    def call(self, y, no, h=None):
        if self._input_domain == "time":
            y = self._ofdm_demodulator(y)
        if self._perfect_csi:
            assert h is not None, "h must be provided for channel_estimator='perfect'"
            if self._input_domain == "time":
                from ..channel import time_to_ofdm_channel
                h = time_to_ofdm_channel(h, self.resource_grid, self._l_min)
            if self._w is not None:
                # effective channel per layer: h_eff[b, r, ra, t, l, s, f] = sum_p h[b, r, ra, t, p, s, f] W[t, p, l]
                h = self._w.effective_channel(h)
            h_hat = h.contiguous()
            err_var = torch.zeros((), dtype=torch.float32, device=h_hat.device)
        elif self._fused is not None and self.fuse_front_end:
            h_hat = None
        else:
            h_hat, err_var = self._channel_estimator(y, no)
        llr = self._fused(y, no) if h_hat is None else self._mimo_detector(y, h_hat, err_var, no)
        llr = self._layer_demapper(llr)
        b_hat, tb_crc_status = self._tb_decoder(llr)
        return (b_hat, tb_crc_status) if self._return_tb_crc_status else b_hat
