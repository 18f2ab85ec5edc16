"""PUSCHTransmitter (reference: src/sionna/phy/nr/pusch_transmitter.py:17-243)."""
import torch
from ..block import Block
from ..mapping import Mapper, BinarySource
from ..ofdm import ResourceGrid, ResourceGridMapper, OFDMModulator
from .pusch_config import PUSCHConfig, check_pusch_configs
from .pusch_pilot_pattern import PUSCHPilotPattern
from .pusch_precoder import PUSCHPrecoder
from .tb_encoder import TBEncoder
from .layer_mapping import LayerMapper
from .config import Config


class PUSCHTransmitter(Block):
    """PUSCHTransmitter(pusch_configs, return_bits=True, output_domain="freq", precision=None, verbose=False)

    Batches of 5G NR PUSCH slots for one or several transmitters: payload bits (random if ``return_bits``) ->
    `TBEncoder` -> `Mapper` -> `LayerMapper` -> `ResourceGridMapper` (data + DMRS) -> optional `PUSCHPrecoder`
    (``precoding="codebook"``) -> optional `OFDMModulator` (``output_domain="time"``).

    ``call(batch_size)`` -> ``(x, b)`` if ``return_bits`` else ``call(b [batch, num_tx, tb_size])`` -> ``x``;
    ``x``: ``[batch, num_tx, num_tx_ant, num_ofdm_symbols, fft_size]`` or ``[batch, num_tx, num_tx_ant, num_time_samples]``.
    """

    def __init__(self, pusch_configs, return_bits=True, output_domain="freq", precision=None, verbose=False, **kwargs):
        super().__init__(precision=precision, **kwargs)
        assert isinstance(return_bits, bool), "return_bits must be bool"
        assert output_domain in ["time", "freq"], "output_domain must be 'time' or 'freq'"
        assert isinstance(verbose, bool), "verbose must be bool"
        self._return_bits, self._output_domain, self._verbose = return_bits, output_domain, verbose
        if isinstance(pusch_configs, PUSCHConfig):
            pusch_configs = [pusch_configs]
        for key, value in check_pusch_configs(pusch_configs).items():
            setattr(self, f"_{key}", value)
        self._pusch_configs = pusch_configs
        if return_bits:
            self._binary_source = BinarySource(precision=self.precision)
        self._tb_encoder = TBEncoder(target_tb_size=self._tb_size, num_coded_bits=self._num_coded_bits,
                                     target_coderate=self._target_coderate,
                                     num_bits_per_symbol=self._num_bits_per_symbol, num_layers=self._num_layers,
                                     n_rnti=self._n_rnti, n_id=self._n_id, channel_type="PUSCH", codeword_index=0,
                                     use_scrambler=True, verbose=verbose, precision=self.precision)
        self._layer_mapper = LayerMapper(num_layers=self._num_layers, precision=self.precision)
        self._mapper = Mapper("qam", self._num_bits_per_symbol, precision=self.precision)
        self._pilot_pattern = PUSCHPilotPattern(self._pusch_configs, precision=self.precision)
        self._resource_grid = ResourceGrid(num_ofdm_symbols=self._num_ofdm_symbols, fft_size=self._num_subcarriers,
                                           subcarrier_spacing=self._subcarrier_spacing, num_tx=self._num_tx,
                                           num_streams_per_tx=self._num_layers,
                                           cyclic_prefix_length=self._cyclic_prefix_length,
                                           pilot_pattern=self._pilot_pattern, precision=self.precision)
        self._resource_grid_mapper = ResourceGridMapper(self._resource_grid, precision=self.precision)
        if self._precoding == "codebook":
            self._precoder = PUSCHPrecoder(self._precoding_matrices, precision=self.precision)
        if self._output_domain == "time":
            self._ofdm_modulator = OFDMModulator(self._cyclic_prefix_length, precision=self.precision)

    resource_grid = property(lambda self: self._resource_grid)
    pilot_pattern = property(lambda self: self._pilot_pattern)

    def show(self):
        self._pusch_configs[0].carrier.show()
        Config.show(self._pusch_configs[0])
        for idx, p in enumerate(self._pusch_configs):
            print(f"---- UE {idx} ----")
            p.dmrs.show()
            p.tb.show()

    def call(self, inputs):
        if self._return_bits:
            b = self._binary_source([int(inputs), self._num_tx, self._tb_size])
        else:
            b = inputs
        c = self._tb_encoder(b)                                # [batch, num_tx, num_coded_bits]
        x_layer = self._layer_mapper(self._mapper(c))          # [batch, num_tx, num_layers, symbols per layer]
        x = self._resource_grid_mapper(x_layer)                # [batch, num_tx, num_layers, num_symbols, fft_size]
        if self._precoding == "codebook":
            x = self._precoder(x)
        if self._output_domain == "time":
            x = self._ofdm_modulator(x)
        return (x, b) if self._return_bits else x
