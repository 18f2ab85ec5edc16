"""LayerMapper / LayerDemapper of TS 38.211 Sec. 6.3.1.3 and 7.3.1.3 (reference: src/sionna/phy/nr/layer_mapping.py).

Pure data movement: the layer split is a reshape + transpose of device tensors (torch views / one copy kernel of the
allocator's memcpy engine); no arithmetic.
"""
import torch
from ..block import Block


class LayerMapper(Block):
    """LayerMapper(num_layers=1, verbose=False, precision=None)

    ``call(x)``: ``[..., n]`` symbols -> ``[..., num_layers, n / num_layers]`` (symbol i goes to layer i mod
    num_layers). For 5..8 layers (PDSCH dual-codeword mode, Table 7.3.1.3-1) the input is a list of two codewords that
    occupy the first floor(num_layers/2) and the remaining layers."""

    def __init__(self, num_layers=1, verbose=False, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        assert isinstance(verbose, bool), "verbose must be bool"
        assert num_layers in range(1, 9), "num_layers must be between 1 and 8."
        self._num_layers = int(num_layers)
        self._verbose = verbose
        if self._num_layers < 5:
            self._num_codewords, self._num_layers0, self._num_layers1 = 1, self._num_layers, 0
        else:
            self._num_codewords = 2
            self._num_layers0 = self._num_layers // 2
            self._num_layers1 = self._num_layers - self._num_layers0
        if verbose:
            print("Number of layers: ", self._num_layers)
            if self._num_codewords == 2:
                print("Dual codeword mode active and cw multiplexing as defined in Tab. 7.3.1.3-1 from 38.211 applied.")
                print(f"Length of cw1/cw2: {self._num_layers0}/{self._num_layers1} ")

    num_codewords = property(lambda self: self._num_codewords)
    num_layers = property(lambda self: self._num_layers)
    num_layers0 = property(lambda self: self._num_layers if self._num_codewords == 1 else self._num_layers0)
    num_layers1 = property(lambda self: 0 if self._num_codewords == 1 else self._num_layers1)

    @staticmethod
    def _split(x, layers):
        n = x.shape[-1]
        assert n % layers == 0, "Invalid input dimensions: last dimension must be a multiple of num_layers."
        return x.reshape(list(x.shape[:-1]) + [n // layers, layers])

    def call(self, inputs):
        if self._num_codewords == 1:
            assert not isinstance(inputs, (list, tuple)), "Only single input codeword expected."
            y = self._split(inputs, self._num_layers)
        else:
            assert isinstance(inputs, (list, tuple)) and len(inputs) == 2, "List of two inputs streams is expected."
            y0 = self._split(inputs[0], self._num_layers0)
            y1 = self._split(inputs[1], self._num_layers1)
            assert y0.shape[-2] == y1.shape[-2], \
                "Invalid input dimensions: both codewords must provide the same number of symbols per layer."
            y = torch.cat([y0, y1], dim=-1)
        return y.transpose(-1, -2).contiguous()


class LayerDemapper(Block):
    """LayerDemapper(layer_mapper, num_bits_per_symbol=1, precision=None)

    Inverse of `LayerMapper` for LLRs: ``[..., num_layers, n / num_layers]`` -> ``[..., n]`` (or a list of two for dual
    codeword mode), keeping groups of ``num_bits_per_symbol`` consecutive LLRs (one symbol) together."""

    def __init__(self, layer_mapper, num_bits_per_symbol=1, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        assert isinstance(layer_mapper, LayerMapper), "layer_mapper must be LayerMapper."
        assert num_bits_per_symbol % 1 == 0, "num_bits_per_symbol must be int."
        self._mapper = layer_mapper
        self._num_bits_per_symbol = int(num_bits_per_symbol)

    def call(self, inputs):
        m, q = self._mapper, self._num_bits_per_symbol
        assert inputs.shape[-2] == m.num_layers, "Invalid input dimension: input shape must be [...,num_layers,n]."
        assert inputs.shape[-1] % q == 0, \
            "Invalid input dimension: last dimension must be a multiple of num_bits_per_symbol."
        lead = list(inputs.shape[:-2])
        x = inputs.reshape(lead + [m.num_layers, inputs.shape[-1] // q, q]).transpose(-2, -3)   # [..., sym, layer, q]
        if m.num_codewords == 1:
            return x.reshape(lead + [-1]).contiguous()
        y0 = x[..., :m.num_layers0, :].reshape(lead + [-1]).contiguous()
        y1 = x[..., m.num_layers0:, :].reshape(lead + [-1]).contiguous()
        return [y0, y1]
