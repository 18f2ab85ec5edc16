"""Pilot patterns (mirror of /root/reference/src/sionna/phy/ofdm/pilot_pattern.py). Host-side NumPy containers."""
import numpy as np
import torch

from ..block import Object
from ..mapping import qam


class PilotPattern(Object):
    """PilotPattern(mask, pilots, normalize=False, precision=None)

    ``mask`` [num_tx, num_streams_per_tx, num_ofdm_symbols, num_effective_subcarriers] marks pilot REs, ``pilots``
    [num_tx, num_streams_per_tx, num_pilots] are mapped onto them in row-major (symbol, subcarrier) order
    (pilot_pattern.py:15-267). With ``normalize`` every pilot sequence is scaled to unit average energy over ALL its
    positions, zeros included (:117-124)."""

    def __init__(self, mask, pilots, normalize=False, precision=None):
        super().__init__(precision=precision)
        self._mask = np.asarray(mask.cpu() if isinstance(mask, torch.Tensor) else mask).astype(np.int32)
        self.pilots = pilots
        self.normalize = normalize
        self._check_settings()

    @property
    def num_tx(self):
        return self._mask.shape[0]

    @property
    def num_streams_per_tx(self):
        return self._mask.shape[1]

    @property
    def num_ofdm_symbols(self):
        return self._mask.shape[2]

    @property
    def num_effective_subcarriers(self):
        return self._mask.shape[3]

    @property
    def num_pilot_symbols(self):
        return int(self._pilots.shape[-1])

    @property
    def num_data_symbols(self):
        return int(self._mask.shape[-1] * self._mask.shape[-2] - self.num_pilot_symbols)

    @property
    def normalize(self):
        return self._normalize

    @normalize.setter
    def normalize(self, value):
        self._normalize = bool(value)

    @property
    def mask(self):
        return self._mask

    @property
    def pilots(self):
        """[num_tx, num_streams_per_tx, num_pilots] complex64 NumPy array (normalised if ``normalize``)."""
        if self._normalize and self._pilots.shape[-1] > 0:
            scale = np.abs(self._pilots) ** 2
            scale = 1 / np.sqrt(np.mean(scale, axis=-1, keepdims=True))
            return (scale.astype(np.complex64) * self._pilots).astype(np.complex64)
        return self._pilots

    @pilots.setter
    def pilots(self, v):
        v = np.asarray(v.cpu() if isinstance(v, torch.Tensor) else v)
        self._pilots = v.astype(np.complex64 if self.precision == "single" else np.complex128)

    def _check_settings(self):
        assert self._mask.ndim == 4, "`mask` must have four dimensions."
        assert self._pilots.ndim == 3, "`pilots` must have three dimensions."
        assert np.array_equal(self._mask.shape[:2], self._pilots.shape[:2]), \
            "The first two dimensions of `mask` and `pilots` must be equal."
        num_pilots = np.sum(self._mask, axis=(-2, -1))
        assert np.min(num_pilots) == np.max(num_pilots), \
            "The number of nonzero elements in the masks for all transmitters and streams must be identical."
        assert self.num_pilot_symbols == np.max(num_pilots), \
            "The shape of the last dimension of `pilots` must equal the number of non-zero entries of `mask`."
        return True


class EmptyPilotPattern(PilotPattern):
    """Pilot pattern without pilots (pilot_pattern.py:229-267)."""

    def __init__(self, num_tx, num_streams_per_tx, num_ofdm_symbols, num_effective_subcarriers, precision=None):
        assert num_tx > 0, "`num_tx` must be positive`."
        assert num_streams_per_tx > 0, "`num_streams_per_tx` must be positive`."
        assert num_ofdm_symbols > 0, "`num_ofdm_symbols` must be positive`."
        assert num_effective_subcarriers > 0, "`num_effective_subcarriers` must be positive`."
        shape = [num_tx, num_streams_per_tx, num_ofdm_symbols, num_effective_subcarriers]
        super().__init__(np.zeros(shape, bool), np.zeros(shape[:2] + [0], np.complex64), normalize=False,
                         precision=precision)


class KroneckerPilotPattern(PilotPattern):
    """Orthogonal comb pilots (pilot_pattern.py:269-375): on every pilot OFDM symbol stream s of transmitter i owns
    every ``num_tx*num_streams_per_tx``-th effective subcarrier starting at ``i*num_streams_per_tx + s`` and is zero on
    the others. Pilot values are random QPSK symbols from NumPy ``default_rng(seed)`` (the reference draws them from a
    TensorFlow generator, whose stream cannot be reproduced without TensorFlow)."""

    def __init__(self, resource_grid, pilot_ofdm_symbol_indices, normalize=True, seed=0, precision=None):
        num_tx = resource_grid.num_tx
        num_streams_per_tx = resource_grid.num_streams_per_tx
        num_ofdm_symbols = resource_grid.num_ofdm_symbols
        num_effective_subcarriers = resource_grid.num_effective_subcarriers
        num_pilot_symbols = len(pilot_ofdm_symbol_indices)
        num_seq = num_tx * num_streams_per_tx
        num_pilots = num_pilot_symbols * num_effective_subcarriers / num_seq
        assert (num_pilots / num_pilot_symbols) % 1 == 0, \
            "`num_effective_subcarriers` must be an integer multiple of `num_tx`*`num_streams_per_tx`."
        num_pilots_per_symbol = int(num_pilots / num_pilot_symbols)
        shape = [num_tx, num_streams_per_tx, num_ofdm_symbols, num_effective_subcarriers]
        mask = np.zeros(shape, bool)
        shape[2] = num_pilot_symbols
        pilots = np.zeros(shape, np.complex64)
        mask[..., pilot_ofdm_symbol_indices, :] = True
        rng = np.random.default_rng(seed)
        pts = qam(2, precision="single")
        for i in range(num_tx):
            for j in range(num_streams_per_tx):
                p = pts[rng.integers(0, 4, size=[num_pilot_symbols, num_pilots_per_symbol])]
                pilots[i, j, :, i * num_streams_per_tx + j::num_seq] = p
        pilots = np.reshape(pilots, [num_tx, num_streams_per_tx, -1])
        super().__init__(mask, pilots, normalize=normalize, precision=precision)
