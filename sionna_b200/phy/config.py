"""Global configuration (mirror of /root/reference/src/sionna/phy/config.py:34-201).

``config.precision`` selects the default real/complex dtypes, ``config.seed`` seeds the Python,
NumPy and device generators. The device generator is a counter-based Philox4x32-10 stream
(``sionna_b200/csrc/rng.cuh``): ``config.next_philox()`` hands every random-drawing kernel
launch a fresh (seed, offset) pair, so results are reproducible for a given seed and launch
order, and per-rank streams differ through ``config.rank_offset`` (set by the multi-GPU driver).
"""
import random
import numpy as np
import torch

dtypes = {
    "single": {"torch": {"cdtype": torch.complex64, "rdtype": torch.float32},
               "np": {"cdtype": np.complex64, "rdtype": np.float32}},
    "double": {"torch": {"cdtype": torch.complex128, "rdtype": torch.float64},
               "np": {"cdtype": np.complex128, "rdtype": np.float64}},
}


class Config:
    """Singleton holding precision, seed and the target device."""
    _instance = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = object.__new__(cls)
        return cls._instance

    def __init__(self):
        self._seed = None
        self._py_rng = None
        self._np_rng = None
        self._philox_seed = None
        self._philox_offset = 0
        self._device = None
        self.rank_offset = 0
        self.precision = "single"

    # ---- device ---------------------------------------------------------------------------
    @property
    def device(self):
        """`torch.device` all blocks allocate on. Raises if no CUDA device exists."""
        if self._device is None:
            if not torch.cuda.is_available():
                raise RuntimeError(
                    "sionna_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
            self._device = torch.device("cuda", torch.cuda.current_device())
        return self._device

    @device.setter
    def device(self, d):
        self._device = torch.device(d)

    # ---- generators -----------------------------------------------------------------------
    @property
    def py_rng(self):
        if self._py_rng is None:
            self._py_rng = random.Random()
        return self._py_rng

    @property
    def np_rng(self):
        if self._np_rng is None:
            self._np_rng = np.random.default_rng()
        return self._np_rng

    def next_philox(self, n_draws=1):
        """Return ``(seed, offset)`` for one kernel launch and advance the stream.

        ``n_draws`` is the number of 128-bit Philox blocks one thread may consume at most;
        the offset advances by that much so launches never overlap."""
        if self._philox_seed is None:
            self._philox_seed = random.SystemRandom().getrandbits(62)
        seed = (self._philox_seed + 0x9E3779B97F4A7C15 * self.rank_offset) & 0x7FFFFFFFFFFFFFFF
        off = self._philox_offset
        self._philox_offset += int(n_draws)
        return seed, off

    @property
    def seed(self):
        return self._seed

    @seed.setter
    def seed(self, seed):
        if seed is not None:
            seed = int(seed)
        self._seed = seed
        self.py_rng.seed(seed)
        self._np_rng = np.random.default_rng(seed)
        self._philox_seed = None if seed is None else (seed * 0x2545F4914F6CDD1D + 0x1234567) & 0x3FFFFFFFFFFFFFFF
        self._philox_offset = 0

    # ---- precision ------------------------------------------------------------------------
    @property
    def precision(self):
        return self._precision

    @precision.setter
    def precision(self, v):
        if v not in ["single", "double"]:
            raise ValueError("Precision must be ``single`` or ``double``.")
        self._precision = v

    @property
    def np_rdtype(self):
        return dtypes[self.precision]["np"]["rdtype"]

    @property
    def np_cdtype(self):
        return dtypes[self.precision]["np"]["cdtype"]

    @property
    def rdtype(self):
        return dtypes[self.precision]["torch"]["rdtype"]

    @property
    def cdtype(self):
        return dtypes[self.precision]["torch"]["cdtype"]

    # names used by reference code that is ported verbatim by users
    tf_rdtype = rdtype
    tf_cdtype = cdtype


config = Config()
