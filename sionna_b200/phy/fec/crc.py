"""CRC encoding / checking (mirror of /root/reference/src/sionna/phy/fec/crc.py), SURVEY.md section 8(f2).

The reference builds a dense [k, crc_length] generator matrix (crc.py:126-156) and evaluates a float matmul mod 2; here the
rows of the same matrix are packed into 32-bit words on the host and ``sb_crc_encode`` XORs the rows selected by the set
bits (one warp per codeword)."""
import numpy as np
import torch

from ..block import Block
from ..._lib import lib, check, ptr, current_stream

_POLYS = {"CRC24A": (24, [24, 23, 18, 17, 14, 11, 10, 7, 6, 5, 4, 3, 1, 0]),
          "CRC24B": (24, [24, 23, 6, 5, 1, 0]),
          "CRC24C": (24, [24, 23, 21, 20, 17, 15, 13, 12, 8, 4, 2, 1, 0]),
          "CRC16": (16, [16, 12, 5, 0]),
          "CRC11": (11, [11, 10, 9, 5, 0]),
          "CRC6": (6, [6, 5, 0])}


class CRCEncoder(Block):
    """CRCEncoder(crc_degree): appends the 38.212 5.1 CRC parity bits ("CRC24A", "CRC24B", "CRC24C", "CRC16", "CRC11",
    "CRC6") to the last dimension: ``[..., k] -> [..., k + crc_length]`` (crc.py:15-215)."""

    def __init__(self, crc_degree, *, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        if not isinstance(crc_degree, str):
            raise TypeError("crc_degree must be str")
        if crc_degree not in _POLYS:
            raise ValueError("Invalid CRC Polynomial")
        self._crc_degree = crc_degree
        self._crc_length, coeffs = _POLYS[crc_degree]
        pol = np.zeros(self._crc_length + 1, int)
        pol[[self._crc_length - c for c in coeffs]] = 1
        self._crc_pol = pol
        self._k = self._n = None
        self._tab = None

    @property
    def crc_degree(self):
        return self._crc_degree

    @property
    def crc_length(self):
        return self._crc_length

    @property
    def crc_pol(self):
        return self._crc_pol

    @property
    def k(self):
        return self._k

    @property
    def n(self):
        return self._n

    def _gen_rows(self, k):
        """Rows of the generator matrix of crc.py:126-156 as integers: row k-1 is the remainder of x^L, every earlier row
        one more polynomial-division step."""
        L = self._crc_length
        pol = int("".join(str(b) for b in self._crc_pol), 2)            # degree-L polynomial, L+1 bits
        rows = np.zeros(k, np.uint32)
        reg = 1 << (L - 1)                                              # x_crc = [1, 0, ..., 0]
        for i in range(k):
            reg <<= 1
            if reg >> L:
                reg ^= pol
            rows[k - i - 1] = reg
        return rows

    def build(self, input_shape):
        k = input_shape[-1]
        assert k is not None, "Shape of last dimension cannot be None."
        self._k, self._n = k, k + self._crc_length
        self._tab = None

    def call(self, bits, /):
        if self.precision != "single":
            raise NotImplementedError("sb_crc_encode is an fp32 kernel.")
        if bits.shape[-1] != self._k:
            self.build(bits.shape)
        dev = self.device
        if self._tab is None or self._tab.device != dev:
            self._tab = torch.from_numpy(self._gen_rows(self._k).view(np.int32)).to(dev)
        x = bits.to(device=dev, dtype=torch.float32).contiguous()
        rows = x.numel() // self._k
        out = torch.empty(list(x.shape[:-1]) + [self._n], dtype=torch.float32, device=dev)
        check(lib().sb_crc_encode(ptr(x), ptr(self._tab), self._k, self._crc_length, ptr(out), rows, current_stream()),
              "sb_crc_encode")
        return out


class CRCDecoder(Block):
    """CRCDecoder(crc_encoder): ``[..., k + crc_length] -> ([..., k] information bits, [..., 1] bool crc_valid)``: the whole
    word is re-encoded and the check passes iff the new parity is all zero (crc.py:217-327)."""

    def __init__(self, crc_encoder, *, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        assert isinstance(crc_encoder, CRCEncoder), "crc_encoder must be a CRCEncoder instance."
        self._encoder = CRCEncoder(crc_encoder.crc_degree, precision=precision)
        self._tab = None

    @property
    def crc_degree(self):
        return self._encoder.crc_degree

    @property
    def encoder(self):
        return self._encoder

    def build(self, input_shape):
        if input_shape[-1] < self._encoder.crc_length:
            raise ValueError("Input length must be greater than or equal to the CRC length.")

    def call(self, x_crc, /):
        """One kernel (``sb_crc_check``): parity of the re-encoded word, validity flag and the information bits."""
        if self.precision != "single":
            raise NotImplementedError("sb_crc_check is an fp32 kernel.")
        enc = self._encoder
        L, n = enc.crc_length, x_crc.shape[-1]
        dev = self.device
        if self._tab is None or self._tab[0] != n or self._tab[1].device != dev:
            self._tab = (n, torch.from_numpy(enc._gen_rows(n).view(np.int32)).to(dev))
        x = x_crc.to(device=dev, dtype=torch.float32).contiguous()
        rows = x.numel() // n
        info = torch.empty(list(x.shape[:-1]) + [n - L], dtype=torch.float32, device=dev)
        valid = torch.empty(list(x.shape[:-1]) + [1], dtype=torch.bool, device=dev)
        check(lib().sb_crc_check(ptr(x), ptr(self._tab[1]), n, L, ptr(info), ptr(valid), rows, current_stream()),
              "sb_crc_check")
        return info.to(x_crc.dtype) if x_crc.dtype.is_floating_point else info, valid
