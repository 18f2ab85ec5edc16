"""5G NR transport-block encoder (mirror of /root/reference/src/sionna/phy/nr/tb_encoder.py:20-435): TB CRC ->
code-block segmentation -> CB CRC -> LDPC encoding + rate matching -> per-CB bit interleaving and concatenation ->
scrambling. Every stage is one of this package's kernels (sb_crc_encode, sb_ldpc5g_encode, sb_gather_rows, sb_scramble)."""
import numpy as np
import torch

from ..block import Block
from ..fec.crc import CRCEncoder
from ..fec.scrambling import TB5GScrambler
from ..fec.ldpc import LDPC5GEncoder
from ..ofdm.resource_grid import gather_rows
from .utils import calculate_tb_size


class TBEncoder(Block):
    """TBEncoder(target_tb_size, num_coded_bits, target_coderate, num_bits_per_symbol, num_layers=1, n_rnti=1, n_id=1, channel_type="PUSCH", codeword_index=0, use_scrambler=True, verbose=False, precision=None)

    ``[..., num_tx, target_tb_size]`` (or ``[..., target_tb_size]``) information bits -> ``[..., num_coded_bits]`` coded,
    interleaved and scrambled bits (38.212 6.2 / 7.2, tb_encoder.py:381-435)."""

    def __init__(self, target_tb_size, num_coded_bits, target_coderate, num_bits_per_symbol, num_layers=1, n_rnti=1, n_id=1,
                 channel_type="PUSCH", codeword_index=0, use_scrambler=True, verbose=False, precision=None, **kwargs):
        kwargs.pop("output_dtype", None)
        super().__init__(precision=precision, **kwargs)
        assert isinstance(use_scrambler, bool), "use_scrambler must be bool."
        assert isinstance(verbose, bool), "verbose must be bool."
        assert channel_type in ("PDSCH", "PUSCH"), "Unsupported channel_type."
        assert target_tb_size % 1 == 0, "target_tb_size must be int."
        assert num_coded_bits % 1 == 0, "num_coded_bits must be int."
        assert 0. < target_coderate <= 948 / 1024, "target_coderate must be in range(0,0.925)."
        assert num_bits_per_symbol % 1 == 0, "num_bits_per_symbol must be int."
        assert num_layers % 1 == 0, "num_layers must be int."
        if channel_type == "PDSCH":
            assert codeword_index in (0, 1), "codeword_index must be 0 or 1."
        else:
            assert codeword_index == 0, 'codeword_index must be 0 for "PUSCH".'
        self._use_scrambler = use_scrambler
        self._target_tb_size, self._num_coded_bits = int(target_tb_size), int(num_coded_bits)
        self._target_coderate = float(target_coderate)
        self._num_bits_per_symbol, self._num_layers = int(num_bits_per_symbol), int(num_layers)
        if isinstance(n_rnti, (list, tuple)):
            assert isinstance(n_id, (list, tuple)), "n_id must be also a list."
            assert len(n_rnti) == len(n_id), "n_id and n_rnti must be of same length."
            self._n_rnti, self._n_id = [int(v) for v in n_rnti], [int(v) for v in n_id]
            multi = True
        else:
            self._n_rnti, self._n_id, multi = [int(n_rnti)], [int(n_id)], False
        self._num_tx = len(self._n_id)
        tb = calculate_tb_size(target_tb_size=self._target_tb_size, num_coded_bits=self._num_coded_bits,
                               target_coderate=self._target_coderate, modulation_order=self._num_bits_per_symbol,
                               num_layers=self._num_layers, verbose=verbose)
        self._tb_size, self._cb_size, self._num_cbs, self._tb_crc_length, self._cb_crc_length, self._cw_lengths = tb
        assert self._tb_size <= self._tb_crc_length + np.sum(self._cw_lengths), "Invalid TB parameters."
        self._k_padding = self._tb_size - self._target_tb_size
        if self._tb_size != self._target_tb_size:
            print(f"Note: actual tb_size={self._tb_size} is slightly different than requested "
                  f"target_tb_size={self._target_tb_size} due to quantization. Internal zero padding will be applied.")
        self._coderate = self._tb_size / self._num_coded_bits
        self._tb_crc_encoder = CRCEncoder("CRC16" if self._tb_crc_length == 16 else "CRC24A", precision=precision)
        self._cb_crc_encoder = CRCEncoder("CRC24B", precision=precision) if self._cb_crc_length == 24 else None
        self._scrambler = None
        if use_scrambler:
            self._scrambler = TB5GScrambler(n_rnti=self._n_rnti if multi else self._n_rnti[0],
                                            n_id=self._n_id if multi else self._n_id[0], binary=True,
                                            channel_type=channel_type, codeword_index=codeword_index, precision=precision)
        n_max, n_min = int(np.max(self._cw_lengths)), int(np.min(self._cw_lengths))
        self._encoder = LDPC5GEncoder(self._cb_size, n_max, num_bits_per_symbol=1, precision=precision)
        short, _ = self._encoder.generate_out_int(n_min, self._num_bits_per_symbol)
        long_, _ = self._encoder.generate_out_int(n_max, self._num_bits_per_symbol)
        perm, punc, pos = [], [], 0
        for l in self._cw_lengths:                                   # tb_encoder.py:262-283
            if l == n_min:
                perm.append(short + pos)
                punc.append(np.arange(pos + n_min, pos + n_max))
                pos += n_max
            elif l == n_max:
                perm.append(long_ + pos)
                pos += l
            else:
                raise ValueError("Invalid cw_lengths.")
        perm_seq = np.concatenate(perm + punc).astype(np.int64)
        self._output_perm = perm_seq
        self._output_perm_inv = np.argsort(perm_seq)
        self._perm_dev = None

    # ---- properties (tb_encoder.py:292-372) ------------------------------------------------------------------------
    tb_size = property(lambda self: self._tb_size)
    k = property(lambda self: self._target_tb_size)
    k_padding = property(lambda self: self._k_padding)
    n = property(lambda self: self._num_coded_bits)
    num_cbs = property(lambda self: self._num_cbs)
    coderate = property(lambda self: self._coderate)
    ldpc_encoder = property(lambda self: self._encoder)
    scrambler = property(lambda self: self._scrambler)
    tb_crc_encoder = property(lambda self: self._tb_crc_encoder)
    cb_crc_encoder = property(lambda self: self._cb_crc_encoder)
    num_tx = property(lambda self: self._num_tx)
    cw_lengths = property(lambda self: self._cw_lengths)
    output_perm_inv = property(lambda self: self._output_perm_inv)

    def build(self, input_shapes):
        assert input_shapes[-1] == self.k, f"Invalid input shape. Expected TB length is {self.k}."

    def call(self, inputs):
        dev = self.device
        shape = list(inputs.shape)
        u = inputs.to(device=dev, dtype=torch.float32)
        if self._k_padding > 0:
            u = torch.cat([u, torch.zeros(shape[:-1] + [self._k_padding], dtype=u.dtype, device=dev)], -1)
        u_crc = self._tb_crc_encoder(u)
        u_cb = u_crc.reshape(-1, self._num_tx, self._num_cbs, self._cb_size - self._cb_crc_length)
        u_cb_crc = self._cb_crc_encoder(u_cb) if self._cb_crc_encoder is not None else u_cb
        c_cb = self._encoder(u_cb_crc)
        n_max = int(np.max(self._cw_lengths))
        c = c_cb.reshape(-1, self._num_cbs * n_max).contiguous()
        total = int(np.sum(self._cw_lengths))
        if self._perm_dev is None or self._perm_dev.device != dev:
            self._perm_dev = torch.from_numpy(self._output_perm[:total].astype(np.int32)[None, :].copy()).to(dev)
        c = gather_rows(c, self._perm_dev, 1, total, 1, self._num_cbs * n_max)          # interleave + concatenate (:414-418)
        c = c.reshape(-1, self._num_tx, total)
        if self._scrambler is not None:
            c = self._scrambler(c)
        return c.to(self.rdtype).reshape(shape[:-1] + [total])
