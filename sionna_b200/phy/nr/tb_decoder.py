"""5G NR transport-block decoder (mirror of /root/reference/src/sionna/phy/nr/tb_decoder.py:20-213): descrambling of the
LLRs, de-interleaving / un-puncturing gather, LDPC BP decoding of all code blocks, CB CRC removal, TB CRC check."""
import numpy as np
import torch

from ..block import Block
from ..fec.crc import CRCDecoder
from ..fec.ldpc import LDPC5GDecoder
from ..ofdm.resource_grid import gather_rows
from .tb_encoder import TBEncoder


class TBDecoder(Block):
    """TBDecoder(encoder, num_bp_iter=20, cn_update="boxplus-phi", vn_update="sum", precision=None)

    ``[..., num_coded_bits]`` LLRs (logits) -> ``(u_hat [..., tb_size], tb_crc_status [...] bool)`` (tb_decoder.py:146-213)."""

    def __init__(self, encoder, num_bp_iter=20, cn_update="boxplus-phi", vn_update="sum", precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        assert isinstance(encoder, TBEncoder), "encoder must be TBEncoder."
        self._tb_encoder = encoder
        self._num_cbs = encoder.num_cbs
        self._decoder = LDPC5GDecoder(encoder.ldpc_encoder, num_iter=num_bp_iter, cn_update=cn_update, vn_update=vn_update,
                                      hard_out=True, return_infobits=True, precision=precision)
        self._tb_crc_decoder = CRCDecoder(encoder.tb_crc_encoder, precision=precision)
        self._cb_crc_decoder = CRCDecoder(encoder.cb_crc_encoder, precision=precision) \
            if encoder.cb_crc_encoder is not None else None
        self._idx = None

    tb_size = property(lambda self: self._tb_encoder.tb_size)
    k = property(lambda self: self._tb_encoder.tb_size)
    n = property(lambda self: self._tb_encoder.n)

    def build(self, input_shapes):
        assert input_shapes[-1] == self.n, f"Invalid input shape. Expected input length is {self.n}."

    def call(self, inputs):
        enc = self._tb_encoder
        dev = self.device
        shape = list(inputs.shape)
        llr = inputs.to(device=dev, dtype=torch.float32).reshape(-1, enc.num_tx, enc.n)
        if enc.scrambler is not None:
            llr = enc.scrambler(llr, binary=False)                             # Descrambler(binary=False) (:73-78)
        n_ldpc_out = enc.ldpc_encoder.n
        total = n_ldpc_out * enc.num_cbs
        if self._idx is None or self._idx.device != dev:
            # position p of the padded, de-interleaved word takes input output_perm_inv[p]; fillers (>= n) read as 0
            src = np.asarray(enc.output_perm_inv).astype(np.int64)
            src = np.where(src < enc.n, src, -1).astype(np.int32)
            self._idx = torch.from_numpy(src[None, :].copy()).to(dev)
        llr_int = gather_rows(llr.reshape(-1, enc.n).contiguous(), self._idx, 1, total, 1, enc.n)
        llr_cb = llr_int.reshape(-1, enc.num_tx, self._num_cbs, n_ldpc_out)
        u_hat_cb = self._decoder(llr_cb)
        if self._cb_crc_decoder is not None:
            u_hat_cb, _ = self._cb_crc_decoder(u_hat_cb)
        u_hat_tb = u_hat_cb.reshape(-1, enc.num_tx, self.tb_size + enc.tb_crc_encoder.crc_length)
        u_hat, crc_ok = self._tb_crc_decoder(u_hat_tb)
        u_hat = u_hat.reshape(shape[:-1] + [self.tb_size])
        crc_ok = crc_ok.reshape(shape[:-1])
        if enc.k_padding > 0:
            u_hat = u_hat[..., :-enc.k_padding]
        return u_hat.to(self.rdtype), crc_ok
