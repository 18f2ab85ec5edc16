"""MIMO utilities (mirror of /root/reference/src/sionna/phy/mimo/utils.py:292-357): noise whitening."""
import torch

from ..config import config
from ..._lib import lib, check, ptr, current_stream


def whiten_channel(y, h, s, return_s=True):
    r"""Whitens ``y = H x + n`` with ``E[n n^H] = S = L L^H``: returns ``L^-1 y``, ``L^-1 H`` (and the identity as the
    whitened covariance if ``return_s``), mimo/utils.py:292-357. One kernel (``sb_mimo_linalg`` mode 1): Cholesky
    factorisation and the two forward substitutions per matrix. y [..., M], h [..., M, K], s [..., M, M]."""
    from ..block import fallback_to_single
    wide = fallback_to_single("whiten_channel", None)
    dev = config.device
    y = torch.as_tensor(y).to(device=dev, dtype=torch.complex64)
    h = torch.as_tensor(h).to(device=dev, dtype=torch.complex64)
    s = torch.as_tensor(s).to(device=dev, dtype=torch.complex64)
    m, k = h.shape[-2], h.shape[-1]
    lead = torch.broadcast_shapes(y.shape[:-1], h.shape[:-2], s.shape[:-2])
    y = y.expand(*lead, m).contiguous()
    h = h.expand(*lead, m, k).contiguous()
    s = s.expand(*lead, m, m).contiguous()
    yw, hw = torch.empty_like(y), torch.empty_like(h)
    check(lib().sb_mimo_linalg(1, ptr(y), ptr(h), ptr(s), ptr(yw), ptr(hw), y.numel() // m, m, k, current_stream()),
          "sb_mimo_linalg")
    if wide:
        yw, hw = yw.to(torch.complex128), hw.to(torch.complex128)
    if return_s:
        sw = torch.eye(m, dtype=yw.dtype, device=dev).expand(*lead, m, m).contiguous()
        return yw, hw, sw
    return yw, hw
