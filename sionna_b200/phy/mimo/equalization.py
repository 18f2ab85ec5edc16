"""MIMO equalisation (mirror of /root/reference/src/sionna/phy/mimo/equalization.py:101-233)."""
import torch

from ..config import config
from ..._lib import lib, check, ptr, current_stream


def lmmse_equalizer(y, h, s, whiten_interference=True, precision=None):
    r"""LMMSE equaliser for ``y = H x + n`` with ``E[n n^H] = S``: returns the unbiased soft symbols
    ``x_hat = diag(G H)^-1 G y`` with ``G = H^H (H H^H + S)^-1`` and the effective noise variances
    ``no_eff = diag(diag(GH)^-1 - I)`` (equalization.py:101-233). The interference is whitened first with the Cholesky
    factor of S, then ``G = (H_w^H H_w + I)^-1 H_w^H`` is obtained by a second Cholesky factorisation
    (``whiten_interference=True``, the reference default; the alternative formulation is not provided).

    y [..., M], h [..., M, K], s [..., M, M] -> x_hat [..., K] complex, no_eff [..., K] real."""
    if precision is None:
        precision = config.precision
    if precision != "single":
        raise NotImplementedError("sb_lmmse_equalize is a complex64 kernel; precision='double' is not available.")
    if not whiten_interference:
        raise NotImplementedError("lmmse_equalizer: only whiten_interference=True is provided.")
    dev = config.device
    y = torch.as_tensor(y).to(device=dev, dtype=torch.complex64)
    h = torch.as_tensor(h).to(device=dev, dtype=torch.complex64)
    s = torch.as_tensor(s).to(device=dev, dtype=torch.complex64)
    m, k = h.shape[-2], h.shape[-1]
    lead = torch.broadcast_shapes(y.shape[:-1], h.shape[:-2], s.shape[:-2])
    y = y.expand(*lead, m).contiguous()
    h = h.expand(*lead, m, k).contiguous()
    s = s.expand(*lead, m, m).contiguous()
    num = y.numel() // m
    x_hat = torch.empty(*lead, k, dtype=torch.complex64, device=dev)
    no_eff = torch.empty(*lead, k, dtype=torch.float32, device=dev)
    check(lib().sb_lmmse_equalize(ptr(y), ptr(h), ptr(s), ptr(x_hat), ptr(no_eff), num, m, k, current_stream()),
          "sb_lmmse_equalize")
    return x_hat, no_eff
