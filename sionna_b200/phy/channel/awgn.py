"""AWGN channel (mirror of /root/reference/src/sionna/phy/channel/awgn.py:10-78)."""
import torch

from ..block import Block
from ..config import config
from ..mapping import _broadcast_inner
from ..._lib import lib, check, ptr, current_stream


class AWGN(Block):
    """AWGN(precision=None): ``y = x + w`` with ``w ~ CN(0, no)`` (``no/2`` per real dimension). ``no`` is a scalar
    or any tensor broadcastable to ``x`` after appending dimensions on the right (awgn.py:63-78)."""

    def __init__(self, *, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)

    def call(self, x, no):
        if self.precision != "single":
            raise NotImplementedError("sb_awgn is a complex64 kernel; precision='double' is not available.")
        dev = self.device
        x = x.to(device=dev, dtype=torch.complex64).contiguous()
        no_t, inner = _broadcast_inner(no, x.shape, dev, torch.float32)
        y = torch.empty_like(x)
        seed, off = config.next_philox()
        check(lib().sb_awgn(ptr(x), ptr(no_t), inner, ptr(y), x.numel(), seed, off, current_stream()), "sb_awgn")
        return y
