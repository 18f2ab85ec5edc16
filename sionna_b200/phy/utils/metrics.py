"""Error metrics (mirror of /root/reference/src/sionna/phy/utils/metrics.py:9-144).

``count_errors`` / ``count_block_errors`` run the fused ``sb_count_errors`` kernel; ``ErrorCounter`` keeps the
four int64 counters {bit errors, block errors, bits, blocks} on the device so that a Monte-Carlo driver can
all-reduce them with one 32-byte NCCL call instead of gathering bit tensors.
"""
import torch

from ..config import config
from ..._lib import lib, check, ptr, current_stream


def _as_rows(b, b_hat):
    b = torch.as_tensor(b)
    b_hat = torch.as_tensor(b_hat)
    dev = b.device if b.is_cuda else (b_hat.device if b_hat.is_cuda else config.device)
    b = b.to(device=dev, dtype=torch.float32)
    b_hat = b_hat.to(device=dev, dtype=torch.float32)
    if b.shape != b_hat.shape:
        b, b_hat = torch.broadcast_tensors(b, b_hat)
    k = b.shape[-1] if b.dim() > 0 else 1
    return b.reshape(-1, k).contiguous(), b_hat.reshape(-1, k).contiguous(), k


class ErrorCounter:
    """Device-resident int64[4] accumulator: bit errors, block errors, bits, blocks."""

    def __init__(self, device=None):
        self.counters = torch.zeros(4, dtype=torch.int64, device=device or config.device)

    def update(self, b, b_hat):
        b2, h2, k = _as_rows(b, b_hat)
        check(lib().sb_count_errors(ptr(b2), ptr(h2), b2.shape[0], k, ptr(self.counters), current_stream()),
              "sb_count_errors")
        return self

    def reset(self):
        self.counters.zero_()

    def values(self):
        """(bit_errors, block_errors, num_bits, num_blocks) as Python ints (synchronises)."""
        return tuple(int(v) for v in self.counters.cpu().tolist())


def count_errors(b, b_hat):
    """Number of positions where ``b`` and ``b_hat`` differ, int64 scalar tensor (metrics.py:94-114)."""
    return ErrorCounter().update(b, b_hat).counters[0]


def count_block_errors(b, b_hat):
    """Number of rows (last dim = block) with at least one difference (metrics.py:116-144)."""
    return ErrorCounter().update(b, b_hat).counters[1]


def compute_ber(b, b_hat, precision="double"):
    """Bit error rate (metrics.py:9-37)."""
    c = ErrorCounter().update(b, b_hat).counters
    dt = torch.float64 if precision == "double" else torch.float32
    return (c[0].to(torch.float64) / c[2].to(torch.float64)).to(dt)


def compute_bler(b, b_hat, precision="double"):
    """Block error rate over the last dimension (metrics.py:66-92)."""
    c = ErrorCounter().update(b, b_hat).counters
    dt = torch.float64 if precision == "double" else torch.float32
    return (c[1].to(torch.float64) / c[3].to(torch.float64)).to(dt)
