"""FEC test utilities (mirror of parts of /root/reference/src/sionna/phy/fec/utils.py needed by the LDPC tests)."""
import os
import numpy as np
import torch

from ..block import Block
from ..config import config
from .._lib_helpers import philox_normal


class GaussianPriorSource(Block):
    """Fake LLRs of an all-zero codeword sent with BPSK over AWGN (fec/utils.py:16-114): logits ~ N(-mu, sigma^2),
    sigma^2 = 4/no, mu = sigma^2/2. ``call(output_shape, no)``."""

    def __init__(self, *, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)

    def __call__(self, output_shape, no=1.0, mi=None):
        return self.call(output_shape, no, mi)

    def call(self, output_shape, no=1.0, mi=None):
        if mi is not None:
            raise NotImplementedError("GaussianPriorSource: the mutual-information parametrisation is not provided.")
        no = float(torch.as_tensor(no).reshape(-1)[0]) if not isinstance(no, (int, float)) else float(no)
        no = max(no, 1e-7)
        sigma_llr = np.sqrt(4.0 / no)
        mu_llr = sigma_llr ** 2 / 2
        shape = [int(s) for s in output_shape]
        return philox_normal(shape, -mu_llr, sigma_llr, self.device).to(self.rdtype)


def load_parity_check_examples(pcm_id, verbose=False):
    """Built-in example parity-check matrices (fec/utils.py:478-531): 0 = (7,4) Hamming, 1 = BCH(63,45),
    2 = BCH(127,106), 3 = (3,6)-regular LDPC n=100, 4 = 802.11n LDPC n=648. Returns ``pcm, k, n, coderate``."""
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ldpc", "codes", "example_pcms.npz")
    with np.load(p) as d:
        pcm = np.array(d[f"pcm{int(pcm_id)}"], dtype=np.int64)
    n = int(pcm.shape[1])
    k = int(n - pcm.shape[0])
    coderate = k / n
    if verbose:
        print(f"\nn: {n}, k: {k}, coderate: {coderate:.3f}")
    return pcm, k, n, coderate
