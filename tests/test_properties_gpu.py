"""Size-independent properties of the CUDA path at BASELINE.json's full sizes (batch 4096, n = 8448) and edge cases
(empty batches, erasures, batch independence), complementing the oracle comparisons that run at small sizes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
K, N, BATCH = 4224, 8448, 4096


@pytest.fixture(scope="module")
def chain(cuda_device):
    from sionna_b200.phy import config
    from sionna_b200.phy.mapping import BinarySource, Mapper, Demapper
    from sionna_b200.phy.channel import AWGN
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    from sionna_b200.phy.utils import ebnodb2no
    config.seed = 99
    enc = LDPC5GEncoder(K, N)
    return {"src": BinarySource(), "enc": enc, "map": Mapper("qam", 2), "demap": Demapper("app", "qam", 2),
            "awgn": AWGN(), "no": lambda e: ebnodb2no(e, 2, K / N),
            "dec": lambda **kw: LDPC5GDecoder(enc, num_iter=20, **kw)}


def test_full_size_round_trip(chain):
    """encode -> QPSK -> AWGN (3 dB) -> demap -> 20 BP iterations recovers all 4096 x 4224 information bits; and the
    encoder output satisfies every parity check (syndrome of the hard-decided noiseless LLRs is decoded in 0 changes)."""
    u = chain["src"]([BATCH, K])
    c = chain["enc"](u)
    assert list(c.shape) == [BATCH, N]
    no = chain["no"](3.0)
    llr = chain["demap"](chain["awgn"](chain["map"](c), no), no)
    for rule in ("boxplus-phi", "minsum"):
        u_hat = chain["dec"](cn_update=rule)(llr)
        assert torch.equal(u_hat, u), rule
    c_hat = chain["dec"](return_infobits=False)(20.0 * (2.0 * c - 1.0))       # noiseless: the codeword is a fixed point
    assert torch.equal(c_hat, c)


@pytest.mark.parametrize("rule", ["boxplus-phi", "boxplus", "minsum", "offset-minsum"])
def test_codeword_symmetry_is_bit_exact(chain, rule):
    """BP is symmetric: flipping the channel LLR signs by a codeword c flips the output LLR signs by c and leaves every
    magnitude untouched, bit for bit (sign handling and sign-symmetric rounding of the kernels)."""
    b = 256
    u = chain["src"]([b, K])
    c = chain["enc"](u)
    no = chain["no"](1.0)                                                     # waterfall: plenty of undecided bits
    zeros = torch.zeros_like(c)
    noise_llr = chain["demap"](chain["awgn"](chain["map"](zeros), no), no)    # all-zero codeword through the channel
    dec = chain["dec"](cn_update=rule, hard_out=False, return_infobits=False)
    out0 = dec(noise_llr)
    s = 1.0 - 2.0 * c                                                         # the reference's logits: bit 1 <-> positive
    out1 = dec(noise_llr * s)
    assert torch.equal(out1, out0 * s)


def test_batch_independence_and_empty_batch(chain):
    """Codeword i's result does not depend on the batch it is decoded in; empty batches pass through every block."""
    u = chain["src"]([300, K])
    no = chain["no"](1.2)
    llr = chain["demap"](chain["awgn"](chain["map"](chain["enc"](u)), no), no)
    dec = chain["dec"](hard_out=False)
    full = dec(llr)
    idx = torch.tensor([0, 7, 150, 299], device=llr.device)
    assert torch.equal(dec(llr[idx].contiguous()), full[idx])
    assert torch.equal(dec(llr.reshape(3, 100, N)).reshape(300, K), full)     # leading dimensions are free
    e = torch.zeros((0, K), device=llr.device)
    c0 = chain["enc"](e)
    assert list(c0.shape) == [0, N]
    x0 = chain["map"](c0)
    assert list(x0.shape) == [0, N // 2]
    l0 = chain["demap"](chain["awgn"](x0, no), no)
    assert list(l0.shape) == [0, N]
    assert list(dec(l0).shape) == [0, K]


def test_erasures_are_recovered(chain):
    """Rate-matching view of erasures: zero LLRs on 15 % of the positions of every codeword are filled in."""
    u = chain["src"]([512, K])
    c = chain["enc"](u)
    llr = 8.0 * (2.0 * c - 1.0)
    g = torch.Generator(device="cpu").manual_seed(1)
    erase = (torch.rand((512, N), generator=g) < 0.15).to(llr.device)
    llr = torch.where(erase, torch.zeros_like(llr), llr)
    assert torch.equal(chain["dec"](return_infobits=False)(llr), c)


def test_error_counter_checksum(chain):
    """count_errors over the full batch equals the sum over its halves (the all-reduce of sim_ber relies on additivity)."""
    from sionna_b200.phy.utils import ErrorCounter
    u = chain["src"]([BATCH, K])
    v = chain["src"]([BATCH, K])
    tot = ErrorCounter(u.device)
    tot.update(u, v)
    halves = ErrorCounter(u.device)
    halves.update(u[:BATCH // 2], v[:BATCH // 2])
    halves.update(u[BATCH // 2:], v[BATCH // 2:])
    assert tot.values() == halves.values()
    assert tot.values()[0] == int((u != v).sum()) and tot.values()[2] == BATCH * K
