"""GPU parity of the config-1 / config-2 chain kernels (through the host classes and the C-ABI) against the oracle:
encoder (bit-exact vs the reference's generator-matrix goldens), Mapper (exact), Demapper (bit-exact vs oracle
kernel-math mode; rtol 1e-4 -- the north-star LLR tolerance -- vs the libm mode), AWGN / BinarySource (moments,
reproducibility), error counters (exact), and the end-to-end links of BASELINE.json configs[0] and configs[1]."""
import os
import numpy as np
import pytest
import torch

from oracle import ldpc as O
from oracle import mapping as M

pytestmark = pytest.mark.gpu


def test_encoder_vs_reference_goldens(cuda_device):
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ldpc_enc_golden.npz"))
    for k, n in g["params"]:
        u = np.unpackbits(g[f"u_{k}_{n}"], axis=1)[:, :k].astype(np.float32)
        c = np.unpackbits(g[f"c_{k}_{n}"], axis=1)[:, :n].astype(np.float32)
        enc = LDPC5GEncoder(int(k), int(n))
        out = enc(torch.from_numpy(u).to(cuda_device))
        assert out.shape == (4, n) and np.array_equal(out.cpu().numpy(), c), f"k={k} n={n}"


def test_encoder_interleaver_multidim_and_checks(cuda_device):
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder
    rng = np.random.default_rng(0)
    for k, n, m in ((300, 720, 6), (64, 192, 2), (1000, 2400, 4)):
        u = rng.integers(0, 2, (3, 5, k)).astype(np.float32)
        enc = LDPC5GEncoder(k, n, num_bits_per_symbol=m)
        c = enc(torch.from_numpy(u).to(cuda_device)).cpu().numpy()
        ref = O.LDPC5GEncoderRef(k, n, num_bits_per_symbol=m)(u)
        assert c.shape == (3, 5, n) and np.array_equal(c, ref)
    with pytest.raises(ValueError):
        LDPC5GEncoder(100, 200)(torch.full((2, 100), 0.5, device=cuda_device))
    with pytest.raises(ValueError):
        LDPC5GEncoder(100, 200)(torch.zeros((2, 99), device=cuda_device))
    # all-zero in -> all-zero out; systematic part
    enc = LDPC5GEncoder(500, 1000)
    assert float(enc(torch.zeros((2, 500), device=cuda_device)).abs().sum()) == 0.0


@pytest.mark.parametrize("m", [2, 4, 6, 8])
def test_mapper_exact(cuda_device, m):
    from sionna_b200.phy.mapping import Mapper
    rng = np.random.default_rng(m)
    bits = rng.integers(0, 2, (4, 7, 12 * m)).astype(np.float32)
    x, idx = Mapper("qam", m, return_indices=True)(torch.from_numpy(bits).to(cuda_device))
    xr, ir = M.mapper(bits, M.qam(m))
    assert np.array_equal(x.cpu().numpy(), xr) and np.array_equal(idx.cpu().numpy(), ir)
    x = Mapper("pam", 3)(torch.from_numpy(bits[..., :12]).to(cuda_device))
    assert np.array_equal(x.cpu().numpy(), M.mapper(bits[..., :12], M.pam(3))[0])


@pytest.mark.gpu
@pytest.mark.parametrize("m", [4, 6, 8, 10])
def test_demapper_app_high_snr_group_underflow(cuda_device, m):
    """`app` demapping when whole bit groups underflow relative to the dimension's largest exponent (|LLR| > 85, noise
    variance down to 1e-4): the kernel's out-of-line per-group path, including its shortcut that returns the group maximum
    without evaluating exp when every other member is below e^-21 of it, equals the oracle's per-group logsumexp bit for
    bit (kernel math) and the libm oracle to 1e-4."""
    from sionna_b200.phy.mapping import Demapper
    rng = np.random.default_rng(70 + m)
    pts = M.qam(m)
    x, _ = M.mapper(rng.integers(0, 2, (4, 7, 24 * m)), pts)
    dem = Demapper("app", "qam", m)
    for no in (2e-2, 2e-3, 1e-4):
        y = (x + (rng.normal(size=x.shape) + 1j * rng.normal(size=x.shape)) * np.sqrt(no / 2)).astype(np.complex64)
        llr = dem(torch.from_numpy(y).to(cuda_device), no).cpu().numpy()
        assert np.isfinite(llr).all() and np.abs(llr).max() > 85
        assert np.array_equal(llr, M.demapper(y, np.float32(no), pts, "app", math_mode=1))
        np.testing.assert_allclose(llr, M.demapper(y, np.float32(no), pts, "app", math_mode=0), rtol=1e-4, atol=2e-5)
    no_sym = (10.0 ** rng.uniform(-4, -1, size=y.shape)).astype(np.float32)
    out = dem(torch.from_numpy(y).to(cuda_device), torch.from_numpy(no_sym).to(cuda_device)).cpu().numpy()
    assert np.array_equal(out, M.demapper(y, no_sym, pts, "app", math_mode=1))



@pytest.mark.parametrize("method", ["app", "maxlog"])
@pytest.mark.parametrize("m", [2, 4, 6, 8, 10])
def test_demapper_vs_oracle(cuda_device, m, method):
    from sionna_b200.phy.mapping import Demapper
    rng = np.random.default_rng(10 + m)
    pts = M.qam(m)
    x, _ = M.mapper(rng.integers(0, 2, (6, 9, 16 * m)), pts)
    no = 0.2
    y = (x + (rng.normal(size=x.shape) + 1j * rng.normal(size=x.shape)) * np.sqrt(no / 2)).astype(np.complex64)
    yd = torch.from_numpy(y).to(cuda_device)
    dem = Demapper(method, "qam", m)
    # scalar no
    llr = dem(yd, no).cpu().numpy()
    assert llr.shape == (6, 9, 16 * m)
    assert np.array_equal(llr, M.demapper(y, np.float32(no), pts, method, math_mode=1))
    np.testing.assert_allclose(llr, M.demapper(y, np.float32(no), pts, method, math_mode=0), rtol=1e-4, atol=2e-5)
    # per-symbol no, per-batch no (broadcast from the left), priors
    no_sym = rng.uniform(0.05, 1.0, size=y.shape).astype(np.float32)
    assert np.array_equal(dem(yd, torch.from_numpy(no_sym).to(cuda_device)).cpu().numpy(),
                          M.demapper(y, no_sym, pts, method, math_mode=1))
    no_b = rng.uniform(0.05, 1.0, size=(6,)).astype(np.float32)
    assert np.array_equal(dem(yd, torch.from_numpy(no_b).to(cuda_device)).cpu().numpy(),
                          M.demapper(y, no_b, pts, method, math_mode=1))
    prior = rng.normal(size=y.shape + (m,)).astype(np.float32) * 2
    out = dem(yd, no, torch.from_numpy(prior).to(cuda_device)).cpu().numpy()
    assert np.array_equal(out, M.demapper(y, np.float32(no), pts, method, prior=prior, math_mode=1))
    np.testing.assert_allclose(out, M.demapper(y, np.float32(no), pts, method, prior=prior, math_mode=0),
                               rtol=1e-4, atol=5e-5)
    p1 = rng.normal(size=(m,)).astype(np.float32)
    assert np.array_equal(dem(yd, no, torch.from_numpy(p1).to(cuda_device)).cpu().numpy(),
                          M.demapper(y, np.float32(no), pts, method, prior=p1, math_mode=1))
    hard = Demapper(method, "qam", m, hard_out=True)(yd, no).cpu().numpy()
    assert np.array_equal(hard, (llr > 0).astype(np.float32))


@pytest.mark.parametrize("method", ["app", "maxlog"])
def test_demapper_generic_constellations_vs_oracle(cuda_device, method):
    """Non-separable constellations (PAM, rotated / custom points) take the 2-D kernel: bit-exact vs the oracle's
    reference-order formula in kernel math; the separable kernel and the 2-D kernel agree to rounding on square QAM."""
    from sionna_b200.phy.mapping import Demapper, Constellation
    rng = np.random.default_rng(77)
    no = np.float32(0.15)
    for name, pts in (("pam3", M.pam(3)), ("rot16", (M.qam(4) * np.exp(0.3j)).astype(np.complex64)),
                      ("qam7bits", (rng.normal(size=128) + 1j * rng.normal(size=128)).astype(np.complex64))):
        m = int(np.log2(len(pts)))
        assert M.separable_levels(pts) is None
        const = Constellation("pam", 3) if name == "pam3" else Constellation("custom", m, points=pts)
        y = (rng.normal(size=(5, 64)) + 1j * rng.normal(size=(5, 64))).astype(np.complex64)
        llr = Demapper(method, constellation=const)(torch.from_numpy(y).to(cuda_device), float(no)).cpu().numpy()
        assert np.array_equal(llr, M.demapper(y, no, const().cpu().numpy(), method, math_mode=1)), name
    # square QAM through BOTH kernels: a custom constellation holding the same points but perturbed by nothing
    pts = M.qam(6)
    y = (rng.normal(size=(4, 100)) + 1j * rng.normal(size=(4, 100))).astype(np.complex64)
    sep = Demapper(method, "qam", 6)(torch.from_numpy(y).to(cuda_device), float(no)).cpu().numpy()
    ref2d = M.demapper(y, no, pts, method, math_mode=0)                       # the reference's 2-D formula, libm
    np.testing.assert_allclose(sep, ref2d, rtol=1e-4, atol=1e-4)


def test_awgn_and_sources(cuda_device):
    from sionna_b200.phy import config
    from sionna_b200.phy.channel import AWGN
    from sionna_b200.phy.mapping import BinarySource
    from sionna_b200.phy.utils import complex_normal
    config.seed = 42
    b1 = BinarySource()([64, 4099])
    config.seed = 42
    b2 = BinarySource()([64, 4099])
    assert torch.equal(b1, b2) and b1.shape == (64, 4099)
    assert abs(float(b1.mean()) - 0.5) < 5e-3 and set(np.unique(b1.cpu().numpy())) == {0.0, 1.0}
    assert not torch.equal(BinarySource()([64, 4099]), b1)            # the stream advances
    s = BinarySource(seed=7)
    assert not torch.equal(s([1000]), s([1000]))
    assert torch.equal(BinarySource(seed=7)([1000]), BinarySource(seed=7)([1000]))
    x = torch.zeros((8, 100001), dtype=torch.complex64, device=cuda_device)
    y = AWGN()(x, 0.5)
    assert abs(float(y.real.var()) - 0.25) < 3e-3 and abs(float(y.imag.var()) - 0.25) < 3e-3
    assert abs(float(y.real.mean())) < 2e-3 and abs(float((y.real * y.imag).mean())) < 2e-3
    k = float(((y.real / 0.5) ** 4).mean())                           # Gaussian kurtosis = 3
    assert abs(k - 3.0) < 0.05
    no = torch.tensor([0.1, 0.2, 0.4, 0.8, 1.0, 2.0, 3.0, 4.0], device=cuda_device)
    y = AWGN()(x, no)                                                 # [8] broadcast from the left
    v = (y.abs() ** 2).mean(dim=1)
    assert torch.allclose(v, no, rtol=0.02)
    z = complex_normal([4, 50000], var=2.0)
    assert abs(float((z.abs() ** 2).mean()) - 2.0) < 0.03


def test_error_counters_exact(cuda_device):
    from sionna_b200.phy.utils import count_errors, count_block_errors, compute_ber, compute_bler, ErrorCounter
    rng = np.random.default_rng(3)
    b = rng.integers(0, 2, (7, 33, 1001)).astype(np.float32)
    bh = b.copy()
    flip = rng.random(b.shape) < 0.001
    bh[flip] = 1 - bh[flip]
    bd, hd = torch.from_numpy(b).to(cuda_device), torch.from_numpy(bh).to(cuda_device)
    assert int(count_errors(bd, hd)) == int(flip.sum())
    assert int(count_block_errors(bd, hd)) == int(flip.any(-1).sum())
    assert float(compute_ber(bd, hd)) == flip.sum() / flip.size
    assert float(compute_bler(bd, hd)) == flip.any(-1).sum() / (7 * 33)
    c = ErrorCounter()
    c.update(bd, hd).update(bd, bd)
    assert c.values() == (int(flip.sum()), int(flip.any(-1).sum()), 2 * flip.size, 2 * 7 * 33)


def test_config0_qpsk_awgn_llr_parity(cuda_device):
    """BASELINE configs[0]: QPSK Mapper -> AWGN -> Demapper("app"), batch 1024, Eb/N0 4 dB: LLR tensor vs the oracle
    fed with the identical received samples (rel <= 1e-4 vs libm mode, bit-exact vs kernel-math mode)."""
    from sionna_b200.phy.mapping import Mapper, Demapper, BinarySource
    from sionna_b200.phy.channel import AWGN
    from sionna_b200.phy.utils import ebnodb2no
    from sionna_b200.phy import config
    config.seed = 1
    no = ebnodb2no(4.0, 2, 1.0)
    b = BinarySource()([1024, 8448])
    x = Mapper("qam", 2)(b)
    y = AWGN()(x, no)
    llr = Demapper("app", "qam", 2)(y, no)
    yh, pts = y.cpu().numpy(), M.qam(2)
    assert np.array_equal(x.cpu().numpy(), M.mapper(b.cpu().numpy(), pts)[0])
    assert np.array_equal(llr.cpu().numpy(), M.demapper(yh, np.float32(no), pts, "app", math_mode=1))
    np.testing.assert_allclose(llr.cpu().numpy(), M.demapper(yh, np.float32(no), pts, "app", math_mode=0),
                               rtol=1e-4, atol=1e-5)
    ber = float(((llr > 0).float() != b).float().mean())
    from scipy.special import erfc
    assert abs(ber - 0.5 * erfc(np.sqrt(10 ** 0.4))) < 2e-3


def test_config1_ldpc_link_ber_matches_oracle(cuda_device):
    """BASELINE configs[1] at reduced batch: BinarySource -> LDPC5GEncoder(4224, 8448) -> QPSK -> AWGN ->
    Demapper("app") -> LDPC5GDecoder(20 it, boxplus-phi). The oracle (libm mode, reference summation order) decodes the
    same LLRs: decoded bits must agree (|dBER| <= 1e-6 on these inputs) and so must the encoder output."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    from sionna_b200.phy.mapping import Mapper, Demapper, BinarySource
    from sionna_b200.phy.channel import AWGN
    from sionna_b200.phy.utils import ebnodb2no
    from sionna_b200.phy import config
    config.seed = 100
    k, n, bs = 4224, 8448, 96
    enc = LDPC5GEncoder(k, n)
    dec = LDPC5GDecoder(enc, num_iter=20)
    enc_r = O.LDPC5GEncoderRef(k, n)
    dec_r = O.LDPC5GDecoderRef(enc_r, num_iter=20)
    tot = {"gpu": 0, "ref": 0, "diff": 0, "bits": 0}
    for ebno in (1.0, 1.5, 2.5):
        no = ebnodb2no(ebno, 2, k / n)
        u = BinarySource()([bs, k])
        c = enc(u)
        assert np.array_equal(c.cpu().numpy(), enc_r(u.cpu().numpy()))
        llr = Demapper("app", "qam", 2)(AWGN()(Mapper("qam", 2)(c), no), no)
        u_hat = dec(llr).cpu().numpy()
        u_ref = dec_r(llr.cpu().numpy())
        uu = u.cpu().numpy()
        tot["gpu"] += int((u_hat != uu).sum()); tot["ref"] += int((u_ref != uu).sum())
        tot["diff"] += int((u_hat != u_ref).sum()); tot["bits"] += uu.size
    assert abs(tot["gpu"] - tot["ref"]) / tot["bits"] <= 1e-6, tot
    assert tot["diff"] / tot["bits"] <= 1e-4, tot
    assert tot["gpu"] > 0                                   # 1 dB is below the waterfall: errors must be present


def test_sim_ber_on_device(cuda_device):
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    from sionna_b200.phy.mapping import Mapper, Demapper, BinarySource
    from sionna_b200.phy.channel import AWGN
    from sionna_b200.phy.utils import ebnodb2no, sim_ber
    from sionna_b200.phy import config
    config.seed = 5
    k, n = 500, 1000
    enc = LDPC5GEncoder(k, n)
    dec = LDPC5GDecoder(enc, num_iter=20, hard_out=False)
    src, mp, dm, ch = BinarySource(), Mapper("qam", 2), Demapper("app", "qam", 2), AWGN()

    def mc_fun(batch_size, ebno_db):
        no = ebnodb2no(ebno_db, 2, k / n)
        u = src([batch_size, k])
        return u, dec(dm(ch(mp(enc(u)), no), no))

    ber, bler = sim_ber(mc_fun, [0.0, 1.0, 2.0, 3.0, 4.0, 6.0], batch_size=500, max_mc_iter=4, soft_estimates=True,
                        num_target_block_errors=200, verbose=False)
    ber = ber.numpy()
    assert ber[0] > 0.05 and ber[0] > ber[1] > ber[2] and ber[3] < 5e-3
    assert ber[-1] == 0                                      # early stop: remaining points stay 0


def test_sim_ber_device_modes_keep_reference_stopping_semantics(cuda_device):
    """CUDA outputs: counters are read asynchronously ("free": once per SNR point; "lag": one batch late with a speculative
    next batch). The COUNTED batches must be those of the reference's rules (misc.py:717-760): a target reached by batch
    ii stops the point with exactly ii+1 batches counted, the speculative batch is discarded."""
    from sionna_b200.phy.utils import sim_ber
    calls = []

    def mc_fun(batch_size, ebno_db):
        calls.append(float(ebno_db))
        b = torch.zeros(batch_size, 10, device=cuda_device)
        b_hat = b.clone()
        if float(ebno_db) < 1.0:
            b_hat[0, 0] = 1.0                                 # exactly one bit / block error per batch
        return b, b_hat

    # "lag": 3 batches reach the target (+1 speculative, not counted); 2 dB: 10 error-free batches -> early stop
    ber, bler = sim_ber(mc_fun, [0.0, 2.0, 4.0], batch_size=5, max_mc_iter=10, num_target_bit_errors=3, verbose=False)
    assert float(ber[0]) == pytest.approx(3 / (3 * 50)) and float(bler[0]) == pytest.approx(3 / 15)
    assert float(ber[1]) == 0 and float(ber[2]) == 0
    assert calls.count(0.0) == 4 and calls.count(2.0) == 10 and calls.count(4.0) == 0
    # target reached exactly by the last allowed batch: no speculation beyond max_mc_iter
    calls.clear()
    ber, _ = sim_ber(mc_fun, [0.0], batch_size=5, max_mc_iter=3, num_target_block_errors=3, verbose=False)
    assert len(calls) == 3 and float(ber[0]) == pytest.approx(3 / 150)
    # "free": no per-batch rule -> every batch counted, totals read once per point
    calls.clear()
    ber, bler = sim_ber(mc_fun, [0.0, 0.5], batch_size=5, max_mc_iter=7, early_stop=False, verbose=True)
    assert len(calls) == 14 and float(ber[0]) == pytest.approx(7 / 350) and float(bler[1]) == pytest.approx(7 / 35)
    # "sync": a callback sees exact per-batch totals
    seen = []

    def cb(mc_iter, snr_idx, ebnos, bit_errors, block_errors, nb_bits, nb_blocks):
        seen.append((mc_iter, int(bit_errors[snr_idx]), int(nb_bits[snr_idx])))
        return sim_ber.CALLBACK_NEXT_SNR if mc_iter == 2 else sim_ber.CALLBACK_CONTINUE
    sim_ber(mc_fun, [0.0], batch_size=5, max_mc_iter=10, callback=cb, verbose=False)
    assert seen == [(0, 1, 50), (1, 2, 100), (2, 3, 150)]


def test_double_precision_falls_back_to_the_single_precision_kernels(cuda_device):
    """precision="double" (reference block.py:25-31): blocks accept / return float64 / complex128 and warn once that the
    arithmetic runs on the fp32 kernels; values equal the single-precision results; pure gathers stay bit exact."""
    import warnings
    from sionna_b200.phy.block import PrecisionWarning
    from sionna_b200.phy.mapping import Mapper, Demapper, BinarySource
    from sionna_b200.phy.channel import AWGN
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    from sionna_b200.phy.mimo import lmmse_equalizer
    from sionna_b200.phy import config
    config.seed = 3
    k, n = 200, 400
    b = BinarySource()([16, k])
    enc_d, enc_s = LDPC5GEncoder(k, n, precision="double"), LDPC5GEncoder(k, n)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        c = enc_d(b.double())
        x = Mapper("qam", 4, precision="double")(c)
        y = AWGN(precision="double")(x, 0.05)
        llr = Demapper("app", "qam", 4, precision="double")(y, 0.05)
        u = LDPC5GDecoder(enc_d, num_iter=10, hard_out=False, precision="double")(llr)
    assert any(issubclass(i.category, PrecisionWarning) for i in w)
    assert c.dtype == torch.float64 and x.dtype == torch.complex128 and llr.dtype == torch.float64 and u.dtype == torch.float64
    assert torch.equal(c.float(), enc_s(b))
    llr_s = Demapper("app", "qam", 4)(y.to(torch.complex64), 0.05)
    # the double-precision constellation is normalised in float64 and then rounded (one ulp from the float32-normalised one)
    assert torch.allclose(llr.float(), llr_s, rtol=2e-6, atol=1e-5)
    assert torch.equal(u.float(), LDPC5GDecoder(enc_s, num_iter=10, hard_out=False)(llr.float()))
    rng = np.random.default_rng(0)
    h = torch.from_numpy(rng.normal(size=(5, 4, 2)) + 1j * rng.normal(size=(5, 4, 2))).to(cuda_device)
    yv = torch.from_numpy(rng.normal(size=(5, 4)) + 1j * rng.normal(size=(5, 4))).to(cuda_device)
    s = torch.eye(4, dtype=torch.complex128, device=cuda_device) * 0.1
    xh, ne = lmmse_equalizer(yv, h, s, precision="double")
    assert xh.dtype == torch.complex128 and ne.dtype == torch.float64
