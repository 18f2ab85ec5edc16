"""CRC and TB scrambling (SURVEY.md section 8(f2) pieces): host logic on CPU, kernels on the GPU."""
import os
import numpy as np
import pytest
import torch

from oracle import nr as R

POLS = ["CRC24A", "CRC24B", "CRC24C", "CRC16", "CRC11", "CRC6"]
GOLD = os.path.join(os.path.dirname(__file__), "golden", "crc_golden.npz")


def test_crc_oracle_and_generator_rows_vs_reference_vectors():
    from sionna_b200.phy.fec.crc import CRCEncoder
    g = np.load(GOLD)
    rng = np.random.default_rng(0)
    for pol in POLS:
        u, x = g[f"u_{pol}"], g[f"x_{pol}"]
        assert np.array_equal(R.crc_parity(u[0], pol), x)                       # oracle vs the reference's KAT
        enc = CRCEncoder(pol)
        assert enc.crc_length == len(x)
        for k in (10, 57, 1000):                                                # product's packed generator rows vs oracle
            rows = enc._gen_rows(k)
            b = rng.integers(0, 2, k)
            acc = np.bitwise_xor.reduce(np.where(b.astype(bool), rows, 0).astype(np.uint32))
            par = (acc >> np.arange(enc.crc_length - 1, -1, -1)) & 1
            assert np.array_equal(par, R.crc_parity(b, pol))


def test_prng_sequence_known_answer_of_the_reference():
    """Gold sequence (product host code and oracle) vs the reference's 100-bit known answer (test_nr_utils.py:280-296);
    invalid arguments are rejected as there (:273-278)."""
    import json
    from sionna_b200.phy.fec.scrambling import generate_prng_seq
    with open(os.path.join(os.path.dirname(__file__), "golden", "prng_golden.json")) as f:
        g = json.load(f)
    c_init = g["n_rnti"] * 2 ** 15 + g["n_id"]
    ref = np.array(g["s_ref"], float)
    assert np.array_equal(generate_prng_seq(g["l"], c_init), ref)
    assert np.array_equal(R.generate_prng_seq(g["l"], c_init), ref)
    assert not np.array_equal(generate_prng_seq(g["l"], c_init + 1), ref)
    for bad in ([-1, 10], [10, -1], [100, 2 ** 32], [10.2, 10], [10, 10.2]):
        with pytest.raises(AssertionError):
            generate_prng_seq(bad[0], bad[1])


def test_prng_sequence_matches_literal_restatement():
    from sionna_b200.phy.fec.scrambling import generate_prng_seq, TB5GScrambler
    for c_init in (0, 1, 1000, 2 ** 31 - 1, 12345678):
        assert np.array_equal(generate_prng_seq(500, c_init), R.generate_prng_seq(500, c_init))
    with pytest.raises(ValueError):
        TB5GScrambler(n_rnti=70000)
    with pytest.raises(ValueError):
        TB5GScrambler(n_id=1024)
    with pytest.raises(TypeError):
        TB5GScrambler(channel_type="PUCCH")


@pytest.mark.gpu
def test_crc_kernels(cuda_device):
    from sionna_b200.phy.fec.crc import CRCEncoder, CRCDecoder
    g = np.load(GOLD)
    rng = np.random.default_rng(1)
    for pol in POLS:
        enc = CRCEncoder(pol)
        x = enc(torch.from_numpy(g[f"u_{pol}"].astype(np.float32)).to(cuda_device)).cpu().numpy()
        assert np.array_equal(x.reshape(-1)[-enc.crc_length:], g[f"x_{pol}"])   # test_crc.py:177-199
        assert enc.k == 10 and enc.n == 10 + enc.crc_length
        dec = CRCDecoder(enc)
        for shape in ([100], [100, 10], [4, 2, 100], [1, 100000]):
            u = rng.integers(0, 2, shape).astype(np.float32)
            xc = enc(torch.from_numpy(u).to(cuda_device))
            assert np.array_equal(xc.cpu().numpy(), R.crc_encode(u, pol)) if np.prod(shape) <= 4000 else True
            u2, ok = dec(xc)
            assert bool(ok.all()) and np.array_equal(u2.cpu().numpy(), u) and ok.shape == tuple(shape[:-1]) + (1,)
        xe = enc(torch.from_numpy(rng.integers(0, 2, (200, 64)).astype(np.float32)).to(cuda_device)).clone()
        pos = torch.from_numpy(rng.integers(0, xe.shape[-1], 200)).to(cuda_device)
        xe[torch.arange(200, device=cuda_device), pos] = 1 - xe[torch.arange(200, device=cuda_device), pos]
        assert not bool(dec(xe)[1].any())                                       # every single-bit error is detected


@pytest.mark.gpu
def test_scrambler_kernel(cuda_device):
    from sionna_b200.phy.fec.scrambling import TB5GScrambler
    rng = np.random.default_rng(2)
    b = rng.integers(0, 2, (5, 3, 700)).astype(np.float32)
    s = TB5GScrambler(n_rnti=77, n_id=300)
    y = s(torch.from_numpy(b).to(cuda_device))
    seq = R.generate_prng_seq(700, 77 * 2 ** 15 + 300).astype(np.float32)
    assert np.array_equal(y.cpu().numpy(), np.abs(b - seq))
    assert np.array_equal(s(y).cpu().numpy(), b)                                # involution
    llr = rng.normal(size=b.shape).astype(np.float32)
    z = s(torch.from_numpy(llr).to(cuda_device), binary=False).cpu().numpy()
    assert np.array_equal(z, llr * (1 - 2 * seq))
    ms = TB5GScrambler(n_rnti=[1, 2, 3], n_id=[5, 6, 7], channel_type="PDSCH", codeword_index=1)
    ym = ms(torch.from_numpy(b).to(cuda_device)).cpu().numpy()
    for i, (nr, ni) in enumerate(zip([1, 2, 3], [5, 6, 7])):
        sq = R.generate_prng_seq(700, nr * 2 ** 15 + 2 ** 14 + ni).astype(np.float32)
        assert np.array_equal(ym[:, i], np.abs(b[:, i] - sq))


TB_GOLD = os.path.join(os.path.dirname(__file__), "golden", "tb_golden.npz")


def _tb_case(g, i):
    k, n, n_id, n_rnti, m, nl = [int(v) for v in g[f"p_{i}"]]
    u = np.unpackbits(g[f"u_{i}"], axis=1)[:, :k]
    c = np.unpackbits(g[f"c_{i}"], axis=1)[:, :n]
    cn = np.unpackbits(g[f"cn_{i}"], axis=1)[:, :n]
    return k, n, n_id, n_rnti, m, nl, float(g[f"r_{i}"]), u, c, cn


@pytest.mark.parametrize("i", [0, 1, 3, 7])
def test_tb_oracle_vs_reference_vectors(i):
    """oracle TB chain == the reference's stored transport-block vectors (test/unit/nr/tb_refs, test_tb_encoder.py:17-63)."""
    g = np.load(TB_GOLD)
    k, n, n_id, n_rnti, m, nl, r, u, c, cn = _tb_case(g, i)
    assert np.array_equal(R.tb_encode(u, n, r, m, nl, n_rnti, n_id, scramble=False), cn)
    assert np.array_equal(R.tb_encode(u, n, r, m, nl, n_rnti, n_id), c)


def test_tb_size_host_logic_matches_oracle():
    from sionna_b200.phy.nr import calculate_tb_size
    rng = np.random.default_rng(0)
    for _ in range(300):
        m = int(rng.choice([2, 4, 6, 8])); nl = int(rng.choice([1, 2, 4]))
        n = int(rng.integers(20, 20000)) * m * nl
        r = float(rng.uniform(0.1, 0.9))
        k = int(rng.integers(24, max(25, int(0.9 * n))))
        a = calculate_tb_size(m, r, target_tb_size=k, num_coded_bits=n, num_layers=nl)
        b = R.tb_params(k, n, r, m, nl)
        assert tuple(int(v) for v in a[:5]) == tuple(int(v) for v in b[:5]) and list(a[5]) == list(b[5])


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(8))
def test_tb_encoder_decoder_vs_reference_vectors(cuda_device, i):
    from sionna_b200.phy.nr import TBEncoder, TBDecoder
    g = np.load(TB_GOLD)
    k, n, n_id, n_rnti, m, nl, r, u, c, cn = _tb_case(g, i)
    enc = TBEncoder(target_tb_size=k, num_coded_bits=n, target_coderate=r, num_bits_per_symbol=m, num_layers=nl,
                    n_rnti=n_rnti, n_id=n_id, channel_type="PUSCH", codeword_index=0, use_scrambler=True)
    ud = torch.from_numpy(u.astype(np.float32)).to(cuda_device)
    cd = enc(ud)
    assert np.array_equal(cd.cpu().numpy(), c.astype(np.float32))
    enc2 = TBEncoder(target_tb_size=k, num_coded_bits=n, target_coderate=r, num_bits_per_symbol=m, num_layers=nl,
                     n_rnti=n_rnti, n_id=n_id, use_scrambler=False)
    assert np.array_equal(enc2(ud).cpu().numpy(), cn.astype(np.float32))
    dec = TBDecoder(enc, cn_update="minsum")                 # min-sum does not need correctly scaled LLRs (test_tb_encoder.py:55)
    u_hat, ok = dec(2 * cd - 1)
    assert np.array_equal(u_hat.cpu().numpy(), u.astype(np.float32)) and bool(ok.all())
    bad = (2 * cd - 1).clone()
    bad[0, : n // 3] *= -1                                    # heavy corruption -> TB CRC must fail
    assert not bool(dec(bad)[1].any())


@pytest.mark.gpu
def test_tb_multi_stream(cuda_device):
    from sionna_b200.phy.nr import TBEncoder, TBDecoder
    rng = np.random.default_rng(3)
    enc = TBEncoder(target_tb_size=6000, num_coded_bits=12000, target_coderate=0.5, num_bits_per_symbol=4,
                    n_rnti=[10, 20, 30], n_id=[1, 2, 3])
    u = rng.integers(0, 2, (4, 3, enc.k)).astype(np.float32)
    c = enc(torch.from_numpy(u).to(cuda_device))
    assert c.shape == (4, 3, 12000)
    for s, (nr, ni) in enumerate(zip([10, 20, 30], [1, 2, 3])):
        assert np.array_equal(c[:, s].cpu().numpy(), R.tb_encode(np.concatenate([u[:, s], np.zeros((4, enc.k_padding))], 1),
                                                                 12000, 0.5, 4, 1, nr, ni))
    u_hat, ok = TBDecoder(enc, cn_update="minsum")(2 * c - 1)
    assert np.array_equal(u_hat.cpu().numpy(), u) and ok.shape == (4, 3) and bool(ok.all())


@pytest.mark.gpu
def test_scrambler_descrambler_and_symbol_sources(cuda_device):
    """Scrambler / Descrambler (scrambling.py:20-579) and the SymbolSource family (mapping.py:1140-1555)."""
    import torch
    from sionna_b200.phy.fec.scrambling import Scrambler, Descrambler, TB5GScrambler
    from sionna_b200.phy.mapping import QAMSource, PAMSource, SymbolInds2Bits, Constellation
    from sionna_b200.phy.channel import RayleighBlockFading
    x = (torch.rand(8, 3, 100, device=cuda_device) > 0.5).float()
    s = Scrambler(seed=42)
    y = s(x)
    assert torch.equal(s(y), x) and not torch.equal(y, x)
    assert 0.4 < float((y != x).float().mean()) < 0.6
    assert torch.equal(Scrambler(seed=42)(x), y)                                 # seed defines the sequence
    assert not torch.equal(Scrambler(seed=43)(x), y)
    llr = torch.randn(8, 3, 100, device=cuda_device)
    d = Descrambler(s, binary=False)
    flipped = s(llr, binary=False)
    assert torch.equal(d(flipped), llr)
    assert torch.equal(flipped.abs(), llr.abs())
    assert torch.equal((flipped != llr), (y != x))                                # signs flip exactly where bits flip
    kb = Scrambler(seed=7, keep_batch_constant=True)
    z = kb(torch.zeros(5, 64, device=cuda_device))
    assert bool((z == z[0]).all())
    seq = np.tile([0, 1], 50)
    e = Scrambler(sequence=seq)(torch.zeros(4, 100, device=cuda_device))
    assert np.array_equal(e.cpu().numpy(), np.tile(seq, (4, 1)))
    s_rand = Scrambler(keep_state=False)
    assert not torch.equal(s_rand(x), s_rand(x))
    assert torch.equal(s_rand(s_rand(x, seed=5), seed=5), x)
    t5 = TB5GScrambler(n_rnti=3, n_id=7)
    assert torch.equal(Descrambler(t5)(t5(x)), x)
    with pytest.raises(TypeError):
        Descrambler("scrambler")
    sym, ind, bits = QAMSource(4, return_indices=True, return_bits=True, seed=1)([6, 50])
    assert list(sym.shape) == [6, 50] and list(ind.shape) == [6, 50] and list(bits.shape) == [6, 50, 4]
    pts = Constellation("qam", 4)().cpu().numpy()
    assert np.allclose(sym.cpu().numpy(), pts[ind.cpu().numpy()])
    assert torch.equal(SymbolInds2Bits(4)(ind), bits)
    p = PAMSource(3)([1000])
    assert p.is_complex() and float(p.imag.abs().max()) == 0.0 and abs(float((p.abs() ** 2).mean()) - 1.0) < 0.1
    a, tau = RayleighBlockFading(1, 4, 2, 1)(256, 14)
    assert list(a.shape) == [256, 1, 4, 2, 1, 1, 14] and list(tau.shape) == [256, 1, 2, 1]
    assert bool((a == a[..., :1]).all()) and abs(float((a.abs() ** 2).mean()) - 1.0) < 0.1


def test_decode_mcs_index_vs_reference_known_answers():
    """decode_mcs_index vs the 11 known-answer lists of the reference's tests (test_nr_utils.py:17-268; TS 38.214 Tables
    5.1.3.1-1..4, 6.1.4.1-1/2 incl. the pi/2-BPSK rows); the first index past each list is invalid."""
    import json
    from sionna_b200.phy.nr import decode_mcs_index
    with open(os.path.join(os.path.dirname(__file__), "golden", "mcs_golden.json")) as f:
        cases = json.load(f)
    assert len(cases) == 11
    for c in cases:
        kw = dict(c["kwargs"])
        variants = (True, False) if kw.get("pi2bpsk") == "both" else (kw["pi2bpsk"],)
        for bpsk in variants:
            kw["pi2bpsk"] = bpsk
            for idx, (q, r) in enumerate(zip(c["qs"], c["rs"])):
                m, rate = decode_mcs_index(mcs_index=idx, **kw)
                assert m == q and float(rate) == np.float32(r / 1024), (kw, idx)
            if len(c["qs"]) < 29:
                with pytest.raises(AssertionError):
                    decode_mcs_index(mcs_index=len(c["qs"]), **kw)
    # vectorised call, mixed tables / channels
    m, r = decode_mcs_index([0, 27, 5], [1, 2, 3], [True, False, True])
    assert list(m) == [2, 8, 2] and np.allclose(r, [120 / 1024, 948 / 1024, 99 / 1024])


def test_calculate_tb_size_consistency_sweep():
    """The structural invariants the reference checks over a parameter sweep (test_nr_utils.py:300-374)."""
    from sionna_b200.phy.nr import calculate_tb_size, decode_mcs_index
    for mcs_index in (0, 4, 16, 20, 27):
        q, r = decode_mcs_index(mcs_index, 2)
        for num_layers in (1, 2, 3, 4):
            for num_prbs in (1, 20, 200, 275):
                for num_ofdm_symbols in (8, 10, 14):
                    for num_dmrs_per_prb in (0, 10, 20):
                        tb, cb, ncb, tb_crc, cb_crc, cw = calculate_tb_size(
                            target_coderate=r, modulation_order=q, num_layers=num_layers, num_prbs=num_prbs,
                            num_ofdm_symbols=num_ofdm_symbols, num_dmrs_per_prb=num_dmrs_per_prb)
                        cw = np.asarray(cw)
                        assert tb == ncb * (cb - cb_crc) - tb_crc and ncb == len(cw)
                        assert cb_crc == (0 if ncb == 1 else 24)
                        assert set(cw.tolist()) <= {int(cw.min()), int(cw.max())}
                        assert tb_crc == (24 if tb > 3824 else 16)
                        n_res = q * num_layers * (12 * num_ofdm_symbols - num_dmrs_per_prb)
                        if n_res <= 156:
                            eff = tb / cw.sum()
                            if tb > 4000:
                                assert abs(eff - r) < 2e-2
                            elif tb > 200:
                                assert abs(eff - r) < 1e-1
