"""On-device channel generation: TDL sum-of-sinusoids taps, CIR -> OFDM channel, uniform RNG."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_uniform_generator_moments_and_range(cuda_device):
    from sionna_b200.phy.channel.tdl import _uniform
    from sionna_b200.phy import config
    config.seed = 5
    u = _uniform([1 << 20], -2.0, 3.0).cpu().numpy()
    assert u.min() >= -2.0 and u.max() < 3.0
    assert abs(u.mean() - 0.5) < 5e-3 and abs(u.var() - 25 / 12) < 2e-2
    v = _uniform([1 << 20], -2.0, 3.0).cpu().numpy()
    assert not np.array_equal(u, v)                                   # the stream advances
    assert abs(np.corrcoef(u, v)[0, 1]) < 5e-3


@pytest.mark.parametrize("model,speed", [("A", 0.0), ("C", 30.0), ("D", 10.0), ("B100", 3.0)])
def test_tdl_taps_equal_oracle_on_identical_draws(cuda_device, model, speed):
    """sb_tdl_sos == float64 NumPy restatement of tdl.py:374-456 fed with the same Doppler / angle / phase draws."""
    from sionna_b200.phy.channel import TDL
    from sionna_b200.phy import config
    from oracle import ofdm as OO
    config.seed = 17
    tdl = TDL(model, 300e-9, 3.5e9, min_speed=speed, max_speed=speed * 1.5, num_rx_ant=2, num_tx_ant=3)
    draws = tdl.draws(8)
    t_steps, fs = 14, 14e3
    a = tdl.synthesize(draws, t_steps, fs).cpu().numpy()
    d = [None if x is None else x.cpu().numpy() for x in draws]
    ref = OO.tdl_sos(d[0], d[1], d[2], d[3], tdl._powers, tdl._los_power, tdl._los_aoa, t_steps, fs)
    assert a.shape == ref.shape == (8, 6, tdl.num_clusters, t_steps)
    assert np.allclose(a, ref, atol=2e-5)
    if speed == 0.0:
        assert np.allclose(a, a[..., :1], atol=1e-6)                   # zero speed: constant over the slot


def test_tdl_statistics_and_cir_to_ofdm(cuda_device):
    """Tap powers follow the power delay profile, the specular component carries the K-factor, the frequency response
    equals the NumPy restatement and has unit average energy after normalisation."""
    from sionna_b200.phy.channel import TDL, cir_to_ofdm_channel, subcarrier_frequencies
    from sionna_b200.phy import config
    from oracle import ofdm as OO
    config.seed = 3
    batch = 4096
    for model in ("B", "E"):
        tdl = TDL(model, 100e-9, 3.5e9, min_speed=20.0, num_rx_ant=2, num_tx_ant=2)
        a, tau = tdl(batch, 14, 15e3)
        assert list(a.shape) == [batch, 1, 2, 1, 2, tdl.num_clusters, 14] and list(tau.shape) == [batch, 1, 1, tdl.num_clusters]
        pw = (a.abs() ** 2).mean(dim=(0, 1, 2, 3, 4, 6)).cpu().numpy()
        assert np.allclose(pw, tdl.mean_powers.numpy(), rtol=0.08, atol=2e-4)
        assert abs(pw.sum() - 1.0) < 0.03
        freqs = subcarrier_frequencies(72, 15e3)
        h = cir_to_ofdm_channel(freqs, a[:16], tau[:16])
        ref = OO.cir_to_ofdm(freqs.numpy(), a[:16].cpu().numpy().astype(np.complex128), tdl.delays.numpy())
        assert np.allclose(h.cpu().numpy(), ref, atol=2e-5)
        hn = cir_to_ofdm_channel(freqs, a, tau, normalize=True)
        assert abs(float((hn.abs() ** 2).mean()) - 1.0) < 1e-3
    # Doppler: the temporal autocorrelation of a tap follows J0(w_d * lag) (Clarke/Jakes), w_d = 2 pi v f_c / c
    tdl = TDL("A", 100e-9, 3.5e9, min_speed=30.0, num_sinusoids=20)
    fs = 14e3
    a, _ = tdl(8192, 64, fs)
    x = a[:, 0, 0, 0, 0, :, :]                                          # [B, P, T]
    lags = np.array([1, 4, 8, 16])
    r = np.array([float((x[..., lag:] * x[..., :-lag].conj()).real.mean() / (x.abs() ** 2).mean()) for lag in lags])
    from scipy.special import j0
    wd = 2 * np.pi * 30.0 / 299792458.0 * 3.5e9
    assert np.allclose(r, j0(wd * lags / fs), atol=0.03)


def test_pusch_link_with_mobility(cuda_device):
    """PUSCH over TDL-C at 30 m/s with two DMRS positions: linear time interpolation tracks the channel."""
    from sionna_b200.phy.nr import PUSCHConfig, PUSCHTransmitter, PUSCHReceiver
    from sionna_b200.phy.channel import ApplyOFDMChannel, TDL, subcarrier_frequencies, cir_to_ofdm_channel
    from sionna_b200.phy import config
    config.seed = 23
    pc = PUSCHConfig()
    pc.carrier.n_size_grid = 12
    pc.carrier.subcarrier_spacing = 30
    pc.dmrs.additional_position = 2
    pc.tb.mcs_index = 8
    tx = PUSCHTransmitter(pc)
    rx = PUSCHReceiver(tx, return_tb_crc_status=True)
    rg = tx.resource_grid
    x, b = tx(256)
    sym_rate = 1.0 / (1.0 / 30e3 + rg.cyclic_prefix_length / (rg.fft_size * 30e3))
    a, tau = TDL("C", 100e-9, 3.5e9, min_speed=30.0, num_rx_ant=4)(256, 14, sym_rate)
    h = cir_to_ofdm_channel(subcarrier_frequencies(rg.fft_size, 30e3), a, tau, normalize=True)
    y = ApplyOFDMChannel()(x, h, 0.01)
    b_hat, crc = rx(y, 0.01)
    assert float((b_hat != b).float().mean()) < 1e-3 and float(crc.float().mean()) > 0.99


def test_time_channel_vs_oracle_and_ofdm_equivalence(cuda_device):
    """cir_to_time_channel / ApplyTimeChannel equal their NumPy restatements, and for a static channel the time-domain
    path (OFDM modulator -> time-variant filter -> OFDM demodulator) reproduces the frequency-domain channel application
    up to the sinc truncation (the reference's time/frequency equivalence check)."""
    from sionna_b200.phy.channel import (TDL, cir_to_time_channel, cir_to_ofdm_channel, ApplyTimeChannel, ApplyOFDMChannel,
                                         subcarrier_frequencies, time_lag_discrete_time_channel, TimeChannel, OFDMChannel)
    from sionna_b200.phy.ofdm import OFDMModulator, OFDMDemodulator, ResourceGrid
    from sionna_b200.phy.utils import complex_normal
    from sionna_b200.phy import config
    from oracle import ofdm as OO
    config.seed = 9
    fft, cp, nsym, scs = 64, 16, 4, 30e3
    bw = fft * scs
    l_min, l_max = time_lag_discrete_time_channel(bw, 300e-9)
    assert l_min == -6 and l_max == int(np.ceil(300e-9 * bw)) + 6
    l_tot = l_max - l_min + 1
    n_time = nsym * (fft + cp)
    tdl = TDL("A", 50e-9, 3.5e9, num_rx_ant=2, num_tx_ant=2)
    a, tau = tdl(6, n_time + l_tot - 1, bw)                                  # zero speed: constant taps
    hm = cir_to_time_channel(bw, a, tau, l_min, l_max)
    ref = OO.cir_to_time(bw, a.cpu().numpy().astype(np.complex128), tdl.delays.numpy(), l_min, l_max)
    assert list(hm.shape) == [6, 1, 2, 1, 2, n_time + l_tot - 1, l_tot]
    assert np.allclose(hm.cpu().numpy(), ref, atol=2e-5)
    x = complex_normal([6, 1, 2, nsym, fft])
    xt = OFDMModulator(cp)(x)                                                # [6, 1, 2, n_time]
    yt = ApplyTimeChannel(n_time, l_tot)(xt, hm)
    yt_ref = OO.apply_time_channel(xt.cpu().numpy()[:, 0], hm.cpu().numpy()[:, 0, :, 0].astype(np.complex128))
    assert np.allclose(yt.cpu().numpy()[:, 0], yt_ref, atol=1e-4)
    y = OFDMDemodulator(fft, l_min, cp)(yt)                                  # [6, 1, 2, nsym, fft]
    hf = cir_to_ofdm_channel(subcarrier_frequencies(fft, scs), a[..., :nsym], tau)
    yf = ApplyOFDMChannel()(x, hf)
    err = float(((y - yf).abs() ** 2).mean() / (yf.abs() ** 2).mean())
    assert err < 1e-2                          # sinc tails cut at l_min = -6 carry (1 / 6 pi)^2 = -25 dB of the energy
    from sionna_b200.phy.channel import time_to_ofdm_channel
    rg0 = ResourceGrid(nsym, fft, scs, cyclic_prefix_length=cp)
    hf_t = time_to_ofdm_channel(hm, rg0, l_min)                               # [6, 1, 2, 1, 2, nsym, fft]
    assert list(hf_t.shape) == list(hf.shape)
    assert float(((hf_t - hf).abs() ** 2).mean() / (hf.abs() ** 2).mean()) < 1e-2
    # convenience blocks: shapes, noise, returned channel
    rg = ResourceGrid(nsym, fft, scs, num_tx=1, num_streams_per_tx=2, cyclic_prefix_length=cp)
    yo, ho = OFDMChannel(tdl, rg, normalize_channel=True, return_channel=True)(x, 0.1)
    assert list(yo.shape) == [6, 1, 2, nsym, fft] and list(ho.shape) == [6, 1, 2, 1, 2, nsym, fft]
    yc, hc = TimeChannel(tdl, bw, n_time, maximum_delay_spread=300e-9, return_channel=True)(xt, 0.1)
    assert list(yc.shape) == [6, 1, 2, n_time + l_tot - 1] and list(hc.shape) == [6, 1, 2, 1, 2, n_time + l_tot - 1, l_tot]


def test_pusch_time_domain_link_over_tdl(cuda_device):
    """PUSCHTransmitter(output_domain="time") -> TimeChannel(TDL-A, 3 m/s) -> PUSCHReceiver(input_domain="time")."""
    from sionna_b200.phy.nr import PUSCHConfig, PUSCHTransmitter, PUSCHReceiver
    from sionna_b200.phy.channel import TDL, TimeChannel, time_lag_discrete_time_channel
    from sionna_b200.phy import config
    config.seed = 31
    pc = PUSCHConfig()
    pc.carrier.n_size_grid = 8
    pc.carrier.subcarrier_spacing = 30
    pc.dmrs.additional_position = 1
    pc.tb.mcs_index = 6
    tx = PUSCHTransmitter(pc, output_domain="time")
    rg = tx.resource_grid
    l_min, l_max = time_lag_discrete_time_channel(rg.bandwidth, 300e-9)
    rx = PUSCHReceiver(tx, input_domain="time", l_min=l_min, return_tb_crc_status=True)
    x, b = tx(64)
    chan = TimeChannel(TDL("A", 30e-9, 3.5e9, min_speed=3.0, num_rx_ant=4), rg.bandwidth, rg.num_time_samples,
                       maximum_delay_spread=300e-9, normalize_channel=True)
    y = chan(x, 0.005)
    b_hat, crc = rx(y, 0.005)
    assert float((b_hat != b).float().mean()) < 1e-3 and float(crc.float().mean()) > 0.98
    # perfect CSI from the time-domain taps (time_to_ofdm_channel inside the receiver)
    chan_h = TimeChannel(TDL("A", 30e-9, 3.5e9, num_rx_ant=4), rg.bandwidth, rg.num_time_samples,
                         maximum_delay_spread=300e-9, normalize_channel=True, return_channel=True)
    y2, h_time = chan_h(x, 0.005)
    rx_p = PUSCHReceiver(tx, channel_estimator="perfect", input_domain="time", l_min=l_min)
    assert float((rx_p(y2, 0.005, h_time) != b).float().mean()) < 1e-3


def test_cir_normalisation_per_link_delays_and_spatial_correlation(cuda_device):
    """channel.cu: (1) normalize=True equals the reference recipe h / sqrt(mean |h|^2 over (rx_ant, tx_ant, t, f)) per
    link computed in float64 from the un-normalised h (channel/utils.py:246-251, :341-348), although the kernel never
    reads h for it (Gram-matrix form); (2) delays that differ from link to link (tau [batch, rx, tx, paths], not a
    broadcast view) take the per-link tables; (3) spatial correlation equals L v / R V T^H in NumPy (tdl.py:466-490)."""
    from sionna_b200.phy.channel import TDL, cir_to_ofdm_channel, cir_to_time_channel, subcarrier_frequencies
    from sionna_b200.phy.utils import complex_normal
    from sionna_b200.phy import config
    from oracle import ofdm as OO
    config.seed = 41
    b, rx, ra, tx, ta, p, t, f = 5, 2, 3, 2, 2, 7, 6, 48
    a = complex_normal([b, rx, ra, tx, ta, p, t])
    a[1] = 0                                                            # an all-zero link must stay zero (no NaN)
    freqs = subcarrier_frequencies(f, 30e3)
    rng = np.random.default_rng(0)
    tau_link = torch.from_numpy((rng.uniform(0, 2e-6, (b, rx, tx, p))).astype(np.float32))
    an = a.cpu().numpy().astype(np.complex128)
    # reference formula with per-link delays
    e = np.exp(-2j * np.pi * tau_link.numpy().astype(np.float64)[..., None] * freqs.numpy().astype(np.float64))  # [b,rx,tx,p,f]
    ref = np.einsum("brmtnpl,brtpf->brmtnlf", an, e)
    h = cir_to_ofdm_channel(freqs, a, tau_link).cpu().numpy()
    assert np.allclose(h, ref, atol=3e-5)
    c = np.sqrt(np.mean(np.abs(ref) ** 2, axis=(2, 4, 5, 6), keepdims=True))
    ref_n = np.where(c > 0, ref / np.where(c > 0, c, 1), 0)
    hn = cir_to_ofdm_channel(freqs, a, tau_link, normalize=True).cpu().numpy()
    assert np.all(np.isfinite(hn)) and np.allclose(hn, ref_n, atol=5e-5)
    # shared delays given as a broadcast view: cached single table, same numbers as the explicit per-link copy
    tau_shared = tau_link[:1, :1, :1].expand(b, rx, tx, p)
    h1 = cir_to_ofdm_channel(freqs, a, tau_shared, normalize=True)
    h2 = cir_to_ofdm_channel(freqs, a, tau_shared.contiguous(), normalize=True)
    assert torch.allclose(h1, h2, atol=1e-6)
    # time-domain taps: sinc table, normalisation = unit mean total tap energy per link
    bw, l_min, l_max = 48 * 30e3, -6, 9
    hm = cir_to_time_channel(bw, a, tau_link, l_min, l_max).cpu().numpy()
    l = np.arange(l_min, l_max + 1)
    g = np.sinc(l - tau_link.numpy().astype(np.float64)[..., None] * bw)                                       # [b,rx,tx,p,L]
    ref_t = np.einsum("brmtnpl,brtpk->brmtnlk", an, g)
    assert np.allclose(hm, ref_t, atol=3e-5)
    ct = np.sqrt(np.mean(np.sum(np.abs(ref_t) ** 2, axis=6, keepdims=True), axis=(2, 4, 5), keepdims=True))
    ref_tn = np.where(ct > 0, ref_t / np.where(ct > 0, ct, 1), 0)
    hmn = cir_to_time_channel(bw, a, tau_link, l_min, l_max, normalize=True).cpu().numpy()
    assert np.allclose(hmn, ref_tn, atol=5e-5)
    # spatial correlation: identical draws with and without the correlation matrices
    n_rx, n_tx = 4, 2
    r_rx = 0.7 ** np.abs(np.subtract.outer(np.arange(n_rx), np.arange(n_rx))) * np.exp(0.3j * np.subtract.outer(np.arange(n_rx), np.arange(n_rx)))
    r_tx = np.array([[1.0, 0.4 - 0.2j], [0.4 + 0.2j, 1.0]])
    config.seed = 77
    a0, _ = TDL("A", 100e-9, 3.5e9, min_speed=5.0, num_rx_ant=n_rx, num_tx_ant=n_tx)(16, 4, 1e3)
    config.seed = 77
    a1, _ = TDL("A", 100e-9, 3.5e9, min_speed=5.0, num_rx_ant=n_rx, num_tx_ant=n_tx, rx_corr_mat=r_rx, tx_corr_mat=r_tx)(16, 4, 1e3)
    lr, lt = np.linalg.cholesky(r_rx), np.linalg.cholesky(r_tx)
    v = a0.cpu().numpy().astype(np.complex128)[:, 0, :, 0]                 # [B, rx_ant, tx_ant, P, T]
    want = np.einsum("ij,bjkpt,lk->bilpt", lr, v, lt.conj())
    assert np.allclose(a1.cpu().numpy()[:, 0, :, 0], want, atol=2e-5)
    config.seed = 77
    full = np.kron(r_rx, r_tx)
    a2, _ = TDL("A", 100e-9, 3.5e9, min_speed=5.0, num_rx_ant=n_rx, num_tx_ant=n_tx, spatial_corr_mat=full)(16, 4, 1e3)
    lf = np.linalg.cholesky(full)
    want2 = np.einsum("ij,bjpt->bipt", lf, v.reshape(16, n_rx * n_tx, v.shape[3], v.shape[4])).reshape(v.shape)
    assert np.allclose(a2.cpu().numpy()[:, 0, :, 0], want2, atol=2e-5)
