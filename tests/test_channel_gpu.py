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
