"""End-to-end links of BASELINE.json configs[2] (SISO OFDM 14x76, 64-QAM, TDL-A, LS + LMMSE) and configs[3] (4 streams x 16
rx antennas MIMO-OFDM, LMMSE LinearDetector + LDPC5G), built like the reference's integration model
(/root/reference/test/integration/test_mimo_ofdm_cdl.py:191-247) from this package's blocks and driven by sim_ber. As in
the reference's integration tests the bar is behavioural (no NaN, BER falls with SNR, error free at high SNR); the
numerical parity of every block is covered in test_ofdm_mimo_gpu.py / test_phy_gpu.py / test_ldpc_decoder_gpu.py. For
configs[2] the demapped LLRs of the whole receive chain are additionally compared with the oracle chain."""
import numpy as np
import pytest
import torch

from oracle import ofdm as F
from oracle import mapping as M

pytestmark = pytest.mark.gpu


class Link:
    def __init__(self, num_streams, num_rx_ant, num_bits_per_symbol, coderate, perfect_csi=False, detector=False,
                 tdl=True, demapping="app"):
        from sionna_b200.phy.ofdm import (ResourceGrid, ResourceGridMapper, LSChannelEstimator, LMMSEEqualizer,
                                          LinearDetector)
        from sionna_b200.phy.mimo import StreamManagement
        from sionna_b200.phy.mapping import Mapper, Demapper, BinarySource
        from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
        from sionna_b200.phy.channel import ApplyOFDMChannel, TDL, subcarrier_frequencies
        self.rg = ResourceGrid(14, 76, 15e3, num_tx=1, num_streams_per_tx=num_streams, cyclic_prefix_length=6,
                               num_guard_carriers=(5, 6), dc_null=True, pilot_pattern="kronecker",
                               pilot_ofdm_symbol_indices=[2, 11])
        self.sm = StreamManagement(np.array([[1]]), num_streams)
        self.m, self.r = num_bits_per_symbol, coderate
        self.n = int(self.rg.num_data_symbols * num_bits_per_symbol)
        self.k = int(self.n * coderate)
        self.src, self.enc = BinarySource(), LDPC5GEncoder(self.k, self.n)
        self.mapper, self.rgm = Mapper("qam", num_bits_per_symbol), ResourceGridMapper(self.rg)
        self.chan = ApplyOFDMChannel()
        self.tdl = TDL("A", 300e-9, 3.5e9, num_rx_ant=num_rx_ant, num_tx_ant=num_streams) if tdl else None
        self.freqs = subcarrier_frequencies(76, 15e3)
        self.est = LSChannelEstimator(self.rg, "nn")
        self.detector = LinearDetector("lmmse", "bit", demapping, self.rg, self.sm, "qam", num_bits_per_symbol) if detector else None
        self.eq, self.demapper = LMMSEEqualizer(self.rg, self.sm), Demapper(demapping, "qam", num_bits_per_symbol)
        self.dec = LDPC5GDecoder(self.enc, hard_out=True, num_iter=20)
        self.num_rx_ant, self.num_streams = num_rx_ant, num_streams

    def channel(self, batch_size):
        from sionna_b200.phy.channel import cir_to_ofdm_channel
        from sionna_b200.phy.utils import complex_normal
        if self.tdl is None:                                   # i.i.d. Rayleigh, flat over the slot
            h = complex_normal([batch_size, 1, self.num_rx_ant, 1, self.num_streams, 1, 1])
            return h.expand(batch_size, 1, self.num_rx_ant, 1, self.num_streams, 14, 76).contiguous()
        a, tau = self.tdl(batch_size, 14, 1.0)
        return cir_to_ofdm_channel(self.freqs, a, tau, normalize=True)

    def front_end(self, batch_size, ebno_db):
        from sionna_b200.phy.utils import ebnodb2no
        no = ebnodb2no(ebno_db, self.m, self.r, self.rg)
        b = self.src([batch_size, 1, self.num_streams, self.k])
        x = self.mapper(self.enc(b))
        h = self.channel(batch_size)
        y = self.chan(self.rgm(x), h, no)
        return b, x, h, y, no

    def __call__(self, batch_size, ebno_db):
        b, x, h, y, no = self.front_end(batch_size, ebno_db)
        h_hat, err_var = self.est(y, no)
        if self.detector is not None:
            llr = self.detector(y, h_hat, err_var, no)
        else:
            x_hat, no_eff = self.eq(y, h_hat, err_var, no)
            llr = self.demapper(x_hat, no_eff)
        return b, self.dec(llr)


def test_config2_siso_ofdm_64qam_tdl_link(cuda_device):
    from sionna_b200.phy.utils import sim_ber
    from sionna_b200.phy import config
    config.seed = 3
    link = Link(num_streams=1, num_rx_ant=1, num_bits_per_symbol=6, coderate=0.5)
    assert (link.rg.num_data_symbols, link.n, link.k) == (768, 4608, 2304)
    ber, bler = sim_ber(link, [0.0, 10.0, 20.0, 35.0], batch_size=256, max_mc_iter=2, verbose=False, early_stop=False)
    ber = ber.numpy()
    assert np.all(np.isfinite(ber)) and ber[0] > 0.1 and ber[0] > ber[1] > ber[2] >= ber[3] and ber[3] < 2e-2
    # receive chain LLRs vs the oracle chain on identical received samples (64-QAM "app")
    b, x, h, y, no = link.front_end(8, 15.0)
    h_hat, ev = link.est(y, no)
    x_hat, no_eff = link.eq(y, h_hat, ev, no)
    llr = link.demapper(x_hat, no_eff).cpu().numpy()
    mask, pil = link.rg.pilot_pattern.mask.astype(bool), link.rg.pilot_pattern.pilots
    eff = F.eff_sc_ind(76, (5, 6), True)
    y_eff = y.cpu().numpy()[..., eff].astype(complex)
    hr, er = F.ls_estimate(y_eff, mask, pil, float(no))
    hr, er = F.nn_interp(hr, mask, pil), F.nn_interp(er, mask, pil)
    xr, nr = F.ofdm_lmmse_equalize(y_eff, hr, er, float(no), mask, F.stream_management([[1]], 1))
    lr = M.demapper(xr.astype(np.complex64), nr.astype(np.float32), M.qam(6), "app")
    # LLRs scale like 1/no_eff; compare where the channel is not in a deep fade and relative to the LLR magnitude
    ok = nr.repeat(6, axis=-1) < 1.0
    np.testing.assert_allclose(llr[ok], lr[ok], rtol=2e-3, atol=2e-3 * np.abs(lr[ok]).max())
    assert np.mean((llr > 0) == (lr > 0)) > 0.9995


def test_config3_mimo_ofdm_lmmse_ldpc_link(cuda_device):
    from sionna_b200.phy.utils import sim_ber
    from sionna_b200.phy import config
    config.seed = 4
    link = Link(num_streams=4, num_rx_ant=16, num_bits_per_symbol=2, coderate=0.5, detector=True, tdl=True)
    ber, bler = sim_ber(link, [-12.0, -8.0, -4.0, 2.0], batch_size=128, max_mc_iter=2, verbose=False, early_stop=False)
    ber, bler = ber.numpy(), bler.numpy()
    assert np.all(np.isfinite(ber)) and ber[0] > 0.05 and ber[0] >= ber[1] >= ber[2] >= ber[3]
    assert ber[3] == 0 and bler[3] == 0                      # error free at high SNR (test_mimo_ofdm_detectors.py:113-127)
    # i.i.d. Rayleigh variant, equaliser + separate demapper route, 16-QAM maxlog
    link2 = Link(num_streams=4, num_rx_ant=16, num_bits_per_symbol=4, coderate=0.5, detector=False, tdl=False,
                 demapping="maxlog")
    ber2, _ = sim_ber(link2, [-6.0, 6.0], batch_size=64, max_mc_iter=1, verbose=False, early_stop=False)
    assert ber2[0] > ber2[1] and ber2[1] == 0


@pytest.mark.parametrize("cfg", ["configs2_siso_2048", "configs3_mimo_1024"])
def test_full_batch_receive_chain_vs_oracle(cuda_device, cfg):
    """The oracle comparison of the receive chain at the configurations' FULL batch (configs[2]: 2048 frames, 64-QAM
    app; configs[3]: 1024 frames, 4 streams x 16 antennas, 16-QAM app through the fused LinearDetector), LLRs of every
    data resource element against the complex128 NumPy chain. LLRs scale with 1/no_eff, whose fp32 evaluation carries the
    cancellation measured in test_lmmse_error_sits_inside_the_reference_fp32_envelope; the bar is therefore relative to
    the largest LLR of the comparison set, and hard decisions must agree except on LLRs that are numerically zero."""
    from sionna_b200.phy import config
    config.seed = 11
    if cfg == "configs2_siso_2048":
        link, batch, ebno, m, streams = Link(1, 1, 6, 0.5), 2048, 18.0, 6, 1
    else:
        link, batch, ebno, m, streams = Link(4, 16, 4, 0.5, detector=True, tdl=True), 1024, 4.0, 4, 4
    b, x, h, y, no = link.front_end(batch, ebno)
    h_hat, ev = link.est(y, no)
    if link.detector is not None:
        llr = link.detector(y, h_hat, ev, no).cpu().numpy()
    else:
        x_hat, no_eff = link.eq(y, h_hat, ev, no)
        llr = link.demapper(x_hat, no_eff).cpu().numpy()
    mask, pil = link.rg.pilot_pattern.mask.astype(bool), link.rg.pilot_pattern.pilots
    eff = F.eff_sc_ind(76, (5, 6), True)
    y_eff = y.cpu().numpy()[..., eff].astype(complex)
    hr, er = F.ls_estimate(y_eff, mask, pil, float(no))
    hr, er = F.nn_interp(hr, mask, pil), F.nn_interp(er, mask, pil)
    xr, nr = F.ofdm_lmmse_equalize(y_eff, hr, er, float(no), mask, F.stream_management([[1]], streams))
    lr = M.demapper(xr.astype(np.complex64), nr.astype(np.float32), M.qam(m), "app")
    assert llr.shape == lr.shape == (batch, 1, streams, link.rg.num_data_symbols * m)
    ok = np.repeat(nr, m, axis=-1) < 1.0                     # not in a deep fade (LLR ~ 0 there, sign is noise)
    scale = np.abs(lr[ok]).max()
    err = np.abs(llr[ok] - lr[ok])
    assert err.max() <= 2e-3 * scale and np.sqrt(np.mean(err ** 2)) <= 1e-4 * scale
    sure = ok & (np.abs(lr) > 1e-3 * scale)
    assert np.array_equal(llr[sure] > 0, lr[sure] > 0)
