"""Receive-chain workloads of BASELINE.json configs[0], [2], [3], [4] for bench.py (`--workload qpsk_awgn | ofdm_siso |
mimo_ofdm | pusch`). Each workload builds its link from the package's public blocks, synthesises TWO alternating input
sets on the device with its own transmit chain + channel (untimed, like the headline LDPC bench), and exposes

  run(i)        one pass of the RECEIVE hot path over batch i & 1, inputs resident in HBM  -> decoded bits / LLRs
  stages(i)     the same pass split into named block calls with their algorithmic bytes (SURVEY.md section 8d figures)
  host_in / out pinned host buffers for the end-to-end leg (H2D of the received samples, D2H of the result)
  cpu_chain(n)  the oracle's restatement of the same pass on the first n frames (NumPy / C, `cpu_baseline`)

A "unit" is what the workload's metric counts (QPSK symbols, data resource elements, coded bits).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class Workload:
    name = metric = unit = desc = ""
    scaling = "weak"                # per-GPU batch fixed (replicas); pusch overrides with "strong" (batch 8192 sharded)
    dtype = "c64/f32"

    def __init__(self, dev, rank, world, batch=None):
        self.dev, self.rank, self.world = dev, rank, world
        self.batch = int(batch or self.default_batch)
        self.inputs, self.truth = [], []

    # --- to be provided ---------------------------------------------------------------------------------------------
    def build(self):
        raise NotImplementedError

    def run(self, i, x=None):
        raise NotImplementedError

    def stages(self, i):
        return []

    def cpu_chain(self, n):
        return None

    # --- helpers ----------------------------------------------------------------------------------------------------
    def host_buffers(self):
        self.host_in = [t.cpu().pin_memory() for t in self.inputs]
        out = self.run(0)
        self.host_out = [torch.empty(out.shape, dtype=out.dtype).pin_memory() for _ in range(2)]
        return out

    @property
    def h2d_bytes(self):
        return self.inputs[0].numel() * self.inputs[0].element_size()

    @property
    def d2h_bytes(self):
        return self.host_out[0].numel() * self.host_out[0].element_size()


# ---------------------------------------------------------------------------------------------------------------------
class QpskAwgn(Workload):
    """configs[0]: QPSK Mapper -> AWGN -> Demapper("app") LLRs, batch 1024 x 8448 coded bits (SURVEY 8d cfg 1)."""
    name, default_batch = "qpsk_awgn", 1024
    metric, unit = "QPSK symbols/s, Demapper('app') LLR compute", "symbols/s"
    n_bits = 8448

    def build(self):
        from sionna_b200.phy.mapping import Mapper, Demapper, BinarySource
        from sionna_b200.phy.channel import AWGN
        from sionna_b200.phy.utils import ebnodb2no
        self.no = float(ebnodb2no(4.0, 2, 1.0))
        self.demapper = Demapper("app", "qam", 2)
        src, mapper, awgn = BinarySource(), Mapper("qam", 2), AWGN()
        # configs[0]'s batch (1024 x 4224 symbols = 35 MB) fits in L2: 8 distinct input sets are rotated (8 x 35 MB of
        # input + 8 x 35 MB of output > 126 MB L2) so that no step finds its data in L2
        self.sets = 8
        for _ in range(self.sets):
            b = src([self.batch, self.n_bits])
            self.truth.append(b)
            self.inputs.append(awgn(mapper(b), self.no))
        self.units_per_step = self.batch * self.n_bits // 2
        self.desc = (f"configs[0]: QPSK Mapper -> AWGN(Eb/N0 4 dB) -> Demapper('app'), batch {self.batch} x {self.n_bits} "
                     f"bits; {self.sets} rotating input sets (> L2)")

    def run(self, i, x=None):
        return self.demapper(self.inputs[i % self.sets] if x is None else x, self.no)

    def stages(self, i):
        s = self.units_per_step
        return [("Demapper app QPSK", lambda: self.run(i), s * 16, "sb_demap_qam: 8 B (y) + 2 x 4 B LLR per symbol")]

    def host_buffers(self):
        out = super().host_buffers()
        self.host_in = self.host_in[:2]
        return out

    def cpu_chain(self, n):
        from oracle import mapping as M
        y = self.inputs[0][:n].cpu().numpy()
        t0 = time.perf_counter()
        llr = M.demapper(y, np.float32(self.no), M.qam(2), "app")
        dt = time.perf_counter() - t0
        got = self.run(0)[:n].cpu().numpy()
        err = float(np.max(np.abs(got - llr) / np.maximum(np.abs(llr), 1e-3)))
        return {"units": y.size, "seconds": dt, "what": f"oracle demapper (libm, OpenMP) on {n} frames",
                "max_rel_diff_vs_gpu": err}


# ---------------------------------------------------------------------------------------------------------------------
def _rg(num_streams):
    from sionna_b200.phy.ofdm import ResourceGrid
    return ResourceGrid(14, 76, 15e3, num_tx=1, num_streams_per_tx=num_streams, cyclic_prefix_length=6,
                        num_guard_carriers=(5, 6), dc_null=True, pilot_pattern="kronecker",
                        pilot_ofdm_symbol_indices=[2, 11])


class OfdmSiso(Workload):
    """configs[2]: OFDM 14 x 76 (CP 6), 64-QAM, TDL-A 300 ns time-domain channel, OFDMDemodulator -> LS(nn) -> LMMSE
    equaliser -> Demapper('app'); batch 2048 frames."""
    name, default_batch = "ofdm_siso", 2048
    metric, unit = "data resource elements/s, OFDM demod + LS + LMMSE + 64-QAM demap", "RE/s"
    ebno_db = 20.0

    def build(self):
        from sionna_b200.phy.ofdm import (ResourceGridMapper, OFDMModulator, OFDMDemodulator, LSChannelEstimator,
                                          LMMSEEqualizer)
        from sionna_b200.phy.mimo import StreamManagement
        from sionna_b200.phy.mapping import Mapper, Demapper, BinarySource
        from sionna_b200.phy.channel import TDL, TimeChannel, time_lag_discrete_time_channel
        from sionna_b200.phy.utils import ebnodb2no
        rg = self.rg = _rg(1)
        sm = StreamManagement(np.array([[1]]), 1)
        m = 6
        self.m = m
        bw = rg.fft_size * rg.subcarrier_spacing
        l_min, l_max = time_lag_discrete_time_channel(bw)
        self.l_min = l_min
        nts = rg.num_ofdm_symbols * (rg.fft_size + rg.cyclic_prefix_length)
        tdl = TDL("A", 300e-9, 3.5e9, num_rx_ant=1, num_tx_ant=1)
        chan = TimeChannel(tdl, bw, nts, l_min=l_min, l_max=l_max, normalize_channel=True)
        self.no = float(ebnodb2no(self.ebno_db, m, 1.0, rg))
        src, mapper, rgm, mod = BinarySource(), Mapper("qam", m), ResourceGridMapper(rg), OFDMModulator(rg.cyclic_prefix_length)
        self.demod = OFDMDemodulator(rg.fft_size, l_min, rg.cyclic_prefix_length)
        self.est, self.eq = LSChannelEstimator(rg, "nn"), LMMSEEqualizer(rg, sm)
        self.demapper = Demapper("app", "qam", m)
        from sionna_b200.phy.ofdm import FusedLSLinearDetector
        self.fused = FusedLSLinearDetector(self.est, rg, sm, "app", "qam", m)       # one launch: LS + nn + LMMSE + demap
        nd = rg.num_data_symbols
        for _ in range(2):
            b = src([self.batch, 1, 1, nd * m])
            self.truth.append(b)
            self.inputs.append(chan(mod(rgm(mapper(b))), self.no))
        self.units_per_step = self.batch * nd
        self.desc = (f"configs[2]: OFDM 14x76 CP 6, 64-QAM, TDL-A 300 ns (time domain), OFDMDemodulator + LS(nn) + "
                     f"LMMSEEqualizer + Demapper(app), batch {self.batch}; 2 alternating input sets")

    def run(self, i, x=None):
        y = self.demod(self.inputs[i & 1] if x is None else x)
        return self.fused(y, self.no)

    def run_separate(self, i):
        y = self.demod(self.inputs[i & 1])
        h_hat, ev = self.est(y, self.no)
        x_hat, no_eff = self.eq(y, h_hat, ev, self.no)
        return self.demapper(x_hat, no_eff)

    def stages(self, i):
        rg, b = self.rg, self.batch
        yt = self.inputs[i & 1]
        y = self.demod(yt)
        h_hat, ev = self.est(y, self.no)
        x_hat, no_eff = self.eq(y, h_hat, ev, self.no)
        nsamp, n_re, nd, f_eff = yt.shape[-1], 14 * 76, rg.num_data_symbols, rg.num_effective_subcarriers
        return [
            ("OFDMDemodulator 76-pt", lambda: self.demod(yt), b * (nsamp + n_re) * 8, "16 B per sample in + out"),
            ("Fused LS(nn)+LMMSE+demap (sb_ofdm_frontend)", lambda: self.fused(y, self.no), b * (n_re * 8 + nd * 24),
             "full grid y in, 6 LLRs per data RE out; nothing else touches HBM"),
            ("[separate] LSChannelEstimator nn", lambda: self.est(y, self.no), b * (n_re * 8 + 14 * f_eff * 12),
             "y in; h_hat c64 + err_var f32 out"),
            ("[separate] LMMSEEqualizer 1x1", lambda: self.eq(y, h_hat, ev, self.no), b * (14 * f_eff * (8 + 8 + 4) + nd * 12),
             "y + h_hat + err_var in, x_hat + no_eff out"),
            ("[separate] Demapper app 64-QAM", lambda: self.demapper(x_hat, no_eff), b * nd * (8 + 4 + 24), "x_hat + no_eff in, 6 LLRs out"),
        ]

    def cpu_chain(self, n):
        from oracle import ofdm as F
        from oracle import mapping as M
        rg = self.rg
        yt = self.inputs[0][:n].cpu().numpy().astype(complex)
        mask, pil = rg.pilot_pattern.mask.astype(bool), rg.pilot_pattern.pilots
        eff = F.eff_sc_ind(76, (5, 6), True)
        t0 = time.perf_counter()
        y = F.ofdm_demodulate(yt, 76, self.l_min, rg.cyclic_prefix_length)[..., eff]
        hr, er = F.ls_estimate(y, mask, pil, self.no)
        hr, er = F.nn_interp(hr, mask, pil), F.nn_interp(er, mask, pil)
        xr, nr = F.ofdm_lmmse_equalize(y, hr, er, self.no, mask, F.stream_management([[1]], 1))
        lr = M.demapper(xr.astype(np.complex64), nr.astype(np.float32), M.qam(6), "app")
        dt = time.perf_counter() - t0
        got = self.run(0, self.inputs[0][:n].contiguous()).cpu().numpy()
        ok = np.repeat(nr, 6, axis=-1) < 1.0
        err = float(np.max(np.abs(got[ok] - lr[ok])) / np.abs(lr[ok]).max())
        return {"units": n * rg.num_data_symbols, "seconds": dt, "what": f"oracle/ofdm.py NumPy complex128 chain on {n} frames",
                "max_llr_diff_over_max_llr": err}


class MimoOfdm(Workload):
    """configs[3]: 4 streams x 16 rx antennas, 14 x 76 grid, TDL-A, LS(nn) + LMMSE LinearDetector('app', 16-QAM) + LDPC5G
    (n = 3072, k = 1536 per stream) BP-20; batch 1024 frames."""
    name, default_batch = "mimo_ofdm", 1024
    metric, unit = "coded bits/s, 4x16 MIMO-OFDM LS + LMMSE detection + LDPC5G BP-20", "coded bits/s"
    ebno_db, streams, rx_ant, m = -2.0, 4, 16, 4

    def build(self):
        from sionna_b200.phy.ofdm import ResourceGridMapper, LSChannelEstimator, LinearDetector
        from sionna_b200.phy.mimo import StreamManagement
        from sionna_b200.phy.mapping import Mapper, BinarySource
        from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
        from sionna_b200.phy.channel import TDL, ApplyOFDMChannel, subcarrier_frequencies, cir_to_ofdm_channel
        from sionna_b200.phy.utils import ebnodb2no, ErrorCounter
        rg = self.rg = _rg(self.streams)
        sm = StreamManagement(np.array([[1]]), self.streams)
        self.n = int(rg.num_data_symbols * self.m)
        self.k = self.n // 2
        enc = LDPC5GEncoder(self.k, self.n)
        self.dec = LDPC5GDecoder(enc, hard_out=True, num_iter=20)
        self.no = float(ebnodb2no(self.ebno_db, self.m, 0.5, rg))
        src, mapper, rgm = BinarySource(), Mapper("qam", self.m), ResourceGridMapper(rg)
        tdl = TDL("A", 300e-9, 3.5e9, num_rx_ant=self.rx_ant, num_tx_ant=self.streams)
        freqs, chan = subcarrier_frequencies(76, 15e3), ApplyOFDMChannel()
        self.est = LSChannelEstimator(rg, "nn")
        self.det = LinearDetector("lmmse", "bit", "app", rg, sm, "qam", self.m)
        from sionna_b200.phy.ofdm import FusedLSLinearDetector
        self.fused = FusedLSLinearDetector(self.est, rg, sm, "app", "qam", self.m)
        self.counter = ErrorCounter(self.dev)
        for _ in range(2):
            b = src([self.batch, 1, self.streams, self.k])
            self.truth.append(b)
            a, tau = tdl(self.batch, 14, 1.0 / rg.ofdm_symbol_duration)
            h = cir_to_ofdm_channel(freqs, a, tau, normalize=True)
            self.inputs.append(chan(rgm(mapper(enc(b))), h, self.no))
        self.units_per_step = self.batch * self.streams * self.n
        self.desc = (f"configs[3]: 4 streams x 16 rx antennas, 14x76 grid, 16-QAM, TDL-A 300 ns, LSChannelEstimator(nn) + "
                     f"LinearDetector(lmmse, bit, app) + LDPC5GDecoder(k={self.k}, n={self.n}, 20 it), batch {self.batch}")

    def run(self, i, x=None):
        y = self.inputs[i & 1] if x is None else x
        llr = self.fused(y, self.no)
        b_hat = self.dec(llr)
        if x is None:
            self.counter.update(self.truth[i & 1], b_hat)
        return b_hat

    def stages(self, i):
        rg, b, y = self.rg, self.batch, self.inputs[i & 1]
        h_hat, ev = self.est(y, self.no)
        llr = self.det(y, h_hat, ev, self.no)
        f_eff, nd, K, M = rg.num_effective_subcarriers, rg.num_data_symbols, self.streams, self.rx_ant
        n_re = b * 14 * f_eff
        e, nv = self.dec.num_edges, self.dec.num_vns
        return [
            ("Fused LS(nn)+LMMSE+demap 4x16 (sb_ofdm_frontend)", lambda: self.fused(y, self.no),
             b * (M * 14 * 76 * 8 + K * nd * self.m * 4), "full grid y (16 antennas) in, LLRs out"),
            ("[separate] LSChannelEstimator nn 4x16", lambda: self.est(y, self.no), b * M * 14 * 76 * 8 + b * M * K * 14 * f_eff * 12,
             "y in; h_hat c64 + err_var f32 out over the grid"),
            ("[separate] LinearDetector lmmse/app 16-QAM", lambda: self.det(y, h_hat, ev, self.no),
             n_re * (M * 8 + M * K * 8 + 4) + b * K * nd * self.m * 4,
             "SURVEY 8d: y 128 B + H 512 B + no 4 B per RE in; LLRs out (x_hat / no_eff stay internal to the detector)"),
            ("LDPC5GDecoder BP-20", lambda: self.dec(llr), b * K * (20 * (8 * e + 4 * nv) + 4 * self.n + 4 * self.k),
             "8 B per edge + 4 B per VN per iteration + I/O (SURVEY 8d formula for this graph)"),
        ]

    def cpu_chain(self, n):
        from oracle import ofdm as F
        from oracle import mapping as M
        from oracle import ldpc as O
        rg = self.rg
        y = self.inputs[0][:n].cpu().numpy().astype(complex)
        mask, pil = rg.pilot_pattern.mask.astype(bool), rg.pilot_pattern.pilots
        eff = F.eff_sc_ind(76, (5, 6), True)
        dec = O.LDPC5GDecoderRef(O.LDPC5GEncoderRef(self.k, self.n), num_iter=20)
        t0 = time.perf_counter()
        ye = y[..., eff]
        hr, er = F.ls_estimate(ye, mask, pil, self.no)
        hr, er = F.nn_interp(hr, mask, pil), F.nn_interp(er, mask, pil)
        xr, nr = F.ofdm_lmmse_equalize(ye, hr, er, self.no, mask, F.stream_management([[1]], self.streams))
        lr = M.demapper(xr.astype(np.complex64), nr.astype(np.float32), M.qam(self.m), "app")
        bh = dec(lr.reshape(-1, self.n), num_threads=os.cpu_count() or 1)
        dt = time.perf_counter() - t0
        got = self.run(0, self.inputs[0][:n].contiguous()).cpu().numpy().reshape(-1, self.k)
        return {"units": n * self.streams * self.n, "seconds": dt,
                "what": f"oracle NumPy complex128 LS/LMMSE + C demapper + C BP decoder (libm) on {n} frames",
                "bit_mismatch_vs_gpu": int((got != bh).sum()), "bits": int(bh.size)}


class Pusch(Workload):
    """configs[4]: 5G NR PUSCH, 2 layers / 2 ports, 8 rx antennas, 16 PRB, MCS 14, TDL-B 100 ns; PUSCHReceiver = LS(lin) +
    CDM de-spreading + LMMSE detection + layer demapping + TB decoding (descrambling, rate recovery, BP-20, CRC).
    Global batch 8192 transport blocks, sharded over the ranks (strong scaling)."""
    name, default_batch = "pusch", 8192
    metric, unit = "coded bits/s, 5G NR PUSCHReceiver (LS + LMMSE + TB decode BP-20)", "coded bits/s"
    scaling = "strong"
    ebno_db = 0.0

    def build(self):
        from sionna_b200.phy.nr import PUSCHConfig, PUSCHTransmitter, PUSCHReceiver
        from sionna_b200.phy.channel import ApplyOFDMChannel, TDL, subcarrier_frequencies, cir_to_ofdm_channel
        from sionna_b200.phy.utils import ebnodb2no, ErrorCounter
        self.global_batch = self.batch
        self.batch = self.batch // self.world                      # this rank's shard
        pc = PUSCHConfig(num_layers=2, num_antenna_ports=2)
        pc.carrier.n_size_grid = 16
        pc.dmrs.additional_position = 1
        pc.tb.mcs_index = 14
        self.pc = pc
        self.tx = PUSCHTransmitter(pc)
        self.rx = PUSCHReceiver(self.tx)
        rg = self.rg = self.tx.resource_grid
        self.tdl = TDL("B", 100e-9, 3.5e9, num_rx_ant=8, num_tx_ant=2)
        self.freqs = subcarrier_frequencies(rg.fft_size, rg.subcarrier_spacing)
        self.chan = ApplyOFDMChannel()
        self.no = float(ebnodb2no(self.ebno_db, pc.tb.num_bits_per_symbol, pc.tb_size / pc.num_coded_bits, rg))
        self.counter = ErrorCounter(self.dev)
        for _ in range(2):
            y, b = self.make_batch()
            self.truth.append(b)
            self.inputs.append(y)
        self.units_per_step = self.batch * int(pc.num_coded_bits)
        self.desc = (f"configs[4]: PUSCH 2 layers / 2 ports, 8 rx antennas, 16 PRB, MCS 14 (tb {int(pc.tb_size)} bits, "
                     f"{int(pc.num_coded_bits)} coded), TDL-B 100 ns, Eb/N0 {self.ebno_db:g} dB, PUSCHReceiver; global batch "
                     f"{self.global_batch} sharded over {self.world} GPU(s)")

    def make_batch(self):
        x, b = self.tx(self.batch)
        a, tau = self.tdl(self.batch, self.rg.num_ofdm_symbols, 1.0 / self.rg.ofdm_symbol_duration)
        from sionna_b200.phy.channel import cir_to_ofdm_channel
        h = cir_to_ofdm_channel(self.freqs, a, tau, normalize=True)
        return self.chan(x, h, self.no), b

    def run(self, i, x=None):
        b_hat = self.rx(self.inputs[i & 1] if x is None else x, self.no)
        if x is None:
            self.counter.update(self.truth[i & 1], b_hat)
        return b_hat

    def link_step(self):
        """The whole Monte-Carlo step (transmitter + channel generation + receiver), what sim_ber runs per batch."""
        y, b = self.make_batch()
        b_hat = self.rx(y, self.no)
        self.counter.update(b, b_hat)
        return b_hat

    def stages(self, i):
        rx, rg, b, y = self.rx, self.rg, self.batch, self.inputs[i & 1]
        h_hat, ev = rx._channel_estimator(y, self.no)
        llr = rx._mimo_detector(y, h_hat, ev, self.no)
        llr_l = rx._layer_demapper(llr)
        dec = rx._tb_decoder._decoder
        ant, lay, f_eff, nd = 8, 2, rg.num_effective_subcarriers, rg.num_data_symbols
        n_re = b * rg.num_ofdm_symbols * f_eff
        m = self.pc.tb.num_bits_per_symbol
        cbs = rx._tb_decoder._num_cbs
        e, nv, n_cb, k_cb = dec.num_edges, dec.num_vns, dec.encoder.n, dec.encoder.k
        return [
            ("Fused PUSCH LS+CDM+lin+LMMSE+demap (sb_ofdm_frontend)", lambda: rx._fused(y, self.no),
             b * (ant * rg.num_ofdm_symbols * rg.fft_size * 8 + lay * nd * m * 4), "full grid y in, LLRs out"),
            ("[separate] PUSCHLSChannelEstimator lin", lambda: rx._channel_estimator(y, self.no),
             n_re * ant * 8 + n_re * ant * lay * 12, "y in; h_hat c64 + err_var f32 out over the grid"),
            ("[separate] LinearDetector lmmse/maxlog", lambda: rx._mimo_detector(y, h_hat, ev, self.no),
             n_re * (ant * 8 + ant * lay * 8 + ant * lay * 4 + 4) + b * lay * nd * m * 4,
             "per RE: y + H + err_var + no in; LLRs out"),
            ("LayerDemapper + TBDecoder BP-20", lambda: rx._tb_decoder(rx._layer_demapper(llr)),
             b * cbs * (20 * (8 * e + 4 * nv) + 4 * n_cb + 4 * k_cb) + b * int(self.pc.num_coded_bits) * 8,
             "LDPC: 8 B per edge + 4 B per VN per iteration + I/O; + descramble/de-interleave copies"),
        ]


WORKLOADS = {w.name: w for w in (QpskAwgn, OfdmSiso, MimoOfdm, Pusch)}
