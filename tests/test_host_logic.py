"""Host-side logic that needs no GPU: code construction, decoder graph plan (jagged-diagonal slot layout built in
C++), rate-recovery maps vs the oracle's literal concat/slice restatement, API argument checks."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import ldpc as O


@pytest.mark.parametrize("k,n", [(12, 20), (64, 128), (292, 500), (300, 500), (3824, 5800), (3825, 5800), (640, 3000),
                                 (4224, 8448), (8448, 9000), (8448, 25344)])
def test_code_construction_matches_oracle(k, n):
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder
    enc, ref = LDPC5GEncoder(k, n), O.LDPC5GEncoderRef(k, n)
    assert (enc._bg, enc.z, enc._i_ls, enc.k_ldpc, enc.n_ldpc) == (ref.bg, ref.z, ref.i_ls, ref.k_ldpc, ref.n_ldpc)
    assert (enc.pcm != ref.pcm).nnz == 0
    # closed-form B^-1 (product) really inverts B over GF(2)
    _, b_inv, _, _ = enc._ru_submatrices()
    prod = (b_inv @ ref.Bm).toarray() % 2
    assert np.array_equal(prod, np.eye(4 * enc.z))


def test_benchmark_graph_dimensions():
    """C=4608, N=8832, E=40320 and the degree histograms of the n=8448 decoding graph."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    dec = LDPC5GDecoder(LDPC5GEncoder(4224, 8448))
    assert (dec.num_cns, dec.num_vns, dec.num_edges) == (4608, 8832, 40320)
    cd, vd = np.bincount(dec._cn_idx), np.bincount(dec._vn_idx)
    assert dict(zip(*np.unique(cd, return_counts=True))) == {3: 192, 5: 384, 6: 1344, 7: 960, 8: 384, 9: 384, 10: 192, 19: 768}
    assert dict(zip(*np.unique(vd, return_counts=True))) == {1: 3840, 3: 384, 4: 960, 5: 576, 6: 384, 7: 192, 8: 960,
                                                            9: 1152, 17: 192, 19: 192}
    assert dec.on_chip


def _check_plan(dec):
    ex = dec._graph.export()
    C_, N_, E_, Lc, Lv = (int(x) for x in ex["dims"][:5])
    cn_idx, vn_idx = np.asarray(dec._cn_idx), np.asarray(dec._vn_idx)
    cdeg, vdeg = np.bincount(cn_idx, minlength=C_), np.bincount(vn_idx, minlength=N_)
    # ranks: degree descending, stable
    assert np.array_equal(ex["cn_order"], np.argsort(-cdeg, kind="stable"))
    assert np.array_equal(ex["vn_order"], np.argsort(-vdeg, kind="stable"))
    crank = np.empty(C_, int); crank[ex["cn_order"]] = np.arange(C_)
    vrank = np.empty(N_, int); vrank[ex["vn_order"]] = np.arange(N_)
    slot = ex["slot_of_edge"]
    assert sorted(slot) == list(range(E_))                       # a permutation of the slots
    for c in np.random.default_rng(0).choice(C_, min(C_, 50), replace=False):
        es = np.nonzero(cn_idx == c)[0]
        es = es[np.argsort(vn_idx[es])]                          # ascending VN
        assert [slot[e] for e in es] == [ex["cn_off"][l] + crank[c] for l in range(len(es))]
    for v in np.random.default_rng(1).choice(N_, min(N_, 50), replace=False):
        es = np.nonzero(vn_idx == v)[0]
        es = es[np.argsort(cn_idx[es])]                          # ascending CN
        assert [ex["vn_slot"][ex["vn_off"][l] + vrank[v]] for l in range(len(es))] == [slot[e] for e in es]


def test_graph_plan_reference_order():
    """sum_order="reference": a CN walks its edges in v2c_perm = np.argsort(cn_idx) order, a VN in ascending edge number
    (the list orders of decoding.py:286, 329); such graphs refuse the QC description."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    enc = LDPC5GEncoder(200, 500)
    dec = LDPC5GDecoder(enc, sum_order="reference")
    assert not dec._graph.is_qc()
    br, bc = np.nonzero(enc._bm >= 0)
    assert not dec._graph.set_qc(enc.z, br, bc, enc._bm[br, bc] % enc.z)
    ex = dec._graph.export()
    cn_idx, vn_idx = dec._cn_idx, dec._vn_idx
    crank, vrank = np.argsort(ex["cn_order"]), np.argsort(ex["vn_order"])
    slot = ex["slot_of_edge"]
    perm = np.argsort(cn_idx)
    assert sorted(slot) == list(range(len(slot)))
    for c in np.random.default_rng(0).choice(dec.num_cns, 50, replace=False):
        es = perm[cn_idx[perm] == c]                             # CN-view order
        assert [slot[e] for e in es] == [ex["cn_off"][l] + crank[c] for l in range(len(es))]
    for v in np.random.default_rng(1).choice(dec.num_vns, 50, replace=False):
        es = np.nonzero(vn_idx == v)[0]                          # ascending edge number
        assert [ex["vn_slot"][ex["vn_off"][l] + vrank[v]] for l in range(len(es))] == [slot[e] for e in es]
    with pytest.raises(ValueError):
        LDPC5GDecoder(enc, sum_order="random")


def test_graph_plan_jds_layout():
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder, LDPCBPDecoder
    from sionna_b200.phy.fec.utils import load_parity_check_examples
    _check_plan(LDPC5GDecoder(LDPC5GEncoder(200, 500)))
    for i in range(5):
        _check_plan(LDPCBPDecoder(load_parity_check_examples(i)[0]))


@pytest.mark.parametrize("k,n,m,prune,info", [(100, 200, None, True, True), (100, 200, None, False, False),
                                               (300, 720, 6, True, False), (64, 180, 4, False, True),
                                               (4224, 8448, 2, True, False)])
def test_rate_recovery_maps_equal_reference_concat_slices(k, n, m, prune, info):
    """The kernel's load/store gather maps reproduce LDPC5GDecoder.call's concat/slice sequence
    (decoding.py:1431-1536), checked with a 0-iteration oracle decode on a ramp input."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    enc = LDPC5GEncoder(k, n, num_bits_per_symbol=m)
    dec = LDPC5GDecoder(enc, prune_pcm=prune, return_infobits=info, hard_out=False)
    in_map, n_in, out_vn, n_out = dec._io_maps()
    enc_r = O.LDPC5GEncoderRef(k, n, num_bits_per_symbol=m)
    ref = O.LDPC5GDecoderRef(enc_r, prune_pcm=prune, return_infobits=info, hard_out=False, num_iter=0, llr_max=1e9)
    assert ref.n_pruned == dec._n_pruned and (ref.pcm != dec.pcm).nnz == 0
    x = (np.arange(n, dtype=np.float32) + 1.0)[None, :]
    vn_vals = np.where(in_map >= 0, x[0, np.clip(in_map, 0, n - 1)], np.where(in_map == -1, 0.0, -1e9))
    assert np.array_equal(vn_vals[out_vn][None, :], ref(x))
    assert n_in == n and n_out == (k if info else n)


def test_layered_schedule_and_pruning_quantisation():
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    enc = LDPC5GEncoder(200, 450)
    dec = LDPC5GDecoder(enc, cn_schedule="layered")
    ref = O.LDPC5GDecoderRef(O.LDPC5GEncoderRef(200, 450), cn_schedule="layered")
    assert dec._n_pruned == ref.n_pruned and np.array_equal(dec._cn_schedule, ref.schedule)
    assert dec._cn_schedule.shape[1] == enc.z and dec.num_cns % enc.z == 0


def test_decoder_argument_checks():
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder, LDPCBPDecoder
    pcm = np.array([[1, 1, 0], [0, 1, 1]], dtype=np.float64)
    with pytest.raises(TypeError):
        LDPCBPDecoder(pcm, hard_out=1)
    with pytest.raises(ValueError):
        LDPCBPDecoder(pcm, num_iter=-1)
    with pytest.raises(TypeError):
        LDPCBPDecoder(pcm, cn_update="nope")
    with pytest.raises(ValueError):
        LDPCBPDecoder(pcm * 2)
    with pytest.raises(TypeError):
        LDPCBPDecoder(pcm, cn_type="boxplus")
    with pytest.raises(ValueError):
        LDPCBPDecoder(pcm, cn_schedule=np.array([[0, 5]]))
    with pytest.raises(TypeError):
        LDPC5GDecoder(pcm)
    with pytest.raises(ValueError):
        LDPC5GEncoder(8449, 10000)
    with pytest.raises(ValueError):
        LDPC5GEncoder(100, 1000)
    d = LDPCBPDecoder(sp.csr_matrix(pcm))
    d.num_iter = 3
    d.llr_max = 10
    assert d.num_iter == 3 and d.llr_max == 10.0 and d.num_edges == 4 and d.coderate == pytest.approx(1 / 3)
    with pytest.raises(AssertionError):
        d.build((5, 4))


def test_constellation_host_side():
    from sionna_b200.phy.mapping import Constellation, qam, pam
    from oracle import mapping as M
    for m in (2, 4, 6, 8):
        assert np.array_equal(qam(m), M.qam(m))
    assert np.array_equal(pam(3), M.pam(3))
    c = Constellation("qam", 4)
    assert c.num_points == 16 and c.points.shape == (16,)
    with pytest.raises(ValueError):
        Constellation("qam", 3)
    with pytest.raises(ValueError):
        Constellation("custom", 2)
    with pytest.raises(ValueError):
        c.points = np.zeros(16)


def test_ebnodb2no_matches_oracle():
    from sionna_b200.phy.utils import ebnodb2no, hard_decisions
    from oracle import mapping as M
    import torch
    for e in (0.0, 2.0, 4.5):
        assert float(ebnodb2no(e, 2, 0.5)) == float(M.ebnodb2no(e, 2, 0.5))
    assert torch.equal(hard_decisions(torch.tensor([-1.0, 0.0, 3.0])), torch.tensor([0.0, 0.0, 1.0]))


def test_qc_description_accepted_and_rejected():
    """sb_ldpc_graph_set_qc verifies the lifted base graph against the edge list (host only, no GPU)."""
    from sionna_b200.phy.fec.ldpc import LDPC5GEncoder, LDPC5GDecoder
    for k, n in ((100, 200), (562, 871), (4224, 8448)):      # (562, 871): pruned size is not a multiple of Z
        enc = LDPC5GEncoder(k, n)
        dec = LDPC5GDecoder(enc)
        assert dec._graph.is_qc()
        br, bc = np.nonzero(enc._bm >= 0)
        sh = enc._bm[br, bc] % enc.z
        bad = sh.copy()
        bad[3] = (bad[3] + 1) % enc.z
        assert not dec._graph.set_qc(enc.z, br, bc, bad)      # wrong shift -> rejected, handle keeps the valid one
        assert dec._graph.is_qc()
        assert not dec._graph.set_qc(enc.z, br[:-1], bc[:-1], sh[:-1]) or dec.num_cns < 46 * enc.z


def test_channel_host_logic():
    """TDL model tables / derived properties and the discrete-time lag rule (no device needed)."""
    from sionna_b200.phy.channel import TDL, time_lag_discrete_time_channel
    from sionna_b200.phy.channel.tdl import subcarrier_frequencies
    t = TDL("A", 300e-9, 3.5e9)
    assert t.num_clusters == 23 and not t.los
    assert abs(float(t.mean_powers.sum()) - 1.0) < 1e-6
    assert abs(float(t.delays[-1]) - 9.6586 * 300e-9) < 1e-12          # TR 38.901 Table 7.7.2-1, last tap
    d = TDL("D", 100e-9, 3.5e9, min_speed=1.0, max_speed=2.0)
    assert d.los and d.num_clusters == 13
    assert abs(10 * np.log10(d.k_factor) - 13.3) < 0.05                  # K-factor of TDL-D
    assert abs(float(d.mean_powers.sum()) - 1.0) < 1e-6
    assert abs(d._doppler(3.0) - 2 * np.pi * 3.0 / 299792458.0 * 3.5e9) < 1e-9
    c = TDL("C300", 10e-9, 3.5e9)                                        # fixed-delay model: delay spread forced
    assert c.delay_spread == 300e-9 and abs(float(c.delays[-1]) - 2595e-9) < 1e-12
    with pytest.raises(AssertionError):
        TDL("Z", 1e-7, 3.5e9)
    with pytest.raises(AssertionError):
        TDL("A", 1e-7, 3.5e9, min_speed=5.0, max_speed=1.0)
    assert time_lag_discrete_time_channel(30.72e6) == (-6, int(np.ceil(3e-6 * 30.72e6)) + 6)
    f = subcarrier_frequencies(5, 15e3).numpy()
    assert np.array_equal(f, np.array([-2, -1, 0, 1, 2]) * 15e3)
    f = subcarrier_frequencies(4, 15e3).numpy()
    assert np.array_equal(f, np.array([-2, -1, 0, 1]) * 15e3)


def test_separable_constellation_detection():
    """The per-dimension demapper is only selected for constellations that factor exactly (oracle twin of the host rule)."""
    from oracle import mapping as M
    for m in (2, 4, 6, 8, 10):
        lev = M.separable_levels(M.qam(m))
        assert lev is not None and len(lev[0]) == 2 ** (m // 2)
        pts = M.qam(m)
        assert M.separable_levels(pts * np.exp(1j * 0.1)) is None          # rotated: not separable
    assert M.separable_levels(M.pam(3)) is None                              # odd number of bits
    pts = M.qam(4).copy()
    pts[5] += 0.01
    assert M.separable_levels(pts) is None                                   # perturbed (e.g. trained) constellation


def test_pilot_pattern_and_resource_grid_host_checks():
    """Host containers mirror the reference's own property tests (test/unit/ofdm/test_pilot_pattern.py:11-95):
    argument validation, pilot / data symbol counts, normalisation, the empty pattern; plus the RE-type bookkeeping of
    ResourceGrid with guards, DC null and a Kronecker pattern."""
    from sionna_b200.phy.ofdm import PilotPattern, EmptyPilotPattern, KroneckerPilotPattern, ResourceGrid
    with pytest.raises(AssertionError):
        PilotPattern(np.zeros([1, 10], bool), np.zeros([1, 10, 20], np.complex64))               # mask rank
    with pytest.raises(AssertionError):
        PilotPattern(np.zeros([4, 2, 10, 46], bool), np.zeros([1, 10, 20, 2], np.complex64))      # pilots rank
    mask = np.zeros([1, 2, 14, 64], bool)
    mask[0, 0, 0, :] = True
    mask[0, 1, 1, :] = True
    with pytest.raises(AssertionError):
        PilotPattern(mask, np.zeros([1, 3, 64], np.complex64))                                     # leading dims differ
    with pytest.raises(AssertionError):
        PilotPattern(mask, np.zeros([1, 2, 65], np.complex64))                                     # wrong pilot count
    bad = mask.copy()
    bad[0, 1, 1:3, :] = True
    with pytest.raises(AssertionError):
        PilotPattern(bad, np.zeros([1, 2, 128], np.complex64))                                     # unequal counts
    pp = PilotPattern(mask, np.zeros([1, 2, 64], np.complex64))
    assert (pp.num_pilot_symbols, pp.num_data_symbols) == (64, 13 * 64)
    m2 = np.zeros([1, 2, 14, 64], bool)
    m2[0, 0, :2, :] = True
    m2[0, 1, 1:3, :] = True
    pp2 = PilotPattern(m2, np.zeros([1, 2, 128], np.complex64))
    assert (pp2.num_pilot_symbols, pp2.num_data_symbols) == (128, 12 * 64)
    ppn = PilotPattern(mask, 3 * np.ones([1, 2, 64], np.complex64), normalize=True)
    assert np.allclose(np.mean(np.abs(ppn.pilots) ** 2, -1), 1.0)
    e = EmptyPilotPattern(4, 2, 14, 55)
    assert (e.num_pilot_symbols, e.num_data_symbols) == (0, 14 * 55)
    rg = ResourceGrid(14, 76, 15e3, num_tx=2, num_streams_per_tx=2, cyclic_prefix_length=6, num_guard_carriers=(5, 6),
                      dc_null=True, pilot_pattern="kronecker", pilot_ofdm_symbol_indices=[2, 11])
    assert rg.num_effective_subcarriers == 64 and rg.num_pilot_symbols == 2 * 64 and rg.num_data_symbols == 12 * 64
    assert rg.num_resource_elements == 14 * 76 and rg.num_time_samples == 14 * 82 and rg.bandwidth == 76 * 15e3
    kp = rg.pilot_pattern
    assert isinstance(kp, KroneckerPilotPattern) and kp.mask.shape == (2, 2, 14, 64)
    nz = np.abs(kp.pilots.reshape(4, 2, 64)) > 0
    assert np.all(nz.sum(axis=0) == 1)                      # the four streams sound disjoint subcarrier combs
    assert np.allclose(np.mean(np.abs(kp.pilots) ** 2, -1), 1.0)
    assert rg.dc_ind == 38 and 38 not in rg.effective_subcarrier_ind and len(rg.effective_subcarrier_ind) == 64


def test_stream_management_equals_reference_outputs():
    """Every derived index array equals what the reference class produced for the same inputs
    (tests/golden/make_stream_management_golden.py ran /root/reference/src/sionna/phy/mimo/stream_management.py)."""
    import json
    import os
    from sionna_b200.phy.mimo import StreamManagement
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "stream_management_golden.json")))
    assert len(cases) >= 10
    for c in cases:
        sm = StreamManagement(np.array(c["rx_tx_association"]), c["num_streams_per_tx"])
        for f, want in c.items():
            if f in ("rx_tx_association",):
                continue
            got = np.asarray(getattr(sm, f))
            assert np.array_equal(got, np.asarray(want)), (c["rx_tx_association"], f)
    with pytest.raises(AssertionError):
        StreamManagement(np.array([[1, 1], [1, 0]]), 1)
    with pytest.raises(AssertionError):
        StreamManagement(np.array([[2]]), 1)


@pytest.mark.parametrize("interp,streams", [("nn", 1), ("lin", 4), ("lin_time_avg", 2)])
def test_frontend_tables_reproduce_ls_plus_interpolation(interp, streams):
    """The fused front-end's host tables (ofdm/frontend.py): sum_i t_w * y[t_idx] must equal LS estimation followed by the
    interpolation (oracle restatement of channel_estimation.py:138-285, 364-734) for random received grids, and e_sum the
    interpolated, floored error variance per unit noise power summed over the streams."""
    from sionna_b200.phy.ofdm import ResourceGrid, LSChannelEstimator, frontend_tables
    from oracle import ofdm as F
    rg = ResourceGrid(14, 76, 15e3, num_tx=1, num_streams_per_tx=streams, cyclic_prefix_length=6, num_guard_carriers=(5, 6),
                      dc_null=True, pilot_pattern="kronecker", pilot_ofdm_symbol_indices=[2, 11])
    est = LSChannelEstimator(rg, interp)
    t = frontend_tables(rg, est)
    assert t is not None and t["num_terms"] <= (1 if interp == "nn" else 4)
    rng = np.random.default_rng(3)
    y = rng.normal(size=(2, 1, 3, 14, 76)) + 1j * rng.normal(size=(2, 1, 3, 14, 76))
    eff = F.eff_sc_ind(76, (5, 6), True)
    mask, pil = rg.pilot_pattern.mask.astype(bool), rg.pilot_pattern.pilots
    h, err = F.ls_estimate(y[..., eff], mask, pil, 1.0)
    if interp == "nn":
        hr, er = F.nn_interp(h, mask, pil), F.nn_interp(err, mask, pil)
    else:
        hr, er = F.lin_interp(h, mask, pil, interp == "lin_time_avg"), F.lin_interp(err, mask, pil, interp == "lin_time_avg").real
    yf = y.reshape(2, 1, 3, -1)
    idx, w = t["t_idx"], t["t_w"].astype(np.complex128)                      # [ts, RE, NT]
    got = np.where(idx >= 0, w * yf[..., np.maximum(idx, 0)], 0).sum(-1)     # [B, rx, ant, ts, RE]
    want = hr.reshape(2, 1, 3, streams, -1)
    assert np.allclose(got, want, atol=1e-6)
    e_sum = np.maximum(er[0, 0, 0].reshape(streams, -1), 0).sum(0)
    assert np.allclose(t["e_sum"], e_sum, rtol=1e-6)
    assert np.array_equal(t["re_full"], (np.arange(14)[:, None] * 76 + eff[None, :]).reshape(-1))


@pytest.mark.parametrize("layers,ports,length,addpos,cdm,interp", [(2, 2, 1, 1, 2, "lin"), (1, 1, 2, 1, 2, "lin"),
                                                                     (4, 4, 1, 0, 2, "nn"), (2, 4, 2, 0, 1, "lin")])
def test_frontend_tables_pusch_cdm_despreading(layers, ports, length, addpos, cdm, interp):
    """PUSCH: the fused front-end's tables also fold the CDM de-spreading of PUSCHLSChannelEstimator
    (nr/pusch_channel_estimation.py:131-167) between the LS division and the interpolation; checked against the oracle
    chain ls_estimate -> pusch_ls_combine -> interpolation on random received grids."""
    from sionna_b200.phy.nr import PUSCHConfig, PUSCHTransmitter
    from sionna_b200.phy.nr.pusch_channel_estimation import PUSCHLSChannelEstimator
    from sionna_b200.phy.ofdm import frontend_tables
    from oracle import ofdm as F
    from oracle import nr as ON
    kw = dict(precoding="codebook", tpmi=1) if ports > layers else {}
    pc = PUSCHConfig(num_layers=layers, num_antenna_ports=ports, **kw)
    pc.carrier.n_size_grid = 4
    pc.dmrs.length = length
    pc.dmrs.additional_position = addpos
    pc.dmrs.num_cdm_groups_without_data = cdm
    tx = PUSCHTransmitter(pc)
    rg = tx.resource_grid
    est = PUSCHLSChannelEstimator(rg, tx._dmrs_length, tx._dmrs_additional_position, tx._num_cdm_groups_without_data,
                                  interpolation_type=interp)
    t = frontend_tables(rg, est)
    assert t is not None and t["num_terms"] <= 16
    rng = np.random.default_rng(9)
    s_n, nf = rg.num_ofdm_symbols, rg.fft_size
    y = rng.normal(size=(2, 1, 2, s_n, nf)) + 1j * rng.normal(size=(2, 1, 2, s_n, nf))
    eff = np.asarray(rg.effective_subcarrier_ind)
    mask, pil = rg.pilot_pattern.mask.astype(bool), rg.pilot_pattern.pilots
    h, err = F.ls_estimate(y[..., eff], mask, pil, 1.0)
    h, err = ON.pusch_ls_combine(h, err, est._num_dmrs_syms, est._dmrs_length, est._num_cdm_groups_without_data)
    if interp == "nn":
        hr, er = F.nn_interp(h, mask, pil), F.nn_interp(err, mask, pil)
    else:
        hr, er = F.lin_interp(h, mask, pil), F.lin_interp(err, mask, pil).real
    ts = hr.shape[3] * hr.shape[4]
    yf = y.reshape(2, 1, 2, -1)
    idx, w = t["t_idx"], t["t_w"].astype(np.complex128)
    got = np.where(idx >= 0, w * yf[..., np.maximum(idx, 0)], 0).sum(-1)
    assert np.allclose(got, hr.reshape(2, 1, 2, ts, -1), atol=2e-6)
    assert np.allclose(t["e_sum"], np.maximum(er[0, 0, 0].reshape(ts, -1), 0).sum(0), rtol=1e-6)


def test_fused_front_end_device_layout_lists_only_data_res():
    """`FusedLSLinearDetector` hands the kernel a compacted RE list (resource elements that carry data for at least one
    stream; pilot-only OFDM symbols are not walked) and term-major tables [streams, terms, listed REs]: the arrays it
    uploads must be exactly that view of `frontend_tables` / the stream-management maps."""
    from sionna_b200.phy.ofdm import ResourceGrid, LSChannelEstimator, FusedLSLinearDetector, frontend_tables
    from sionna_b200.phy.ofdm.equalization import _sm_tables
    from sionna_b200.phy.mimo import StreamManagement
    rg = ResourceGrid(14, 76, 15e3, num_tx=1, num_streams_per_tx=2, cyclic_prefix_length=6, num_guard_carriers=(5, 6),
                      dc_null=True, pilot_pattern="kronecker", pilot_ofdm_symbol_indices=[2, 11])
    sm = StreamManagement(np.array([[1]]), 2)
    est = LSChannelEstimator(rg, "lin")
    fused = FusedLSLinearDetector(est, rg, sm, "maxlog", "qam", 4)
    t = frontend_tables(rg, est)
    _, _, _, data_pos = _sm_tables(rg, sm)
    data_pos = np.asarray(data_pos)
    keep = np.nonzero((data_pos >= 0).any(0))[0]
    assert len(keep) == 12 * 64 == fused._num_listed                        # 2 of 14 symbols carry only pilots
    d = fused._np
    assert d["t_idx"].shape == (2, t["num_terms"], len(keep)) and d["t_idx"].flags["C_CONTIGUOUS"]
    assert np.array_equal(d["t_idx"], t["t_idx"][:, keep, :].transpose(0, 2, 1))
    assert np.array_equal(d["t_w"], t["t_w"][:, keep, :].transpose(0, 2, 1))
    assert np.array_equal(d["re_full"], t["re_full"][keep]) and np.array_equal(d["e_sum"], t["e_sum"][keep])
    assert np.array_equal(d["data_pos"], data_pos[:, keep]) and (d["data_pos"] >= 0).any(0).all()
    # every stream's data symbols are all reachable through the list, each exactly once
    for q in range(2):
        assert np.array_equal(np.sort(d["data_pos"][q][d["data_pos"][q] >= 0]), np.arange(rg.pilot_pattern.num_data_symbols))
    assert fused._lev[0].dtype == np.float32 and len(fused._lev[0]) == 4


def test_block_call_protocol_and_double_precision_fallback():
    """Block.__call__ (reference block.py:82-155): float / complex arguments are cast to the block precision, ints and
    Python scalars are left alone, build() runs once with the argument shapes; a block set to precision="double" keeps the
    float64 / complex128 I/O contract while its call() sees single-precision tensors, and says so once per class. The
    protocol is pure host logic, so a dummy block on the CPU device exercises it (kernels never run here)."""
    import warnings
    import torch
    from sionna_b200.phy import config
    from sionna_b200.phy.block import Block, PrecisionWarning

    class Probe(Block):
        def __init__(self, **kw):
            super().__init__(**kw)
            self.builds, self.seen = [], []

        def build(self, *shapes, **kw_shapes):
            self.builds.append((shapes, kw_shapes))

        def call(self, x, idx, scale=None):
            self.seen.append((x.dtype, idx.dtype, None if scale is None else scale.dtype, self.precision))
            return x * 2, {"idx": idx, "z": x.to(torch.complex64) if not x.dtype.is_complex else x}

    old = config._device
    try:
        config.device = "cpu"
        p = Probe()
        out, extra = p(np.ones((2, 3), np.float64), torch.arange(3), scale=torch.ones(3, dtype=torch.float64))
        assert out.dtype == torch.float32 and p.seen[0] == (torch.float32, torch.int64, torch.float32, "single")
        assert p.builds == [(((2, 3), (3,)), {"scale": (3,)})] and p.built
        p(torch.ones(4, 3), torch.arange(3))
        assert len(p.builds) == 1                                            # build() once
        with pytest.raises(ValueError):
            Probe(precision="half")
        d = Probe(precision="double")
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            out, extra = d(torch.ones(2, 3, dtype=torch.float32), torch.arange(3))
            d(torch.ones(2, 3), torch.arange(3))
        assert sum(issubclass(i.category, PrecisionWarning) for i in w) <= 1  # once per class (0 if another test warned first)
        assert d.seen[0][0] == torch.float32 and d.seen[0][3] == "single"     # call() ran in single precision ...
        assert d.precision == "double" and d.rdtype == torch.float64          # ... and the block is double again afterwards
        assert out.dtype == torch.float64 and extra["z"].dtype == torch.complex128 and extra["idx"].dtype == torch.int64
    finally:
        config._device = old
