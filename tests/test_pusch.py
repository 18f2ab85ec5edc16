"""PUSCH (BASELINE.json configs[4]): host configuration logic and the CPU oracle against the
reference's golden vectors (CPU), then the CUDA chain against the oracle and the same vectors (gpu)."""
import json
import os
import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pusch_golden.npz")


@pytest.fixture(scope="module")
def gold():
    with np.load(GOLD) as z:
        d = {k: z[k] for k in z.files}
    d["configs"] = json.loads(bytes(d["configs_json"]).decode())
    return d


def make_config(c):
    """PUSCHConfig of a golden case, set up like the reference's test (test_pusch_transmitter.py:17-49)."""
    from sionna_b200.phy.nr import PUSCHConfig
    pc = PUSCHConfig()
    pc.carrier.n_cell_id = c["carrier"]["n_cell_id"]
    pc.carrier.slot_number = c["carrier"]["slot_number"]
    p, d = c["pusch"], c["pusch"]["dmrs"]
    pc.n_size_bwp = p["n_size_bwp"]
    pc.symbol_allocation = p["symbol_allocation"]
    pc.n_rnti = p["n_rnti"]
    pc.num_antenna_ports = p["num_antenna_ports"]
    pc.num_layers = p["num_layers"]
    pc.precoding = p["precoding"]
    if pc.precoding == "codebook":
        pc.tpmi = p["tpmi"]
    pc.dmrs.length = d["length"]
    pc.dmrs.config_type = d["config_type"]
    pc.dmrs.additional_position = d["additional_position"]
    pc.dmrs.num_cdm_groups_without_data = d["num_cdm_groups_without_data"]
    pc.dmrs.dmrs_port_set = d["dmrs_port_set"]
    pc.dmrs.n_scid = d["n_scid"]
    pc.dmrs.n_id = d["n_id"]
    pc.tb.mcs_index = p["tb"]["mcs_index"]
    pc.tb.mcs_table = p["tb"]["mcs_table"]
    return pc


def projections(n, seed, nproj):
    rng = np.random.default_rng(seed)                        # same construction as tests/golden/make_pusch_golden.py
    return rng.standard_normal((nproj, n)) + 1j * rng.standard_normal((nproj, n))


def oracle_cfg(pc):
    s0, sl = pc.symbol_allocation
    return {"m": pc.tb.num_bits_per_symbol, "target_coderate": float(pc.tb.target_coderate),
            "num_layers": pc.num_layers, "n_rnti": pc.n_rnti,
            "n_id": pc.carrier.n_cell_id if pc.tb.n_id is None else pc.tb.n_id,
            "dmrs_mask": pc.dmrs_mask[:, s0:s0 + sl], "dmrs_grid": pc.dmrs_grid[:, :, s0:s0 + sl],
            "w": pc.precoding_matrix}


def check_grid(gold, i, grid):
    """grid [ports, symbols, subcarriers] vs the stored fingerprint (all cases) and the full grid (subset)."""
    g = np.transpose(grid, (2, 1, 0))                        # the reference stores [subcarriers, symbols, ports]
    assert list(g.shape) == list(gold[f"shape_{i}"])
    proj = projections(g.size, 1000 + i, int(gold["nproj"])) @ g.reshape(-1)
    assert np.allclose(proj, gold[f"proj_{i}"], rtol=1e-4, atol=1e-3 * np.sqrt(g.size))
    if f"grid_{i}" in gold:
        assert np.allclose(g, gold[f"grid_{i}"], atol=1e-5)


# ---- CPU: host logic + oracle --------------------------------------------------------------------------------------
def test_dmrs_sequences_match_reference_vectors(gold):
    """PUSCHConfig.dmrs_grid vs reference_dmrs_{1,2}.npy (the reference's test_pusch_config.py:19-64)."""
    from sionna_b200.phy.nr import PUSCHConfig
    for k, n_grid in ((1, 1), (2, 4)):
        pc = PUSCHConfig()
        pc.carrier.n_size_grid = n_grid
        pc.dmrs.config_type = 2
        pc.dmrs.num_cdm_groups_without_data = 3
        pc.dmrs.additional_position = 1
        pc.dmrs.length = 2
        pc.dmrs.n_id = [4, 4]
        cols = []
        for n_cell_id in [0, 1, 10, 24, 99, 1006]:
            for slot in [0, 1, 5, 9]:
                for port in [0, 3, 4, 9, 11]:
                    pc.carrier.n_cell_id = n_cell_id
                    pc.carrier.slot_number = slot
                    pc.dmrs.dmrs_port_set = [port]
                    a = pc.dmrs_grid
                    pil = np.concatenate([a[0, :, 2], a[0, :, 3], a[0, :, 10], a[0, :, 11]])
                    cols.append(pil[np.where(pil)] / np.sqrt(3))
        assert np.allclose(np.transpose(np.array(cols)), gold[f"reference_dmrs_{k}"], atol=1e-6)


def test_config_defaults_and_validation():
    from sionna_b200.phy.nr import PUSCHConfig, CarrierConfig, PUSCHDMRSConfig, TBConfig, decode_mcs_index
    pc = PUSCHConfig()
    assert (pc.num_subcarriers, pc.dmrs_symbol_indices, pc.num_coded_bits, pc.tb_size) == (48, [2], 2496, 1352)
    assert pc.dmrs.allowed_dmrs_ports == [0, 1, 2, 3] and pc.dmrs.beta == pytest.approx(np.sqrt(2))
    with pytest.raises(AssertionError):
        pc.num_layers = 5
    with pytest.raises(AssertionError):
        CarrierConfig(subcarrier_spacing=20)
    with pytest.raises(AssertionError):
        PUSCHConfig(num_layers=2)                             # non-codebook needs num_layers == num_antenna_ports
    with pytest.raises(AssertionError):
        PUSCHDMRSConfig(length=2, additional_position=2)
    c = CarrierConfig(subcarrier_spacing=30, slot_number=19)
    assert (c.mu, c.num_slots_per_frame, c.num_symbols_per_slot) == (1, 20, 14)
    assert CarrierConfig(subcarrier_spacing=60, cyclic_prefix="extended").num_symbols_per_slot == 12
    assert decode_mcs_index(14)[0] == 4 and decode_mcs_index(14)[1] == pytest.approx(553 / 1024)
    assert decode_mcs_index(27, 2)[0] == 8
    assert decode_mcs_index(0, 1, True, True, True) == (1, pytest.approx(240 / 1024))
    with pytest.raises(AssertionError):
        decode_mcs_index(28, 2)
    t = TBConfig(mcs_index=3, mcs_table=3)
    assert t.num_bits_per_symbol == 2 and t.target_coderate == pytest.approx(64 / 1024)
    clone = pc.clone()
    clone.carrier.n_cell_id = 7
    assert pc.carrier.n_cell_id == 1
    w = PUSCHConfig(num_layers=2, num_antenna_ports=4, precoding="codebook", tpmi=7).precoding_matrix
    assert w.shape == (4, 2) and np.allclose(np.sum(np.abs(w) ** 2), 1.0)


def test_dmrs_ports_are_orthogonal():
    """LS estimation on a block-constant channel separates all ports (the reference's test_pusch_config.py:66-169)."""
    from sionna_b200.phy.nr import PUSCHConfig
    from oracle import nr as ON
    rng = np.random.default_rng(3)
    for config_type, length, groups, ports in ((2, 2, 3, [1, 2, 5, 11]), (2, 1, 3, [2, 3, 4, 5]), (1, 1, 2, [0, 1, 2, 3])):
        pc = PUSCHConfig(num_layers=4, num_antenna_ports=4)
        pc.carrier.n_size_grid = 4
        pc.dmrs.config_type, pc.dmrs.length, pc.dmrs.num_cdm_groups_without_data = config_type, length, groups
        pc.dmrs.dmrs_port_set = ports
        for add in range(2):
            pc.dmrs.additional_position = add
            a = pc.dmrs_grid                                   # [ports, subcarriers, symbols]
            chan = rng.standard_normal(4) + 1j * rng.standard_normal(4)
            y = np.einsum("p,pks->ks", chan, a)
            m = pc.dmrs_mask.T
            n_sym = len(pc.dmrs_symbol_indices)
            for j in range(4):
                pil = a[j].T[m]
                with np.errstate(divide="ignore", invalid="ignore"):
                    h = np.where(pil == 0, 0, y.T[m] / pil)
                h, _ = ON.pusch_ls_combine(h, np.zeros_like(h, float), n_sym, length, groups)
                assert np.allclose(h[pil != 0], chan[j])


@pytest.mark.parametrize("case", list(range(83)))
def test_oracle_transmitter_matches_reference_vectors(gold, case):
    """Host configuration + CPU oracle chain == the reference's stored transmit grids
    (all 83 cases of test_pusch_transmitter.py:tests_against_reference)."""
    from oracle import nr as ON
    pc = make_config(gold["configs"][case])
    b = np.unpackbits(gold[f"b_{case}"])[:int(gold[f"nb_{case}"])]
    assert pc.tb_size == b.size
    grid = ON.pusch_transmit(oracle_cfg(pc), b[None].astype(np.float32), pc.tb_size, pc.num_coded_bits)[0]
    check_grid(gold, case, grid)


# ---- GPU: CUDA chain vs the reference vectors and the oracle ---------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", list(range(83)))
def test_transmitter_matches_reference_vectors(gold, case, cuda_device):
    import torch
    from sionna_b200.phy.nr import PUSCHTransmitter
    pc = make_config(gold["configs"][case])
    b = np.unpackbits(gold[f"b_{case}"])[:int(gold[f"nb_{case}"])].astype(np.float32)
    tx = PUSCHTransmitter(pc, return_bits=False)
    x = tx(torch.from_numpy(b[None, None]).to(cuda_device))                # [1, 1, ports, symbols, subcarriers]
    assert list(x.shape) == [1, 1, pc.num_antenna_ports, 14, pc.num_subcarriers]
    check_grid(gold, case, x[0, 0].cpu().numpy())


@pytest.mark.gpu
def test_layer_mapper_and_precoder(cuda_device):
    import torch
    from sionna_b200.phy.nr import LayerMapper, LayerDemapper, PUSCHPrecoder, PUSCHConfig
    from oracle import nr as ON
    rng = np.random.default_rng(0)
    for nl in (1, 2, 3, 4):
        x = (rng.standard_normal((3, 2, 24 * nl)) + 1j * rng.standard_normal((3, 2, 24 * nl))).astype(np.complex64)
        lm = LayerMapper(nl)
        y = lm(torch.from_numpy(x).to(cuda_device))
        assert np.array_equal(y.cpu().numpy(), ON.layer_map(x, nl))
        llr = rng.standard_normal((3, 2, nl, 24 * 4)).astype(np.float32)
        z = LayerDemapper(lm, 4)(torch.from_numpy(llr).to(cuda_device))
        assert np.array_equal(z.cpu().numpy(), ON.layer_demap(llr, 4))
    lm8 = LayerMapper(7)                                                    # dual codeword mode, 3 + 4 layers
    x0, x1 = rng.standard_normal((2, 30)).astype(np.float32), rng.standard_normal((2, 40)).astype(np.float32)
    y = lm8([torch.from_numpy(x0).to(cuda_device), torch.from_numpy(x1).to(cuda_device)])
    assert list(y.shape) == [2, 7, 10]
    z0, z1 = LayerDemapper(lm8, 1)(y)
    assert np.array_equal(z0.cpu().numpy(), x0) and np.array_equal(z1.cpu().numpy(), x1)
    for layers, ports, tpmi in ((1, 2, 3), (2, 4, 13), (3, 4, 2), (4, 4, 4)):
        ws = [PUSCHConfig(num_layers=layers, num_antenna_ports=ports, precoding="codebook", tpmi=t).precoding_matrix
              for t in (tpmi, 0)]
        x = (rng.standard_normal((5, 2, layers, 3, 12)) + 1j * rng.standard_normal((5, 2, layers, 3, 12))).astype(np.complex64)
        y = PUSCHPrecoder(ws)(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
        ref = np.einsum("tpl,btlsf->btpsf", np.stack(ws), x)
        assert np.allclose(y, ref, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("config_type,length,groups,add,layers", [(1, 1, 2, 1, 2), (1, 2, 2, 1, 4), (2, 1, 3, 0, 4),
                                                                   (2, 2, 3, 1, 2), (1, 1, 1, 2, 1)])
def test_pusch_ls_channel_estimator_vs_oracle(cuda_device, config_type, length, groups, add, layers):
    """PUSCHLSChannelEstimator (LS + CDM de-spreading + interpolation) == NumPy restatement; exact channel recovery on a
    noiseless block-constant channel."""
    import torch
    from sionna_b200.phy.nr import PUSCHConfig, PUSCHTransmitter, PUSCHLSChannelEstimator
    from oracle import ofdm as OO, nr as ON
    rng = np.random.default_rng(1)
    pc = PUSCHConfig(num_layers=layers, num_antenna_ports=layers)
    pc.carrier.n_size_grid = 3
    pc.dmrs.config_type, pc.dmrs.length = config_type, length
    pc.dmrs.num_cdm_groups_without_data, pc.dmrs.additional_position = groups, add
    tx = PUSCHTransmitter(pc)
    rg = tx.resource_grid
    x, _ = tx(4)                                                             # [4, 1, layers, 14, 36]
    xn = x.cpu().numpy()
    n_ant = 3
    chan = rng.standard_normal((4, 1, n_ant, 1, layers)) + 1j * rng.standard_normal((4, 1, n_ant, 1, layers))
    y = np.einsum("bratl,btlsf->brasf", chan, xn)
    no = np.float32(0.05)
    y_noisy = (y + np.sqrt(no / 2) * (rng.standard_normal(y.shape) + 1j * rng.standard_normal(y.shape))).astype(np.complex64)
    n_sym = len(pc.dmrs_symbol_indices)
    mask, pilots = rg.pilot_pattern.mask.astype(bool), rg.pilot_pattern.pilots
    for interp in ("nn", "lin"):
        est = PUSCHLSChannelEstimator(rg, length, n_sym // length - 1, groups, interpolation_type=interp)
        h, err = est(torch.from_numpy(y_noisy).to(cuda_device), float(no))
        h_r, e_r = OO.ls_estimate(y_noisy, mask, pilots, no)
        h_r, e_r = ON.pusch_ls_combine(h_r, e_r, n_sym, length, groups)
        if interp == "nn":
            h_i, e_i = OO.nn_interp(h_r, mask, pilots), OO.nn_interp(e_r, mask, pilots)
        else:
            h_i, e_i = OO.lin_interp(h_r, mask, pilots), np.maximum(OO.lin_interp(e_r, mask, pilots), 0)
        assert np.allclose(h.cpu().numpy(), h_i, rtol=1e-4, atol=1e-5)
        assert np.allclose(err.cpu().numpy(), e_i, rtol=1e-4, atol=1e-7)
        h0, _ = est(torch.from_numpy(y.astype(np.complex64)).to(cuda_device), 0.0)     # noiseless: exact channel
        want = np.broadcast_to(chan[..., None, None], h0.shape)
        assert np.allclose(h0.cpu().numpy(), want, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("scenario", ["siso_awgn", "mimo_2x2layers_8ant", "codebook_perfect_csi", "time_domain",
                                      "tdl_large_batch"])
def test_pusch_link_end_to_end(cuda_device, scenario):
    """BASELINE.json configs[4]: PUSCHTransmitter -> channel -> PUSCHReceiver recovers every transport block at high SNR
    (reference: test_pusch_receiver.py)."""
    import torch
    from sionna_b200.phy.nr import PUSCHConfig, PUSCHTransmitter, PUSCHReceiver
    from sionna_b200.phy.channel import AWGN
    from sionna_b200.phy.ofdm import OFDMDemodulator
    from sionna_b200.phy import config
    config.seed = 11
    g = torch.Generator(device="cpu").manual_seed(5)

    def crandn(*shape):
        return torch.complex(torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)).to(cuda_device) / np.sqrt(2)

    batch = 16
    if scenario == "siso_awgn":
        pc = PUSCHConfig()
        pc.carrier.n_size_grid = 8
        tx = PUSCHTransmitter(pc)
        rx = PUSCHReceiver(tx, return_tb_crc_status=True)
        x, b = tx(batch)
        y = AWGN()(x, 0.01)
        b_hat, crc = rx(y, 0.01)
    elif scenario == "time_domain":
        pc = PUSCHConfig()
        pc.carrier.n_size_grid = 6
        pc.tb.mcs_index = 5
        tx = PUSCHTransmitter(pc, output_domain="time")
        rx = PUSCHReceiver(tx, input_domain="time", l_min=0, return_tb_crc_status=True)
        x, b = tx(batch)                                                     # [batch, 1, 1, time samples]
        y = AWGN()(x, 0.002)
        b_hat, crc = rx(y, 0.002)
    elif scenario == "tdl_large_batch":
        # the configuration of tools/pusch_sim.py at a batch that exceeds one wave of every element-wise kernel
        from sionna_b200.phy.channel import ApplyOFDMChannel, TDL, subcarrier_frequencies, cir_to_ofdm_channel
        batch = 2048
        pc = PUSCHConfig(num_layers=2, num_antenna_ports=2)
        pc.carrier.n_size_grid = 16
        pc.dmrs.additional_position = 1
        tx = PUSCHTransmitter(pc)
        rx = PUSCHReceiver(tx, return_tb_crc_status=True)
        rg = tx.resource_grid
        x, b = tx(batch)
        a, tau = TDL("B", 100e-9, 3.5e9, num_rx_ant=8, num_tx_ant=2)(batch, 14, 1.0)
        h = cir_to_ofdm_channel(subcarrier_frequencies(rg.fft_size, rg.subcarrier_spacing), a, tau, normalize=True)
        y = ApplyOFDMChannel()(x, h, 0.002)
        b_hat, crc = rx(y, 0.002)
    elif scenario == "mimo_2x2layers_8ant":
        pcs = []
        for u in range(2):
            pc = PUSCHConfig(num_layers=2, num_antenna_ports=2)
            pc.carrier.n_size_grid = 6
            pc.dmrs.dmrs_port_set = [2 * u, 2 * u + 1]
            pc.dmrs.additional_position = 1
            pc.n_rnti = 10 + u
            pc.tb.mcs_index = 9
            pcs.append(pc)
        tx = PUSCHTransmitter(pcs)
        rx = PUSCHReceiver(tx, return_tb_crc_status=True)
        x, b = tx(batch)                                                     # [batch, 2, 2, 14, 72]
        h = crandn(batch, 1, 8, 2, 2, 1, 1)
        y = torch.einsum("bratpsf,btpsf->brasf", h.expand(-1, -1, -1, -1, -1, 14, 72), x)
        y = AWGN()(y, 0.001)
        b_hat, crc = rx(y, 0.001)
    else:
        pc = PUSCHConfig(num_layers=2, num_antenna_ports=4, precoding="codebook", tpmi=9)
        pc.carrier.n_size_grid = 6
        pc.tb.mcs_index = 9
        tx = PUSCHTransmitter(pc)
        rx = PUSCHReceiver(tx, channel_estimator="perfect", return_tb_crc_status=True)
        x, b = tx(batch)                                                     # [batch, 1, 4, 14, 72]
        h = crandn(batch, 1, 8, 1, 4, 1, 1).expand(-1, -1, -1, -1, -1, 14, 72).contiguous()
        y = torch.einsum("bratpsf,btpsf->brasf", h, x)
        y = AWGN()(y, 0.001)
        b_hat, crc = rx(y, 0.001, h)
    assert b_hat.shape == b.shape
    assert torch.equal(b_hat, b)
    assert bool(crc.all())


def test_oracle_pusch_ls_combine_closed_form():
    """oracle.nr.pusch_ls_combine on hand-made inputs: two ports of one CDM group with w_f = (+,+) / (+,-) are separated by
    the pairwise average; double-symbol DMRS additionally averages the two symbols; zero entries stay zero."""
    from oracle import nr as ON
    h0, h1 = 1.0 + 2.0j, -0.5 + 0.25j                        # the channels of port 0 and port 1
    # single-symbol DMRS, 2 CDM groups without data -> groups of n = 4 masked REs, ports occupy entries (0, 2) of each
    y = np.array([h0 + h1, 0, h0 - h1, 0, h0 + h1, 0, h0 - h1, 0])        # LS estimates seen through port 0's pilots
    out, err = ON.pusch_ls_combine(y[None], np.ones((1, 8)), 1, 1, 2)
    assert np.allclose(out[0], [h0, 0, h0, 0, h0, 0, h0, 0]) and np.allclose(err, 0.5)
    # port 1's pilots carry w_f = (+, -): the LS division turns the second entry into -(h0 - h1) ... = h1 - h0
    y1 = np.array([h0 + h1, 0, -(h0 - h1), 0])
    out1, _ = ON.pusch_ls_combine(y1[None], np.ones((1, 4)), 1, 1, 2)
    assert np.allclose(out1[0], [h1, 0, h1, 0])
    # double-symbol DMRS: two symbols of 4 REs each are averaged first, error variance / 4 in total
    ys = np.concatenate([y[:4] + 0.1, y[:4] - 0.1])
    ys[[1, 3, 5, 7]] = 0
    out2, err2 = ON.pusch_ls_combine(ys[None], np.ones((1, 8)), 2, 2, 2)
    assert np.allclose(out2[0], [h0, 0, h0, 0, h0, 0, h0, 0]) and np.allclose(err2, 0.25)


def test_precoded_dmrs_match_reference_vectors(gold):
    """PUSCHConfig.dmrs_grid_precoded for every TPMI of the six codebooks (TS 38.211 Tables 6.3.1.5-1..7) vs the
    reference's stored grids (test_pusch_config.py:169-230): pins the extracted codebook tables and the DMRS port mapping."""
    from sionna_b200.phy.nr import PUSCHConfig
    pc = PUSCHConfig()
    pc.carrier.n_size_grid = 1
    pc.carrier.slot_number = 1
    pc.dmrs.additional_position = 0
    pc.dmrs.config_type = 2
    pc.dmrs.num_cdm_groups_without_data = 3
    pc.dmrs.length = 2
    pc.dmrs.n_id = [8, 8]
    pc.precoding = "codebook"
    total = 0
    for layers, ports in ((1, 2), (1, 4), (2, 2), (2, 4), (3, 4), (4, 4)):
        if ports >= pc.num_layers:                               # keep num_layers <= num_antenna_ports at every step
            pc.num_antenna_ports = ports
            pc.num_layers = layers
        else:
            pc.num_layers = layers
            pc.num_antenna_ports = ports
        ref = gold[f"dmrs_precoded_{layers}_{ports}"]
        for tpmi in range(ref.shape[0]):
            pc.tpmi = tpmi
            assert np.allclose(pc.dmrs_grid_precoded / np.sqrt(3), ref[tpmi], atol=1e-6), (layers, ports, tpmi)
            total += 1
    assert total == 6 + 28 + 3 + 22 + 7 + 5


def test_layer_mapper_known_answers():
    """LayerMapper for 1..8 layers vs the reference's predefined sequences (test_layer_mapper.py:14-203): the oracle and
    the product's `call` (pure tensor reshapes, run here on CPU tensors); LayerDemapper inverts it."""
    import json
    import torch
    from oracle import nr as ON
    from sionna_b200.phy.nr import LayerMapper, LayerDemapper
    with open(os.path.join(os.path.dirname(__file__), "golden", "layer_mapper_golden.json")) as f:
        cases = json.load(f)
    assert [c["num_layers"] for c in cases] == list(range(1, 9))
    for c in cases:
        nl, want = c["num_layers"], np.array(c["out"])
        ins = [np.array(v) for v in c["inputs"]]
        got = ON.layer_map(ins[0] if nl <= 4 else ins, nl)
        assert np.array_equal(got, want), nl
        lm = LayerMapper(nl)
        t_in = torch.from_numpy(ins[0]).float() if nl <= 4 else [torch.from_numpy(v).float() for v in ins]
        y = lm.call(t_in)
        assert np.array_equal(y.numpy(), want), nl
        back = LayerDemapper(lm, 1).call(y)
        if nl <= 4:
            assert np.array_equal(back.numpy(), ins[0])
        else:
            assert np.array_equal(back[0].numpy(), ins[0]) and np.array_equal(back[1].numpy(), ins[1])


def _pusch_estimation_case(num_layers, length, additional_position, config_type, groups, rng, no=None):
    """One configuration of the reference's estimator sweep (test_channel_estimation.py:67-111) on the CPU: host
    configuration + oracle LS / CDM de-spreading / nearest-neighbour interpolation on a block-constant channel."""
    from sionna_b200.phy.nr import PUSCHConfig, PUSCHPilotPattern
    from oracle import ofdm as OO, nr as ON
    pc = PUSCHConfig(num_layers=num_layers, num_antenna_ports=num_layers)
    pc.n_size_bwp = 4
    pc.dmrs.length = length
    pc.dmrs.additional_position = additional_position
    pc.dmrs.config_type = config_type
    pc.dmrs.num_cdm_groups_without_data = groups
    pp = PUSCHPilotPattern(pc)
    mask, pilots = pp.mask.astype(bool), pp.pilots
    n_ant, batch = 3, 4
    chan = rng.standard_normal((batch, 1, n_ant, 1, num_layers)) + 1j * rng.standard_normal((batch, 1, n_ant, 1, num_layers))
    x = (rng.integers(0, 2, (batch, 1, num_layers, 14, 48)) * 2 - 1) / np.sqrt(2) + 0j      # data REs
    for l in range(num_layers):
        x[:, 0, l][:, mask[0, l]] = pilots[0, l]
    y = np.einsum("bratl,btlsf->brasf", chan, x)
    if no is not None:
        y = y + np.sqrt(no / 2) * (rng.standard_normal(y.shape) + 1j * rng.standard_normal(y.shape))
    n_sym = len(pc.dmrs_symbol_indices)
    h, e = OO.ls_estimate(y, mask, pilots, 0.0 if no is None else no)
    h, e = ON.pusch_ls_combine(h, e, n_sym, length, groups)
    h_hat, e_hat = OO.nn_interp(h, mask, pilots), OO.nn_interp(e, mask, pilots)
    want = np.broadcast_to(chan[..., None, None], h_hat.shape)
    return want, h_hat, e_hat


def test_pusch_estimator_sweep_block_constant_channel():
    """Noiseless: exact recovery for every DMRS configuration of the reference's sweep; AWGN: the empirical error
    variance equals the predicted one (test_channel_estimation.py:56-62, atol 1e-2)."""
    rng = np.random.default_rng(12)
    count = 0
    for num_layers in (1, 2, 4):
        for length in (1, 2):
            for additional_position in range(0, (3 if length == 1 else 1) + 1):
                for config_type in (1, 2):
                    for groups in range(1 if num_layers < 4 else 2, (2 if config_type == 1 else 3) + 1):
                        want, h_hat, _ = _pusch_estimation_case(num_layers, length, additional_position, config_type,
                                                                groups, rng)
                        assert np.allclose(want, h_hat, atol=1e-6), (num_layers, length, additional_position, config_type, groups)
                        count += 1
    assert count == 78
    for cfg in ((1, 1, 0, 1, 2), (2, 2, 1, 2, 3), (4, 1, 2, 1, 2)):
        want, h_hat, e_hat = _pusch_estimation_case(*cfg, rng, no=0.01)
        assert abs(np.var(want - h_hat) - e_hat.mean()) < 1e-2 and abs(np.var(want - h_hat) / e_hat.mean() - 1) < 0.2


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["2layers_8ant_lin", "1layer_4ant_double_symbol", "4layers_nn"])
def test_pusch_receiver_fused_front_end_equals_separate_blocks(cuda_device, cfg):
    """PUSCHReceiver's default front-end runs as ONE launch (PUSCHLSChannelEstimator incl. CDM de-spreading + linear
    interpolation + LMMSE + max-log demapping, ofdm/frontend.py). Switching the fusion off runs the three blocks of the
    reference's receiver one after the other: same LLRs to fp32 rounding, same decoded transport blocks."""
    import torch
    from sionna_b200.phy.nr import PUSCHConfig, PUSCHTransmitter, PUSCHReceiver
    from sionna_b200.phy.nr.pusch_channel_estimation import PUSCHLSChannelEstimator
    from sionna_b200.phy.channel import ApplyOFDMChannel, TDL, subcarrier_frequencies, cir_to_ofdm_channel
    from sionna_b200.phy import config
    config.seed = 77
    if cfg == "2layers_8ant_lin":
        pc, ant = PUSCHConfig(num_layers=2, num_antenna_ports=2), 8
        pc.dmrs.additional_position = 1
    elif cfg == "1layer_4ant_double_symbol":
        pc, ant = PUSCHConfig(), 4
        pc.dmrs.length = 2
        pc.dmrs.additional_position = 1
    else:
        pc, ant = PUSCHConfig(num_layers=4, num_antenna_ports=4), 16
        pc.dmrs.config_type = 2
        pc.dmrs.num_cdm_groups_without_data = 2
    pc.carrier.n_size_grid = 6
    pc.tb.mcs_index = 12
    tx = PUSCHTransmitter(pc)
    rx = PUSCHReceiver(tx)
    if cfg == "4layers_nn":
        est = PUSCHLSChannelEstimator(tx.resource_grid, tx._dmrs_length, tx._dmrs_additional_position,
                                      tx._num_cdm_groups_without_data, interpolation_type="nn")
        rx = PUSCHReceiver(tx)
        from sionna_b200.phy.ofdm.frontend import FusedLSLinearDetector
        rx._channel_estimator = est
        rx._fused = FusedLSLinearDetector(est, tx.resource_grid, rx._stream_management, "maxlog",
                                          constellation=rx._mimo_detector._constellation)
    assert rx._fused is not None
    rg = tx.resource_grid
    x, b = tx(32)
    a, tau = TDL("B", 100e-9, 3.5e9, min_speed=3.0, num_rx_ant=ant, num_tx_ant=pc.num_antenna_ports)(32, rg.num_ofdm_symbols,
                                                                                           1.0 / rg.ofdm_symbol_duration)
    h = cir_to_ofdm_channel(subcarrier_frequencies(rg.fft_size, rg.subcarrier_spacing), a, tau, normalize=True)
    no = 0.02
    y = ApplyOFDMChannel()(x, h, no)
    llr_f = rx._fused(y, no)
    h_hat, ev = rx._channel_estimator(y, no)
    llr_u = rx._mimo_detector(y, h_hat, ev, no)
    scale = float(llr_u.abs().max())
    assert llr_f.shape == llr_u.shape and float((llr_f - llr_u).abs().max()) <= 3e-4 * scale
    b_f = rx(y, no)
    rx.fuse_front_end = False
    b_u = rx(y, no)
    assert torch.equal(b_f, b_u) and float((b_f != b).float().mean()) < 1e-3
