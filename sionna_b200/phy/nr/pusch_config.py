"""PUSCHConfig: slot-level configuration of a 5G NR PUSCH transmission (reference: src/sionna/phy/nr/pusch_config.py).

Derived quantities follow TS 38.211 Sec. 6.3.1.5 (codebooks), 6.4.1.1 (DMRS sequence, mapping and positions, Tables
6.4.1.1.3-3/-4) and TS 38.214 Sec. 6.1.4.2 (transport-block size). Not supported, as in the reference: frequency
hopping, transform precoding, DMRS lambda-bar != 0.
"""
import numpy as np
from .config import Config, Param, one_of
from .carrier_config import CarrierConfig
from .pusch_dmrs_config import PUSCHDMRSConfig
from .tb_config import TBConfig
from .utils import calculate_tb_size, nr_tables
from ..fec.scrambling import generate_prng_seq


def _opt_range(lo, hi, msg):
    def check(_, value):
        if value is not None:
            assert lo <= value < hi, msg
        return value
    return check


def _two(_, value):
    assert len(value) == 2, "symbol_allocation must have two elements"
    return value


def _is_bool(_, value):
    assert isinstance(value, bool), "transform_precoding must be bool"
    return value


# Additional DMRS positions l-bar (beyond l_0) of TS 38.211 Tables 6.4.1.1.3-3 (single-symbol) and -4 (double-symbol),
# without frequency hopping: {mapping type: {dmrs length: [(minimum l_d, [extra positions for additional_position
# = 1, 2, 3])]}}; the last row whose minimum duration is reached applies.
_EXTRA_POS = {
    "A": {1: [(8, [[7], [7], [7]]), (10, [[9], [6, 9], [6, 9]]), (12, [[9], [6, 9], [5, 8, 11]]),
              (13, [[11], [7, 11], [5, 8, 11]])],
          2: [(10, [[8]]), (13, [[10]])]},
    "B": {1: [(5, [[4], [4], [4]]), (8, [[6], [3, 6], [3, 6]]), (10, [[8], [4, 8], [3, 6, 9]]),
              (12, [[10], [5, 10], [3, 6, 9]])],
          2: [(8, [[5]]), (10, [[7]]), (12, [[9]])]},
}
# shortest duration that carries any DMRS at all
_MIN_DURATION = {("A", 1): 4, ("A", 2): 4, ("B", 1): 1, ("B", 2): 5}
_NUM_TPMI = {(1, 2): 6, (1, 4): 28, (2, 2): 3, (2, 4): 22, (3, 4): 7, (4, 4): 5}


class PUSCHConfig(Config):
    """PUSCHConfig(carrier_config=None, pusch_dmrs_config=None, tb_config=None, **kwargs)

    Settable: n_size_bwp (None = carrier.n_size_grid), n_start_bwp (0), num_layers (1), num_antenna_ports (1),
    mapping_type ("A"), symbol_allocation ([0, 14]), n_rnti (1), precoding ("non-codebook"), transform_precoding
    (False), tpmi (0), and the child configurations ``carrier``, ``dmrs``, ``tb``.
    """
    _name = "PUSCH Configuration"

    n_size_bwp = Param(None, _opt_range(1, 276, "n_size_bwp must be in the range from 1 to 275"))
    n_start_bwp = Param(0, one_of(range(2474), "n_start_bwp must be in the range from 0 to 2473"))
    num_layers = Param(1, one_of((1, 2, 3, 4), "num_layers must be in [1,...,4]"))
    num_antenna_ports = Param(1, one_of((1, 2, 4), "num_antenna_ports must be in [1,2,4]"))
    mapping_type = Param("A", one_of(("A", "B"), "mapping_type must be A or B"))
    symbol_allocation = Param([0, 14], _two)
    n_rnti = Param(1, _opt_range(0, 65536, "n_rnti must be in [0, 65535]"))
    precoding = Param("non-codebook", one_of(("codebook", "non-codebook"), "Unknown value for precoding"))
    transform_precoding = Param(False, _is_bool)
    tpmi = Param(0, one_of(range(28), "tpmi must be in [0,...,27]"))

    def __init__(self, carrier_config=None, pusch_dmrs_config=None, tb_config=None, **kwargs):
        super().__init__(**kwargs)
        self.carrier = carrier_config
        self.dmrs = pusch_dmrs_config
        self.tb = tb_config
        self.check_config()

    # ---- child configurations ----------------------------------------------------------------------------------
    @property
    def carrier(self):
        return self._carrier

    @carrier.setter
    def carrier(self, value):
        if value is None:
            value = CarrierConfig()
        assert isinstance(value, CarrierConfig), "carrier must be an instance of CarrierConfig"
        self._carrier = value

    @property
    def dmrs(self):
        return self._dmrs

    @dmrs.setter
    def dmrs(self, value):
        if value is None:
            value = PUSCHDMRSConfig()
        assert isinstance(value, PUSCHDMRSConfig), "pusch_dmrs_config must be an instance of PUSCHDMRSConfig"
        self._dmrs = value

    @property
    def tb(self):
        return self._tb

    @tb.setter
    def tb(self, value):
        if value is None:
            value = TBConfig(channel_type="PUSCH")
        assert isinstance(value, TBConfig), "tb must be an instance of TBConfig"
        assert value.channel_type == "PUSCH", 'TBConfig must be configured for "PUSCH"'
        self._tb = value

    # ---- derived: DMRS time positions --------------------------------------------------------------------------
    frequency_hopping = property(lambda self: "neither")

    @property
    def l_0(self):
        """First DMRS symbol relative to l_ref."""
        return self.dmrs.type_a_position if self.mapping_type == "A" else 0

    @property
    def l_d(self):
        return self.symbol_allocation[1]

    @property
    def l_ref(self):
        return 0 if self.mapping_type == "A" else self.symbol_allocation[0]

    @property
    def l_prime(self):
        return list(range(self.dmrs.length))

    @property
    def l_bar(self):
        """DMRS positions l-bar for the allocation length (Tables 6.4.1.1.3-3/-4)."""
        key = (self.mapping_type, self.dmrs.length)
        l_d = max(self.l_d, 3)                                # rows "< 4" share the first table row
        if l_d < _MIN_DURATION[key] and not (key == ("B", 1)):
            return []
        if key == ("A", 1) and l_d < 4:
            return []
        pos = [self.l_0]
        add = self.dmrs.additional_position
        if add > 0:
            extra = []
            for min_ld, per_add in _EXTRA_POS[self.mapping_type][self.dmrs.length]:
                if l_d >= min_ld:
                    extra = per_add[min(add, len(per_add)) - 1]
            pos += extra
        return pos

    @property
    def l(self):
        return [lb + lp for lb in self.l_bar for lp in self.l_prime]

    @property
    def n(self):
        per_prb = 3 if self.dmrs.config_type == 1 else 2
        return list(range(self.num_resource_blocks * per_prb))

    @property
    def dmrs_symbol_indices(self):
        return [v + self.l_ref for v in self.l]

    # ---- derived: sizes ----------------------------------------------------------------------------------------
    @property
    def num_resource_blocks(self):
        return self.carrier.n_size_grid if self.n_size_bwp is None else self.n_size_bwp

    @property
    def num_subcarriers(self):
        return 12 * self.num_resource_blocks

    @property
    def num_res_per_prb(self):
        """Data REs per PRB in the allocation."""
        n_dmrs = len(self.dmrs_symbol_indices)
        per_group = 6 if self.dmrs.config_type == 1 else 4
        free_in_dmrs_symbol = 12 - per_group * self.dmrs.num_cdm_groups_without_data
        return 12 * (self.symbol_allocation[1] - n_dmrs) + n_dmrs * free_in_dmrs_symbol

    def _cdm_group_subcarriers(self, group):
        """Subcarriers (within a PRB) of CDM group `group`."""
        if self.dmrs.config_type == 1:
            return np.arange(group, 12, 2)
        return np.array([0, 1, 6, 7]) + 2 * group

    @property
    def dmrs_mask(self):
        """bool [num_subcarriers, num_symbols_per_slot]: REs that carry no data (all CDM groups without data)."""
        mask = np.zeros([self.num_subcarriers, self.carrier.num_symbols_per_slot], bool)
        sc = np.concatenate([self._cdm_group_subcarriers(g) for g in range(self.dmrs.num_cdm_groups_without_data)])
        rows = (sc[None, :] + 12 * np.arange(self.num_resource_blocks)[:, None]).reshape(-1)
        for sym in self.dmrs_symbol_indices:
            mask[rows, sym] = True
        return mask

    def c_init(self, l):
        """Gold-sequence seed of DMRS symbol l (TS 38.211 6.4.1.1.1.1, lambda-bar = 0)."""
        n_scid = self.dmrs.n_scid
        n_id = self.carrier.n_cell_id if self.dmrs.n_id is None else self.dmrs.n_id[n_scid]
        sym = self.carrier.num_symbols_per_slot * self.carrier.slot_number + l + 1
        return int(((sym * (2 * n_id + 1) << 17) + 2 * n_id + n_scid) % (1 << 31))

    @property
    def dmrs_grid(self):
        """complex [num_dmrs_ports, num_subcarriers, num_symbols_per_slot]: unprecoded DMRS of every port,
        a(k, l) = beta * w_f(k') w_t(l') r(2n + k'), k = 4n + 2k' + Delta (type 1) or 6n + k' + Delta (type 2)."""
        self.check_config()
        ports = self.dmrs.dmrs_port_set if len(self.dmrs.dmrs_port_set) else list(range(self.num_layers))
        saved = self.dmrs.dmrs_port_set
        self.dmrs.dmrs_port_set = ports
        try:
            deltas, w_f, w_t, beta = self.dmrs.deltas, self.dmrs.w_f, self.dmrs.w_t, self.dmrs.beta
        finally:
            self.dmrs.dmrs_port_set = saved
        grid = np.zeros([len(ports), self.num_subcarriers, self.carrier.num_symbols_per_slot], complex)
        n = np.arange(len(self.n))
        stride, kp_step = (4, 2) if self.dmrs.config_type == 1 else (6, 1)
        for lb in self.l_bar:
            for lp in self.l_prime:
                c = generate_prng_seq(2 * self.num_subcarriers, self.c_init(lb + lp)).astype(float)
                r = ((1 - 2 * c[0::2]) + 1j * (1 - 2 * c[1::2])) / np.sqrt(2)
                for j in range(len(ports)):
                    for kp in (0, 1):
                        k = stride * n + kp_step * kp + deltas[j]
                        grid[j, k, self.l_ref + lb + lp] = r[2 * n + kp] * (w_f[kp][j] * w_t[lp][j])
        return beta * grid

    @property
    def precoding_matrix(self):
        """W [num_antenna_ports, num_layers] of TS 38.211 Tables 6.3.1.5-1..7 for the configured tpmi, or None."""
        if self.precoding == "non-codebook" or self.num_antenna_ports == 1:
            return None
        w = nr_tables().get(f"w_{self.num_layers}_{self.num_antenna_ports}")
        return None if w is None else w[self.tpmi]

    @property
    def dmrs_grid_precoded(self):
        if self.precoding == "non-codebook":
            return None
        return np.einsum("pl,lks->pks", self.precoding_matrix, self.dmrs_grid)

    num_ov = property(lambda self: 0)

    @property
    def num_coded_bits(self):
        n_re = (self.num_res_per_prb - self.num_ov) * self.num_resource_blocks
        return int(self.tb.tb_scaling * self.tb.num_bits_per_symbol * self.num_layers * n_re)

    @property
    def tb_size(self):
        n_re = min(156, self.num_res_per_prb - self.num_ov) * self.num_resource_blocks
        target = int(self.tb.target_coderate * self.tb.tb_scaling * n_re * self.tb.num_bits_per_symbol
                     * self.num_layers)
        return int(calculate_tb_size(target_tb_size=target, num_coded_bits=self.num_coded_bits,
                                     target_coderate=self.tb.target_coderate,
                                     modulation_order=self.tb.num_bits_per_symbol, verbose=False)[0])

    def show(self):
        self.carrier.show()
        Config.show(self)
        self.dmrs.show()
        self.tb.show()

    def check_config(self):
        self.carrier.check_config()
        self.dmrs.check_config()
        start, length = self.symbol_allocation
        if self.precoding == "codebook":
            if len(self.dmrs.dmrs_port_set) > 0:
                assert len(self.dmrs.dmrs_port_set) == self.num_layers, \
                    "num_layers must be equal to the number of dmrs ports"
            assert self.num_layers <= self.num_antenna_ports, "num_layers must be <= num_antenna_ports"
            assert self.num_antenna_ports >= 2, "precoding requires two or more antenna ports"
        else:
            assert self.num_layers == self.num_antenna_ports, "num_layers must be == num_antenna_ports"
        if self.dmrs.length == 1:
            if self.mapping_type == "A":
                assert length >= 4, "Symbol allocation is too short"
        else:
            assert self.dmrs.additional_position < 2, "dmrs.additional_position must be <2 for this dmrs.length"
            assert length >= 4, "Symbol allocation too short"
            if self.mapping_type == "B":
                assert length >= 5, "Symbol allocation is too short"
        if self.mapping_type == "A" and self.dmrs.additional_position == 3:
            assert self.dmrs.type_a_position == 2, "additional_position=3 only allowed for type_a_position=2"
        count = _NUM_TPMI.get((self.num_layers, self.num_antenna_ports))
        if count is not None:
            assert self.tpmi in range(count), f"tpmi must be in [0,...,{count - 1}]"
        max_length = self.carrier.num_symbols_per_slot
        if self.mapping_type == "A":
            assert start == 0, "symbol_allocation[0] must be 0 for mapping_type A"
            assert 4 <= length <= max_length, "symbol_allocation[1] must be in [4, 14 (or 12)]"
        else:
            assert 0 <= start <= 13, "symbol_allocation[0] must be in [0,13] for mapping_type B"
            assert 1 <= length <= max_length, "symbol_allocation[1] must be in [1, 14 (or 12)]"
            if self.dmrs.length == 2:
                assert length >= 5, "symbol_allocation[1] must be >=5 for dmrs.length==2"
        assert start + length <= max_length, "symbol_allocation[0]+symbol_allocation[1] must be < 14 (or 12)"
        self._revalidate(("n_size_bwp", "n_start_bwp", "num_layers", "mapping_type", "symbol_allocation", "n_rnti",
                          "precoding", "transform_precoding", "tpmi"))
        assert self.tb.channel_type == "PUSCH", 'TB_config must be configured for "PUSCH" transmission.'
        if len(self.dmrs.dmrs_port_set) > 0:
            assert self.num_layers == len(self.dmrs.dmrs_port_set), "num_layers must equal the number of DMRS ports"
        return True


def check_pusch_configs(pusch_configs):
    """Validates a list of PUSCHConfig (one per transmitter) and returns the parameters shared by the transmitter and
    the receiver (reference: pusch_config.py:1012-1065)."""
    assert isinstance(pusch_configs, list), "pusch_configs must be a Sequence of instances of PUSCHConfig"
    for pc in pusch_configs:
        assert isinstance(pc, PUSCHConfig), "All elements of pusch_configs must be instances of PUSCHConfig"
        pc.check_config()
    pc = pusch_configs[0]
    scs = pc.carrier.subcarrier_spacing * 1e3
    params = {
        "num_bits_per_symbol": pc.tb.num_bits_per_symbol, "num_tx": len(pusch_configs), "num_layers": pc.num_layers,
        "num_subcarriers": pc.num_subcarriers, "num_ofdm_symbols": pc.symbol_allocation[1],
        "subcarrier_spacing": scs, "num_antenna_ports": pc.num_antenna_ports, "precoding": pc.precoding,
        "precoding_matrices": [], "pusch_config": pc, "carrier_config": pc.carrier,
        "num_coded_bits": pc.num_coded_bits, "target_coderate": pc.tb.target_coderate, "n_id": [], "n_rnti": [],
        "tb_size": pc.tb_size, "dmrs_length": pc.dmrs.length,
        "dmrs_additional_position": pc.dmrs.additional_position,
        "num_cdm_groups_without_data": pc.dmrs.num_cdm_groups_without_data,
    }
    params["bandwidth"] = params["num_subcarriers"] * scs
    params["cyclic_prefix_length"] = int(np.ceil(pc.carrier.cyclic_prefix_length * params["bandwidth"]))
    for c in pusch_configs:
        if params["precoding"] == "codebook":
            params["precoding_matrices"].append(c.precoding_matrix)
        params["n_id"].append(c.carrier.n_cell_id if c.tb.n_id is None else c.tb.n_id)
        params["n_rnti"].append(c.n_rnti)
    return params
