"""PUSCHDMRSConfig: DMRS parameters of TS 38.211 Sec. 6.4.1.1 (reference: src/sionna/phy/nr/pusch_dmrs_config.py)."""
from collections.abc import Sequence
import numpy as np
from .config import Config, Param, one_of


def _port_set(_, value):
    if isinstance(value, (int, np.integer)):
        return [int(value)]
    if isinstance(value, Sequence) or isinstance(value, np.ndarray):
        return list(value)
    raise ValueError("dmrs_port_set must be an integer or list")


def _n_id(_, value):
    if value is None:
        return None
    if isinstance(value, (int, np.integer)):
        assert 0 <= value < 65536, "n_id must be in [0, 65535]"
        return [int(value), int(value)]
    assert len(value) == 2, "n_id must be either [] or a two-tuple"
    for e in value:
        assert 0 <= e < 65536, "Each element of n_id must be in [0, 65535]"
    return value


class PUSCHDMRSConfig(Config):
    """PUSCHDMRSConfig(**kwargs): config_type (1), type_a_position (2), additional_position (0), length (1),
    dmrs_port_set ([] = ports 0..num_layers-1), n_id (None = n_cell_id), n_scid (0), num_cdm_groups_without_data (2)."""
    _name = "PUSCH DMRS Configuration"

    config_type = Param(1, one_of((1, 2), "config_type must be in [1,2]"))
    type_a_position = Param(2, one_of((2, 3), "type_a_position must be in [2,3]"))
    additional_position = Param(0, one_of((0, 1, 2, 3), "additional_position must be in [0,1,2,3]"))
    length = Param(1, one_of((1, 2), "Invalid DMRS length"))
    dmrs_port_set = Param([], _port_set)
    n_id = Param(None, _n_id)
    n_scid = Param(0, one_of((0, 1), "n_scid must be 0 or 1"))
    num_cdm_groups_without_data = Param(2, one_of((1, 2, 3), "num_cdm_groups_without_data must be in [1,2,3]"))

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.check_config()

    # Tables 6.4.1.1.3-1 (type 1: 2 CDM groups, comb 2) and 6.4.1.1.3-2 (type 2: 3 CDM groups, pairs of subcarriers):
    # port p -> CDM group, frequency shift Delta, w_f(k'), w_t(l')
    def _port_table(self):
        groups = 2 if self.config_type == 1 else 3
        ports = 4 * groups                                    # with double-symbol DMRS
        p = np.arange(ports)
        lam = (p % (2 * groups)) // 2
        delta = lam if self.config_type == 1 else 2 * lam
        w_f = np.stack([np.ones(ports, int), 1 - 2 * (p % 2)])
        w_t = np.stack([np.ones(ports, int), 1 - 2 * (p // (2 * groups))])
        return lam, delta, w_f, w_t

    @property
    def allowed_dmrs_ports(self):
        """Nominal ports for config_type / length / num_cdm_groups_without_data (4, 6, 8 or 12 ports at most)."""
        groups = 2 if self.config_type == 1 else 3
        g = min(self.num_cdm_groups_without_data, groups)
        first = list(range(2 * g))
        if self.length == 1:
            return first
        return first + [q + 2 * groups for q in first]

    @property
    def cdm_groups(self):
        return [int(self._port_table()[0][p]) for p in self.dmrs_port_set]

    @property
    def deltas(self):
        return [int(self._port_table()[1][p]) for p in self.dmrs_port_set]

    @property
    def w_f(self):
        return self._port_table()[2][:, self.dmrs_port_set]

    @property
    def w_t(self):
        return self._port_table()[3][:, self.dmrs_port_set]

    @property
    def beta(self):
        """PUSCH-to-DMRS EPRE ratio, TS 38.214 Table 6.2.2-1."""
        n = self.num_cdm_groups_without_data
        if n == 3 and self.config_type != 2:
            return None
        return float(np.sqrt(n)) if n > 1 else 1.0

    def check_config(self):
        if self.length == 2:
            assert self.additional_position in (0, 1), "additional_position must be in [0, 1] for length==2"
        for p in self.dmrs_port_set:
            assert p in self.allowed_dmrs_ports, f"Unallowed DMRS port {p}. Not in {self.allowed_dmrs_ports}."
        if self.config_type == 1:
            assert self.num_cdm_groups_without_data in (1, 2), \
                "num_cdm_groups_without_data must be in [1,2] for config_type 1"
        self._revalidate(("config_type", "type_a_position", "additional_position", "length", "dmrs_port_set", "n_id",
                          "n_scid", "num_cdm_groups_without_data"))
