"""TBConfig: MCS selection of a transport block (reference: src/sionna/phy/nr/tb_config.py:12-409)."""
from .config import Config, Param, one_of
from .utils import decode_mcs_index


def _tb_n_id(_, value):
    if value is not None:
        assert value in range(1024), "n_id must be in [0, 1023]"
    return value


class TBConfig(Config):
    """TBConfig(**kwargs): mcs_index (14), mcs_table (1), channel_type ("PUSCH"), n_id (None = n_cell_id); derived:
    target_coderate, num_bits_per_symbol, tb_scaling (1.0)."""
    _name = "Transport Block Configuration"

    mcs_index = Param(14, one_of(range(29), "mcs_index must be in range from 0 to 28"))
    mcs_table = Param(1, one_of(range(1, 5), "mcs_table must be in range from 1 to 4"))
    channel_type = Param("PUSCH", one_of(("PUSCH", "PDSCH"), 'Only "PUSCH" and "PDSCH" are supported'))
    n_id = Param(None, _tb_n_id)

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.check_config()

    name = property(lambda self: "Transport Block Configuration")

    def _mcs(self):
        return decode_mcs_index(self.mcs_index, self.mcs_table, is_pusch=self.channel_type == "PUSCH")

    @property
    def target_coderate(self):
        return self._mcs()[1]

    @property
    def num_bits_per_symbol(self):
        return self._mcs()[0]

    tb_scaling = property(lambda self: 1.)

    def check_config(self):
        self._revalidate(("mcs_index", "mcs_table", "channel_type", "n_id"))
