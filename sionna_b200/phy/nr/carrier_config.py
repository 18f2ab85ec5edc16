"""CarrierConfig: OFDM numerology of TS 38.211 Sec. 4 (reference: src/sionna/phy/nr/carrier_config.py:11-277)."""
from .config import Config, Param, one_of

_SCS = (15, 30, 60, 120, 240, 480, 960)


def _slot_number(cfg, value):
    assert 0 <= value < cfg.num_slots_per_frame, "slot_number cannot exceed the number of slots per frame-1"
    return value


class CarrierConfig(Config):
    """CarrierConfig(**kwargs): n_cell_id (1), cyclic_prefix ("normal"), subcarrier_spacing (15, kHz), n_size_grid (4),
    n_start_grid (0), slot_number (0), frame_number (0) and the numerology derived from them."""
    _name = "Carrier Configuration"

    n_cell_id = Param(1, one_of(range(1008), "n_cell_id must be in the range from 0 to 1007"))
    cyclic_prefix = Param("normal", one_of(("normal", "extended"), "Invalid cyclic prefix"))
    subcarrier_spacing = Param(15, one_of(_SCS, "Invalid subcarrier spacing"))
    n_size_grid = Param(4, one_of(range(1, 276), "n_size_grid must be in the range from 1 to 275"))
    n_start_grid = Param(0, one_of(range(2200), "n_start_grid must be in the range from 0 to 2199"))
    slot_number = Param(0, _slot_number)
    frame_number = Param(0, one_of(range(1024), "frame_number must be in [0, 1023]"))

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.check_config()

    # ---- derived, read-only -----------------------------------------------------------------------------------
    @property
    def mu(self):
        return _SCS.index(self.subcarrier_spacing)

    @property
    def num_symbols_per_slot(self):
        return 14 if self.cyclic_prefix == "normal" else 12

    @property
    def num_slots_per_subframe(self):
        return 1 << self.mu

    @property
    def num_slots_per_frame(self):
        return 10 * self.num_slots_per_subframe

    frame_duration = property(lambda self: 10e-3)
    sub_frame_duration = property(lambda self: 1e-3)
    t_c = property(lambda self: 1 / (480e3 * 4096))
    t_s = property(lambda self: 1 / (15e3 * 2048))
    kappa = property(lambda self: 64.)

    @property
    def cyclic_prefix_length(self):
        """CP duration N_CP,l * T_c in seconds (TS 38.211 5.3.1); the long CP is applied when the slot starts a
        half-subframe (slot 0 or 7 * 2^mu), as the reference does (:246-258)."""
        if self.cyclic_prefix == "extended":
            n_cp = 512 * self.kappa / 2 ** self.mu
        else:
            n_cp = 144 * self.kappa / 2 ** self.mu
            if self.slot_number in (0, 7 * 2 ** self.mu):
                n_cp += 16 * self.kappa
        return n_cp * self.t_c

    def check_config(self):
        if self.cyclic_prefix == "extended":
            assert self.subcarrier_spacing == 60, "Extended cyclic prefix only valid for 60kHz subcarrier spacing"
        self._revalidate(("n_cell_id", "cyclic_prefix", "subcarrier_spacing", "n_size_grid", "slot_number",
                          "frame_number"))
