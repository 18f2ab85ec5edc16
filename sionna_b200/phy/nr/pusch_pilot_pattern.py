"""PUSCHPilotPattern: DMRS of a set of PUSCH transmitters as an OFDM pilot pattern
(reference: src/sionna/phy/nr/pusch_pilot_pattern.py:13-94)."""
import warnings
from collections.abc import Sequence
import numpy as np
from ..ofdm.pilot_pattern import PilotPattern
from .pusch_config import PUSCHConfig


class PUSCHPilotPattern(PilotPattern):
    """PUSCHPilotPattern(pusch_configs, precision=None): one PUSCHConfig per transmitter. The mask covers every RE of
    the CDM groups without data on the DMRS symbols of the allocation; the pilots of stream j are DMRS port j's values
    on those REs (zero on the REs of the other CDM groups)."""

    def __init__(self, pusch_configs, precision=None):
        if isinstance(pusch_configs, PUSCHConfig):
            pusch_configs = [pusch_configs]
        elif isinstance(pusch_configs, Sequence):
            for c in pusch_configs:
                assert isinstance(c, PUSCHConfig), "Each element of pusch_configs must be a valide PUSCHConfig"
        else:
            raise ValueError("Invalid value for pusch_configs")
        first = pusch_configs[0]
        num_streams, num_sc, num_sym = first.num_layers, first.num_subcarriers, first.l_d
        num_pilots = int(np.sum(first.dmrs_mask))
        seen = []
        masks, pilots = [], []
        for cfg in pusch_configs:
            assert cfg.num_layers == num_streams, "All pusch_configs must have the same number of layers"
            assert cfg.num_subcarriers == num_sc, "All pusch_configs must have the same number of subcarriers"
            assert cfg.l_d == num_sym, "All pusch_configs must have the same number of OFDM symbols"
            assert cfg.precoding == first.precoding, "All pusch_configs must have a the same precoding method"
            dmrs_mask = cfg.dmrs_mask
            assert int(np.sum(dmrs_mask)) == num_pilots, "All pusch_configs must have a the same number of masked REs"
            for port in cfg.dmrs.dmrs_port_set:
                if port in seen:
                    warnings.warn(f"DMRS port {port} used by multiple transmitters")
            seen += list(cfg.dmrs.dmrs_port_set)
            start = cfg.symbol_allocation[0]
            m = dmrs_mask[:, start:start + num_sym].T              # [symbols, subcarriers]
            grid = cfg.dmrs_grid[:, :, start:start + num_sym]       # [ports, subcarriers, symbols]
            masks.append(np.broadcast_to(m, (num_streams,) + m.shape))
            pilots.append(np.stack([grid[j].T[m] for j in range(num_streams)]))
        super().__init__(np.stack(masks), np.stack(pilots), normalize=False, precision=precision)
