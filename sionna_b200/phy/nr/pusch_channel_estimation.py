"""PUSCHLSChannelEstimator (reference: src/sionna/phy/nr/pusch_channel_estimation.py:13-169)."""
from ..ofdm.channel_estimation import LSChannelEstimator
from ..._lib import lib, check, ptr, current_stream


class PUSCHLSChannelEstimator(LSChannelEstimator):
    """PUSCHLSChannelEstimator(resource_grid, dmrs_length, dmrs_additional_position, num_cdm_groups_without_data, interpolation_type="nn", interpolator=None, precision=None)

    LS estimation at the DMRS REs followed by the separation of the ports that share a CDM group: averaging over the
    two symbols of a double-symbol DMRS and over pairs of adjacent DMRS subcarriers (kernel ``sb_pusch_ls_combine``),
    then interpolation as in `LSChannelEstimator`. Inputs / outputs as `LSChannelEstimator`."""

    def __init__(self, resource_grid, dmrs_length, dmrs_additional_position, num_cdm_groups_without_data,
                 interpolation_type="nn", interpolator=None, precision=None, **kwargs):
        super().__init__(resource_grid, interpolation_type, interpolator, precision=precision, **kwargs)
        self._dmrs_length = int(dmrs_length)
        self._dmrs_additional_position = int(dmrs_additional_position)
        self._num_cdm_groups_without_data = int(num_cdm_groups_without_data)
        # as the reference (:109-113); the pilot pattern may hold fewer DMRS symbols than this nominal count when the
        # allocation is short, in which case the constructor arguments must describe the actual pattern
        self._num_dmrs_syms = self._dmrs_length * (self._dmrs_additional_position + 1)
        self._num_pilots_per_dmrs_sym = int(self._pilot_pattern.pilots.shape[-1] / self._num_dmrs_syms)

    def estimate_at_pilot_locations(self, y_eff_flat, no):
        h, err = super().estimate_at_pilot_locations(y_eff_flat, no)          # y / p, no / |p|^2, 0 where p == 0
        rows, p = h.shape[0] * h.shape[1], h.shape[2]
        check(lib().sb_pusch_ls_combine(ptr(h), ptr(err), rows, p, self._num_pilots_per_dmrs_sym, self._dmrs_length,
                                        2 * self._num_cdm_groups_without_data, current_stream()), "sb_pusch_ls_combine")
        return h, err
