"""Channel models and channel application (mirror of sionna.phy.channel, SURVEY.md section 8 rows a / f3)."""
from .awgn import AWGN
from .apply_ofdm_channel import ApplyOFDMChannel
from .tdl import TDL, cir_to_ofdm_channel, subcarrier_frequencies
from .time_channel import (time_lag_discrete_time_channel, cir_to_time_channel, time_to_ofdm_channel, ApplyTimeChannel, GenerateOFDMChannel,
                           OFDMChannel, GenerateTimeChannel, TimeChannel)
from .rayleigh_block_fading import RayleighBlockFading
