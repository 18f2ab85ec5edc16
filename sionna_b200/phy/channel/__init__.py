"""Channel models on the hot path (mirror of sionna.phy.channel): AWGN, frequency-domain channel application."""
from .awgn import AWGN
from .apply_ofdm_channel import ApplyOFDMChannel
from .tdl import TDL, cir_to_ofdm_channel, subcarrier_frequencies
