"""Channel models on the hot path (mirror of sionna.phy.channel): AWGN."""
from .awgn import AWGN
