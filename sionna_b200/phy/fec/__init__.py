"""Forward error correction (mirror of sionna.phy.fec): LDPC codes and test utilities."""
from . import ldpc
from . import utils
