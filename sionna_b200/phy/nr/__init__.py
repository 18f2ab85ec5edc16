"""5G NR transport-block chain (mirror of parts of sionna.phy.nr)."""
from .utils import calculate_tb_size
from .tb_encoder import TBEncoder
from .tb_decoder import TBDecoder
