"""LDPC encoding / decoding (mirror of sionna.phy.fec.ldpc)."""
from .encoding import LDPC5GEncoder
from .decoding import LDPCBPDecoder, LDPC5GDecoder
from .utils import RaggedMessages, EXITCallback, DecoderStatisticsCallback, WeightedBPCallback
