"""5G NR PUSCH link (mirror of sionna.phy.nr)."""
from .utils import calculate_tb_size, decode_mcs_index
from .config import Config
from .carrier_config import CarrierConfig
from .pusch_dmrs_config import PUSCHDMRSConfig
from .tb_config import TBConfig
from .pusch_config import PUSCHConfig, check_pusch_configs
from .pusch_pilot_pattern import PUSCHPilotPattern
from .layer_mapping import LayerMapper, LayerDemapper
from .pusch_precoder import PUSCHPrecoder
from .tb_encoder import TBEncoder
from .tb_decoder import TBDecoder
from .pusch_channel_estimation import PUSCHLSChannelEstimator
from .pusch_transmitter import PUSCHTransmitter
from .pusch_receiver import PUSCHReceiver
