"""``sionna_b200.phy`` -- mirror of ``sionna.phy`` for the link-level hot path."""
from .config import config, dtypes
from .block import Block, Object
from . import utils
from . import mapping
from . import channel
from . import fec
from . import mimo
from . import ofdm
from . import signal
