"""sionna_b200 -- B200-native (sm_100a) implementation of the Sionna PHY link-level hot path.

The public surface mirrors ``sionna.phy`` (``/root/reference/src/sionna/phy``): the same
``Block.__call__ -> build(shapes) -> call()`` protocol, constructor arguments, tensor shapes and
sign conventions, with ``torch.Tensor`` (CUDA) as the device-memory container and every kernel a
hand-written CUDA kernel reached through the C-ABI declared in ``include/sionna_b200.h``.
There is no CPU fallback: calling a block without a CUDA device / without the built
``libsionna_b200.so`` raises.
"""
__version__ = "0.1.0"
from . import phy  # noqa: F401
