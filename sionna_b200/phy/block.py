"""``Object`` / ``Block`` call protocol (mirror of /root/reference/src/sionna/phy/block.py:13-155).

``Block.__call__`` casts every floating / complex tensor or ndarray argument to the block's
precision (ints and Python scalars are left alone, block.py:122-131), moves it to the CUDA
device, runs ``build(*arg_shapes, **kwarg_shapes)`` exactly once (block.py:144-155) and then
``call``. Tensors are ``torch.Tensor`` on the device; they are containers for device memory that
the C-ABI kernels read and write through ``data_ptr()``.
"""
from abc import ABC, abstractmethod
import numpy as np
import torch
from .config import config, dtypes


class Object(ABC):
    """Base class carrying the precision (block.py:13-80)."""

    def __init__(self, *args, precision=None, **kwargs):
        if precision is None:
            self._precision = config.precision
        elif precision in ["single", "double"]:
            self._precision = precision
        else:
            raise ValueError("'precision' must be 'single' or 'double'")

    @property
    def precision(self):
        return self._precision

    @property
    def cdtype(self):
        return dtypes[self.precision]["torch"]["cdtype"]

    @property
    def rdtype(self):
        return dtypes[self.precision]["torch"]["rdtype"]

    @property
    def device(self):
        return config.device

    def _cast_or_check_precision(self, v):
        """Cast a tensor / array / scalar to the block precision on the device (block.py:54-80)."""
        if not isinstance(v, torch.Tensor):
            v = torch.as_tensor(np.asarray(v))
        if v.dtype.is_complex:
            v = v.to(device=self.device, dtype=self.cdtype)
        else:
            v = v.to(device=self.device, dtype=self.rdtype)
        return v


def _map_structure(fn, obj):
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map_structure(fn, o) for o in obj)
    if isinstance(obj, dict):
        return {k: _map_structure(fn, o) for k, o in obj.items()}
    return fn(obj)


class Block(Object):
    """Processing block with one-time ``build`` (block.py:82-155)."""

    def __init__(self, *args, precision=None, **kwargs):
        super().__init__(precision=precision, **kwargs)
        self._built = False

    @property
    def built(self):
        return self._built

    def build(self, *arg_shapes, **kwarg_shapes):
        pass

    @abstractmethod
    def call(self, *args, **kwargs):
        raise NotImplementedError("Subclasses must implement this method.")

    def _convert_to_tensor(self, v):
        if isinstance(v, np.ndarray):
            v = torch.from_numpy(np.ascontiguousarray(v))
        if isinstance(v, torch.Tensor):
            if v.dtype.is_floating_point:
                v = v.to(device=self.device, dtype=self.rdtype)
            elif v.dtype.is_complex:
                v = v.to(device=self.device, dtype=self.cdtype)
            else:
                v = v.to(device=self.device)
        return v

    @staticmethod
    def _get_shape(v):
        if hasattr(v, "shape"):
            return tuple(v.shape)
        if isinstance(v, (list, tuple)):
            if any(hasattr(e, "shape") and not isinstance(e, np.ndarray) for e in v):
                return [Block._get_shape(e) for e in v]       # list of tensors: list of shapes
            try:
                return tuple(np.shape(v))
            except ValueError:
                return ()
        return ()

    def __call__(self, *args, **kwargs):
        args, kwargs = _map_structure(self._convert_to_tensor, [list(args), kwargs])
        if not self._built:
            shapes = [[self._get_shape(a) for a in args], {k: self._get_shape(v) for k, v in kwargs.items()}]
            self.build(*shapes[0], **shapes[1])
            self._built = True
        return self.call(*args, **kwargs)
