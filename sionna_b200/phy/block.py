"""``Object`` / ``Block`` call protocol (mirror of /root/reference/src/sionna/phy/block.py:13-155).

``Block.__call__`` casts every floating / complex tensor or ndarray argument to the block's
precision (ints and Python scalars are left alone, block.py:122-131), moves it to the CUDA
device, runs ``build(*arg_shapes, **kwarg_shapes)`` exactly once (block.py:144-155) and then
``call``. Tensors are ``torch.Tensor`` on the device; they are containers for device memory that
the C-ABI kernels read and write through ``data_ptr()``.
"""
from abc import ABC, abstractmethod
import warnings
import numpy as np
import torch
from .config import config, dtypes


class PrecisionWarning(UserWarning):
    """A block configured with precision="double" ran on the single-precision kernels."""


_warned_double = set()


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

    # precision="double": every kernel of this package computes in fp32 / complex64. A block set to double precision
    # keeps the reference's I/O contract (float64 / complex128 in and out, block.py:25-31, 122-131) but FALLS BACK to the
    # single-precision kernels for the arithmetic and says so once per block class (PrecisionWarning); pure data-movement
    # blocks (resource-grid gathers) stay exact. Results therefore carry fp32 accuracy.
    def _call_double_fallback(self, args, kwargs):
        cls = type(self).__name__
        if cls not in _warned_double:
            _warned_double.add(cls)
            warnings.warn(f"{cls}: precision='double' falls back to the single-precision (fp32 / complex64) kernels; "
                          "inputs and outputs are float64 / complex128, the arithmetic is not.", PrecisionWarning,
                          stacklevel=3)

        def narrow(v):
            if isinstance(v, torch.Tensor):
                if v.dtype == torch.float64:
                    return v.to(torch.float32)
                if v.dtype == torch.complex128:
                    return v.to(torch.complex64)
            return v

        def widen(v):
            if isinstance(v, torch.Tensor):
                if v.dtype == torch.float32:
                    return v.to(torch.float64)
                if v.dtype == torch.complex64:
                    return v.to(torch.complex128)
            return v
        args, kwargs = _map_structure(narrow, [list(args), kwargs])
        self._precision = "single"
        try:
            out = self.call(*args, **kwargs)
        finally:
            self._precision = "double"
        return _map_structure(widen, out)

    def __call__(self, *args, **kwargs):
        args, kwargs = _map_structure(self._convert_to_tensor, [list(args), kwargs])
        if not self._built:
            shapes = [[self._get_shape(a) for a in args], {k: self._get_shape(v) for k, v in kwargs.items()}]
            self.build(*shapes[0], **shapes[1])
            self._built = True
        return self._invoke(*args, **kwargs)

    def _invoke(self, *args, **kwargs):
        """``call`` with the double-precision fallback; blocks that override ``__call__`` go through here as well."""
        if self._precision == "double" and not getattr(self, "_native_double", False):
            return self._call_double_fallback(list(args), kwargs)
        return self.call(*args, **kwargs)


def fallback_to_single(name, precision):
    """For free functions with a ``precision`` argument: True if the caller asked for double precision (the result is
    then computed by the single-precision kernel and widened; warns once per function)."""
    if (precision or config.precision) != "double":
        return False
    if name not in _warned_double:
        _warned_double.add(name)
        warnings.warn(f"{name}: precision='double' falls back to the single-precision (fp32 / complex64) kernel.",
                      PrecisionWarning, stacklevel=3)
    return True
