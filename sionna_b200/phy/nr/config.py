"""Base class of the NR configuration objects (reference: src/sionna/phy/nr/config.py:13-53).

Configurable properties are declared with the `Param` descriptor (default value + validator) instead of hand-written
property pairs; behaviour is the reference's: keyword arguments that name a property are applied in order, unknown
keywords are ignored, every assignment is validated (AssertionError / ValueError), `clone()` copies, `show()` prints.
"""
import copy
import numpy as np


class Param:
    """Validated attribute with a default. `check(obj, value)` returns the (possibly normalised) value or raises."""

    def __init__(self, default, check=None, doc=""):
        self.default, self.check, self.__doc__ = default, check, doc

    def __set_name__(self, owner, name):
        self.name, self.slot = name, "_" + name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        if self.slot not in obj.__dict__:
            obj.__dict__[self.slot] = copy.copy(self.default)
        return obj.__dict__[self.slot]

    def __set__(self, obj, value):
        obj.__dict__[self.slot] = self.check(obj, value) if self.check else value


def one_of(choices, msg):
    def check(_, value):
        assert value in choices, msg
        return value
    return check


class Config:
    _name = "Configuration"
    _hidden = ("show", "name", "check_config", "check_config_precoded", "clone", "c_init", "dmrs", "tb", "carrier")

    def __init__(self, **kwargs):
        for key, value in kwargs.items():
            if key in dir(self):
                setattr(self, key, value)

    @classmethod
    def _params(cls):
        return [n for n in dir(cls) if isinstance(getattr(cls, n, None), Param)]

    def _revalidate(self, names):
        for n in names:
            setattr(self, n, getattr(self, n))

    def clone(self, deep=True):
        return copy.deepcopy(self) if deep else copy.copy(self)

    def check_config(self):
        pass

    def show(self):
        self.check_config()
        print(self._name)
        print("=" * len(self._name))
        for a in dir(self):
            if a[0] == "_" or a in self._hidden:
                continue
            val = getattr(self, a)
            if callable(val) and not isinstance(val, (list, np.ndarray)):
                continue
            if a in ("dmrs_grid", "dmrs_grid_precoded", "dmrs_mask", "n"):
                print(f"{a} : shape {np.array(val).shape}")
            else:
                print(f"{a} : {val}")
        print("\r")
