"""Minimal `dm_env` stand-in (TimeStep / StepType / specs) used when the real package is absent
(it is absent in this image).  The real `dm_env` is preferred when importable."""
import collections
import enum

import numpy as np

try:  # pragma: no cover - not installed here
    import dm_env as _real
    from dm_env import specs as _real_specs
    TimeStep, StepType = _real.TimeStep, _real.StepType
    Array, BoundedArray = _real_specs.Array, _real_specs.BoundedArray
    HAVE_DM_ENV = True
except Exception:
    HAVE_DM_ENV = False

    class StepType(enum.IntEnum):
        FIRST = 0
        MID = 1
        LAST = 2

    class TimeStep(collections.namedtuple('TimeStep', ['step_type', 'reward', 'discount', 'observation'])):
        __slots__ = ()

        def first(self):
            return np.all(self.step_type == StepType.FIRST)

        def mid(self):
            return np.all(self.step_type == StepType.MID)

        def last(self):
            return np.all(self.step_type == StepType.LAST)

    class Array:
        def __init__(self, shape, dtype, name=None):
            self.shape, self.dtype, self.name = tuple(shape), np.dtype(dtype), name

        def __repr__(self):
            return f'Array(shape={self.shape}, dtype={self.dtype}, name={self.name!r})'

    class BoundedArray(Array):
        def __init__(self, shape, dtype, minimum, maximum, name=None):
            super().__init__(shape, dtype, name)
            self.minimum = np.broadcast_to(np.asarray(minimum, dtype=dtype), shape).copy()
            self.maximum = np.broadcast_to(np.asarray(maximum, dtype=dtype), shape).copy()
