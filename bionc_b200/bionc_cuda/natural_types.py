"""Light-weight mirrors of the reference's array types for natural coordinates
(``bionc/bionc_numpy/natural_coordinates.py:11-83, :191-242``, ``natural_velocities.py``,
``natural_accelerations.py``): ndarray subclasses with the per-segment slot layout
``[u(0:3) rp(3:6) rd(6:9) w(9:12)]`` and ``v = rp - rd``, batch aware (the segment axis is the last one).

They are conveniences for authoring states; every ``bionc_cuda`` method equally accepts plain numpy
arrays, torch tensors and the reference's own ``NaturalCoordinates`` objects.
"""

from __future__ import annotations

import numpy as np


class _Segment12(np.ndarray):
    _names = ("u", "rp", "rd", "w")

    def __new__(cls, input_array):
        obj = np.asarray(input_array, dtype=np.float64).view(cls)
        if obj.shape[-1] != 12 and not (obj.ndim == 2 and obj.shape == (12, 1)):
            raise ValueError("a segment needs 12 natural coordinates")
        return obj

    @classmethod
    def from_components(cls, *parts, **named):
        """``from_components(u, rp, rd, w)`` (reference natural_coordinates.py:27-49); keyword names of the
        velocity / acceleration variants (``udot``, ``rpdot``, ... ) are accepted too."""
        if named:
            parts = tuple(named[k] for k in named)
        if len(parts) != 4:
            raise ValueError("four 3-vectors (u, rp, rd, w) are required")
        return cls(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 3) if np.ndim(p) > 1 else np.asarray(p, dtype=np.float64) for p in parts], axis=-1))

    def _slot(self, k):
        a = np.asarray(self)
        a = a[:, 0] if a.shape == (12, 1) else a
        return a[..., 3 * k : 3 * k + 3]

    u = property(lambda self: self._slot(0))
    rp = property(lambda self: self._slot(1))
    rd = property(lambda self: self._slot(2))
    w = property(lambda self: self._slot(3))

    @property
    def v(self):
        return self.rp - self.rd

    def to_uvw(self):
        return self.u, self.v, self.w

    def to_array(self):
        return np.asarray(self)


class SegmentNaturalCoordinates(_Segment12):
    pass


class SegmentNaturalVelocities(_Segment12):
    pass


class SegmentNaturalAccelerations(_Segment12):
    pass


class _Natural(np.ndarray):
    _segment_cls = _Segment12

    def __new__(cls, input_array):
        a = np.asarray(input_array, dtype=np.float64)
        if a.ndim == 1:
            a = a[:, np.newaxis]  # the reference stores one system as a column (n, 1)
        obj = a.view(cls)
        n = obj.shape[0] if (obj.ndim == 2 and obj.shape[1] == 1) else obj.shape[-1]
        if n % 12 != 0:
            raise ValueError("the number of natural coordinates must be a multiple of 12")
        return obj

    @classmethod
    def from_qi(cls, segments):
        """Stack per-segment coordinates (reference natural_coordinates.py:205-221)."""
        return cls(np.concatenate([np.asarray(s, dtype=np.float64).reshape(-1) for s in segments]))

    from_qdoti = from_qi
    from_qddoti = from_qi

    def _flat(self):
        a = np.asarray(self)
        return a[:, 0] if (a.ndim == 2 and a.shape[1] == 1) else a

    def nb_qi(self) -> int:
        return self._flat().shape[-1] // 12

    nb_qdoti = nb_qi
    nb_qddoti = nb_qi

    def vector(self, segment_idx: int):
        return self._segment_cls(self._flat()[..., 12 * segment_idx : 12 * segment_idx + 12])

    def u(self, i):
        return self.vector(i).u

    def rp(self, i):
        return self.vector(i).rp

    def rd(self, i):
        return self.vector(i).rd

    def w(self, i):
        return self.vector(i).w

    def v(self, i):
        return self.vector(i).v

    def to_array(self):
        return self._flat()


class NaturalCoordinates(_Natural):
    _segment_cls = SegmentNaturalCoordinates


class NaturalVelocities(_Natural):
    _segment_cls = SegmentNaturalVelocities


class NaturalAccelerations(_Natural):
    _segment_cls = SegmentNaturalAccelerations
