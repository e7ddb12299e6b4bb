"""Import shim (test infrastructure only)."""


class Markers:
    def __init__(self, *a, **k):
        raise RuntimeError("pyomeca is only stubbed in this container")

    @classmethod
    def from_c3d(cls, *a, **k):
        raise RuntimeError("pyomeca is only stubbed in this container")
