"""Import shim (test infrastructure only)."""


class Rotation:
    def __init__(self, *a, **k):
        raise RuntimeError("biorbd is only stubbed in this container")
