"""Import shim (test infrastructure only)."""


class c3d:
    def __init__(self, *a, **k):
        raise RuntimeError("ezc3d is only stubbed in this container")
