"""Import shim (test infrastructure only)."""


class TRCData:
    def __init__(self, *a, **k):
        raise RuntimeError("trc is only stubbed in this container")
