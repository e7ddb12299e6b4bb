"""Import shim (test infrastructure only): lets the reference `bionc` package be imported in a
container without CasADi.  Every symbol raises if it is actually *called*; the numpy backend of
the reference only uses these names in type hints."""


def _missing(name):
    def _raise(*args, **kwargs):
        raise RuntimeError(f"casadi.{name} called, but casadi is only stubbed in this container")

    _raise.__name__ = name
    return _raise


class _Stub:
    def __init__(self, *args, **kwargs):
        raise RuntimeError(f"casadi.{type(self).__name__} instantiated, but casadi is only stubbed here")


class MX(_Stub):
    pass


class SX(_Stub):
    pass


class DM(_Stub):
    pass


class Function(_Stub):
    pass


for _n in (
    "nlpsol transpose horzcat vertcat solve reshape sum1 sum2 cross inv dot cos sin sqrt sumsqr "
    "jacobian norm_1 norm_2 exp evalf fabs fmax fmin if_else mtimes tril triu diag trace acos atan2 "
    "gradient hessian substitute qpsol rootfinder"
).split():
    globals()[_n] = _missing(_n)


def __getattr__(name):  # any other casadi name: hand out a raising callable
    if name.startswith("__"):
        raise AttributeError(name)
    return _missing(name)
