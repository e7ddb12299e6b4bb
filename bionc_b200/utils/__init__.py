"""``RK4`` / ``forward_integration`` (reference ``bionc/utils/ode_solver.py``), workload generators and
roofline counts.  ``RK4`` and ``forward_integration`` are imported lazily (they need torch)."""


def __getattr__(name):
    if name in ("RK4", "forward_integration"):
        from . import ode_solver

        return getattr(ode_solver, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
