"""bionc_b200: a B200-native (sm_100a, FP64) backend for bioNC's constrained forward dynamics.

Layout mirrors the reference's backend convention (``bionc/bionc_numpy``, ``bionc/bionc_casadi``):
``bionc_b200.bionc_cuda`` exports the same names as ``bionc.bionc_numpy`` for the forward-dynamics
path, with a leading batch axis; ``bionc_b200.utils`` holds ``RK4`` / ``forward_integration``
(reference ``bionc/utils/ode_solver.py``).
"""

__version__ = "0.1.0"
