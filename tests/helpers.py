"""Shared helpers for the parity tests."""
import os

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(REPO, "tests", "golden")


def load_case(name):
    with np.load(os.path.join(GOLDEN, "cases", name + ".npz"), allow_pickle=False) as z:
        return {k: np.array(z[k]) for k in z.files}


def model_path(name):
    return os.path.join(GOLDEN, "models", name + ".npz")


def fext_list(case):
    """(segment, kind, param) triples as the oracle wants them."""
    return [(int(s), int(k), p) for s, k, p in zip(case["fext_seg"], case["fext_kind"], case["fext_param"])]


def rel_err(a, b):
    """max |a - b| / max(1, max|b|): relative on the vector's scale (lambda ~ 0 in free fall)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b))))) if a.size else 0.0


def nvrtc_available() -> bool:
    """The model-specialised kernels need libnvrtc at run time (opened with dlopen by the library)."""
    import ctypes

    for name in ("libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"):
        try:
            ctypes.CDLL(name)
            return True
        except OSError:
            continue
    return False
