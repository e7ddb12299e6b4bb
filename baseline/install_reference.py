#!/usr/bin/env python
"""Install the UNMODIFIED reference (Ipuch/bioNC) into ``baseline/_ref`` so that it travels to the GPU box.

``baseline/_ref`` is git-ignored (never part of this repository's history) but not gpurun-ignored.  The bench's
reference arm (``bench.py --impl reference``), its ``cpu_baseline`` leg and the live-reference GPU tests import
``bionc`` from there; nothing under ``bionc_b200/`` does.

The contract's install command

    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
        --target baseline/_ref /root/reference

fails in this image: the reference's build backend is ``hatchling`` (pyproject.toml:1-3), which is neither installed nor
in the wheelhouse.  The reference is pure Python and its wheel holds exactly the ``bionc`` package directory
(pyproject.toml: ``[tool.hatch.build.targets.wheel] packages = ["bionc"]``), so the fallback below produces what the
wheel would: a copy of that directory.  Next to it go the data files the reference's own examples and tests load
(``.nmod`` / ``.nc`` model pickles) and the one example module that builds the knee model (config 4).

Third-party packages the reference imports at module level but that are absent here (casadi, ezc3d, pyomeca, trc,
biorbd -- none of them on the numpy forward-dynamics path) are satisfied by the import shims in ``oracle/stubs``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
DEST = os.path.join(HERE, "_ref")
DATA = [  # (path below the reference root, path below baseline/_ref/data)
    ("pendulum.nmod", "pendulum.nmod"),
    ("pendulum_with_force.nmod", "pendulum_with_force.nmod"),
    ("examples/forward_dynamics/pendulum.nmod", "examples/forward_dynamics/pendulum.nmod"),
    ("examples/forward_dynamics/double_universal_pendulum.nmod", "examples/forward_dynamics/double_universal_pendulum.nmod"),
    ("examples/models/lower_limb.nc", "examples/models/lower_limb.nc"),
    ("examples/knee_parallel_mechanism/knee_feikes.py", "examples/knee_parallel_mechanism/knee_feikes.py"),
]


def install(reference: str = "/root/reference", verbose: bool = True) -> str:
    """Returns how the install went: "pip", "copy" or "present" (already installed, reference tree not available)."""
    marker = os.path.join(DEST, "bionc", "__init__.py")
    if not os.path.isdir(reference):
        if os.path.exists(marker):
            return "present"
        raise RuntimeError(f"{reference} not found and baseline/_ref is empty")
    os.makedirs(DEST, exist_ok=True)
    how = "copy"
    pip = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", "/opt/wheelhouse",
           "--no-deps", "--upgrade", "--target", DEST]
    tmp = "/tmp/bionc_reference_copy"  # the build writes into the source tree and /root/reference is read-only
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(reference, tmp, ignore=shutil.ignore_patterns(".git", "docs"))
    res = subprocess.run(pip + [tmp], capture_output=True, text=True)
    if res.returncode == 0 and os.path.exists(marker):
        how = "pip"
    else:
        if verbose:
            print("pip install of the reference failed (%s); copying the package directory instead"
                  % (res.stderr.strip().splitlines() or ["?"])[-1])
        shutil.rmtree(os.path.join(DEST, "bionc"), ignore_errors=True)
        shutil.copytree(os.path.join(reference, "bionc"), os.path.join(DEST, "bionc"), ignore=shutil.ignore_patterns("__pycache__"))
    shutil.rmtree(tmp, ignore_errors=True)
    for src, dst in DATA:
        out = os.path.join(DEST, "data", dst)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(reference, src), out)
    with open(os.path.join(DEST, "INSTALLED.txt"), "w") as f:
        f.write(f"bioNC reference installed by baseline/install_reference.py via {how} from {reference}\n")
    return how


if __name__ == "__main__":
    print(install(*(sys.argv[1:2])))
