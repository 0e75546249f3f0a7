"""Shared test plumbing.

``-m "not gpu"`` : oracle vs the reference's goldens, host logic, C-ABI surface (runs without a GPU).
``-m gpu``       : parity of the CUDA path against the oracle and the goldens, through the C ABI.
"""
import glob
import json
import os
import subprocess
import sys

import pytest

# The group tests run up to eight ranks of a sharded solve on ONE device, and their exchange kernels wait for each other:
# every rank's stream needs a hardware queue of its own (the default of 8 connections lets streams alias).  Read by the
# driver when the CUDA context is created, i.e. after this line.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "slow: long-running CPU oracle cases")


def _ensure_built():
    lib = os.path.join(ROOT, "stcsp_solver_b200", "libstcsp_b200.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    sem = os.path.join(ROOT, "oracle", "libsemantic.so")
    if not (os.path.exists(lib) and os.path.exists(orc) and os.path.exists(sem)):
        subprocess.check_call(["make", "-C", ROOT, "-j8"], stdout=subprocess.DEVNULL)


_ensure_built()


def load_goldens():
    out = {}
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.json"))):
        with open(p) as f:
            out[os.path.basename(p)[:-5]] = json.load(f)
    return out


GOLDENS = load_goldens()


def golden_text(g):
    """Model text of a golden case: probes carry it, benchmark instances are regenerated."""
    from stcsp_solver_b200 import instances
    return g["model"] if "model" in g else instances.by_name(g["name"])


def golden_flags(g):
    return [g["flags"]] if g.get("flags") else []
