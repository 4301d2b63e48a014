"""Import the reference's own pure-Python host code (util.py) behind a stub TensorFlow.

TEST INFRASTRUCTURE, build-container only: /root/reference is not present on the GPU box,
so nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` may call this.  TensorFlow 2.1 is
not installable here (no network); ``util.py`` needs the name ``tf`` only for an import
line and two type annotations (util.py:7, util.py:14, util.py:297, util.py:437), so a
MagicMock module is enough to run DataLoader / Sampler / Evaluator.results /
ExemplarGenerator.__init__ / ExemplarGenerator.herding unmodified.
"""
import importlib
import os
import sys
from unittest import mock

REFERENCE_DIR = os.environ.get("ADER_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "util.py"))


def load_reference_util():
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_DIR)
    for name in ("tensorflow", "tensorflow.compat", "tensorflow.compat.v1"):
        sys.modules.setdefault(name, mock.MagicMock())
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    return importlib.import_module("util")
