"""TEST INFRASTRUCTURE ONLY -- headless loader for the *unmodified* reference file.

Imports /root/reference/microsound_0.2.1/main_v2.py with stub modules standing in
for the GUI / file-IO dependencies it names at import time (soundfile, PyQt6,
pyqtgraph).  Nothing from the reference is copied: the file
is executed where it lies.  /root/reference only exists in the build container,
so everything that must run on the GPU box uses `oracle/microsound_np.py` (the
numpy restatement, verified against this loader in tests/test_oracle_vs_reference.py)
and the committed fixtures under tests/golden/.

Only tests/, oracle/make_golden.py and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import wave
from unittest.mock import MagicMock

import numpy as np

REFERENCE_ROOT = os.environ.get("MS_REFERENCE_ROOT", "/root/reference")
REFERENCE_FILE = os.path.join(REFERENCE_ROOT, "microsound_0.2.1", "main_v2.py")

_cached = None


def available() -> bool:
    return os.path.isfile(REFERENCE_FILE)


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return MagicMock()


def load():
    """Return the reference module object (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise FileNotFoundError(f"reference not present at {REFERENCE_FILE}")
    saved = {}
    names = ["soundfile", "PyQt6", "PyQt6.QtCore", "PyQt6.QtWidgets", "pyqtgraph"]
    for name in names:
        saved[name] = sys.modules.get(name)
        sys.modules[name] = _Stub(name)
    qc, qw = sys.modules["PyQt6.QtCore"], sys.modules["PyQt6.QtWidgets"]
    qc.QObject = type("QObject", (), {})
    qw.QMainWindow = type("QMainWindow", (), {})
    qc.pyqtSignal = lambda *a, **k: None
    qc.pyqtSlot = lambda *a, **k: (lambda f: f)
    sys.modules["PyQt6"].QtCore, sys.modules["PyQt6"].QtWidgets = qc, qw
    try:
        spec = importlib.util.spec_from_file_location("_ms_reference_main_v2", REFERENCE_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for name in names:
            if saved[name] is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = saved[name]
    _cached = mod
    return mod


def load_ir_wav(path: str) -> np.ndarray:
    """What on_load_ir does (main_v2.py:1401-1413) with the stdlib wave reader:
    int16 -> float64, mean over channels, normalise to 0.9 peak."""
    with wave.open(path, "rb") as w:
        ch, sw, nfr = w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(nfr)
    assert sw == 2, "shipped IRs are 16-bit"
    a = np.frombuffer(raw, dtype="<i2").astype(np.float64) / 32768.0
    if ch > 1:
        a = a.reshape(-1, ch).mean(axis=1)
    m = float(np.max(np.abs(a))) if a.size else 0.0
    return a if m <= 0 else a * (0.9 / m)
