"""TEST INFRASTRUCTURE: a `device` for audio_suite_b200.engine backed by host memory and the
block-emulator build of the kernels (libms_emul.so).  Lets the CPU test-suite drive the real
engine + real kernel bodies end to end.  Never imported by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

from audio_suite_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libms_emul.so")


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB


class EmulDevice:
    def __init__(self):
        build()
        self.lib = _abi.load_library(LIB)

    def stream_ptr(self):
        return C.c_void_p(None)

    def empty(self, n, dtype):
        return np.zeros(max(1, int(n)), dtype=dtype)

    zeros = empty

    def upload(self, arr):
        a = np.ascontiguousarray(arr)
        return a.view(np.uint8).reshape(-1).copy() if a.size else np.zeros(1, np.uint8)

    def ptr(self, buf):
        return C.c_void_p(buf.ctypes.data)

    def download(self, buf, offset, count):
        return np.array(buf[offset:offset + count], copy=True)

    def synchronize(self):
        pass
