"""The C-ABI shared library loads and exports every symbol include/microsound_b200.h declares.
No compute call is made here (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from audio_suite_b200 import _abi
from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "microsound_b200.h")


def _declared():
    text = open(HEADER).read()
    names = set(re.findall(r"^(?:int|size_t|const char\*|void|unsigned long long)\s+(ms_\w+)\(", text, re.M))
    macro = text[text.index("#define MS_DECLARE_API"):text.index("MS_DECLARE_API(_f32, float)")]
    staged = set(re.findall(r"(ms_\w+)##SFX", macro))
    assert len(staged) >= 17
    return sorted(names | {f"{n}_{p}" for n in staged for p in ("f32", "f64")})


def test_binding_covers_header():
    assert _declared() == _abi.exported_symbols()


def test_cuda_library_exports_every_symbol():
    if not os.path.isfile(_abi.LIB_PATH):
        pytest.fail(f"{_abi.LIB_PATH} missing: run __graft_entry__.build()")
    lib = C.CDLL(_abi.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    l2 = _abi.load_library(_abi.LIB_PATH)
    assert l2.ms_version() == 1 and l2.ms_is_cuda_build() == 1 and l2.ms_last_error() is not None


def test_struct_sizes_match_the_header():
    import subprocess, tempfile
    src = '#include "%s"\n#include <stdio.h>\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n",' \
          'sizeof(ms_band_edge),sizeof(ms_spec_op),sizeof(ms_spec_job),sizeof(ms_synth_evt),sizeof(ms_ola_render),' \
          'sizeof(ms_ola_evt),sizeof(ms_fir_render),sizeof(ms_post_render));}' % HEADER
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "s")], text=True).split()]
    want = [C.sizeof(t) for t in (_abi.BandEdge, _abi.SpecOp, _abi.SpecJob, _abi.SynthEvt, _abi.OlaRender,
                                  _abi.OlaEvt, _abi.FirRender, _abi.PostRender)]
    assert got == want


def test_product_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from audio_suite_b200 import engine, configs
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        engine.render(configs.canonical("C1b"))
