"""audio_suite_b200 -- B200-native drop-in for the Microsound offline render path of
maetyu-d/audio-suite (`microsound_0.2.1/main_v2.py: render`).

    from audio_suite_b200 import render, render_batch, FACTORY_DEFAULTS

`render(params, progress=None) -> (float64[out_n, 2], meta)` keeps the reference signature.
All compute runs in hand-written sm_100a CUDA kernels behind a C ABI (include/microsound_b200.h);
there is no CPU fallback.
"""
from .configs import FACTORY_DEFAULTS, with_defaults, canonical, c5_params, synth_ir  # noqa: F401


def __getattr__(name):
    if name in ("render", "render_batch", "BatchRenderer"):
        from . import engine
        return getattr(engine, name)
    if name in ("load_preset", "batch_render", "batch_name", "write_wav_float32"):
        from . import frontend
        return getattr(frontend, name)
    raise AttributeError(name)
