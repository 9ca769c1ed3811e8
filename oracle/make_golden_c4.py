"""Writes tests/golden/c4_full.npz: the FULL long-form render (BASELINE.json configs[3], "C4": 96 kHz, 600 s,
x500 unfold clipped to 30 MHz, 1222 events of 300000 samples, x2.5 stretch, reflection cloud, 10 s IR -> 8192
taps, stereo diffusion) rendered by the UNMODIFIED reference in the build container (about four minutes and
5 GB on one core), reduced to what travels: every 997th frame (57773 frames), three contiguous windows, per-
channel sums and the peak.  997 is prime, so the decimated set visits every position modulo the FFT tile sizes.

Run:  python oracle/make_golden_c4.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader  # noqa: E402
from audio_suite_b200 import configs  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "c4_full.npz")
STEP = 997
WINDOWS = ((0, 4096), (28_800_000, 4096), (57_600_000 - 4096, 4096))


def main():
    ref = ref_loader.load()
    t = time.perf_counter()
    audio, meta = ref.render(configs.canonical("C4"))
    dt = time.perf_counter() - t
    print("reference C4 render: %.1f s, shape %s" % (dt, audio.shape))
    out = dict(numpy_version=np.array(np.__version__), step=np.array(STEP), decimated=audio[::STEP].astype(np.float32),
               decimated_f64=audio[::STEP], sums=audio.sum(axis=0), abs_sums=np.abs(audio).sum(axis=0),
               peak=np.array(np.max(np.abs(audio))), seconds_reference_1core=np.array(dt),
               design_sr_base=np.array(meta["design_sr_base"]), out_sr=np.array(meta["out_sr"]),
               grain_last=meta["grain_last"][::37].astype(np.float32))
    for k, (a, n) in enumerate(WINDOWS):
        out["win%d" % k] = audio[a:a + n]
        out["win%d_at" % k] = np.array(a)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
