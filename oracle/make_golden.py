"""Writes tests/golden/*.npz from the UNMODIFIED reference (oracle/ref_loader.py) in the build container.

The reference ships no tests or golden vectors, so these fixtures are what pins
the oracle (oracle/microsound_np.py) and, through it, the CUDA path on machines where
/root/reference does not exist.  numpy's Generator streams are stable per numpy version only; the
version is recorded in every file.

Run:  python oracle/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader  # noqa: E402
from audio_suite_b200 import configs  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    ver = np.array(np.__version__)

    # --- end-to-end renders (full for the small one, every 8th frame for the rest)
    renders = {}
    for name in ("C1b", "C1", "C2", "C3"):
        audio, meta = ref.render(configs.canonical(name))
        step = 1 if name == "C1b" else 8
        renders[name + "_audio"] = audio[::step]
        renders[name + "_step"] = np.array(step)
        renders[name + "_grain_last"] = meta["grain_last"][::step]
        renders[name + "_micro_last"] = meta["micro_last"][::step]
    for i in (0, 3, 5, 7, 11):
        audio, meta = ref.render(configs.c5_params(i))
        renders[f"C5_{i}_audio"] = audio[::8]
        renders[f"C5_{i}_step"] = np.array(8)
    np.savez_compressed(os.path.join(OUT, "renders.npz"), numpy_version=ver, **renders)

    # --- stage vectors on small seeded inputs
    rng = np.random.default_rng(2026)
    x = rng.standard_normal(1000)
    st = dict(numpy_version=ver, x=x)
    st["lowpass_roll"] = ref.lowpass_fft(x, 96000.0, 18000.0, roll=2500.0)
    st["lowpass_brick"] = ref.lowpass_fft(x, 96000.0, 18000.0, roll=0.0)
    st["bandpass"] = ref.bandpass_fft(x, 96000.0, 4000.0, 16000.0, roll=2000.0)
    st["bandpass_brick"] = ref.bandpass_fft(x, 96000.0, 4000.0, 16000.0, roll=0.0)
    st["stretch_4"] = ref.fft_partial_stretch(x, 4.0)
    st["stretch_0p3"] = ref.fft_partial_stretch(x, 0.3)
    xo = rng.standard_normal(999)
    st["x_odd"] = xo
    st["stretch_odd_2p5"] = ref.fft_partial_stretch(xo, 2.5)
    st["lowpass_odd"] = ref.lowpass_fft(xo, 48000.0 * 33.3, 18000.0 * 33.3, roll=2500.0)
    st["multiband"] = ref.unfold_multiband(x, 1_200_000.0, 48000, [(0, 2000.0), (2000.0, 8000.0), (8000.0, 20000.0)],
                                           [35.0, 20.0, 12.0], roll_hz=2000.0)
    st["adsr"] = ref.make_adsr(5000, 48000, 20.0, 30.0, 0.65, 40.0, 1.8)
    st["adsr_short"] = ref.make_adsr(2000, 48000, 20.0, 250.0, 0.65, 1800.0, 1.8)
    st["er_cloud"] = ref.early_reflection_cloud(x, 8000, taps=40, max_ms=45, seed=7)
    ir = rng.standard_normal((300, 2))
    st["ir"] = ir
    st["conv_ir"] = ref.convolve_ir_short(x, ir)
    st["stereo_even"] = ref.spectral_diffusion_stereo(x, 48000, width=0.65)
    st["stereo_odd"] = ref.spectral_diffusion_stereo(xo, 48000, width=0.65)
    st["soft_clip"] = ref.soft_clip(x * 3.0, drive=1.7)
    st["normalize"] = ref.normalize(np.column_stack([x, -2 * x]), peak=0.98)
    for mode in configs.BASIC_MODES:
        st["gen_" + mode.replace(" ", "_")] = ref.gen_basic(1_200_000, 1.5, 4242, mode, 0.02, -3.0, 4200.0, 12.0)
    np.savez_compressed(os.path.join(OUT, "stages.npz"), **st)

    # --- scalar / integer decisions
    ev = {}
    for proc in ("Single", "Poisson", "Clustered", "Hawkes"):
        ev["times_" + proc] = np.array(ref.generate_event_times(proc, 3.0, 18.0, 12345, 6, 25.0, 0.6, 0.25))
    pts = ref.parse_breakpoints("0:18, 4:40, junk, 8:14, 2:x, :3")
    ev["bp_points"] = np.array(pts)
    ev["bp_eval"] = np.array([ref.eval_breakpoints(pts, t, 7.0) for t in (-1.0, 0.0, 1.0, 4.0, 6.5, 8.0, 9.0)])
    np.savez_compressed(os.path.join(OUT, "events.npz"), numpy_version=ver, **ev)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
