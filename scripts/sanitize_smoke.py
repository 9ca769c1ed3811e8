"""Small end-to-end workload for compute-sanitizer (one tool per gpurun call): the fused FIR phases, the warp-local and
block-wide FFT tiles (direct, two-pass, in-tile Bluestein, global Bluestein), partial lock (atomics + bisection),
cepstral warp, imprint, resonator / waveguide, event feedback, every generator, the post passes, the decimation
extension -- each checked against the oracle so a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from audio_suite_b200 import configs, engine  # noqa: E402
from oracle import microsound_np as O  # noqa: E402


def main():
    dev = engine.CudaDevice(0)
    W = configs.with_defaults
    ir = configs.synth_ir(0.2, 48000, 3)
    cases = []
    p = configs.c5_params(3)
    p["out_dur_s"] = 0.3
    cases.append(("sweep member (fused FIR, in-tile Bluestein)", p))
    p = configs.c5_params(7)
    p.update(out_dur_s=1.3, event_process="Poisson", grains_per_sec=6.0)
    cases.append(("sweep member, two block pairs", p))
    cases.append(("C1b", configs.canonical("C1b")))
    for mode in configs.BASIC_MODES:
        cases.append((mode, W(gen_mode=mode, event_process="Poisson", out_dur_s=0.3, grains_per_sec=25.0, time_unfold=40.0, micro_ms=2.0,
                              partial_stretch=1.7, space_ir_on=True, _ir_audio=ir)))
    cases.append(("partial lock + warp + multiband + imprint", W(partial_lock_on=True, partial_stretch=1.7, nl_warp_on=True,
                  unfold_mode="Multi-band unfold", spectral_imprint_on=True, gen_mode="Wavelet atoms", event_process="Poisson",
                  out_dur_s=0.3, grains_per_sec=25.0, time_unfold=60.0, micro_ms=3.0, bandlimit_on=False)))
    cases.append(("cepstral warp", W(cep_warp_on=True, cep_factor=1.3, bandlimit_on=False, partial_stretch=1.2, event_process="Poisson",
                  out_dur_s=0.3, grains_per_sec=25.0, time_unfold=60.0, micro_ms=3.0)))
    cases.append(("stick-slip + waveguide + lock", W(gen_mode="Stick–slip friction", wg_on=True, wg_lines=6, wg_max_ms=1.5, partial_lock_on=True,
                  partial_stretch=1.18, event_process="Poisson", out_dur_s=0.3, grains_per_sec=25.0, micro_ms=2.0)))
    cases.append(("micro-chaos + resonator + feedback + imprint", W(gen_mode="Micro-chaos", res_bank_on=True, event_feedback_on=True,
                  spectral_imprint_on=True, bandlimit_on=False, event_process="Poisson", out_dur_s=0.3, grains_per_sec=25.0, micro_ms=2.0)))
    cases.append(("odd length stereo (FFT rotation), prime grain length", W(event_process="Clustered", out_dur_s=0.10003, base_sr=44100,
                  grains_per_sec=30.0, bp_unfold="0:20, 0.1:33.3")))
    worst = 0.0
    for name, p in cases:
        out, _ = engine.render(p, device=dev)
        ref, _ = O.render(p)
        err = float(np.max(np.abs(out - ref)))
        worst = max(worst, err)
        print("%-55s max-abs %.2e" % (name, err), flush=True)
        assert err < 1e-5 + 4.0 * O.rounding_noise_floor(p), name
    from audio_suite_b200 import decimate as D
    from scipy import signal
    x = np.random.default_rng(1).standard_normal(5000)
    assert np.max(np.abs(D.decimate(x, 16, dev) - signal.resample_poly(x, 1, 16))) < 1e-11
    print("decimation extension ok; worst render error %.2e" % worst)


if __name__ == "__main__":
    main()
