"""Writes tests/golden/next_rows.npz: vectors from the UNMODIFIED reference for the SURVEY 8(f) rows that have been
accelerated since the first fixtures -- the wavelet-atom generator (main_v2.py:317-331, 165-170) and
SpectralImprint (565-581) -- plus decimated renders of parameter sets shaped like the shipped presets that need
only those rows.  Run:  python oracle/make_golden_next.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
from oracle import ref_loader  # noqa: E402
import kernel_checks as K  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "next_rows.npz")


def main():
    ref = ref_loader.load()
    st = dict(numpy_version=np.array(np.__version__))
    st["wavelet"] = ref.gen_wavelet_atoms(1_200_000, 1.6, 4242, base_hz=1800, count=10, spread=0.9)
    st["wavelet_floor"] = ref.gen_wavelet_atoms(1_200_000, 0.11, 7, base_hz=2400, count=3, spread=0.6)     # 132 samples
    rng = np.random.default_rng(31)
    imp = ref.SpectralImprint()
    seq = [rng.standard_normal(n) for n in (400, 400, 400, 401, 401, 63, 400)]
    st["imprint_in"] = np.concatenate(seq)
    st["imprint_out"] = np.concatenate([imp.apply(x.copy(), amount=0.35, smooth=0.9) for x in seq])
    for name in K.PRESET_LIKE:
        p = K.preset_like(name)
        p["out_dur_s"] = 2.0
        audio, meta = ref.render(p)
        st["render_" + name] = audio[::4]
    np.savez_compressed(OUT, **st)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
