"""TEST INFRASTRUCTURE.  Writes tests/golden/presets.npz: all 27 shipped presets (microsound_0.2.1/presets/*.json, merged
over the factory defaults the way on_load_preset does, main_v2.py:1286-1291) rendered for 2 s by the UNMODIFIED reference
(oracle/ref_loader.py imports main_v2.py where it lies), with the shipped impulse responses loaded the way on_load_ir
does (main_v2.py:1401-1413) wherever a preset uses one, and a seeded synthetic grey image for the scan-line preset.

Per preset the fixture holds: the merged parameter dict (JSON), the name of the IR, every 8th output frame (float32),
the (pct, msg) sequence the reference's progress callback received (main_v2.py:599-600, 757-758, 783-784), and the
oracle's rounding-noise floor (how much of the reference's own output a 1e-15 jitter of the grains moves: cepstral warp /
imprint / resonator sign keep the phase of bins that hold only rounding noise).  IR arrays and the image ride along, so
the GPU box needs nothing from /root/reference.      Run:  python oracle/make_golden_presets.py"""
import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from audio_suite_b200 import frontend  # noqa: E402
from oracle import microsound_np as O, ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "presets.npz")
STEP = 8
DUR = 2.0


def main():
    ref = ref_loader.load()
    base = os.path.join(ref_loader.REFERENCE_ROOT, "microsound_0.2.1")
    ir_files = sorted(glob.glob(os.path.join(base, "irs", "*.wav")))
    img = np.random.default_rng(5).integers(0, 256, (40, 300)).astype(np.uint8)
    st = dict(numpy_version=np.array(np.__version__), step=np.array(STEP), img_gray=img)
    names = []
    for k, f in enumerate(sorted(glob.glob(os.path.join(base, "presets", "*.json")))):
        name = os.path.splitext(os.path.basename(f))[0]
        p = frontend.load_preset(f)                                  # {**factory defaults, **json} (tested against get_params' keys)
        p["out_dur_s"] = DUR
        ir_name = ""
        if p["space_ir_on"] or p["gen_mode"] == "IR fragment":
            path = ir_files[k % len(ir_files)]
            ir_name = os.path.basename(path)
            # on_load_ir through the reference's own normalize(); the product loader must give the same array
            a = ref_loader.load_ir_wav(path)
            assert np.array_equal(a, frontend.load_ir_wav(path)), path
            st["ir_" + ir_name] = a
            p["_ir_audio"] = a
        if p["gen_mode"] == "Image scanline":
            p["_img_gray"] = img
        msgs = []
        audio, meta = ref.render(p, progress=lambda pct, msg: msgs.append((int(pct), str(msg))))
        floor = float(O.rounding_noise_floor(p))
        st["audio_" + name] = audio[::STEP].astype(np.float32)
        st["floor_" + name] = np.array(floor)
        st["params_" + name] = np.array(json.dumps({k2: v for k2, v in p.items() if not k2.startswith("_")}))
        st["ir_of_" + name] = np.array(ir_name)
        st["progress_" + name] = np.array(json.dumps(msgs))
        names.append(name)
        print(f"{name:32s} events msgs {len(msgs):3d}  peak {np.max(np.abs(audio)):.3f}  floor {floor:.3e}  ir {ir_name}", flush=True)
    st["names"] = np.array(json.dumps(names))
    np.savez_compressed(OUT, **st)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
