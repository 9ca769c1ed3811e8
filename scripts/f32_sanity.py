import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import kernel_checks as K
from audio_suite_b200 import configs, engine
from oracle import microsound_np as O
dev = engine.CudaDevice(0)
for name in K.PRESET_LIKE:
    p = K.preset_like(name); p["out_dur_s"] = 1.5
    out, _ = engine.render(p, device=dev, precision="f32")
    ref, _ = O.render(p)
    print("%-28s f32 finite %s max-abs %.2e (floor %.1e)" % (name, bool(np.isfinite(out).all()), np.max(np.abs(out - ref)), O.rounding_noise_floor(p)))
