"""A/B timing of engine builds (development aid; not a test).

    python tests/gpu_ab.py ab_libs/v0.so ab_libs/v1.so ... [--reps 4] [--large-only]

Every library runs the BASELINE configs[1] clip (large-v2 + 10 Medusa-Linear heads, golden fixture
`large_linear_k10`) in the persistent (ring) mode: tokens are checked against the fixture, the decode
time is the device time of the loop (best and median of the repetitions).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from whisper_medusa_b200 import WhisperMedusaModel, _lib  # noqa: E402
from whisper_medusa_b200.synthetic import preset_config, synthetic_audio, synthetic_state_dict  # noqa: E402


def main():
    libs = [a for a in sys.argv[1:] if a.endswith(".so")]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 4
    g = np.load(os.path.join(ROOT, "tests", "golden", "large_linear_k10.npz"))
    seed, stream, max_len, heads, is_block = [int(v) for v in g["meta"]]
    cfg = preset_config("large-v2", heads=heads)
    sd = synthetic_state_dict(cfg, seed=seed)
    pcm = synthetic_audio(float(g["audio_seconds"]), stream_id=stream)
    pen = None if g["penalty"][0] < 0 else (int(g["penalty"][0]), float(g["penalty"][1]))
    kw = dict(language="en", max_length=max_len, exponential_decay_length_penalty=pen,
              medusa_temperature=float(g["temperature"]))
    for path in libs:
        _lib._lib = None
        _lib.LIB_PATH = os.path.abspath(path)
        try:
            model = WhisperMedusaModel(cfg, sd).to("cuda:0")
            model.set_decode_mode("persistent")
            for kv in [o for o in os.environ.get("WM_AB_OPTS", "").split(",") if o]:      # e.g. WM_AB_OPTS=iters_per_launch=1
                k, v = kv.split("=")
                model.set_option(k, int(v))
            times, ok = [], True
            for r in range(reps + 1):
                out = model.generate_from_pcm(pcm, **kw)[0].tolist()
                ok = ok and out == g["tokens"].tolist() and model.last_trace.accept_lengths == g["accept_lengths"].tolist()
                if r > 0:
                    times.append(model.last_trace.ms_decode / max(1, model.last_trace.iterations))
            it = model.last_trace.iterations
            print(f"{os.path.basename(path):<12} tokens {'OK ' if ok else 'BAD'} iterations {it:3d}  ms/iter best {min(times):.4f} "
                  f"median {float(np.median(times)):.4f}  encoder {model.last_trace.ms_encoder:.2f} ms", flush=True)
            model.close()
        except Exception as e:  # keep going: one broken variant must not lose the others
            print(f"{os.path.basename(path):<12} FAILED: {type(e).__name__}: {e}", flush=True)


if __name__ == "__main__":
    main()
