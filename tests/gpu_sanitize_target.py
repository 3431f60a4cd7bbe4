"""Target of tests/run_sanitizer.sh: one micro clip (and optionally one tiny.en clip) through the default (persistent
ring) decode mode, tokens checked against the committed goldens.  Small on purpose: compute-sanitizer slows the
kernels down 10-100x."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from whisper_medusa_b200 import WhisperMedusaModel  # noqa: E402
from whisper_medusa_b200.synthetic import preset_config, synthetic_audio, synthetic_state_dict  # noqa: E402


def run(name: str, preset: str, max_iters: int):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    seed, stream, max_len, heads, is_block = [int(v) for v in g["meta"]]
    cfg = preset_config(preset, heads=heads, heads_type="medusa_block" if is_block else "base_head")
    m = WhisperMedusaModel(cfg, synthetic_state_dict(cfg, seed=seed)).to("cuda:0")
    pen = None if g["penalty"][0] < 0 else (int(g["penalty"][0]), float(g["penalty"][1]))
    m.generate_from_pcm(synthetic_audio(float(g["audio_seconds"]), stream_id=stream), max_length=max_len, max_iters=max_iters,
                        exponential_decay_length_penalty=pen, medusa_temperature=float(g["temperature"]))
    tr = m.last_trace
    n = tr.iterations
    ok = tr.accept_lengths == g["accept_lengths"].tolist()[:n] and tr.sequences[: len(tr.sequences)] == g["sequences"].tolist()[: len(tr.sequences)]
    print(f"{name}: {n} iterations, {tr.launches_decode} ring-kernel launches, tokens {'OK' if ok else 'MISMATCH'}", flush=True)
    m.close()
    return ok


if __name__ == "__main__":
    ok = run("micro_linear_k4", "micro", int(os.environ.get("WM_SAN_ITERS", "6")))
    ok = run("micro_block_k10", "micro", int(os.environ.get("WM_SAN_ITERS", "6"))) and ok
    if "--tiny" in sys.argv:
        ok = run("tiny_linear_k4", "tiny.en", 3) and ok
    sys.exit(0 if ok else 1)
