"""Per-stage timeline of the persistent decode kernel (device %globaltimer, CTA 0 and last CTA).

    python tests/gpu_stage_profile.py [--preset large-v2] [--heads 10] [--block] [--iters 6] [--tc]

Prints, per stage type, the mean body time and barrier-wait time over the layers of the LAST
iteration, and the critical-path sum.  Used to decide what to optimise next (profiles/*.txt).
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from whisper_medusa_b200 import WhisperMedusaModel, _lib  # noqa: E402
from whisper_medusa_b200.synthetic import preset_config, synthetic_audio, synthetic_state_dict  # noqa: E402

STAGES = ["EMBED", "QKV", "SELF_ATTN", "OPROJ", "CROSS_Q", "CROSS_ATTN", "CROSS_O", "FC1", "FC2", "FINAL_LN",
          "COPY_HIDDEN", "TAIL_SEED", "HEADS", "VOCAB", "SELECT1", "SELECT2", "SELECT_FIN", "ACCEPT"]
MODES = ["A", "B", "TAIL"]


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def main():
    preset = arg("--preset", "large-v2")
    heads = arg("--heads", 10)
    iters = arg("--iters", 6)
    cfg = preset_config(preset, heads=heads, heads_type="medusa_block" if "--block" in sys.argv else "base_head")
    model = WhisperMedusaModel(cfg, synthetic_state_dict(cfg, seed=0)).to("cuda:0")
    model.set_decode_mode("persistent")
    if "--tc" in sys.argv:
        model.set_option("enc_gemm", 1)
    if "--no-prof" in sys.argv:     # plain run (e.g. under ncu): no timeline
        pcm = synthetic_audio(30.0 if preset == "large-v2" else 5.0)
        model.generate_from_pcm(pcm, language="en" if cfg.is_multilingual else None, max_iters=iters)
        print(f"decode {model.last_trace.ms_decode:.3f} ms for {model.last_trace.iterations} iterations")
        model.close()
        return
    model.set_option("profile", 1)
    pcm = synthetic_audio(30.0 if preset == "large-v2" else 5.0)
    lang = "en" if cfg.is_multilingual else None
    model.generate_from_pcm(pcm, language=lang, max_iters=iters)
    tr = model.last_trace
    lib = _lib.load()
    cap = 4096
    buf = (C.c_int64 * (cap * 16))()
    n = C.c_int32(0)
    rc = lib.wm_get_stage_profile(model._handle, buf, cap, C.byref(n))
    assert rc == 0, lib.wm_last_error(model._handle)
    rows = np.frombuffer(buf, dtype=np.int64)[: n.value * 16].reshape(-1, 16)
    agg = OrderedDict()
    for row in rows:
        st, mode, layer, b0, w0, b1, w1 = row[:7]
        a = agg.setdefault((STAGES[st], MODES[mode]), [0] * 13)
        a[0] += 1; a[1] += b0; a[2] += w0; a[3] += b1; a[4] += w1
        for k in range(8):
            a[5 + k] += max(row[7 + k], 0)
    print(f"{preset} K={heads} iterations={tr.iterations} decode {tr.ms_decode:.3f} ms "
          f"({tr.ms_decode / max(1, tr.iterations):.3f} ms/iter), encoder {tr.ms_encoder:.3f} ms, mel {tr.ms_mel:.3f} ms")
    sub = ["desc", "x-load", "ln-stat", "x-loop", "x-sync", "w-ready", "mma", "epi"]
    print(f"{'stage':<12}{'mode':<5}{'n':>3}{'body0':>7}{'wait0':>7}{'bodyN':>7}{'waitN':>7}{'total':>8}"
          + "".join(f"{x:>8}" for x in sub) + "   (us; sub-phases = offsets from stage begin, CTA 0)")
    tot = 0.0
    for (st, mode), a in agg.items():
        cnt = a[0]
        t = (a[1] + a[2]) / 1e3
        tot += t
        print(f"{st:<12}{mode:<5}{cnt:>3}{a[1] / cnt / 1e3:>7.2f}{a[2] / cnt / 1e3:>7.2f}{a[3] / cnt / 1e3:>7.2f}"
              f"{a[4] / cnt / 1e3:>7.2f}{t:>8.1f}" + "".join(f"{a[5 + k] / cnt / 1e3:>8.2f}" for k in range(8)))
    print(f"sum over stages of the last iteration (CTA 0): {tot / 1e3:.3f} ms")
    model.close()


if __name__ == "__main__":
    main()
