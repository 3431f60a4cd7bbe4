"""Generate the golden fixtures under tests/golden/ from the CPU oracle (run in the authoring
container; the fixtures travel, /root/reference and long oracle runs do not).

    python -m oracle.make_golden [case ...]

Every fixture records the seeds that reproduce its inputs, the oracle outputs in the "engine"
regime (token ids, accept lengths, sampled mel / encoder / logits values) and how well-conditioned
the discrete decisions were along the path (smallest top-2 logit gap, smallest relative margin of
the typical-acceptance comparison): seeds are picked so no decision sits within fp32 noise.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import medusa_ref as M  # noqa: E402
from oracle import whisper_ref as W  # noqa: E402
from whisper_medusa_b200.synthetic import preset_config, synthetic_audio, synthetic_state_dict  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (preset, heads, heads_type, weight seed, audio seconds, stream id, max_length, penalty, medusa temperature
#          [, posterior_alpha, posterior_threshold])
# The synthetic large-v2 model accepts every candidate under the default typical-acceptance constants (flat
# posteriors: p_candidate * e^H ~ 200-300 against alpha = 0.3).  ``posterior_alpha`` -- a generation-config field of
# the reference (medusa_utils.py:14-18) -- is the acceptance knob: alpha = 100 fails the weaker chain positions, which
# gives mixed accept lengths.
CASES = {
    "micro_linear_k4": ("micro", 4, "base_head", 2, 5.0, 0, 120, None, 1.0),
    "micro_block_k10": ("micro", 10, "medusa_block", 9, 5.0, 1, 120, None, 1.0),
    "micro_linear_k4_t0": ("micro", 4, "base_head", 1, 5.0, 0, 100, (20, 1.05), 0.0),
    "tiny_linear_k4": ("tiny.en", 4, "base_head", 3, 5.0, 0, 200, None, 1.0),
    "tiny_block_k4": ("tiny.en", 4, "medusa_block", 1, 5.0, 0, 200, None, 1.0),
    "large_linear_k10": ("large-v2", 10, "base_head", 0, 30.0, 0, 448, None, 1.0),
    "large_block_k10": ("large-v2", 10, "medusa_block", 0, 30.0, 0, 120, None, 1.0),   # short budget: ~10 iterations
    # mixed accept lengths + the length penalty of the reference's eval script (eval_whisper_medusa.py:61-65) + EOS stop
    "large_linear_k10_mixed": ("large-v2", 10, "base_head", 0, 30.0, 0, 448, (140, 1.01), 1.0, 100.0, 0.09),
    "large_linear_k6_mixed": ("large-v2", 6, "base_head", 0, 30.0, 1, 130, None, 1.0, 100.0, 0.09),
    "large_linear_k4_mixed": ("large-v2", 4, "base_head", 0, 30.0, 2, 110, None, 1.0, 100.0, 0.09),
    "large_linear_k2_mixed": ("large-v2", 2, "base_head", 0, 30.0, 3, 90, None, 1.0, 100.0, 0.09),
    # Block heads of the synthetic model are uncorrelated with the base model (ratio p*e^H ~ 0.02 / 0.015 / 0.004 ...):
    # alpha = 0.016 accepts the first one or two positions => the block's carry / tail path with accept >= 1 at d = 1280
    "large_block_k10_mixed": ("large-v2", 10, "medusa_block", 0, 30.0, 0, 110, None, 1.0, 0.016, 0.09),
    # alpha = 230 also fails the first chain position about half of the time: accept-0 iterations (two emitted
    # tokens, the extra one-token sweep) interleaved with accepting ones
    # (stream 5 of a small search over streams 4-7 x alpha 210/230/250: the only one with accept-0 iterations whose
    # tightest acceptance comparison is still 2.4e-3 away from its threshold)
    "large_linear_k10_a0mix": ("large-v2", 10, "base_head", 0, 30.0, 5, 150, None, 1.0, 230.0, 0.09),
}

TOPN = 16


def _topn(rows: torch.Tensor):
    v, i = torch.topk(rows, TOPN, dim=-1)
    return v.numpy().astype(np.float32), i.numpy().astype(np.int32)


def make_case(name: str) -> dict:
    preset, heads, htype, seed, secs, stream, max_len, pen, temp = CASES[name][:9]
    alpha, thr = (CASES[name][9:] + (0.3, 0.09))[:2] if len(CASES[name]) > 9 else (0.3, 0.09)
    cfg = preset_config(preset, heads=heads, heads_type=htype)
    t0 = time.time()
    sd = synthetic_state_dict(cfg, seed=seed)
    w = W.RefWeights(sd)
    pcm = synthetic_audio(secs, stream_id=stream)
    mel = W.log_mel_spectrogram(pcm)
    melt = torch.from_numpy(mel)
    language = "en" if cfg.is_multilingual else None
    prompt = M.init_tokens(cfg, language)
    gp = M.gen_params(cfg, prompt, pen, max_len, temperature=temp, posterior_alpha=alpha, posterior_threshold=thr)
    out = {}
    for regime in ("engine", "fp32"):
        enc = W.encoder_forward(w, cfg, melt, regime)
        tr = M.medusa_greedy_search(w, cfg, enc, prompt, gp, regime, capture_logits=2)
        toks = M.strip_output(tr.sequences, len(prompt), gp)
        tag = "" if regime == "engine" else "_fp32"
        out["tokens" + tag] = np.array(toks, dtype=np.int32)
        out["sequences" + tag] = np.array(tr.sequences, dtype=np.int32)
        out["accept_lengths" + tag] = np.array(tr.accept_lengths, dtype=np.int32)
        out["enc_sample" + tag] = enc[::50].numpy().astype(np.float32)
        out["min_top2_gap" + tag] = np.float32(tr.min_top2_gap)
        out["min_accept_margin" + tag] = np.float32(tr.min_accept_margin)
        for it in range(len(tr.passA_logits)):
            for ab, rows in (("A", tr.passA_logits[it]), ("B", tr.passB_logits[it])):
                v, i = _topn(rows)
                out[f"logits{ab}{it}_topv{tag}"] = v
                out[f"logits{ab}{it}_topi{tag}"] = i
                out[f"logits{ab}{it}_strided{tag}"] = rows[:, ::97].numpy().astype(np.float32)
        print(f"  {name} [{regime}] {len(toks)} tokens, {tr.iters} iters, accept hist "
              f"{np.bincount(np.array(tr.accept_lengths), minlength=heads + 1).tolist()}, "
              f"min gap {tr.min_top2_gap:.5f}, min accept margin {tr.min_accept_margin:.5f} ({time.time() - t0:.0f}s)",
              flush=True)
    out["mel_sample"] = mel[:, ::8].astype(np.float32)
    out["prompt"] = np.array(prompt, dtype=np.int32)
    out["meta"] = np.array([seed, stream, max_len, heads, int(htype == "medusa_block")], dtype=np.int64)
    out["audio_seconds"] = np.float32(secs)
    out["temperature"] = np.float32(temp)
    out["posterior"] = np.array([alpha, thr], dtype=np.float64)
    out["penalty"] = np.array(pen if pen is not None else (-1, 1.0), dtype=np.float64)
    return out


def main(argv):
    names = argv or [n for n in CASES if not n.startswith("large")]
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for n in names:
        print(f"generating {n}", flush=True)
        np.savez_compressed(os.path.join(GOLDEN, n + ".npz"), **make_case(n))


if __name__ == "__main__":
    main(sys.argv[1:])
