"""The committed golden fixtures (tests/golden/*.npz, made by oracle/make_golden.py) pin the oracle:
re-running it on the recorded seeds must reproduce them bit-for-bit on the token level.  CPU only,
small configurations (the large-v2 fixture is checked on the GPU box against the engine)."""
import os

import numpy as np
import pytest
import torch

from oracle import medusa_ref as M
from oracle import whisper_ref as W
from _wm_paths import GOLDEN
from whisper_medusa_b200.synthetic import preset_config, synthetic_audio, synthetic_state_dict

CASES = ["micro_linear_k4", "micro_block_k10", "micro_linear_k4_t0", "tiny_linear_k4", "tiny_block_k4"]


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    seed, stream, max_len, heads, is_block = [int(v) for v in g["meta"]]
    preset = {"micro": "micro", "tiny": "tiny.en", "large": "large-v2"}[name.split("_")[0]]
    cfg = preset_config(preset, heads=heads, heads_type="medusa_block" if is_block else "base_head")
    pen = None if g["penalty"][0] < 0 else (int(g["penalty"][0]), float(g["penalty"][1]))
    return g, cfg, seed, stream, max_len, pen, float(g["temperature"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    g, cfg, seed, stream, max_len, pen, temp = load_case(name)
    w = W.RefWeights(synthetic_state_dict(cfg, seed=seed))
    pcm = synthetic_audio(float(g["audio_seconds"]), stream_id=stream)
    mel = W.log_mel_spectrogram(pcm)
    assert np.abs(mel[:, ::8] - g["mel_sample"]).max() < 1e-5
    toks, tr = M.generate(w, cfg, torch.from_numpy(mel), language="en" if cfg.is_multilingual else None,
                          exponential_decay_length_penalty=pen, regime="engine", max_length=max_len, temperature=temp)
    assert toks == g["tokens"].tolist()
    assert tr.accept_lengths == g["accept_lengths"].tolist()
    # invariants of SURVEY.md 3.3: tokens emitted per iteration = accept+1 (accept>=1) or 2 (accept=0)
    n_new = sum(a + 1 if a > 0 else 2 for a in tr.accept_lengths)
    assert len(tr.sequences) == len(g["prompt"]) + n_new


def test_golden_cases_cover_the_decision_space():
    seen = set()
    for name in CASES:
        g, cfg, *_ = load_case(name)
        seen |= set(g["accept_lengths"].tolist())
        assert float(g["min_top2_gap"]) > 5e-4, name        # no argmax decided inside fp32 noise
    assert {0, 1, 2, 3, 4} <= seen
    g = np.load(os.path.join(GOLDEN, "micro_linear_k4_t0.npz"))
    assert len(g["tokens"]) + len(g["prompt"]) < int(g["meta"][2]) - 12, "the EOS path must be exercised"
