"""Seeded synthetic checkpoints and audio (no hub / dataset access in this environment).

The state dict uses the reference's checkpoint key layout (attribute
names in reference ``model.py:213-256``):

    whisper_model.model.encoder.* / whisper_model.model.decoder.* / whisper_model.proj_out.weight
    medusa_heads.{i}.{l}.linear.{weight,bias}            (reference model.py:235-246)
    medusa_block.*  (a WhisperDecoderLayer)              (reference model.py:248-256)

All tensors are fp16 (an "fp16 checkpoint"); both the oracle and the CUDA engine consume
exactly these values.  Everything is random (including LayerNorm affine and biases) so a
dropped bias or a mis-packed matrix cannot hide behind an identity initialisation.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .config import MedusaConfig

SAMPLE_RATE = 16000


def _sinusoids(length: int, channels: int) -> torch.Tensor:
    """Whisper encoder position table (HF ``modeling_whisper.py`` ``sinusoids``)."""
    inc = math.log(10000.0) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
    t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([t.sin(), t.cos()], dim=1)


def _decoder_layer_keys(prefix: str, d: int, ffn: int):
    yield f"{prefix}.self_attn_layer_norm", "ln", (d,)
    for p in ("q_proj", "k_proj", "v_proj", "out_proj"):
        yield f"{prefix}.self_attn.{p}", "lin" if p != "k_proj" else "lin_nobias", (d, d)
    yield f"{prefix}.encoder_attn_layer_norm", "ln", (d,)
    for p in ("q_proj", "k_proj", "v_proj", "out_proj"):
        yield f"{prefix}.encoder_attn.{p}", "lin" if p != "k_proj" else "lin_nobias", (d, d)
    yield f"{prefix}.final_layer_norm", "ln", (d,)
    yield f"{prefix}.fc1", "lin", (ffn, d)
    yield f"{prefix}.fc2", "lin", (d, ffn)


def synthetic_state_dict(config: MedusaConfig, seed: int = 0, enc_gain: float = 1.0,
                         dec_gain: float = 2.0, head_gain: float = 0.8,
                         logit_std: float = 1.75) -> Dict[str, torch.Tensor]:
    """Deterministic fp16 state dict of the exact shapes of ``config`` (CPU tensors).

    Linear weights are N(0, gain^2 / fan_in).  The gains are chosen (empirically, see
    DESIGN.md "synthetic checkpoints") so that a random model does not collapse onto one
    repeated token: ``dec_gain`` 2 makes the decoder a strongly non-linear function of the
    previous token, ``logit_std`` sets the spread of the vocabulary logits
    (= std(embedding) * sqrt(d)), and ``head_gain`` makes the Medusa heads disagree with the
    base head often enough that every accept length 0..K occurs under typical acceptance.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    d, V = config.d_model, config.vocab_size
    sd: Dict[str, torch.Tensor] = {}
    gain = {"v": enc_gain}

    def randn(*shape, scale):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * scale).to(torch.float16)

    def add(name: str, kind: str, shape):
        if kind == "ln":
            sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(*shape, generator=g)).to(torch.float16)
            sd[name + ".bias"] = randn(*shape, scale=0.1)
        elif kind in ("lin", "lin_nobias"):
            sd[name + ".weight"] = randn(*shape, scale=gain["v"] / math.sqrt(shape[1]))
            if kind == "lin":
                sd[name + ".bias"] = randn(shape[0], scale=0.05)
        else:
            raise AssertionError(kind)

    enc = "whisper_model.model.encoder"
    sd[f"{enc}.conv1.weight"] = randn(d, config.num_mel_bins, 3, scale=enc_gain / math.sqrt(3 * config.num_mel_bins))
    sd[f"{enc}.conv1.bias"] = randn(d, scale=0.05)
    sd[f"{enc}.conv2.weight"] = randn(d, d, 3, scale=enc_gain / math.sqrt(3 * d))
    sd[f"{enc}.conv2.bias"] = randn(d, scale=0.05)
    sd[f"{enc}.embed_positions.weight"] = _sinusoids(config.max_source_positions, d).to(torch.float16)
    for i in range(config.encoder_layers):
        p = f"{enc}.layers.{i}"
        add(f"{p}.self_attn_layer_norm", "ln", (d,))
        for q in ("q_proj", "k_proj", "v_proj", "out_proj"):
            add(f"{p}.self_attn.{q}", "lin" if q != "k_proj" else "lin_nobias", (d, d))
        add(f"{p}.final_layer_norm", "ln", (d,))
        add(f"{p}.fc1", "lin", (config.encoder_ffn_dim, d))
        add(f"{p}.fc2", "lin", (d, config.encoder_ffn_dim))
    add(f"{enc}.layer_norm", "ln", (d,))

    dec = "whisper_model.model.decoder"
    gain["v"] = dec_gain
    sd[f"{dec}.embed_tokens.weight"] = randn(V, d, scale=logit_std / math.sqrt(d))
    sd[f"{dec}.embed_positions.weight"] = randn(config.max_target_positions, d, scale=0.5 * logit_std / math.sqrt(d))
    for i in range(config.decoder_layers):
        for name, kind, shape in _decoder_layer_keys(f"{dec}.layers.{i}", d, config.decoder_ffn_dim):
            add(name, kind, shape)
    add(f"{dec}.layer_norm", "ln", (d,))
    sd["whisper_model.proj_out.weight"] = sd[f"{dec}.embed_tokens.weight"]  # tied

    n_heads = config.medusa_num_heads + (0 if config.is_block else 1)  # reference model.py:235-256
    gain["v"] = head_gain
    for i in range(n_heads):
        for l in range(config.medusa_num_layers):
            add(f"medusa_heads.{i}.{l}.linear", "lin", (config.medusa_hidden_size, d))
    gain["v"] = dec_gain
    if config.is_block:
        for name, kind, shape in _decoder_layer_keys("medusa_block", d, config.decoder_ffn_dim):
            add(name, kind, shape)
    return sd


def synthetic_audio(seconds: float, stream_id: int = 0, seed: int = 1234) -> np.ndarray:
    """Speech-like 16 kHz mono f32 clip: a few AM-modulated harmonic
    stacks plus low-level noise, peak-normalised to 0.5."""
    rng = np.random.default_rng(seed + stream_id)
    n = int(round(seconds * SAMPLE_RATE))
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    x = np.zeros(n, dtype=np.float64)
    for _ in range(int(rng.integers(3, 6))):
        f0 = rng.uniform(90.0, 260.0)
        am = 0.5 * (1.0 + np.sin(2 * np.pi * rng.uniform(1.5, 6.0) * t + rng.uniform(0, 2 * np.pi)))
        for h in range(1, 9):
            x += am * (rng.uniform(0.2, 1.0) / h) * np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 2 * np.pi))
    x += 0.02 * rng.standard_normal(n)
    x *= 0.5 / max(np.abs(x).max(), 1e-9)
    return x.astype(np.float32)


def preset_config(name: str, heads: int = 10, heads_type: str = "base_head", **kw) -> MedusaConfig:
    """BASELINE.json configs: ``large-v2`` (cfg 2-5), ``tiny.en`` (cfg 1), ``micro`` (tests)."""
    table = {
        "large-v2": "openai/whisper-large-v2",
        "tiny.en": "openai/whisper-tiny.en",
        "micro": "synthetic/whisper-micro",
    }
    wname = table.get(name, name)
    from .config import WHISPER_PRESETS

    d_model = WHISPER_PRESETS[wname]["d_model"]
    return MedusaConfig(
        medusa_num_heads=heads, medusa_num_layers=1, medusa_hidden_size=d_model,
        whisper_model_name=wname, medusa_choices=[1] * (heads + 1), medusa_heads_type=heads_type, **kw)
