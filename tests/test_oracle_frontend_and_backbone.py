"""Pins the oracle's restatement of the third-party arithmetic (transformers Whisper) against the
installed HF implementation on the same seeded weights / audio.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import whisper_ref as W
from whisper_medusa_b200.synthetic import preset_config, synthetic_audio, synthetic_state_dict


def test_mel_filters_match_hf():
    from transformers import WhisperFeatureExtractor

    fe = WhisperFeatureExtractor()
    assert np.abs(fe.mel_filters - W.mel_filter_bank()).max() < 1e-7


@pytest.mark.parametrize("seconds", [5.0, 30.0, 0.37])
def test_log_mel_matches_hf_feature_extractor(seconds):
    """HF feature_extraction_whisper.py:189-342 (__call__) on the same PCM."""
    from transformers import WhisperFeatureExtractor

    pcm = synthetic_audio(seconds, stream_id=3)
    hf = WhisperFeatureExtractor()(pcm, sampling_rate=16000, return_tensors="np").input_features[0]
    ours = W.log_mel_spectrogram(pcm)
    assert ours.shape == (80, 3000)
    assert np.abs(hf - ours).max() < 1e-5


def test_log_mel_edge_cases():
    # empty clip and over-long clip (truncated to 30 s) -- the extractor pads / truncates
    z = W.log_mel_spectrogram(np.zeros(0, dtype=np.float32))
    assert z.shape == (80, 3000) and np.allclose(z, z[0, 0])
    long = synthetic_audio(31.0)
    assert np.array_equal(W.log_mel_spectrogram(long), W.log_mel_spectrogram(long[:480000]))


def _hf_model(cfg, sd):
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    hc = WhisperConfig(
        vocab_size=cfg.vocab_size, num_mel_bins=cfg.num_mel_bins, d_model=cfg.d_model,
        encoder_layers=cfg.encoder_layers, encoder_attention_heads=cfg.encoder_attention_heads,
        decoder_layers=cfg.decoder_layers, decoder_attention_heads=cfg.decoder_attention_heads,
        encoder_ffn_dim=cfg.encoder_ffn_dim, decoder_ffn_dim=cfg.decoder_ffn_dim,
        max_source_positions=cfg.max_source_positions, max_target_positions=cfg.max_target_positions,
        pad_token_id=cfg.pad_token_id, bos_token_id=cfg.bos_token_id, eos_token_id=cfg.eos_token_id,
        decoder_start_token_id=cfg.decoder_start_token_id, attn_implementation="eager")
    m = WhisperForConditionalGeneration(hc).eval()
    hsd = {k[len("whisper_model."):]: v.float() for k, v in sd.items() if k.startswith("whisper_model.")}
    missing, unexpected = m.load_state_dict(hsd, strict=False)
    assert not unexpected, unexpected
    assert all("proj_out" in k or "embed_positions" in k for k in missing), missing
    return m


@pytest.mark.parametrize("preset", ["micro", "tiny.en"])
def test_encoder_and_decoder_match_hf(preset):
    """Oracle (fp32 regime) vs HF WhisperModel: encoder states, decoder hidden states for a prompt
    pass, a one-token cached pass and a K+1-token pass with explicit position ids (the verify pass
    of medusa_utils.py:494-516)."""
    cfg = preset_config(preset, heads=4)
    sd = synthetic_state_dict(cfg, seed=11)
    w = W.RefWeights(sd)
    hf = _hf_model(cfg, sd)
    mel = torch.from_numpy(W.log_mel_spectrogram(synthetic_audio(5.0)))
    with torch.no_grad():
        enc_hf = hf.model.encoder(mel[None]).last_hidden_state[0]
    enc = W.encoder_forward(w, cfg, mel, "fp32")
    assert (enc - enc_hf).abs().max() < 2e-4 * max(1.0, float(enc_hf.abs().max()))

    cache = W.new_cache(cfg)
    prompt = [cfg.decoder_start_token_id, 7, 9, 11]
    h0 = W.decoder_forward(w, cfg, prompt, [0, 1, 2, 3], enc, cache, "fp32")
    h1 = W.decoder_forward(w, cfg, [13], [4], enc, cache, "fp32")
    tree = [21, 22, 23, 24, 25]
    h2 = W.decoder_forward(w, cfg, tree, [5, 6, 7, 8, 9], enc, cache, "fp32")
    with torch.no_grad():
        o0 = hf.model.decoder(input_ids=torch.tensor([prompt]), encoder_hidden_states=enc_hf[None], use_cache=True)
        o1 = hf.model.decoder(input_ids=torch.tensor([[13]]), encoder_hidden_states=enc_hf[None],
                              past_key_values=o0.past_key_values, use_cache=True)
        o2 = hf.model.decoder(input_ids=torch.tensor([tree]), encoder_hidden_states=enc_hf[None],
                              past_key_values=o1.past_key_values, use_cache=True,
                              position_ids=torch.tensor([[5, 6, 7, 8, 9]]))
    for ours, theirs in ((h0, o0), (h1, o1), (h2, o2)):
        ref = theirs.last_hidden_state[0]
        assert (ours - ref).abs().max() < 5e-4 * max(1.0, float(ref.abs().max()))
    # proj_out is the tied embedding (HF modeling_whisper.py:964-1100)
    lg = W.proj_out(w, h1)
    with torch.no_grad():
        lg_hf = hf.proj_out(o1.last_hidden_state[0])
    assert (lg - lg_hf).abs().max() < 1e-3


def test_engine_regime_is_close_to_fp32_regime():
    cfg = preset_config("micro", heads=4)
    w = W.RefWeights(synthetic_state_dict(cfg, seed=5))
    mel = torch.from_numpy(W.log_mel_spectrogram(synthetic_audio(5.0)))
    a = W.encoder_forward(w, cfg, mel, "fp32")
    b = W.encoder_forward(w, cfg, mel, "engine")
    assert (a - b).abs().max() < 3e-2 and (a - b).abs().mean() < 3e-3
