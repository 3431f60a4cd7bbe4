"""CPU ORACLE SUPPORT (test infrastructure -- never imported by the product path).

Executable pin of the reference's speculative loop.  The reference package cannot be imported
under the installed transformers (5.5.0; it pins 4.49.0), but the code of the
loop itself is plain torch.  This module loads -- at RUN TIME, from the read-only checkout, nothing
is copied into the repository -- the verbatim source of

    WhisperMedusaModel._medusa_greedy_search          model.py:404-835
    WhisperMedusaModel.forward                        model.py:1223-1347
    WhisperMedusaModel._forward_medusa_block          model.py:1349-1417
    WhisperMedusaModel._update_medusa_outputs         model.py:317-402
    WhisperMedusaModel.prepare_inputs_for_medusa_tree_generation / get_medusa_choice
    MedusaResBlock                                    model.py:180-210
    medusa_utils.py (whole module: generate_medusa_buffers, generate_candidates, tree_decoding,
                     evaluate_posterior, update_inference_inputs)

and binds those functions, unmodified, onto an adapter class that supplies the handful of
transformers-4.49 hooks they call, implemented on the INSTALLED Whisper modules:

    whisper_model.medusa_forward      -> installed ``WhisperModel`` encoder / decoder, legacy-tuple KV in and out
                                         (4.49 ``WhisperDecoder.forward``: tuple in => ``from_legacy_cache`` ...
                                         ``to_legacy_cache`` out; reference comments model.py:379-381, 1373-1380)
    whisper_model.prepare_inputs_for_generation      4.49 ``modeling_whisper.py``: drop the cached prefix,
                                                     ``past_length = past_key_values[0][0].shape[2]``
    whisper_model._update_model_kwargs_for_generation 4.49 ``generation/utils.py``: past_key_values + cache_position
    whisper_model._has_unfinished_sequences          ``not this_peer_finished``
    EncoderDecoderCache.from_legacy_cache / to_legacy_cache / DynamicCache   (names the block path uses)
    medusa_block(hidden, attention_mask=, encoder_hidden_states=, layer_head_mask=, cross_attn_layer_head_mask=,
                 past_key_value=, output_attentions=, use_cache=) -> tuple      (4.49 ``WhisperDecoderLayer`` call form)

The three hooks in the middle are re-stated from the 4.49 sources as the reference's own comments describe them
(they cannot be read offline); everything the loop *decides* -- candidate generation, the verify pass, typical
acceptance, which KV rows survive, token append, stop rules, post-EOS fill -- runs from the reference's files.

Use: ``RefLoop.available()``; ``RefLoop(cfg, state_dict).generate(mel, ...)`` -> (tokens, sequences, accept_lengths).
"""
from __future__ import annotations

import ast
import importlib.util
import os
import textwrap
import warnings
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("WM_REFERENCE_ROOT", "/root/reference")
REF_MODEL = os.path.join(REF_ROOT, "whisper_medusa", "models", "model.py")
REF_UTILS = os.path.join(REF_ROOT, "whisper_medusa", "models", "medusa_utils.py")

_METHODS = ("_medusa_greedy_search", "forward", "_forward_medusa_block", "_update_medusa_outputs",
            "prepare_inputs_for_medusa_tree_generation", "get_medusa_choice")


def available() -> bool:
    return os.path.isfile(REF_MODEL) and os.path.isfile(REF_UTILS)


def _load_medusa_utils():
    spec = importlib.util.spec_from_file_location("ref_medusa_utils", REF_UTILS)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


# ----------------------------------------------------------------------------------------------
# 4.49-style cache shims on top of the installed cache classes
# ----------------------------------------------------------------------------------------------
def _make_cache_shims():
    from transformers.cache_utils import Cache, DynamicCache
    from transformers.cache_utils import EncoderDecoderCache as _EDC

    class EncoderDecoderCache(_EDC):
        """Installed cache + the two legacy converters of transformers 4.49 (``cache_utils.py``)."""

        @classmethod
        def from_legacy_cache(cls, past_key_values):
            sa, ca = DynamicCache(), DynamicCache()          # lazily growing: the medusa block appends layer N
            for i, layer in enumerate(past_key_values or ()):
                sa.update(layer[0], layer[1], i)
                if len(layer) > 2:
                    ca.update(layer[2], layer[3], i)
            return cls(sa, ca)

        def to_legacy_cache(self):
            out = ()
            sa, ca = self.self_attention_cache, self.cross_attention_cache
            for i in range(len(sa.layers)):
                layer = (sa.layers[i].keys, sa.layers[i].values)
                if i < len(ca.layers) and ca.layers[i].get_seq_length() > 0:
                    layer += (ca.layers[i].keys, ca.layers[i].values)
                out += (layer,)
            return out

    return Cache, DynamicCache, EncoderDecoderCache


class _BlockAdapter(nn.Module):
    """The medusa block called the 4.49 way (kwargs ``past_key_value`` / ``layer_head_mask``, tuple result)."""

    def __init__(self, layer: nn.Module):
        super().__init__()
        self.layer = layer

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, layer_head_mask=None,
                cross_attn_layer_head_mask=None, past_key_value=None, output_attentions=None, use_cache=None):
        assert layer_head_mask is None and cross_attn_layer_head_mask is None and not output_attentions
        # 4.49 ``WhisperDecoder`` hands every layer the causal mask it built; the reference passes the
        # ENCODER attention_mask here (None on the generate path, model.py:1383) => no mask at all: the block
        # attends over its whole cache plus all new rows.  With T == 1 (pass A) that is the causal result; the
        # verify pass result of the block is never read (disable_medusa: only its KV rows are kept).
        out = self.layer(hidden_states, attention_mask=attention_mask, encoder_hidden_states=encoder_hidden_states,
                         past_key_values=past_key_value, use_cache=use_cache)
        return (out,) if torch.is_tensor(out) else out


class _WhisperAdapter(nn.Module):
    """``self.whisper_model`` of the reference wrapper, on the installed ``WhisperForConditionalGeneration``."""

    def __init__(self, hf, edc_cls):
        super().__init__()
        self.hf = hf
        self.proj_out = hf.proj_out
        self.config = hf.config
        self._EDC = edc_cls

    # -- reference model.py:54-131 (Whisper2MedusaHeadsConditionalGeneration.medusa_forward -> self.model(...)) --
    def medusa_forward(self, input_features=None, attention_mask=None, decoder_input_ids=None,
                       decoder_attention_mask=None, head_mask=None, decoder_head_mask=None, cross_attn_head_mask=None,
                       encoder_outputs=None, past_key_values=None, decoder_inputs_embeds=None,
                       decoder_position_ids=None, labels=None, use_cache=None, output_attentions=None,
                       output_hidden_states=None, return_dict=None, **kwargs):
        from transformers.modeling_outputs import Seq2SeqModelOutput

        assert labels is None and decoder_inputs_embeds is None and decoder_attention_mask is None
        if encoder_outputs is None:
            encoder_outputs = self.hf.model.encoder(input_features)
        enc = encoder_outputs[0]
        legacy_in = past_key_values
        cache = self._EDC.from_legacy_cache(legacy_in) if not isinstance(legacy_in, self._EDC) else legacy_in
        pos = decoder_position_ids
        if pos is not None and pos.dim() == 1:
            pos = pos[None, :]
        dec = self.hf.model.decoder(input_ids=decoder_input_ids, encoder_hidden_states=enc, past_key_values=cache,
                                    use_cache=True, position_ids=pos)
        return Seq2SeqModelOutput(last_hidden_state=dec.last_hidden_state, past_key_values=cache.to_legacy_cache(),
                                  encoder_last_hidden_state=enc)

    # -- transformers 4.49 modeling_whisper.py WhisperForConditionalGeneration.prepare_inputs_for_generation --
    def prepare_inputs_for_generation(self, decoder_input_ids, past_key_values=None, use_cache=None,
                                      encoder_outputs=None, attention_mask=None, decoder_attention_mask=None,
                                      cache_position=None, **kwargs):
        past_length = 0
        if past_key_values is not None:
            past_length = past_key_values[0][0].shape[2]          # legacy tuple branch
            if decoder_input_ids.shape[1] > past_length:
                remove_prefix_length = past_length
            else:
                remove_prefix_length = decoder_input_ids.shape[1] - 1
            decoder_input_ids = decoder_input_ids[:, remove_prefix_length:]
        if cache_position is None:
            cache_position = torch.arange(past_length, past_length + decoder_input_ids.shape[1])
        elif use_cache:
            cache_position = cache_position[-decoder_input_ids.shape[1]:]
        return {"encoder_outputs": encoder_outputs, "past_key_values": past_key_values,
                "decoder_input_ids": decoder_input_ids.contiguous(), "use_cache": use_cache,
                "decoder_attention_mask": decoder_attention_mask, "decoder_position_ids": None,
                "cache_position": cache_position}

    # -- transformers 4.49 generation/utils.py GenerationMixin._update_model_kwargs_for_generation --
    def _update_model_kwargs_for_generation(self, outputs, model_kwargs, is_encoder_decoder=False, num_new_tokens=1):
        model_kwargs["past_key_values"] = outputs.past_key_values
        if model_kwargs.get("use_cache", True):
            model_kwargs["cache_position"] = model_kwargs["cache_position"][-1:] + num_new_tokens
        return model_kwargs

    def _has_unfinished_sequences(self, this_peer_finished, synced_gpus, device):
        return not this_peer_finished

    def validate_stopping_criteria(self, stopping_criteria, max_length):
        return stopping_criteria


class _Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _build_ref_class():
    """exec the verbatim method sources of the reference onto a fresh class."""
    from transformers.generation.logits_process import LogitsProcessorList
    from transformers.generation.stopping_criteria import StoppingCriteriaList
    from transformers.generation.utils import (GenerateDecoderOnlyOutput, GenerateEncoderDecoderOutput,
                                               GenerateNonBeamOutput)
    from transformers.modeling_outputs import Seq2SeqLMOutput
    from transformers.utils import ModelOutput
    import logging
    import typing

    src = open(REF_MODEL).read()
    tree = ast.parse(src)
    by_name = {n.name: n for n in tree.body if isinstance(n, ast.ClassDef)}
    wrapper = by_name["WhisperMedusaModel"]
    segs = []
    for n in wrapper.body:
        if isinstance(n, ast.FunctionDef) and n.name in _METHODS:
            lines = src.splitlines()[n.lineno - 1 - len(n.decorator_list): n.end_lineno]
            segs.append("\n".join(lines))
    found = [n.name for n in wrapper.body if isinstance(n, ast.FunctionDef) and n.name in _METHODS]
    assert set(found) == set(_METHODS), found
    resblock = "\n".join(src.splitlines()[by_name["MedusaResBlock"].lineno - 1: by_name["MedusaResBlock"].end_lineno])
    Cache, DynamicCache, EDC = _make_cache_shims()
    mu = _load_medusa_utils()

    class _Logger:
        def warning_once(self, *a, **k):
            pass

        info = warning = warning_once

    ns = dict(torch=torch, nn=nn, warnings=warnings, medusa_utils=mu, LogitsProcessorList=LogitsProcessorList,
              StoppingCriteriaList=StoppingCriteriaList, GenerateDecoderOnlyOutput=GenerateDecoderOnlyOutput,
              GenerateEncoderDecoderOutput=GenerateEncoderDecoderOutput, GenerateNonBeamOutput=GenerateNonBeamOutput,
              Seq2SeqLMOutput=Seq2SeqLMOutput, ModelOutput=ModelOutput, Cache=Cache, DynamicCache=DynamicCache,
              EncoderDecoderCache=EDC, logger=_Logger(), MedusaCrossEntropyLoss=None, MedusaKLDivLoss=None,
              logging=logging)
    for k in ("Optional", "Union", "List", "Tuple", "Dict", "Any", "Callable"):
        ns[k] = getattr(typing, k)
    exec(compile(resblock, REF_MODEL, "exec"), ns)
    body = "class _RefMethods:\n" + "\n\n".join(segs) + "\n"
    exec(compile(body, REF_MODEL, "exec"), ns)
    return ns["_RefMethods"], ns["MedusaResBlock"], EDC, mu


class RefLoop(nn.Module):
    """The reference's loop code running on the installed Whisper modules (batch 1, CPU, fp32)."""

    def __init__(self, cfg, state_dict):
        super().__init__()
        from transformers import WhisperConfig, WhisperForConditionalGeneration
        from transformers.models.whisper.modeling_whisper import WhisperDecoderLayer

        methods, ResBlock, EDC, mu = _build_ref_class()
        for name in _METHODS:
            setattr(RefLoop, name, getattr(methods, name))   # the reference's functions, unmodified
        self.medusa_utils = mu
        hc = WhisperConfig(
            vocab_size=cfg.vocab_size, num_mel_bins=cfg.num_mel_bins, d_model=cfg.d_model,
            encoder_layers=cfg.encoder_layers, encoder_attention_heads=cfg.encoder_attention_heads,
            decoder_layers=cfg.decoder_layers, decoder_attention_heads=cfg.decoder_attention_heads,
            encoder_ffn_dim=cfg.encoder_ffn_dim, decoder_ffn_dim=cfg.decoder_ffn_dim,
            max_source_positions=cfg.max_source_positions, max_target_positions=cfg.max_target_positions,
            pad_token_id=cfg.pad_token_id, bos_token_id=cfg.bos_token_id, eos_token_id=cfg.eos_token_id,
            decoder_start_token_id=cfg.decoder_start_token_id, attn_implementation="eager")
        hf = WhisperForConditionalGeneration(hc).eval()
        hsd = {k[len("whisper_model."):]: v.float() for k, v in state_dict.items() if k.startswith("whisper_model.")}
        missing, unexpected = hf.load_state_dict(hsd, strict=False)
        assert not unexpected, unexpected
        self.whisper_model = _WhisperAdapter(hf, EDC)
        n_heads = cfg.medusa_num_heads + (0 if cfg.is_block else 1)          # reference model.py:235-256
        self.medusa_heads = nn.ModuleList(
            nn.Sequential(*[ResBlock(cfg.d_model, cfg.medusa_hidden_size) for _ in range(cfg.medusa_num_layers)])
            for _ in range(n_heads))
        self.medusa_heads.load_state_dict(
            {k[len("medusa_heads."):]: v.float() for k, v in state_dict.items() if k.startswith("medusa_heads.")})
        if cfg.is_block:
            layer = WhisperDecoderLayer(hc, layer_idx=cfg.decoder_layers)      # reference model.py:248-256
            layer.load_state_dict(
                {k[len("medusa_block."):]: v.float() for k, v in state_dict.items() if k.startswith("medusa_block.")})
            self.medusa_block = _BlockAdapter(layer)
        self.eval()
        self.mcfg = cfg
        self.config = _Cfg(medusa_heads_type=cfg.medusa_heads_type, medusa_num_heads=cfg.medusa_num_heads,
                           medusa_choices=list(cfg.medusa_choices), output_whisper_original=False,
                           is_encoder_decoder=True, use_cache=True, medusa_loss_on_original=False)
        self.generation_config = _Cfg(pad_token_id=cfg.pad_token_id, eos_token_id=cfg.eos_token_id, output_scores=False,
                                      output_attentions=False, output_hidden_states=False,
                                      return_dict_in_generate=False, max_length=448)

    @property
    def base_model(self):
        return _Cfg(device=torch.device("cpu"))

    # ------------------------------------------------------------------------------------------
    def run_loop(self, enc: torch.Tensor, prompt: Sequence[int], *, suppress_tokens=None, begin_suppress_tokens=None,
                 exponential_decay_length_penalty=None, max_length: int = 448, temperature: float = 1.0,
                 posterior_threshold: float = 0.09, posterior_alpha: float = 0.3) -> Tuple[List[int], List[int]]:
        """``_medusa_greedy_search`` called the way ``_multi_heads_generate`` (model.py:1130-1150) calls it.
        Returns (full sequences incl. prompt, accept lengths)."""
        from transformers.generation.logits_process import (ExponentialDecayLengthPenalty, LogitsProcessorList,
                                                            SuppressTokensAtBeginLogitsProcessor,
                                                            SuppressTokensLogitsProcessor)
        from transformers.generation.stopping_criteria import (EosTokenCriteria, MaxLengthCriteria,
                                                               StoppingCriteriaList)
        from transformers.modeling_outputs import BaseModelOutput

        procs = LogitsProcessorList()          # order of 4.49 ``_get_logits_processor``
        if exponential_decay_length_penalty is not None:
            procs.append(ExponentialDecayLengthPenalty(tuple(exponential_decay_length_penalty), self.mcfg.eos_token_id,
                                                       len(prompt)))
        if suppress_tokens:
            procs.append(SuppressTokensLogitsProcessor(list(suppress_tokens)))
        if begin_suppress_tokens:
            procs.append(SuppressTokensAtBeginLogitsProcessor(list(begin_suppress_tokens), len(prompt)))
        stop = StoppingCriteriaList([MaxLengthCriteria(max_length=max_length),
                                     EosTokenCriteria(eos_token_id=self.mcfg.eos_token_id)])
        self.generation_config.max_length = max_length
        # the loop records accept lengths in a local list only: tap evaluate_posterior (pure observer)
        accepts: List[int] = []
        mu = self.medusa_utils
        orig = mu.evaluate_posterior

        def tap(*a, **k):
            best, acc = orig(*a, **k)
            accepts.append(int(acc))
            return best, acc

        mu.evaluate_posterior = tap
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ids = self._medusa_greedy_search(
                    torch.tensor([list(prompt)], dtype=torch.long), logits_processor=procs, stopping_criteria=stop,
                    pad_token_id=self.mcfg.pad_token_id, eos_token_id=self.mcfg.eos_token_id, output_scores=False,
                    output_logits=False, return_dict_in_generate=False, temperature=temperature,
                    posterior_threshold=posterior_threshold, posterior_alpha=posterior_alpha, synced_gpus=False,
                    streamer=None, encoder_outputs=BaseModelOutput(last_hidden_state=enc[None]), use_cache=True)
        finally:
            mu.evaluate_posterior = orig
        return ids[0].tolist(), accepts

    def encode(self, mel: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            return self.whisper_model.hf.model.encoder(mel[None].float()).last_hidden_state[0]
