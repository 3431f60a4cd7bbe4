"""CPU ORACLE (test infrastructure -- never imported by the product path).

Restatement of the reference's Medusa speculative decode loop and its helpers, for batch 1:

* ``generate_medusa_buffers`` / ``generate_candidates`` / ``evaluate_posterior``
      reference ``whisper_medusa/models/medusa_utils.py:305-421 / 424-458 / 526-588``
      (pinned against the reference file itself in ``tests/test_oracle_medusa_utils.py``).
* ``medusa_greedy_search``
      reference ``whisper_medusa/models/model.py:404-835`` (loop), ``:317-402`` (which KV
      rows survive), ``medusa_utils.py:461-523`` (verify pass), ``:591-671`` (token append).
* ``generate``
      reference ``model.py:1419-1779`` + ``:1842-2013``: prompt tokens, processors,
      ``temperature := 1.0`` (``:1878-1881``), prompt / EOS stripping (``:1929-1973``).
* logits processors: HF ``generation/logits_process.py:1893-1901`` (suppress),
  ``:1847-1862`` (begin-suppress), ``:1742-1772`` (EOS exponential decay).

Pinning: the reference's package cannot be imported under the installed transformers, but
``oracle/ref_harness.py`` executes the VERBATIM source of its loop (``_medusa_greedy_search``, ``forward``,
``_forward_medusa_block``, ``_update_medusa_outputs``, all of ``medusa_utils.py``) on the installed Whisper modules;
``tests/test_ref_loop_pin.py`` holds this restatement to those outputs bit-for-bit on 120 unselected streams (chains and
branching trees, Linear and Block heads, both acceptance rules, length penalty / EOS), frozen in
``tests/golden/ref_loop_streams.npz``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import whisper_ref as W


# --------------------------------------------------------------------------------------
# medusa_utils restatements
# --------------------------------------------------------------------------------------
def generate_medusa_buffers(medusa_choices: Sequence[int]) -> dict:
    """Index tables of the candidate tree (reference ``medusa_utils.py:305-421``).

    ``medusa_choices[i]`` = branching factor at depth i.  Nodes are numbered level by level;
    level i has ``prod(choices[:i+1])`` nodes.  Returns the three tables the loop consumes
    (``medusa_attn_mask`` / ``list_indices`` are built but never read by the reference --
    SURVEY.md 3.3 -- and are omitted).
    """
    ch = [int(c) for c in medusa_choices]
    level_sizes, offsets = [], []
    prod, cum = 1, 0
    for c in ch:
        prod *= c
        level_sizes.append(prod)
    n_cand = level_sizes[-1]
    # tree_indices: node -> index into the flat candidate list [c0 | top-k of head 1 | ...]
    tree_indices: List[int] = []
    flat_off = 0
    for i, c in enumerate(ch):
        reps = level_sizes[i] // c
        tree_indices += list(range(flat_off, flat_off + c)) * reps
        flat_off += c
    position_ids: List[int] = []
    for i, s in enumerate(level_sizes):
        position_ids += [i] * s
    # retrieve_indices[cand, depth] = node id of that candidate's ancestor at depth
    retrieve = torch.zeros(n_cand, len(ch), dtype=torch.long)
    start = 0
    for i, s in enumerate(level_sizes):
        rep = n_cand // s
        retrieve[:, i] = torch.arange(start, start + s).repeat_interleave(rep)
        start += s
    return {
        "tree_indices": torch.tensor(tree_indices, dtype=torch.long),
        "medusa_position_ids": torch.tensor(position_ids, dtype=torch.long),
        "retrieve_indices": retrieve,
    }


def generate_candidates(medusa_rows: torch.Tensor, base_row: torch.Tensor, medusa_topk: Sequence[int],
                        tree_indices: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Reference ``medusa_utils.py:424-458`` for batch 1 / last position.

    ``base_row`` ``[V]`` (processed base logits), ``medusa_rows`` ``[H, V]``.
    Returns ``candidates [n_cand, H+1]`` (cartesian product, first factor varies slowest)
    and ``tree_candidates [n_tree]``.
    """
    cands = [torch.argmax(base_row).reshape(1)]
    for i in range(medusa_rows.shape[0]):
        cands.append(torch.topk(medusa_rows[i], int(medusa_topk[i])).indices)
    flat = torch.cat(cands)
    candidates = torch.cartesian_prod(*cands)
    if candidates.dim() == 1:
        candidates = candidates[None, :]
    return candidates, flat[tree_indices]


def evaluate_posterior(logits: torch.Tensor, candidates: torch.Tensor, temperature: float,
                       posterior_threshold: float, posterior_alpha: float) -> Tuple[int, int]:
    """Reference ``medusa_utils.py:526-588``.  ``logits [n_cand, H+1, V]``.

    temperature == 0: longest prefix matching the verify argmax.  Otherwise typical
    acceptance: ``p(cand) > min(threshold, alpha * exp(-entropy))`` with
    ``entropy = -sum p log(p + 1e-5)``; ties on length broken by summed log-likelihood.
    """
    if temperature == 0:
        mask = (candidates[:, 1:] == torch.argmax(logits[:, :-1], dim=-1)).int()
        lens = torch.cumprod(mask, dim=1).sum(dim=1)
        accept = int(lens.max())
        best = 0 if accept == 0 else int(torch.argmax(lens))
        return best, accept
    prob = torch.softmax(logits[:, :-1] / temperature, dim=-1)
    cprob = torch.gather(prob, -1, candidates[:, 1:].unsqueeze(-1)).squeeze(-1)
    entropy = -torch.sum(prob * torch.log(prob + 1e-5), dim=-1)
    thr = torch.minimum(torch.full_like(entropy, posterior_threshold), torch.exp(-entropy) * posterior_alpha)
    mask = cprob > thr
    lens = torch.cumprod(mask, dim=1).sum(dim=1)
    accept = int(lens.max())
    if accept == 0:
        return 0, 0
    best_set = torch.where(lens == accept)[0]
    like = torch.sum(torch.log(cprob[best_set, :accept]), dim=-1)
    return int(best_set[torch.argmax(like)]), accept


# --------------------------------------------------------------------------------------
# logits processors
# --------------------------------------------------------------------------------------
@dataclass
class GenParams:
    """Everything ``_medusa_greedy_search`` reads from generation config + kwargs."""

    eos_token_id: int
    pad_token_id: int
    suppress_tokens: Optional[List[int]] = None
    begin_suppress_tokens: Optional[List[int]] = None
    begin_index: int = 0                      # len(prompt), model.py:1641-1644
    # (start_index, factor) of ExponentialDecayLengthPenalty or None
    exponential_decay_length_penalty: Optional[Tuple[int, float]] = None
    prompt_len: int = 0                       # input_ids_seq_length given to the penalty
    max_length: int = 448
    temperature: float = 1.0                  # generate() forces 1.0 (model.py:1878-1881)
    posterior_threshold: float = 0.09
    posterior_alpha: float = 0.3
    medusa_choices: List[int] = field(default_factory=list)


def process_logits(rows: torch.Tensor, cur_len: int, gp: GenParams) -> torch.Tensor:
    """Apply the three processors to ``rows [R, V]`` with ``cur_len = input_ids.shape[1]``
    (the same ``input_ids`` is passed for every row -- reference ``model.py:653-665,689-694``)."""
    rows = rows.clone()
    if gp.exponential_decay_length_penalty is not None:
        start, factor = gp.exponential_decay_length_penalty
        reg_start = start + gp.prompt_len
        if cur_len > reg_start:
            idx = cur_len - reg_start
            pen = torch.abs(rows[:, gp.eos_token_id]) * (pow(factor, idx) - 1)
            rows[:, gp.eos_token_id] = rows[:, gp.eos_token_id] + pen
    if gp.suppress_tokens:
        rows[:, torch.tensor(gp.suppress_tokens)] = float("-inf")
    if gp.begin_suppress_tokens and cur_len == gp.begin_index:
        rows[:, torch.tensor(gp.begin_suppress_tokens)] = float("-inf")
    return rows


# --------------------------------------------------------------------------------------
# the loop
# --------------------------------------------------------------------------------------
@dataclass
class LoopTrace:
    sequences: List[int]
    accept_lengths: List[int] = field(default_factory=list)
    iters: int = 0
    # per-iteration diagnostics used to pick well-conditioned golden seeds
    min_top2_gap: float = float("inf")
    min_accept_margin: float = float("inf")
    passA_logits: List[torch.Tensor] = field(default_factory=list)  # optional capture
    passB_logits: List[torch.Tensor] = field(default_factory=list)
    candidates: List[List[int]] = field(default_factory=list)


def medusa_greedy_search(w: W.RefWeights, cfg, enc: torch.Tensor, prompt: Sequence[int], gp: GenParams,
                         regime: str = "fp32", capture_logits: int = 0,
                         max_iters: Optional[int] = None, tree_attention: bool = False) -> LoopTrace:
    """Batch-1 restatement of reference ``model.py:404-835``; returns the full ``input_ids``
    (prompt included) after the post-EOS fill (``:798-810``).

    ``tree_attention=True`` is NOT reference behaviour (the reference builds ``medusa_attn_mask`` and never applies it,
    SURVEY.md 3.3): every verify row then attends to the cache and to its own ancestors only -- the mode the engine offers
    as ``generate(tree_attention=True)`` for branching ``medusa_choices``."""
    H = cfg.medusa_num_heads
    choices = list(gp.medusa_choices) if gp.medusa_choices else list(cfg.medusa_choices)
    bufs = generate_medusa_buffers(choices)                      # model.py:615-630
    topk = choices[1:]
    tree_idx, pos_tbl, retrieve = bufs["tree_indices"], bufs["medusa_position_ids"], bufs["retrieve_indices"]
    input_ids = [int(t) for t in prompt]
    cache = W.new_cache(cfg)
    tr = LoopTrace(sequences=input_ids)
    unfinished = True
    allowed = None
    if tree_attention:
        n_tree = int(pos_tbl.shape[0])
        allowed = torch.eye(n_tree, dtype=torch.bool)
        for c in range(retrieve.shape[0]):
            path = retrieve[c].tolist()
            for j, node in enumerate(path):
                allowed[node, path[: j + 1]] = True
    while True:
        L = len(input_ids)
        kv = cache.length
        # A0: tokens not yet cached; after iteration 1 this is exactly one token
        new = input_ids[kv:]
        assert (tr.iters == 0 and kv == 0) or len(new) == 1, (L, kv)
        hidden = W.decoder_forward(w, cfg, new, list(range(kv, L)), enc, cache, regime)   # pass A
        # heads only matter at the last position (generate_candidates reads logits[:, -1]); the
        # block type still needs every new position for its KV slot, so all rows go through.
        rows_raw = W.medusa_logits(w, cfg, hidden, enc, cache, False, regime)[:, -1, :]  # [H+1, V]
        rows = process_logits(rows_raw, L, gp)                                            # A1
        candidates, tree_candidates = generate_candidates(rows[1:], rows[0], topk, tree_idx)  # A2
        # B: verify pass on a *copy* of the cache (tree_outputs.past_key_values, model.py:383-401)
        vcache = cache.clone()
        pos_b = (pos_tbl + L).tolist()
        hidden_b = W.decoder_forward(w, cfg, tree_candidates.tolist(), pos_b, enc, vcache, regime, allowed)
        logits_b_raw = W.medusa_logits(w, cfg, hidden_b, enc, vcache, True, regime)[0]    # [n_tree, V]
        logits_b = process_logits(logits_b_raw, L, gp)                                    # B1
        vlog = logits_b[retrieve]                                                         # [n_cand, H+1, V]
        best, accept = evaluate_posterior(vlog, candidates, gp.temperature, gp.posterior_threshold,
                                          gp.posterior_alpha)                             # C
        # diagnostics (not part of the algorithm)
        top2 = torch.topk(rows, 2, dim=-1).values
        tr.min_top2_gap = min(tr.min_top2_gap, float((top2[:, 0] - top2[:, 1]).min()))
        if gp.temperature != 0:
            prob = torch.softmax(vlog[best, :-1] / gp.temperature, dim=-1)
            cp = prob.gather(-1, candidates[best, 1:, None])[:, 0]
            ent = -(prob * torch.log(prob + 1e-5)).sum(-1)
            thr = torch.minimum(torch.full_like(ent, gp.posterior_threshold), torch.exp(-ent) * gp.posterior_alpha)
            upto = min(accept + 1, cp.shape[0])
            if upto > 0:
                tr.min_accept_margin = min(tr.min_accept_margin, float(((cp - thr).abs() / thr)[:upto].min()))
        if accept == 0:
            t2 = torch.topk(vlog[best, 0], 2).values
            tr.min_top2_gap = min(tr.min_top2_gap, float(t2[0] - t2[1]))
        if tr.iters < capture_logits:
            tr.passA_logits.append(rows_raw.clone())     # raw (pre-processor) rows, like wm_last_logits
            tr.passB_logits.append(logits_b_raw.clone())
        tr.candidates.append(candidates[best].tolist())
        tr.accept_lengths.append(accept)
        use_base = accept == 0                                                            # D
        # E: tokens (medusa_utils.py:630-652)
        next_tokens = candidates[best, : accept + 1].tolist()
        if use_base:
            next_tokens.append(int(torch.argmax(vlog[best, 0])))
        if not unfinished:
            next_tokens = [gp.pad_token_id] * len(next_tokens)
        input_ids = input_ids + next_tokens
        # E: KV rows (model.py:383-401): first `accept` (or accept+1 if use_base) tree rows
        keep = accept + 1 if use_base else accept
        sel = retrieve[best, : accept + 1].tolist()[:keep]
        cache = vcache
        cache.keep_rows(L, [L + s for s in sel])
        assert cache.length == len(input_ids) - 1, (cache.length, len(input_ids))
        tr.iters += 1
        # G: stop rules (model.py:774-793)
        if gp.eos_token_id in next_tokens:
            unfinished = False
        if len(input_ids) >= gp.max_length:            # MaxLengthCriteria
            unfinished = False
        if (not unfinished) or len(input_ids) + H >= gp.max_length:
            break
        if max_iters is not None and tr.iters >= max_iters:
            break
    # post-EOS fill (model.py:798-810)
    if gp.eos_token_id in input_ids:
        j = input_ids.index(gp.eos_token_id)
        input_ids = input_ids[: j + 1] + [gp.eos_token_id] * (len(input_ids) - j - 1)
    tr.sequences = input_ids
    return tr


def strip_output(sequences: Sequence[int], prompt_len: int, gp: GenParams) -> List[int]:
    """Reference ``model.py:1929`` (drop prompt) and ``:1950-1973`` (drop trailing pad/EOS)."""
    seq = list(sequences[prompt_len:])
    if seq and seq[-1] == gp.pad_token_id:
        n_pad = sum(1 for t in seq if t == gp.pad_token_id)
        if gp.pad_token_id == gp.eos_token_id:
            n_pad -= 1
        if n_pad != 0:
            seq = seq[:-n_pad]
    if seq and seq[-1] == gp.eos_token_id:
        seq = seq[:-1]
    return seq


def init_tokens(cfg, language: Optional[str] = None, task: Optional[str] = None) -> List[int]:
    """Prompt ids (HF ``generation_whisper.py:1455-1608`` for the supported cases): multilingual
    -> ``[sot, <|lang|>, <|task|>, <|notimestamps|>]``; English-only -> ``[sot, <|notimestamps|>]``."""
    toks = [int(cfg.decoder_start_token_id)]
    if getattr(cfg, "is_multilingual", False):
        lang = language or "en"
        key = lang if lang.startswith("<|") else f"<|{lang}|>"
        toks.append(int(cfg.lang_to_id[key]))
        toks.append(int(cfg.task_to_id[task or "transcribe"]))
    toks.append(int(cfg.no_timestamps_token_id))
    return toks


def gen_params(cfg, prompt: Sequence[int], exponential_decay_length_penalty=None,
               max_length: Optional[int] = None, **over) -> GenParams:
    gp = GenParams(
        eos_token_id=int(cfg.eos_token_id), pad_token_id=int(cfg.pad_token_id),
        suppress_tokens=list(cfg.suppress_tokens) if getattr(cfg, "suppress_tokens", None) else None,
        begin_suppress_tokens=list(cfg.begin_suppress_tokens) if getattr(cfg, "begin_suppress_tokens", None) else None,
        begin_index=len(prompt), prompt_len=len(prompt),
        exponential_decay_length_penalty=exponential_decay_length_penalty,
        max_length=int(max_length if max_length is not None else getattr(cfg, "max_length", 448)),
        medusa_choices=list(cfg.medusa_choices),
    )
    for k, v in over.items():
        setattr(gp, k, v)
    return gp


def generate(w: W.RefWeights, cfg, mel: torch.Tensor, language: Optional[str] = None,
             exponential_decay_length_penalty=None, regime: str = "fp32", max_length: Optional[int] = None,
             **over) -> Tuple[List[int], LoopTrace]:
    """Oracle of ``WhisperMedusaModel.generate(input_features, language=...)`` for one
    <=30 s clip: returns (generated ids without prompt / trailing EOS, trace)."""
    enc = W.encoder_forward(w, cfg, mel, regime)
    prompt = init_tokens(cfg, language)
    gp = gen_params(cfg, prompt, exponential_decay_length_penalty, max_length, **over)
    tr = medusa_greedy_search(w, cfg, enc, prompt, gp, regime)
    return strip_output(tr.sequences, len(prompt), gp), tr
