"""Pins the oracle's restatement of medusa_utils.py against the reference file itself (loaded by
path; only where /root/reference exists, i.e. the authoring container) and against known answers
recorded from it."""
import importlib.util
import os

import pytest
import torch

from oracle import medusa_ref as M

REF = "/root/reference/whisper_medusa/models/medusa_utils.py"


def _ref():
    if not os.path.isfile(REF):
        pytest.skip("reference checkout not present on this box")
    spec = importlib.util.spec_from_file_location("ref_medusa_utils", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_buffers_known_answer():
    # probed from the reference: generate_medusa_buffers([1, 2, 2])
    b = M.generate_medusa_buffers([1, 2, 2])
    assert b["tree_indices"].tolist() == [0, 1, 2, 3, 4, 3, 4]
    assert b["medusa_position_ids"].tolist() == [0, 1, 1, 2, 2, 2, 2]
    assert b["retrieve_indices"].tolist() == [[0, 1, 3], [0, 1, 4], [0, 2, 5], [0, 2, 6]]
    c = M.generate_medusa_buffers([1] * 11)
    assert c["tree_indices"].tolist() == list(range(11))
    assert c["medusa_position_ids"].tolist() == list(range(11))
    assert c["retrieve_indices"].tolist() == [list(range(11))]


@pytest.mark.parametrize("choices", [[1, 1, 1, 1, 1], [1] * 11, [1, 2, 2], [1, 3, 2, 1], [1, 6, 5, 4, 3]])
def test_buffers_match_reference(choices):
    r = _ref().generate_medusa_buffers(choices, device="cpu")
    o = M.generate_medusa_buffers(choices)
    for k in ("tree_indices", "medusa_position_ids", "retrieve_indices"):
        assert torch.equal(r[k].long(), o[k].long()), k


@pytest.mark.parametrize("choices", [[1, 1, 1, 1, 1], [1, 2, 2], [1, 3, 2, 1]])
def test_candidates_match_reference(choices):
    ref = _ref()
    g = torch.Generator().manual_seed(7)
    H, V, T = len(choices) - 1, 300, 3
    med = torch.randn(H, 1, T, V, generator=g)
    base = torch.randn(1, T, V, generator=g)
    bufs = M.generate_medusa_buffers(choices)
    rc, rt = ref.generate_candidates(med, base, choices[1:], bufs["tree_indices"])
    oc, ot = M.generate_candidates(med[:, 0, -1], base[0, -1], choices[1:], bufs["tree_indices"])
    assert torch.equal(rc, oc)
    assert torch.equal(rt[0], ot)


@pytest.mark.parametrize("temperature", [0.0, 1.0, 0.7])
@pytest.mark.parametrize("choices", [[1, 1, 1, 1, 1], [1, 2, 2]])
def test_evaluate_posterior_matches_reference(choices, temperature):
    ref = _ref()
    g = torch.Generator().manual_seed(3)
    bufs = M.generate_medusa_buffers(choices)
    n_cand, depth = bufs["retrieve_indices"].shape
    V = 200
    for trial in range(40):
        logits = torch.randn(n_cand, depth, V, generator=g) * (0.5 + trial % 4)
        cands = torch.randint(0, V, (n_cand, depth), generator=g)
        if trial % 3 == 0:  # make some prefixes match the argmax so accepts happen
            cands[:, 1:] = logits[:, :-1].argmax(-1)
            cands[:, 1 + trial % depth:] = 0
        rb, ra = ref.evaluate_posterior(logits, cands, temperature, 0.09, 0.3)
        ob, oa = M.evaluate_posterior(logits, cands, temperature, 0.09, 0.3)
        assert int(ra) == oa and int(rb) == ob


def test_processors_match_hf():
    from transformers.generation.logits_process import (ExponentialDecayLengthPenalty,
                                                        SuppressTokensAtBeginLogitsProcessor,
                                                        SuppressTokensLogitsProcessor)

    g = torch.Generator().manual_seed(1)
    V, eos = 100, 50
    gp = M.GenParams(eos_token_id=eos, pad_token_id=eos, suppress_tokens=[1, 5, 9], begin_suppress_tokens=[2, eos],
                     begin_index=4, exponential_decay_length_penalty=(6, 1.3), prompt_len=4)
    procs = [ExponentialDecayLengthPenalty((6, 1.3), eos, 4), SuppressTokensLogitsProcessor([1, 5, 9]),
             SuppressTokensAtBeginLogitsProcessor([2, eos], 4)]
    for cur_len in (4, 5, 10, 11, 17):
        rows = torch.randn(3, V, generator=g)
        ids = torch.zeros(3, cur_len, dtype=torch.long)
        want = rows.clone()
        for p in procs:
            want = p(ids, want)
        got = M.process_logits(rows, cur_len, gp)
        assert torch.equal(torch.nan_to_num(want, neginf=-1e30), torch.nan_to_num(got, neginf=-1e30))


def test_strip_output_rules():
    gp = M.GenParams(eos_token_id=9, pad_token_id=9)
    assert M.strip_output([1, 2, 5, 6, 9, 9, 9], 2, gp) == [5, 6]      # post-EOS fill removed, EOS removed
    assert M.strip_output([1, 2, 5, 6], 2, gp) == [5, 6]                # no EOS
    assert M.strip_output([1, 2, 9, 9], 2, gp) == []                    # only EOS
    gp2 = M.GenParams(eos_token_id=9, pad_token_id=8)
    assert M.strip_output([1, 2, 5, 9, 8, 8], 2, gp2) == [5]
