"""WER / CER of the reference's evaluation, without the `jiwer` dependency.

The reference scores transcripts with `jiwer==3.0.3` (reference `requirements.txt:1`, call sites
`whisper_medusa/utils/metrics.py:5-71`): a chain of text transforms followed by a word- (or
character-) level Levenshtein alignment; the corpus score is

    (substitutions + deletions + insertions) / (substitutions + deletions + hits)
  =  edit_distance(reference, hypothesis) / len(reference)          summed over the utterances,

(`metrics.py:33-37, 66-70`) and the per-utterance list holds the same ratio per pair.  `jiwer` is
not installed in this image, so this module restates the published behaviour of the transforms
the reference composes (jiwer 3.0.3 `transforms.py`: ToLowerCase, ExpandCommonEnglishContractions,
RemoveKaldiNonWords, RemoveWhiteSpace(replace_by_space=True), RemoveMultipleSpaces,
RemovePunctuation, Strip, ReduceToListOfListOfWords / ...OfChars) and computes the edit distance
with a plain dynamic programme.  Only the edit distance and the reference length enter the scores,
and both are unique (the split of an optimal alignment into S / D / I is not, and is not used).

Host-side text processing: nothing here touches the GPU path.
"""
from __future__ import annotations

import re
import unicodedata
from typing import List, Sequence, Tuple

__all__ = ["compute_wer", "compute_cer", "wer_standardize", "cer_standardize", "edit_distance"]

# ---- transforms (jiwer 3.0.3 semantics) --------------------------------------------------------

_CONTRACTIONS = (
    # specific words first, then the general attachments -- the order of jiwer's
    # ExpandCommonEnglishContractions.process_string
    (r"won't", "will not"), (r"can\'t", "can not"), (r"let\'s", "let us"),
    (r"n\'t", " not"), (r"\'re", " are"), (r"\'s", " is"), (r"\'d", " would"),
    (r"\'ll", " will"), (r"\'t", " not"), (r"\'ve", " have"), (r"\'m", " am"),
)
_KALDI_NON_WORDS = re.compile(r"[<\[][^>\]]*[>\]]")     # [laughter], <unk>, ...
_WHITESPACE = "\t\n\r\x0b\x0c "


def _expand_contractions(s: str) -> str:
    for pat, sub in _CONTRACTIONS:
        s = re.sub(pat, sub, s)
    return s


def _remove_punctuation(s: str) -> str:
    # jiwer: every code point whose Unicode category starts with "P"
    return "".join(ch for ch in s if not unicodedata.category(ch).startswith("P"))


def _whitespace_to_space(s: str) -> str:
    return "".join(" " if ch in _WHITESPACE else ch for ch in s)


def _collapse_spaces(s: str) -> str:
    return re.sub(r"\s\s+", " ", s)


def wer_standardize(s: str) -> List[str]:
    """reference metrics.py:6-17 -> list of words"""
    s = s.lower()
    s = _expand_contractions(s)
    s = _KALDI_NON_WORDS.sub("", s)
    s = _whitespace_to_space(s)
    s = _collapse_spaces(s)
    s = _remove_punctuation(s)
    s = s.strip()
    return [w for w in s.split(" ") if len(w) >= 1]


def cer_standardize(s: str) -> List[str]:
    """reference metrics.py:41-50 -> list of characters (spaces included, as jiwer's
    ReduceToListOfListOfChars keeps them)"""
    s = s.lower()
    s = _whitespace_to_space(s)
    s = _collapse_spaces(s)
    s = _remove_punctuation(s)
    s = s.strip()
    return list(s)


# ---- alignment ---------------------------------------------------------------------------------

def edit_distance(ref: Sequence, hyp: Sequence) -> int:
    """Levenshtein distance (unit costs) between two token sequences."""
    if len(ref) < len(hyp):
        # the distance is symmetric; keep the inner row short
        ref, hyp = hyp, ref
    prev = list(range(len(hyp) + 1))
    for i, r in enumerate(ref, 1):
        cur = [i] + [0] * len(hyp)
        for j, h in enumerate(hyp, 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (r != h))
        prev = cur
    return prev[-1]


def _score(predictions: Sequence[str], references: Sequence[str], standardize) -> Tuple[float, List[float]]:
    incorrect = 0
    total = 0
    per_utt: List[float] = []
    for prediction, reference in zip(predictions, references):
        ref = standardize(reference)
        hyp = standardize(prediction)
        # reference metrics.py:23-26 / 56-59: an utterance that normalises to nothing scores as the word "EMPTY"
        if not ref:
            ref = standardize("EMPTY")
        if not hyp:
            hyp = standardize("EMPTY")
        d = edit_distance(ref, hyp)
        per_utt.append(d / len(ref))
        incorrect += d
        total += len(ref)
    return incorrect / total, per_utt


def compute_wer(predictions: Sequence[str], references: Sequence[str]) -> Tuple[float, List[float]]:
    """Corpus WER and the per-utterance WERs (reference `compute_wer`, metrics.py:5-38)."""
    return _score(predictions, references, wer_standardize)


def compute_cer(predictions: Sequence[str], references: Sequence[str]) -> Tuple[float, List[float]]:
    """Corpus CER and the per-utterance CERs (reference `compute_cer`, metrics.py:41-71)."""
    return _score(predictions, references, cer_standardize)
