"""DPR-side helpers of the ANN refresh (BASELINE config 5): checkpoint loading, id mappings and the
answer-string filter.  Same behaviour as the pieces of the reference that
drivers/run_ann_data_gen_dpr.py uses:

  * ``load_states_from_checkpoint`` / ``CheckpointState`` / ``get_model_obj``   utils/dpr_utils.py:23-25,58-59,74-78
  * ``load_mapping``                                                          data/DPR_data.py:132-144
  * ``has_answer`` / ``SimpleTokenizer``                                      utils/dpr_utils.py:241-306

The answer filter is CPU string work (regex tokenisation + token-sequence matching); it stays on the
host.  ``AnswerMatcher`` caches the tokenised passages and answers because the same
passage is tested against many questions during one refresh.
"""
from __future__ import annotations

import collections
import os
import unicodedata
from typing import Dict, List, Sequence, Tuple

import regex
import torch

CheckpointState = collections.namedtuple(
    "CheckpointState", ["model_dict", "optimizer_dict", "scheduler_dict", "offset", "epoch", "encoder_params"])


def load_states_from_checkpoint(model_file: str) -> CheckpointState:
    """utils/dpr_utils.py:74-78: a torch-saved dict with exactly the CheckpointState fields."""
    state_dict = torch.load(model_file, map_location="cpu", weights_only=False)
    return CheckpointState(**state_dict)


def get_model_obj(model):
    return model.module if hasattr(model, "module") else model


def load_mapping(data_dir: str, out_name: str) -> Tuple[Dict[int, int], Dict[int, int]]:
    """data/DPR_data.py:132-144: TSV `pid \\t offset`."""
    pid2offset, offset2pid = {}, {}
    with open(os.path.join(data_dir, out_name), "r") as f:
        for line in f.readlines():
            a = line.split("\t")
            pid2offset[int(a[0])] = int(a[1])
            offset2pid[int(a[1])] = int(a[0])
    return pid2offset, offset2pid


_ALPHA_NUM = r"[\p{L}\p{N}\p{M}]+"
_NON_WS = r"[^\p{Z}\p{C}]"
_TOKEN_RE = regex.compile("(%s)|(%s)" % (_ALPHA_NUM, _NON_WS), flags=regex.IGNORECASE + regex.UNICODE + regex.MULTILINE)


def tokenize_uncased(text: str) -> List[str]:
    """SimpleTokenizer().tokenize(_normalize(text)).words(uncased=True) (utils/dpr_utils.py:253-257,267-306,331-338)."""
    return [m.group().lower() for m in _TOKEN_RE.finditer(unicodedata.normalize("NFD", text))]


def has_answer(answers: Sequence[str], text, tokenizer=None) -> bool:
    """utils/dpr_utils.py:241-264: True iff any answer's token sequence occurs contiguously in the text."""
    if text is None:
        return False
    words = tokenize_uncased(text)
    for single_answer in answers:
        a = tokenize_uncased(single_answer)
        for i in range(0, len(words) - len(a) + 1):
            if a == words[i:i + len(a)]:
                return True
    return False


class AnswerMatcher:
    """has_answer with the tokenisations cached (identical results): one refresh tests the same passage against many
    questions, and most answers' first word does not occur in the passage at all (set test before the scan)."""

    def __init__(self, passages):
        self.passages = passages          # offset -> (text, title), as load_data builds it
        self._ptok: Dict[int, tuple] = {}
        self._atok: Dict[str, List[str]] = {}

    def _p(self, doc_id: int):
        t = self._ptok.get(doc_id)
        if t is None:
            text = self.passages[doc_id][0]
            if text is None:
                t = (None, None)
            else:
                words = tokenize_uncased(text)
                t = (words, frozenset(words))
            self._ptok[doc_id] = t
        return t

    def _a(self, ans: str):
        t = self._atok.get(ans)
        if t is None:
            t = self._atok[ans] = tokenize_uncased(ans)
        return t

    def has_answer(self, answers: Sequence[str], doc_id: int) -> bool:
        words, wset = self._p(doc_id)
        if words is None:
            return False
        for ans in answers:
            a = self._a(ans)
            n = len(a)
            if n == 0:
                return True  # the reference's loop: an empty answer matches at i = 0 when the text is non-empty-ranged
            first = a[0]
            if first not in wset:
                continue
            for i in range(0, len(words) - n + 1):
                if words[i] == first and words[i:i + n] == a:
                    return True
        return False
