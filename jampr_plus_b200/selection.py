"""Host-side sample selection of ``RPEnv.export_sol(mode="similarity")`` (reference env.py:498-545 with
lib/utils/seq_match.py:8-43): keep the ``num_best`` cheapest samples and add the ``num_other`` samples whose tour plans
are the least similar to them, by difflib sequence matching on the flattened plans, picked cyclically per best sample."""
import difflib
import itertools as it
from typing import List, Sequence, Tuple

import numpy as np

__all__ = ["plan_to_string_seq", "similarity_scores", "select_similarity"]


def plan_to_string_seq(plan: Sequence[Sequence[int]]) -> str:
    return " ".join(str(t) for t in it.chain.from_iterable(plan))


def similarity_scores(anchor: str, candidates: Sequence[str]) -> List[float]:
    matcher = difflib.SequenceMatcher(isjunk=lambda x: x == " ", a=anchor)
    out = []
    for c in candidates:
        matcher.set_seq2(c)
        out.append(matcher.ratio())
    return out


def select_similarity(samples: List[List[List[int]]], costs: Sequence[float], best_idx: Sequence[int],
                      num_other: int) -> Tuple[List[List[List[int]]], List[float]]:
    """``samples``: the tour plans of all samples of one instance, ``costs`` their totals, ``best_idx`` the indices
    of the cheapest ones (ascending cost).  Returns (selected plans, their costs), best first."""
    best_idx = [int(i) for i in best_idx]
    num_best = len(best_idx)
    best_smps = [samples[i] for i in best_idx]
    best_costs = [costs[i] for i in best_idx]
    rest = [i for i in range(len(samples)) if i not in best_idx]
    other_smps = [samples[i] for i in rest]
    other_costs = [costs[i] for i in rest]
    b = [plan_to_string_seq(p) for p in other_smps]
    order = [np.argsort(similarity_scores(plan_to_string_seq(s), b)) for s in best_smps]
    picked: List[int] = []
    i, j = 0, [0] * num_best
    while len(picked) < num_other:
        cand = int(order[i][j[i]])
        if cand in picked:
            j[i] += 1
            continue
        picked.append(cand)
        i = i + 1 if i < num_best - 1 else 0
    return best_smps + [other_smps[q] for q in picked], best_costs + [other_costs[q] for q in picked]
