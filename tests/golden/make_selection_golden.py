#!/usr/bin/env python
"""Golden vector for export_sol(mode="similarity"): runs the reference's own lib/utils/seq_match.py functions and the
selection loop of lib/routing/env.py:498-545 (restated inline around them) on seeded random tour plans.
Build container only (needs /root/reference); writes tests/golden/selection_similarity.json."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from lib.utils.seq_match import get_similarity_scores, plan_to_string_seq  # noqa: E402  (reference)


def random_plan(rng, n=40, k=6):
    perm = rng.permutation(np.arange(1, n + 1)).tolist()
    cuts = sorted(rng.choice(np.arange(1, n), size=k - 1, replace=False).tolist())
    return [perm[a:b] for a, b in zip([0] + cuts, cuts + [n])]


def main():
    rng = np.random.default_rng(42)
    cases = []
    for P, num_best, num_other in ((20, 2, 3), (12, 3, 3), (8, 1, 4)):
        smps = [random_plan(rng) for _ in range(P)]
        # make some samples near-duplicates of others so that the similarity ranking is non-trivial
        smps[3] = [list(t) for t in smps[0]]
        smps[5] = [list(t) for t in smps[1][:-1]] + [smps[1][-1][::-1]]
        cst = rng.random(P).round(4).tolist()
        bst = np.argsort(cst, kind="stable")[:num_best].tolist()
        best_smps = [smps[i] for i in bst]
        other_smps = [smps[i] for i in range(P) if i not in bst]
        other_costs = [cst[i] for i in range(P) if i not in bst]
        b = [plan_to_string_seq(p) for p in other_smps]
        div = [np.argsort(get_similarity_scores(anchor=plan_to_string_seq(s), candidates=b)) for s in best_smps]
        idx, i, j = [], 0, [0] * num_best
        while len(idx) < num_other:            # env.py:527-539
            d = int(div[i][j[i]])
            if d not in idx:
                idx.append(d)
            else:
                j[i] += 1
                continue
            i = i + 1 if i < num_best - 1 else 0
        cases.append({"samples": smps, "costs": cst, "best_idx": bst, "num_other": num_other,
                      "tours": best_smps + [other_smps[q] for q in idx],
                      "sel_costs": [cst[q] for q in bst] + [other_costs[q] for q in idx]})
    json.dump(cases, open(os.path.join(HERE, "selection_similarity.json"), "w"))
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
