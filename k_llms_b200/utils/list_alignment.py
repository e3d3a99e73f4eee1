"""List alignment pre-pass (host): make the n candidate lists position-compatible before the element-wise vote.

Restates, in its own structure, what reference consensus_utils.py ("cu") and majority_sorting.py ("ms") do:

    dynamic similarity threshold          cu:185-252   (+ low-end outlier cut cu:152-182)
    reference list from support groups    cu:255-333   (group representative re-elected by the similarity medoid)
    Hungarian assignment to the reference cu:336-380   (scipy.optimize.linear_sum_assignment)
    prune low-support columns             cu:109-149
    order columns by pairwise majority    ms:8-112     (Condorcet-style topological order, ties by mean position)
    recursion + key-path bookkeeping      cu:550-613

SURVEY.md §8f-3 ranks this as a later row of the hot-path table: it is O(n^2 L^2) similarity work per list field
and stays on the host for now.  Behavioural quirks that matter for parity are kept on purpose and marked QUIRK.
"""
from __future__ import annotations

import heapq
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
from scipy.optimize import linear_sum_assignment

from . import similarity

Index = Tuple[int, int]
BASE_THRESHOLD = 0.5


class _PairSims:
    """Symmetric memo of element similarities, addressed by (list, position) pairs (cu:81-106)."""

    def __init__(self, sim_fn: Callable[[Any, Any], float], lists: List[list]):
        self.sim_fn, self.lists, self.memo = sim_fn, lists, {}

    def get(self, a: Index, b: Index) -> float:
        hit = self.memo.get((a, b))
        if hit is None:
            hit = self.memo.get((b, a))
        if hit is None:
            hit = self.sim_fn(self.lists[a[0]][a[1]], self.lists[b[0]][b[1]])
            self.memo[(a, b)] = self.memo[(b, a)] = hit
        return hit


def _low_cutoff(scores: List[float]) -> float:
    """cu:152-174: cut the low tail only if the bottom 20 % shows a jump > 3x the median gap."""
    if len(scores) == 0:
        return 0.0
    s = np.sort(scores)
    cut = s[0]
    gaps = np.diff(s[: int(0.2 * len(s))])
    if len(gaps) > 0:
        jump = np.median(gaps) * 3
        at = int(np.argmax(gaps > jump))
        if gaps[at] > jump:
            cut = s[at + 1] + 0.0001
    return float(cut)


def dynamic_threshold(sims: _PairSims) -> float:
    """cu:185-252: greedy best partner of every element among LATER lists; 0.95 x the lowest representative score."""
    lists = sims.lists
    if not lists or len(lists) < 2:
        return BASE_THRESHOLD
    best_scores: List[float] = []
    for i, li in enumerate(lists):
        if not li:
            continue
        taken: Dict[int, set] = {j: set() for j in range(len(lists)) if j != i}
        for ki in range(len(li)):
            top, partner = BASE_THRESHOLD, None
            for j in range(i + 1, len(lists)):
                for kj in range(len(lists[j])):
                    if kj in taken[j]:
                        continue
                    s = sims.get((i, ki), (j, kj))
                    if s > top:
                        top, partner = s, (j, kj)
            if partner is not None and top > 0:
                best_scores.append(top)
                taken[partner[0]].add(partner[1])
    best_scores.sort()
    floor = _low_cutoff(best_scores)
    kept = [x for x in best_scores if x >= floor]
    if not kept:
        return BASE_THRESHOLD
    return max(BASE_THRESHOLD, 0.95 * kept[0])


def _elect(members: List[Index]) -> Index:
    """Representative of a support group: the reference runs consensus_as_primitive over the (list, pos) TUPLES
    (cu:306-311), i.e. the similarity medoid of index pairs — QUIRK: positions compared as numbers with the 1 %
    tolerance of numerical_similarity, 0 and 0 counting as equal."""
    if len(members) == 1:
        return members[0]
    value, _ = similarity.medoid(list(members), "embeddings", lambda texts: [[0.0] * 10 for _ in texts], 1.0)
    return value


def build_reference(sims: _PairSims, min_support_ratio: float, threshold: float) -> List[Index]:
    """cu:255-333: cluster all elements into support groups (one member per list), keep the well-supported ones."""
    lists = sims.lists
    groups: Dict[Index, List[Index]] = {}
    group_lists: Dict[Index, set] = {}
    for li, lst in enumerate(lists):
        for pos in range(len(lst)):
            cand = (li, pos)
            best, home = -1, None
            for rep, used in group_lists.items():  # dict order matters: the first of equally similar groups wins
                if li in used:
                    continue
                s = sims.get(cand, rep)
                if s >= threshold and s > best:
                    best, home = s, rep
            if home is None:
                groups[cand] = [cand]
                group_lists[cand] = {li}
                continue
            groups[home].append(cand)
            group_lists[home].add(li)
            new_rep = _elect(groups[home])
            if new_rep != home:  # re-keyed groups move to the END of the dicts, exactly like the reference's del + insert
                groups[new_rep] = groups.pop(home)
                group_lists[new_rep] = group_lists.pop(home)
    ratios = {rep: len(members) / len(lists) for rep, members in groups.items()}
    ratios = {rep: r for rep, r in ratios.items() if r >= min_support_ratio}
    return [rep for rep, _ in sorted(ratios.items(), key=lambda kv: (-kv[1], kv[0]))]


def assign_to_reference(sims: _PairSims, reference: List[Index], threshold: float) -> List[list]:
    """cu:336-380: per list, a min-cost assignment of its elements to the reference slots."""
    lists = sims.lists
    aligned = [[None for _ in reference] for _ in lists]
    if not reference:
        return aligned
    for li, lst in enumerate(lists):
        if not lst:
            continue
        sim = np.full((len(reference), len(lst)), -np.inf)
        for r, ref in enumerate(reference):
            for pos in range(len(lst)):
                sim[r, pos] = 1.0 if (li, pos) == ref else sims.get((li, pos), ref)
        rows, cols = linear_sum_assignment(1.0 - sim)
        for r, pos in zip(rows, cols):
            if sim[r, pos] >= threshold and aligned[li][r] is None:
                aligned[li][r] = lst[pos]
    return aligned


def prune_low_support(aligned: List[list], min_support_ratio: float) -> List[list]:
    """cu:109-149: drop columns fewer than min_support_ratio of the lists fill (keep the best if none qualifies)."""
    if not aligned:
        return aligned
    widths = {len(row) for row in aligned}
    if len(widths) != 1:
        return aligned
    width = widths.pop()
    if width == 0:
        return aligned
    support = [sum(1 for row in aligned if row[c] is not None) / len(aligned) for c in range(width)]
    bar = min(min_support_ratio, max(support)) if max(support) < min_support_ratio else min_support_ratio
    keep = [c for c, s in enumerate(support) if s >= bar]
    return [[row[c] for c in keep] for row in aligned]


def original_positions(aligned: List[list], originals: List[list]) -> List[List[Optional[int]]]:
    """ms:8-23: where each aligned cell sat in its own list — QUIRK: matched by object identity (id)."""
    out: List[List[Optional[int]]] = [[None] * len(aligned[0]) for _ in aligned]
    for r, (row, orig) in enumerate(zip(aligned, originals)):
        where = {id(obj): k for k, obj in enumerate(orig)}
        for c, cell in enumerate(row):
            if cell is not None:
                out[r][c] = where.get(id(cell))
    return out


def order_by_majority(aligned: List[list], originals: List[list]):
    """ms:78-112: column a precedes column b when more lists had a's element before b's; cycles and ties fall back
    to the mean original position."""
    if not aligned:
        return aligned, [[None for _ in row] for row in aligned]
    pos = original_positions(aligned, originals)
    width = len(pos[0])
    wins = [[0] * width for _ in range(width)]
    for row in pos:
        seen = [(c, k) for c, k in enumerate(row) if k is not None]
        for a, ka in seen:
            for b, kb in seen:
                if ka < kb:
                    wins[a][b] += 1
    after: List[set] = [set() for _ in range(width)]
    indeg = [0] * width
    for a in range(width):
        for b in range(width):
            if a != b and wins[a][b] > wins[b][a]:
                after[a].add(b)
                indeg[b] += 1
    total, count = [0.0] * width, [0] * width
    for row in pos:
        for c, k in enumerate(row):
            if k is not None:
                total[c] += k
                count[c] += 1
    mean_pos = [total[c] / count[c] if count[c] else float("inf") for c in range(width)]
    heap = [(mean_pos[c], c) for c in range(width) if indeg[c] == 0]
    heapq.heapify(heap)
    order: List[int] = []
    while heap:
        _, u = heapq.heappop(heap)
        order.append(u)
        for v in after[u]:
            indeg[v] -= 1
            if indeg[v] == 0:
                heapq.heappush(heap, (mean_pos[v], v))
    if len(order) < width:  # columns caught in a Condorcet cycle
        order.extend(sorted((c for c in range(width) if c not in order), key=lambda c: mean_pos[c]))
    return [[row[c] for c in order] for row in aligned], [[row[c] for c in order] for row in pos]


def lists_alignment(lists: List[list], sim_fn: Callable[[Any, Any], float], min_support_ratio: float = 0.5,
                    reference_list_idx: Optional[int] = None):
    """cu:383-430.  Returns (aligned lists of equal width, original position of every aligned cell)."""
    if not lists or all(not lst for lst in lists):
        return [[] for _ in lists], [[None for _ in range(len(lst))] for lst in lists]
    sims = _PairSims(sim_fn, lists)
    if reference_list_idx is None:
        thr = dynamic_threshold(sims)
        reference = build_reference(sims, min_support_ratio, thr)
        aligned = assign_to_reference(sims, reference, 0.95 * thr)
        aligned = prune_low_support(aligned, min_support_ratio)
        return order_by_majority(aligned, lists)
    reference = [(reference_list_idx, i) for i in range(len(lists[reference_list_idx]))]
    aligned = assign_to_reference(sims, reference, 0.0)
    return aligned, original_positions(aligned, lists)


def align_list_values(values: list, string_similarity_method: str, embed, client, min_support_ratio: float,
                      max_novelty_ratio: float, current_path: str, reference_idx: Optional[int], recurse: Callable):
    """The list branch of recursive_list_alignments (cu:550-613); `values` is already a private deep copy."""
    key_mappings: Dict[str, list] = {}
    lists_only = [(lst if isinstance(lst, list) else []) for lst in values]
    positions: List[List[Optional[int]]] = [[None for _ in lst] for lst in lists_only]
    if any(lst for lst in lists_only):
        def sim_fn(a, b):
            return similarity.generic_similarity(a, b, string_similarity_method, embed)

        aligned, positions = lists_alignment(lists_only, sim_fn, min_support_ratio, reference_idx)
        for i, row in enumerate(aligned):
            values[i] = row
    else:
        for i in range(len(values)):
            values[i] = []
    if len(values) > 0:
        width = len(values[0])
        if width > 0:
            for c in range(width):
                column, sub_map = recurse([row[c] for row in values], string_similarity_method, embed, client,
                                          min_support_ratio, max_novelty_ratio=max_novelty_ratio, current_path="",
                                          reference_idx=reference_idx)
                for r, cell in enumerate(column):
                    values[r][c] = cell
                for key, sub_values in sub_map.items():  # rebuild the paths against the ORIGINAL positions
                    path = f"{current_path}.{c}" if current_path else str(c)
                    path = f"{path}.{key}" if key else path
                    mapped = []
                    for r, v in enumerate(sub_values):
                        src = positions[r][c]
                        if src is None or v is None:
                            mapped.append(None)
                        else:
                            origin = f"{current_path}.{src}" if current_path else src
                            mapped.append(f"{origin}.{v}" if v else origin)
                    key_mappings[path] = mapped
        elif current_path:
            key_mappings[current_path] = [current_path] * len(values)
    return values, key_mappings
