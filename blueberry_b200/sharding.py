"""Partitioning of contact records across the GPUs of one box.

Every pair is scored independently once (S, possible[], observed[]) are global, so the unit of
sharding is a chromosome (or a diagonal band of a chromosome bigger than a fair share); the only
exchange is the all-reduce of the per-distance table after K1 (engine.reduce_distance_stats).
"""


def lpt_assign(loads, world_size):
    """Longest-processing-time bin packing: returns rank_of_unit (list) minimising the max load greedily."""
    order = sorted(range(len(loads)), key=lambda i: -loads[i])
    totals = [0] * world_size
    owner = [0] * len(loads)
    for i in order:
        r = min(range(world_size), key=lambda k: (totals[k], k))
        owner[i] = r
        totals[r] += loads[i]
    return owner


def split_bands(n_bins, K, parts):
    """Split one chromosome's pairs (i, i+d), 0 <= d <= K, into `parts` contiguous row blocks of i with
    (nearly) equal record counts.  Returns [(row_lo, row_hi), ...]; no halo is needed (pairs are independent)."""
    K = min(K, n_bins - 1)
    total = (K + 1) * n_bins - K * (K + 1) // 2

    def rows_upto(target):          # smallest r with pairs(rows < r) >= target
        lo, hi = 0, n_bins
        while lo < hi:
            mid = (lo + hi) // 2
            full = min(mid, max(n_bins - K, 0))
            tail = mid - full
            pairs = full * (K + 1) + tail * K - tail * (tail - 1) // 2 if tail > 0 else full * (K + 1)
            if pairs >= target:
                hi = mid
            else:
                lo = mid + 1
        return lo

    cuts = [0] + [rows_upto(total * p // parts) for p in range(1, parts)] + [n_bins]
    return [(cuts[i], cuts[i + 1]) for i in range(parts)]
