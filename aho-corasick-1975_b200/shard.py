"""Sharding of one text across ranks (one process per GPU, torch.distributed for the plumbing).

The scan path shards into independent units: rank g owns positions [a_g, b_g) and scans
[a_g - lead_g, b_g) from state 0 with lead_g = min(a_g, Lmax-1); occurrences ending inside the lead are dropped by the
library, so every occurrence is reported by exactly one rank.  The only collective is the all-gather of one match count
per rank (exclusive offsets + global total); match lists never cross GPUs, the host concatenates them.
"""
import numpy as np


def plan_shards(nb_symbols, world_size, max_keyword_length, align=16):
    """[(first_owned, end_owned, lead)] for every rank; owned ranges tile [0, nb_symbols); starts are multiples of `align`."""
    per = -(-nb_symbols // world_size)
    per = -(-per // align) * align
    plan = []
    for g in range(world_size):
        a, b = min(g * per, nb_symbols), min((g + 1) * per, nb_symbols)
        lead = min(a, max(max_keyword_length - 1, 0))
        lead = min(a, -(-lead // align) * align)  # keep the shard pointer aligned; a longer lead is harmless
        plan.append((a, b, lead))
    return plan


def sharded_scan(scan_fn, nb_symbols, max_keyword_length, rank, world_size, all_gather_counts):
    """Runs this rank's shard.

    scan_fn(first_symbol, nb_symbols, lead, base) -> records of that window (numpy structured array, end = absolute position);
    all_gather_counts(local_count) -> list of every rank's count (NCCL / gloo all_gather of one uint64 per rank).
    Returns (local_records, exclusive_offset_of_this_rank, global_total).
    """
    a, b, lead = plan_shards(nb_symbols, world_size, max_keyword_length)[rank]
    recs = scan_fn(a - lead, b - (a - lead), lead, a - lead) if b > a else np.zeros(0, dtype=[("end", "<u8"), ("id", "<u4"), ("len", "<u4")])
    counts = [int(c) for c in all_gather_counts(len(recs))]
    return recs, sum(counts[:rank]), sum(counts)
