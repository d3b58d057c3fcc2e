"""Scene sharding across the GPUs of one box: independent units, no data-path collective."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple


def rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when not launched by torchrun."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_round_robin(n_units: int, rank: int, world_size: int) -> List[int]:
    """Indices of the scenes / reference views owned by ``rank`` (round-robin, as SURVEY.md §8e prescribes)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size %d/%d" % (rank, world_size))
    return list(range(rank, n_units, world_size))


def shard_pairs(pairs: Sequence, rank: int, world_size: int):
    """Rows of a filter pair list owned by ``rank``: every GPU holds the full depth stack, reference views are split."""
    idx = shard_round_robin(len(pairs), rank, world_size)
    return [pairs[i] for i in idx], idx
