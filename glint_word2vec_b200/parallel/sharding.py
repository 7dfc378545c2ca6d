"""Column (hidden-dimension) sharding of the two embedding matrices.

The reference's core idea: both ``syn0`` and ``syn1neg`` are partitioned by
COLUMN over ``numParameterServers`` shards; shard s holds ``d/S`` columns of
every vocabulary row (README.md:69, MLLIB:207-212).  Here one
shard = one GPU = one process.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class ColumnShard:
    rank: int
    world: int
    vector_size: int        # d (logical columns)
    cols: int               # K: padded columns held by EVERY rank (multiple of 4 -> 16 B rows)
    col_start: int          # first logical column of this shard
    real_cols: int          # logical columns actually owned (<= cols; rest are zero padding)

    @property
    def padded_vector_size(self) -> int:
        return self.cols * self.world


def shard_cols(vector_size: int, world: int) -> int:
    """Columns per rank: ceil(d / S) rounded up to a multiple of 4 floats so
    every row slice is 16-byte aligned (float4 / TMA bulk-copy granularity)."""
    k = -(-vector_size // world)
    return -(-k // 4) * 4


def make_shard(vector_size: int, world: int, rank: int) -> ColumnShard:
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    k = shard_cols(vector_size, world)
    start = min(rank * k, vector_size)
    real = max(0, min(vector_size, (rank + 1) * k) - start)
    return ColumnShard(rank, world, vector_size, k, start, real)
