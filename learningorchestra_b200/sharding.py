"""Row-range sharding of a table across the GPUs of one box + the single count-matrix all-reduce.

The reference has no multi-device path (one mongod pipeline / three single-core Spark executors,
``projection_image/server.py:58-60``).  Projection and cast are row-independent and the histogram is a
commutative integer sum over rows (``histogram_image/histogram.py:31-36``), so rank r owns the
contiguous rows ``[r*N/W, (r+1)*N/W)`` — boundaries rounded to 32 rows so every shard's slabs keep
their 128-byte alignment — runs the fused kernel on them, and ONE ``all_reduce(SUM)`` of the
``k x nbins`` uint64 count matrix (NCCL over NVLink when the tensor lives on the GPU) yields the
global histogram on every rank.  Integer sums are order independent, so the result is bit-exact
for any world size.  The projected fp32 output stays sharded the same way (the reference's reader
pages by ``_id`` range, ``database_api_image/utils.py:17-23``).
"""
from __future__ import annotations

ROW_ALIGN = 32


def shard_bounds(total_rows: int, world_size: int, rank: int, align: int = ROW_ALIGN) -> tuple[int, int]:
    """Rows [begin, end) owned by ``rank``; every interior boundary is a multiple of ``align``."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    if total_rows < 0:
        raise ValueError("total_rows < 0")

    def cut(r: int) -> int:
        if r >= world_size:
            return total_rows
        return min(total_rows, (total_rows * r // world_size) // align * align)

    return cut(rank), cut(rank + 1)


def all_shard_bounds(total_rows: int, world_size: int, align: int = ROW_ALIGN) -> list[tuple[int, int]]:
    return [shard_bounds(total_rows, world_size, r, align) for r in range(world_size)]


def allreduce_counts(counts_tensor, group=None):
    """In-place SUM all-reduce of a count matrix held in a torch int64 tensor (the uint64 counts' bit
    patterns: two's-complement addition is the same operation).  CUDA tensor + NCCL backend = one
    ncclAllReduce over NVLink on the current stream; CPU tensor + gloo = the host-logic test path."""
    import torch
    import torch.distributed as dist

    if counts_tensor.dtype != torch.int64:
        raise TypeError("counts must be viewed as int64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts_tensor, op=dist.ReduceOp.SUM, group=group)
    return counts_tensor
