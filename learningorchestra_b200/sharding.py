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


class PeerReduce:
    """Histogram merge FUSED INTO THE KERNEL'S FLUSH over NVLink peer memory (reduce-to-root).

    The NCCL path is: every rank accumulates a local count matrix, then one all-reduce.  Here there is no
    separate collective: all ranks' fused kernels flush their per-tile bin sums with system-scope RED.64
    straight into ONE count matrix that lives in the root GPU's HBM (mapped into the other processes with
    CUDA IPC) — the reduction is done by the owner's L2 atomic units while the kernels are still streaming.
    What remains is bookkeeping with two 8-byte flags per rank:

      step i, every rank :  [i >= 2: wait local.clean >= i-1]  ->  kernel(counts = root.buf[i%2])
                            ->  release-add root.arrived[i%2] += 1
      step i, root only  :  wait arrived[i%2] >= W*(i//2+1)  ->  copy buf[i%2] to the result  ->  zero buf[i%2]
                            ->  release-add every peer's clean += 1

    Two count buffers alternate so ranks may run up to two steps ahead of the root without waiting; the
    arrival counter is per buffer, so increments of a rank that is already one step ahead can never be
    mistaken for a slower rank's arrival (nobody can start step i+2 before the root finished step i).
    Waits are bounded (``timeout_ms``): a lost peer raises ``timed_out`` instead of hanging the GPU.
    """

    FLAG_BYTES = 256          # u64 slots: [0] arrived(even steps) [1] arrived(odd steps) [2] clean [3] timed_out
    _ARRIVED, _CLEAN, _TIMED_OUT = 0, 16, 24

    def __init__(self, engine, k: int, nbins: int, group=None, root: int = 0, timeout_ms: int = 2000):
        import torch.distributed as dist
        self.engine, self.k, self.nbins, self.root, self.timeout_ms = engine, int(k), int(nbins), root, timeout_ms
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n = self.k * self.nbins
        self.step = 0
        self.flags = engine.dev_alloc(self.FLAG_BYTES)
        self.local_counts = engine.dev_alloc(2 * self.n * 8) if self.rank == root else 0
        self.result = engine.dev_alloc(self.n * 8) if self.rank == root else 0
        mine = {"flags": engine.ipc_export(self.flags),
                "counts": engine.ipc_export(self.local_counts) if self.rank == root else None}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._opened = []
        if self.rank == root:
            self.counts_base = self.local_counts
            self.root_flags = self.flags
            self.peer_clean = []
            for r, h in enumerate(everyone):
                if r != root:
                    p = engine.ipc_open(h["flags"])
                    self._opened.append(p)
                    self.peer_clean.append(p + self._CLEAN)
        else:
            self.counts_base = engine.ipc_open(everyone[root]["counts"])
            self.root_flags = engine.ipc_open(everyone[root]["flags"])
            self._opened += [self.counts_base, self.root_flags]
        dist.barrier(group=group)

    # -- one step -------------------------------------------------------------------------------------
    def counts_for_step(self):
        """DeviceCounts (possibly peer memory) the kernel of the current step must flush into."""
        buf = self.counts_base + (self.step % 2) * self.n * 8
        return self.engine.wrap_counts(self.k, self.nbins, buf)

    def before_kernel(self, stream=None):
        if self.rank != self.root and self.step >= 2:
            self.engine.flag_wait(self.flags + self._CLEAN, self.step - 1, self.flags + self._TIMED_OUT,
                                  self.timeout_ms, stream)

    def after_kernel(self, stream=None):
        eng = self.engine
        parity = self.step % 2
        eng.flag_add(self.root_flags + self._ARRIVED + 8 * parity, 1, stream)    # arrived[parity] += 1 on the root
        if self.rank == self.root:
            buf = self.counts_base + parity * self.n * 8
            # one launch: wait for W arrivals, move the merged counts to `result`, re-zero, signal "clean"
            eng.peer_root_epilogue(self.flags + self._ARRIVED + 8 * parity, self.world * (self.step // 2 + 1),
                                   self.flags + self._TIMED_OUT, buf, self.result, self.n, self.peer_clean,
                                   self.timeout_ms, stream)
        self.step += 1

    # -- results ----------------------------------------------------------------------------------------
    def result_numpy(self, stream=None):
        """Merged counts [k, nbins] of the last finished step (root only)."""
        if self.rank != self.root:
            return None
        return self.engine.read_u64(self.result, self.n, stream).reshape(self.k, self.nbins)

    def timed_out(self, stream=None) -> int:
        return int(self.engine.read_u64(self.flags + self._TIMED_OUT, 1, stream)[0])

    def close(self):
        for p in self._opened:
            try:
                self.engine.ipc_close(p)
            except Exception:
                pass
        self._opened = []
        for p in (self.flags, self.local_counts, self.result):
            if p:
                self.engine.dev_free(p)
        self.flags = self.local_counts = self.result = 0
