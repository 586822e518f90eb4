"""Multi-GPU: reads are independent, so chunks shard across ranks with no data-path exchange;
the only collective is ONE all-reduce (NCCL over NVLink) of the final int64 histogram --
the device-level form of the reference's ``streamable(sum)`` / ``bincount_reduce``
(bionumpy/streams/decorators.py:78-110, bionumpy/streams/reductions.py:6-14)."""
import numpy as np
import torch
import torch.distributed as dist

NEWLINE = 10


def shard_records(n_records: int, world_size: int, rank: int):
    """Contiguous, balanced record ranges: rank g takes [first, first + count)."""
    base, rem = divmod(n_records, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def find_fastq_record_start(buf: np.ndarray, pos: int) -> int:
    """Smallest p >= pos such that buf[p:] starts a FASTQ record: the line at p starts with '@',
    the line two below starts with '+', and the sequence and quality lines have equal length
    (a quality line may itself start with '@', so '@' alone is not enough).  -1 if none."""
    n = buf.size
    if pos == 0:
        return 0
    nl = np.flatnonzero(buf[pos - 1:min(n, pos + (1 << 20))] == NEWLINE) + pos - 1
    for i in range(len(nl) - 4):
        p = int(nl[i]) + 1
        if p >= n or buf[p] != ord("@"):
            continue
        l1, l2, l3, l4 = (int(nl[i + j]) for j in (1, 2, 3, 4))
        if buf[l2 + 1] == ord("+") and (l2 - l1) == (l4 - l3):
            return p
    return -1


def shard_byte_ranges(buf: np.ndarray, world_size: int):
    """Split a FASTQ byte buffer into ``world_size`` ranges that start on record boundaries."""
    n = buf.size
    cuts = [0]
    for g in range(1, world_size):
        p = find_fastq_record_start(buf, (n * g) // world_size)
        cuts.append(n if p < 0 else p)
    cuts.append(n)
    cuts = list(np.maximum.accumulate(cuts))
    return [(cuts[g], cuts[g + 1]) for g in range(world_size)]


def all_reduce_histogram(hist: torch.Tensor, group=None) -> torch.Tensor:
    """SUM all-reduce in place (int64).  A no-op outside a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist
