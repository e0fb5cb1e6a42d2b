"""Row sharding of the bank over the GPUs of one box and the candidate exchange between them.

New in the build — the reference is single-device (/root/reference/main.py:58-61).  The bank is cut into contiguous
row blocks (rank r holds rows [r*ceil(N/G), min(N, (r+1)*ceil(N/G)))); every rank scans its block for the SAME query
batch, then one all-gather of the [B, k+skip] sortable u64 candidates (8 bytes each — latency-bound, ≤ 1 MiB) feeds
the k-way merge kernel.  Contiguous blocks make (rank, local row) order equal global row order, so the
"lower index wins ties" rule survives the merge and every rank ends with the identical list.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of the rows rank ``rank`` owns."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    per = -(-n_rows // world_size) if n_rows > 0 else 0
    begin = min(n_rows, rank * per)
    return begin, min(n_rows, begin + per)


def owner_of_row(row: int, n_rows: int, world_size: int) -> int:
    per = -(-n_rows // world_size)
    return row // per


class CandidateExchange:
    """All-gather of each rank's candidate keys ``[B, kk]`` (int64 view of u64) into ``[world, B, kk]``.

    On GPUs this is one NCCL all-gather over NVLink 5 / NVSwitch on the caller's stream; on CPU (gloo) the same
    call path is used by the world_size-2 tests.
    """

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world_size = dist.get_world_size(group)
        else:
            self.rank, self.world_size = 0, 1
        self._buf = None

    def gather(self, keys: torch.Tensor) -> torch.Tensor:
        assert keys.dtype == torch.int64 and keys.dim() == 2 and keys.is_contiguous()
        if self.world_size == 1:
            return keys.unsqueeze(0)
        b, kk = keys.shape
        shape = (self.world_size * b, kk)      # dim-0 concatenation: the layout both NCCL and gloo accept
        if self._buf is None or tuple(self._buf.shape) != shape or self._buf.device != keys.device:
            self._buf = torch.empty(shape, dtype=torch.int64, device=keys.device)
        dist.all_gather_into_tensor(self._buf, keys, group=self.group)
        return self._buf.view(self.world_size, b, kk)
