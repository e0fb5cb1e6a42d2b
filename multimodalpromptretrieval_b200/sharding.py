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


def plan_shard_reads(file_ranges, begin: int, end: int):
    """Which slices of which shard files cover rows [begin, end)?  ``file_ranges`` = [(row_begin, row_end), ...] of
    the files in a shard directory (any world size); returns [(file_index, first_row_in_file, n_rows, dst_offset)]."""
    out = []
    for i, (fb, fe) in enumerate(file_ranges):
        lo, hi = max(begin, fb), min(end, fe)
        if lo < hi:
            out.append((i, lo - fb, hi - lo, lo - begin))
    covered = sum(n for _, _, n, _ in out)
    if covered != max(0, end - begin):
        raise ValueError(f"shard files cover {covered} of the {end - begin} rows [{begin}, {end})")
    return out


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


class P2PExchange:
    """Peer-memory candidate exchange: one symmetric-memory buffer per rank, mapped into every peer.  The exchange itself
    runs inside the retrieval kernel's tail (csrc/tail.cuh): a warp stores its query's candidates straight into every
    peer's buffer over NVLink as epoch-tagged 8-byte words, polls its own buffer for the peers' words and merges — no
    fence, no flag, no collective library on
    the data path, epoch kept on the device.  This class only owns the buffer and the table of peer pointers the C ABI
    takes (``mpr_retrieve_args.peer_bufs``).  ``cap`` (>= b * kk of any search) is fixed at construction: growing it
    would be an implicit collective (rendezvous + barrier) in the middle of a search.

    With ``world_size == 1`` it degenerates to a self-exchange through an ordinary device buffer (single-GPU tests)."""

    def __init__(self, device: torch.device, cap: int, group: Optional[dist.ProcessGroup] = None,
                 world_size: Optional[int] = None, rank: int = 0):
        import ctypes as C
        from . import kernels as K
        self.cap = int(cap)
        if world_size is not None:                       # explicit (tests that play a multi-rank exchange on one GPU)
            self.rank, self.world_size = int(rank), int(world_size)
            distributed = False
        elif dist.is_available() and dist.is_initialized():
            self.rank, self.world_size = dist.get_rank(group), dist.get_world_size(group)
            distributed = self.world_size > 1
        else:
            self.rank, self.world_size = 0, 1
            distributed = False
        nbytes = K.exchange_bytes(self.world_size, self.cap)
        if nbytes == 0:
            raise ValueError(f"bad exchange geometry world={self.world_size} cap={self.cap}")
        if not distributed:
            self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
            self.peer_ptrs = [self.buf.data_ptr()] * self.world_size
            self._hdl = None
        else:
            import torch.distributed._symmetric_memory as symm_mem
            self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
            self.buf.zero_()
            self._hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
            self.peer_ptrs = [int(p) for p in self._hdl.buffer_ptrs]
            torch.cuda.synchronize(device)
            dist.barrier(group)                 # nobody pushes before every buffer is zeroed
        self.c_ptrs = (C.c_void_p * self.world_size)(*[C.c_void_p(p) for p in self.peer_ptrs])

    def fill_args(self, args) -> None:
        """Point a ``RetrieveArgs`` block at this exchange."""
        args.rank, args.world, args.xchg_cap = self.rank, self.world_size, self.cap
        args.peer_bufs = self.c_ptrs
