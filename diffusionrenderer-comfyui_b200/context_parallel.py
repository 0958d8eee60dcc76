"""Context parallelism of one video's token sequence over the GPUs of a box (csrc/cp.cu).

Tokens are split contiguously along the latent time axis over P ranks (one process per GPU).  Every operator of the
GeneralDIT is token-local except self-attention; the Ulysses exchange around it (tokens-sharded <-> heads-sharded) is
fused into the producing kernels as P2P stores over NVLink into peer-mapped buffers:

    qk_norm_rope_scatter  ->  [device barrier]  ->  attention over H/P heads x all S tokens, epilogue stores each output
    row to the GPU that owns the token  ->  [device barrier]  ->  out-projection on the local tokens

Two exchange schemes share the plumbing (`mode`):
  "ulysses" (default)  heads are split for the attention; the exchange is fused into the QKV GEMM epilogue and the
                       attention epilogue as P2P stores (4 x S/P x D x 2 B per block and rank);
  "ring"               heads stay whole, every rank keeps its query rows and visits the K/V blocks of all ranks in ring
                       order (r, r-1, ...): the next block is pulled from its owner's memory over NVLink on a side stream
                       while the current one is attended to, and the attention kernel's ring epilogue merges the blocks
                       into an fp32 running state (2 x (P-1)/P x S x D x 2 B per block and rank — 3.5x the Ulysses traffic
                       at P = 8, but no constraint that P divide the head count).

torch.distributed (NCCL) is plumbing only: it carries the 64-byte CUDA IPC handles of the peer buffers at start-up and
the final all-gather of the (tiny) latent.  `EmulatedGroup` runs the same kernels for P virtual ranks inside one process
on one GPU (no barrier needed: the stages are issued rank after rank on one stream) — used by the single-GPU tests.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

BF16 = torch.bfloat16


def shard_frames(total_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """[t0, t1) of the latent frames owned by `rank`; the split must be even (every rank runs the same kernels)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    if total_frames % world:
        raise ValueError(f"context parallelism over {world} GPUs needs the {total_frames} latent frames to split evenly")
    n = total_frames // world
    return rank * n, (rank + 1) * n


def head_owner(head: int, num_heads: int, world: int) -> Tuple[int, int]:
    """(owning rank, local head index) of a head: contiguous blocks of H/P heads"""
    if num_heads % world:
        raise ValueError(f"{num_heads} heads do not split over {world} ranks")
    per = num_heads // world
    return head // per, head % per


def gather_frames(x_local: torch.Tensor, group=None) -> torch.Tensor:
    """[..., T/P, H, W] per rank -> [..., T, H, W] on every rank, frames in rank order (3.6 MB per pass at 57x704x1280,
    once per sampler run)"""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [torch.empty_like(x_local) for _ in range(world)]
    dist.all_gather(parts, x_local.contiguous(), group=group)
    return torch.cat(parts, dim=-3)


class _RawTensor:
    """a device allocation we own, exposed to torch through __cuda_array_interface__"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerBuffer:
    """`nbytes` of zeroed device memory on every rank of the group, each mapped into every other rank (CUDA IPC)."""

    def __init__(self, nbytes: int, group=None):
        import torch.distributed as dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nbytes = nbytes
        p = ctypes.c_void_p()
        _lib.check(_lib.load().drb_peer_alloc(nbytes, ctypes.byref(p)), "drb_peer_alloc")
        self.local_ptr = p.value
        handle = (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES)()
        _lib.check(_lib.load().drb_peer_export(self.local_ptr, handle), "drb_peer_export")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device="cuda")
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=group)
        self.ptrs: List[int] = []
        self._imported: List[int] = []
        for r, h in enumerate(allh):
            if r == self.rank:
                self.ptrs.append(self.local_ptr)
                continue
            raw = (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES)(*h.cpu().tolist())
            q = ctypes.c_void_p()
            _lib.check(_lib.load().drb_peer_import(raw, ctypes.byref(q)), "drb_peer_import")
            self.ptrs.append(q.value)
            self._imported.append(q.value)
        self._holders = [_RawTensor(p, nbytes) for p in self.ptrs]
        self.peer_bytes = [torch.as_tensor(h, device="cuda") for h in self._holders]     # every rank's buffer, mapped here
        self.bytes_tensor = self.peer_bytes[self.rank]

    def view(self, shape, dtype, rank: Optional[int] = None) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        src = self.bytes_tensor if rank is None else self.peer_bytes[rank]
        return src[: n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)

    def close(self) -> None:
        lib = _lib.load()
        for p in self._imported:
            lib.drb_peer_close(p)
        self._imported = []
        if self.local_ptr:
            lib.drb_peer_free(self.local_ptr)
            self.local_ptr = 0


class ContextParallel:
    """Per-rank handle of a context-parallel group over torch.distributed (real multi-GPU)."""

    def __init__(self, group=None, mode: str = "ulysses"):
        import torch.distributed as dist
        if mode not in ("ulysses", "ring"):
            raise ValueError("mode must be 'ulysses' or 'ring'")
        if not dist.is_initialized():
            raise RuntimeError("context parallelism needs an initialised torch.distributed process group (backend nccl)")
        self.mode = mode
        self.copy_stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.CP_MAX_RANKS:
            raise ValueError(f"at most {_lib.CP_MAX_RANKS} ranks")
        self._buffers: Dict[str, PeerBuffer] = {}
        self._flags = PeerBuffer(256, group)
        self._epoch = 0

    def alloc(self, name: str, shape, dtype=BF16) -> Tuple[torch.Tensor, List[int]]:
        """collective: every rank calls it with the same arguments in the same order"""
        nbytes = torch.empty((), dtype=dtype).element_size()
        for s in shape:
            nbytes *= s
        old = self._buffers.get(name)
        if old is None or old.nbytes < nbytes:
            if old is not None:
                old.close()
            self._buffers[name] = old = PeerBuffer(nbytes, self.group)
        return old.view(shape, dtype), list(old.ptrs)

    def alloc_views(self, name: str, shape, dtype=BF16) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        """like alloc, but returns every rank's buffer as a tensor (P2P-mapped): what the ring pulls K/V blocks from"""
        local, _ = self.alloc(name, shape, dtype)
        buf = self._buffers[name]
        return local, [buf.view(shape, dtype, r) for r in range(self.world)]

    def barrier(self) -> None:
        """device-side: enqueues one tiny kernel on the current stream; the host does not wait"""
        self._epoch += 1
        _lib.call("drb_cp_barrier", _lib.ptr_array(self._flags.ptrs), self.rank, self.world, self._epoch,
                  torch.cuda.current_stream().cuda_stream)

    def all_gather_frames(self, x_local: torch.Tensor) -> torch.Tensor:
        return gather_frames(x_local, self.group)

    def close(self) -> None:
        for b in self._buffers.values():
            b.close()
        self._buffers = {}
        self._flags.close()


class EmulatedGroup:
    """P virtual ranks in one process on one GPU: the same kernels and pointer tables, ordinary torch allocations, and
    no barrier (the caller issues every stage for all ranks before the next stage).  Test scaffolding for the sharding
    arithmetic and the two fused-exchange kernels; not a performance path."""

    def __init__(self, world: int, mode: str = "ulysses"):
        self.world, self.mode = world, mode
        self._tensors: Dict[str, List[torch.Tensor]] = {}
        self.copy_stream = torch.cuda.Stream()
        self.ranks = [_EmulatedRank(self, r) for r in range(world)]

    def _alloc(self, name, shape, dtype):
        ts = self._tensors.get(name)
        if ts is None or tuple(ts[0].shape) != tuple(shape):
            ts = self._tensors[name] = [torch.zeros(*shape, device="cuda", dtype=dtype) for _ in range(self.world)]
        return ts


class _EmulatedRank:
    def __init__(self, group: EmulatedGroup, rank: int):
        self._g, self.rank, self.world, self.mode, self.copy_stream = group, rank, group.world, group.mode, group.copy_stream

    def alloc_views(self, name: str, shape, dtype=BF16):
        ts = self._g._alloc(name, shape, dtype)
        return ts[self.rank], list(ts)

    def alloc(self, name: str, shape, dtype=BF16):
        ts = self._g._alloc(name, shape, dtype)
        return ts[self.rank], [t.data_ptr() for t in ts]

    def barrier(self) -> None:
        pass
