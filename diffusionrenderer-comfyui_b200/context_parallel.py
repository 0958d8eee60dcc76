"""Context parallelism of one video's token sequence over the GPUs of a box (csrc/cp.cu).

Tokens are split contiguously along the latent time axis over P ranks (one process per GPU).  Every operator of the
GeneralDIT is token-local except self-attention; the Ulysses exchange around it (tokens-sharded <-> heads-sharded) is
fused into the producing kernels as P2P stores over NVLink into peer-mapped buffers:

    QKV GEMM whose epilogue normalises / rotates q, k and stores every head's rows into the GPU that owns the head
      -> attention over H/P heads x all S tokens, epilogue stores each output row to the GPU that owns the token
      -> out-projection on the local tokens

The ordering between a producer's P2P stores and the consumer on another GPU is folded into the kernels themselves
(`fused_sync`, csrc/cp_sync.cuh): the producer kernel's last CTA publishes a flag to every peer, the consumer kernel's
TMA-producer thread waits for all ranks' flags right before its first operand load.  `barrier()` — one tiny stand-alone
kernel between the stages — is the A/B reference of that and what the ring mode uses.

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

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

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
        self._flags = PeerBuffer(256, group)       # uint32 [CP_FLAG_SLOTS][CP_MAX_RANKS] + status word, zeroed
        self._flag_array = _lib.ptr_array(self._flags.ptrs)
        self._counters = torch.zeros(_lib.CP_FLAG_SLOTS, dtype=torch.int32, device="cuda")   # per-slot CTA counters
        self._epochs = [0] * _lib.CP_FLAG_SLOTS
        self._epoch = 0
        # True: flag waits inside the kernels instead of barrier kernels between them.  Measured slower on B200 (cp8, five
        # batched passes: 68.4 vs 67.2 ms per step; cp2: 268 vs 263): every attention CTA must fence its P2P stores at
        # system scope before it may count itself done — a NVLink round trip at the tail of each of 2 200 CTAs that hold
        # the whole SM — whereas the barrier kernel fences once per launch (profiles/r02_cp8_batch_and_sync_sweep.log).
        self.fused_sync = False
        self.timeout_ms = 60000
        import weakref
        self._finalizer = weakref.finalize(self, ContextParallel._release, self._buffers, self._flags)

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

    def next_epoch(self, slot: int) -> int:
        """the next epoch of a flag slot (every rank calls this in the same order: SPMD)"""
        self._epochs[slot] += 1
        return self._epochs[slot]

    def sync(self, signal=None, wait=None) -> "_lib.CpSync":
        """drb_cp_sync descriptor for one kernel launch: `signal` / `wait` = (slot, epoch) or None"""
        d = _lib.CpSync()
        d.flag_ptrs = ctypes.cast(self._flag_array, ctypes.POINTER(ctypes.c_void_p))
        d.world, d.rank, d.timeout_ms = self.world, self.rank, self.timeout_ms
        d.counter = None
        if signal is not None:
            d.signal_slot, d.signal_epoch = signal
            d.counter = self._counters.data_ptr() + 4 * signal[0]
        if wait is not None:
            d.wait_slot, d.wait_epoch = wait
        return d

    def check(self) -> None:
        """Synchronises the device and raises if an in-kernel wait (or the barrier kernel) gave up on a peer."""
        torch.cuda.synchronize()
        status = self._flags.view((64,), torch.int32)[_lib.CP_STATUS_WORD]
        if int(status.item()) != 0:
            raise RuntimeError("context parallelism: a device-side wait for a peer GPU timed out (a rank died or fell "
                               f"more than {self.timeout_ms / 1000:.0f} s behind); results since then are invalid")

    def all_gather_frames(self, x_local: torch.Tensor) -> torch.Tensor:
        return gather_frames(x_local, self.group)

    @staticmethod
    def _release(buffers, flags) -> None:
        for b in buffers.values():
            b.close()
        buffers.clear()
        flags.close()

    def close(self) -> None:
        """Frees the peer buffers and IPC mappings (also runs when the object is garbage-collected)."""
        self._finalizer()


class EmulatedGroup:
    """P virtual ranks in one process on one GPU: the same kernels and pointer tables, ordinary torch allocations, and
    no barrier (the caller issues every stage for all ranks before the next stage).  Test scaffolding for the sharding
    arithmetic and the two fused-exchange kernels; not a performance path."""

    def __init__(self, world: int, mode: str = "ulysses"):
        self.world, self.mode = world, mode
        self.fused_sync = False
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
        self.fused_sync = False

    def alloc_views(self, name: str, shape, dtype=BF16):
        ts = self._g._alloc(name, shape, dtype)
        return ts[self.rank], list(ts)

    def alloc(self, name: str, shape, dtype=BF16):
        ts = self._g._alloc(name, shape, dtype)
        return ts[self.rank], [t.data_ptr() for t in ts]

    def barrier(self) -> None:
        pass
