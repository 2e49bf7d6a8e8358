"""Batch sharding across the GPUs of one box: independent ciphertexts, contiguous split, prepared keys
replicated once per GPU, no per-operation collective.  `torch.distributed` is only plumbing (barrier, max over ranks)."""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [start, stop) of `total` items owned by `rank`; the first `total % world` ranks get one extra item."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} out of range for world size {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (device time) over the process group; identity without a group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else _collective_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _collective_device():
    """Where the tensors of a collective must live: the current CUDA device under NCCL (which cannot move host tensors), the host under gloo."""
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_shards(local, total: int):
    """All-gather variable-length shards (numpy arrays split along axis 0) back into one array on every rank; used only when a
    caller wants the results collocated -- the hot path itself never communicates."""
    import numpy as np
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size()
    outs = [None] * world
    dist.all_gather_object(outs, local)
    res = np.concatenate(outs, axis=0)
    assert res.shape[0] == total
    return res


class _DeviceBytes:
    """__cuda_array_interface__ view of a raw device allocation (lets torch.distributed move it without a copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def broadcast_prepared(module, buf, src: int = 0):
    """Replicate prepared key material (a hal.DevBuf: VmpPMat / SvpPPol data) from rank `src` to every rank of the process group, in
    place, over NCCL (NVLink broadcast at setup -- SURVEY.md section 8e; the hot path itself never communicates).  The prepared layout is
    plain bytes, identical on every GPU, so the key is prepared once and not once per rank.  No-op without a process group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    module.sync()  # the prepare kernels run on the module's stream, the collective on torch's
    t = torch.as_tensor(_DeviceBytes(buf.ptr, buf.nbytes), device=torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(t, src=src)
    torch.cuda.synchronize()


def replicate_prepared(prepare, broadcast=None, src: int = 0) -> str:
    """Prepare key material on rank `src` and replicate it, with every rank taking the SAME branch whatever happens on `src`:

        rank src runs prepare() and catches its failure -> all ranks all-reduce(MIN) a status flag ->
        flag 1: every rank runs broadcast() (the collective that ships the prepared bytes)   -> "broadcast"
        flag 0: nobody enters the broadcast; every rank runs prepare() locally (rank src again: its error, if it persists, is raised
                there and nowhere is a rank left waiting inside a collective)                   -> "local"

    Without a process group: prepare() only -> "local".  (the previous bench code let rank 0 skip the broadcast on an
    exception while the other ranks blocked in it.)"""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        prepare()
        return "local"
    ok = 1
    if dist.get_rank() == src:
        try:
            prepare()
        except Exception as e:  # noqa: BLE001 -- reported through the flag; raised again below on the local path
            import sys

            print(f"[sharding] prepare failed on rank {src}: {e!r}; every rank prepares locally", file=sys.stderr)
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=_collective_device())
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 1:
        if broadcast is not None:
            broadcast()
        return "broadcast"
    prepare()
    return "local"
