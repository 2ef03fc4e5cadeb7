"""Host-side multi-GPU plumbing: independent work items (frames of a 2D+t stack, label crops) are sharded over the
ranks of one torchrun launch, item i -> rank i mod world; there is NO collective on the data path.
Results are assembled on rank 0, which alone writes the output file, exactly one writer as in the reference's loop
(infer_script_local.py:115-165 fills ``results_array[T,H,W]`` and writes one TIFF).

Assembly: ranks of one box write their rows straight into a shared-memory array (``/dev/shm``: zero copy, no
serialisation); ranks on different hosts send their rows to rank 0 over the CPU (gloo) backend.
"""
import os
import socket
import uuid

import numpy as np
import torch


def dist_info():
    """(rank, world_size) of the initialised default process group, (0, 1) without one."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env():
    """Initialise torch.distributed from a torchrun environment (RANK / WORLD_SIZE / LOCAL_RANK), one process per GPU.
    Returns (rank, world, local_rank).  No-op (0, 1, 0) for a plain ``python`` launch."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
    return rank, world, local


def shard_indices(n, rank, world):
    """Work items of ``rank``: i = rank (mod world)."""
    return list(range(rank, n, world))


def _object_device():
    import torch.distributed as dist
    # object collectives stage through tensors on the backend's device
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _barrier():
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        dist.barrier(device_ids=[torch.cuda.current_device()])
    else:
        dist.barrier()


def run_sharded(fn, n, specs, transport="auto"):
    """Run ``fn(indices, outs)`` on this rank's shard and assemble the results on rank 0.

    ``specs``: list of (row_shape, dtype); ``outs[k]`` has shape (n,) + row_shape and ``fn`` fills rows ``indices`` of
    every array.  Returns the list of full arrays on rank 0 and None on the other ranks.  Without a process group
    (or world_size 1) ``fn`` simply gets all indices."""
    import torch.distributed as dist
    rank, world = dist_info()
    if world == 1:
        outs = [np.zeros((n,) + tuple(s), dtype=d) for s, d in specs]
        fn(list(range(n)), outs)
        return outs
    mine = shard_indices(n, rank, world)
    if transport == "auto":
        hosts = [None] * world
        dist.all_gather_object(hosts, socket.gethostname())
        transport = "shm" if len(set(hosts)) == 1 and os.path.isdir("/dev/shm") else "p2p"
    if transport == "shm":
        names = [None]
        if rank == 0:
            names = [[f"/dev/shm/mbseg_{os.getpid()}_{uuid.uuid4().hex}_{k}" for k in range(len(specs))]]
            for path, (s, d) in zip(names[0], specs):
                np.lib.format.open_memmap(path, mode="w+", dtype=np.dtype(d), shape=(n,) + tuple(s))
        dist.broadcast_object_list(names, src=0, device=_object_device())
        outs = [np.load(path, mmap_mode="r+") for path in names[0]]
        try:
            fn(mine, outs)
            for o in outs:
                o.flush()
            _barrier()                                   # every rank's rows are in the shared arrays
            # rank 0 keeps the mapping (no copy of the gathered data); the names are removed below, the pages live as
            # long as the returned arrays do
            result = outs if rank == 0 else None
        finally:
            del outs
            _barrier()
            if rank == 0:
                for path in names[0]:
                    try:
                        os.unlink(path)
                    except OSError:
                        pass
        return result
    # point to point over the CPU backend: rank r sends its rows (in shard order) to rank 0
    local = [np.zeros((len(mine),) + tuple(s), dtype=d) for s, d in specs]

    class _View:
        """maps global row indices onto the rank-local buffers"""

        def __init__(self, buf):
            self.buf, self.pos = buf, {t: j for j, t in enumerate(mine)}

        def __setitem__(self, t, v):
            self.buf[self.pos[int(t)]] = v

        def __getitem__(self, t):
            return self.buf[self.pos[int(t)]]

    fn(mine, [_View(b) for b in local])
    if rank != 0:
        for b in local:
            if b.size:
                dist.send(torch.from_numpy(b.view(np.uint8).reshape(-1)), dst=0)
        return None
    outs = [np.zeros((n,) + tuple(s), dtype=d) for s, d in specs]
    for k, b in enumerate(local):
        outs[k][mine] = b
    for r in range(1, world):
        theirs = shard_indices(n, r, world)
        for k, (s, d) in enumerate(specs):
            if not theirs:
                continue
            buf = np.zeros((len(theirs),) + tuple(s), dtype=d)
            dist.recv(torch.from_numpy(buf.view(np.uint8).reshape(-1)), src=r)
            outs[k][theirs] = buf
    return outs
