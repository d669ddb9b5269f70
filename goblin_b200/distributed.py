"""Multi-GPU rendering: samples-per-pixel sharded over one process per GPU.

The reference has one collective-like step, Film::mergeTile (src/GoblinFilm.cpp:
140-153): every worker thread's full-frame (colour, weight) tile is summed into
the film under a mutex.  Across GPUs that is an all-reduce(sum) of the 4 W H
float film buffer.  The library does it itself (gb_film_allreduce: NCCL on the
context's stream, include/goblin_b200.h); torch.distributed is only the launcher's
channel that carries the 128-byte communicator id from rank 0 to the other ranks
(init_film_comm), and the stand-in collective of the CPU (gloo) tests, where
there is no device film.  Geometry is never partitioned: every rank holds a full
scene replica and camera samples are independent, so the film sum is the only
data-path exchange.
"""
import numpy as np


def spp_shard(spp_total, rank, world):
    """Sample-index range [begin, end) of `rank`: contiguous, disjoint, covering [0, spp_total).
    Philox is keyed on (seed, pixel, sample index), so the union over ranks is exactly the
    1-GPU sample set."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    return spp_total * rank // world, spp_total * (rank + 1) // world


class DeviceFilm:
    """The context's device film exposed to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ctx):
        ptr, n_floats = ctx.film_device_ptr()
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3}

    def tensor(self, device):
        import torch
        return torch.as_tensor(self, device=device)


def init_film_comm(ctx, rank, world, group=None):
    """Create the library's own NCCL communicator for `ctx` (gb_comm_init_rank).  The id is made on rank 0
    and broadcast over the already initialised torch.distributed group (any backend)."""
    import torch
    import torch.distributed as dist
    from . import api
    if world < 2:
        return
    if dist.get_backend(group) == "nccl":
        t = torch.zeros(api.COMM_ID_BYTES, dtype=torch.uint8, device=torch.device("cuda", ctx.device))
    else:
        t = torch.zeros(api.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, src=0, group=group)
    ctx.comm_init_rank(bytes(t.cpu().numpy().tobytes()), world, rank)


def allreduce_film(film, group=None):
    """Sum the (r, g, b, weight) film over all ranks, in place.  `film` is a torch tensor (the
    device film on the GPU path, a host tensor in the gloo tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(film, op=dist.ReduceOp.SUM, group=group)
    return film


def render_sharded(ctx, scene, seed, rank, world, spp=None, **render_kw):
    """Render this rank's share of the samples into the context's device film and all-reduce it.
    Asynchronous on the context's own stream, on the context's own device.  With a library communicator
    (init_film_comm) the all-reduce is gb_film_allreduce; otherwise torch.distributed does it on a torch
    ExternalStream wrapping the context's stream, so that it is ordered after the render kernels whatever
    torch's current device / stream are."""
    import torch
    spp_total = scene.spp_squared(spp)
    begin, end = spp_shard(spp_total, rank, world)
    device = torch.device("cuda", ctx.device)
    if ctx.comm_size() == world and world > 1:
        ctx.film_clear()
        ctx.render(seed=seed, spp_total=spp_total, spp_begin=begin, spp_end=end, **render_kw)
        ctx.film_allreduce()
        return DeviceFilm(ctx).tensor(device)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
    with torch.cuda.device(device), torch.cuda.stream(stream):
        ctx.film_clear()
        ctx.render(seed=seed, spp_total=spp_total, spp_begin=begin, spp_end=end, **render_kw)
        film = DeviceFilm(ctx).tensor(device)
        allreduce_film(film)
    return film


def normalize(film_rgbw):
    """Film::writeImage: colour / weight (src/GoblinFilm.cpp:164-173)."""
    film_rgbw = np.asarray(film_rgbw, np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.nan_to_num(film_rgbw[..., :3] / film_rgbw[..., 3:4])
