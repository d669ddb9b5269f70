"""The N > 1 path: spp sharding + film all-reduce.

CPU (gloo, world_size 2): the host-side logic of goblin_b200.distributed with the oracle port
standing in for the device renderer (there is no CPU renderer in the product).
GPU (marked gpu, needs >= 2 devices): the real thing over NCCL."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from goblin_b200 import api, distributed
from tests import oracle_port as op
from tests import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_spp_shard_is_a_partition():
    for spp in (1, 4, 16, 100, 4096):
        for world in (1, 2, 3, 4, 8):
            got = [distributed.spp_shard(spp, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == spp
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [e - b for b, e in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        distributed.spp_shard(4, 2, 2)


WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from goblin_b200 import api, distributed
from tests import oracle_port as op, util
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
scene = api.Scene(util.TINY_PT)
spp = 16
b, e = distributed.spp_shard(spp, rank, world)
film, counters, _ = op.render(scene, seed=5, spp_total=spp, spp_begin=b, spp_end=e, threads=2)
t = torch.from_numpy(film.reshape(-1))
distributed.allreduce_film(t)
n = torch.tensor([counters["camera_samples"]], dtype=torch.int64)
dist.all_reduce(n)
if rank == 0:
    np.save({out!r}, film)
    open({out!r} + ".n", "w").write(str(int(n.item())))
dist.destroy_process_group()
"""


def test_gloo_world2_film_allreduce(built, tmp_path):
    out = str(tmp_path / "film.npy")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=util.ROOT, out=out))
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=300)
        assert p.returncode == 0, o
    scene = api.Scene(util.TINY_PT)
    whole, counters, _ = op.render(scene, seed=5, spp_total=16, threads=2)
    merged = np.load(out)
    assert int(open(out + ".n").read()) == counters["camera_samples"]
    assert np.allclose(merged, whole, rtol=1e-5, atol=1e-6)


GPU_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from goblin_b200 import api, distributed
from tests import util
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
scene = api.Scene(util.TINY_PT)
ctx = api.Context(rank)
ctx.upload_scene(scene)
# first through torch.distributed's all-reduce on the context's stream ...
film = distributed.render_sharded(ctx, scene, seed=5, rank=rank, world=world, spp=16)
ctx.synchronize()
torch.cuda.synchronize()
via_torch = film.cpu().numpy().copy()
# ... then through the library's own NCCL communicator (gb_comm_init_rank + gb_film_allreduce)
distributed.init_film_comm(ctx, rank, world)
assert ctx.comm_size() == world
film = distributed.render_sharded(ctx, scene, seed=5, rank=rank, world=world, spp=16)
ctx.synchronize()
via_lib = ctx.film_download().reshape(-1)
if rank == 0:
    np.save({out!r}, via_lib)
    np.save({out!r} + ".torch.npy", via_torch)
ctx.comm_destroy()
dist.destroy_process_group()
"""


@pytest.mark.gpu
def test_nccl_world2_matches_single_gpu(built, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "film.npy")
    script = tmp_path / "worker.py"
    script.write_text(GPU_WORKER.format(root=util.ROOT, out=out))
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        o, _ = p.communicate(timeout=600)
        assert p.returncode == 0, o
    scene = api.Scene(util.TINY_PT)
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    ctx.film_clear()
    ctx.render(seed=5, spp_total=16)
    whole = ctx.film_download().reshape(-1)
    assert np.allclose(np.load(out), whole, rtol=1e-4, atol=1e-5)             # gb_film_allreduce
    assert np.allclose(np.load(out + ".torch.npy"), whole, rtol=1e-4, atol=1e-5)  # torch.distributed on the same stream


@pytest.mark.gpu
def test_single_process_two_contexts_film_allreduce(built):
    """g_ray --gpus N's merge: one process, one context per GPU, gb_comm_init_all + gb_film_allreduce_all
    (Film::mergeTile, src/GoblinFilm.cpp:140-153).  Every GPU ends up holding the 1-GPU film."""
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    scene = api.Scene(util.TINY_PT)
    ctxs = [api.Context(g) for g in range(2)]
    for c in ctxs:
        c.upload_scene(scene)
    api.comm_init_all(ctxs)
    assert api.nccl_version() > 20000 and all(c.comm_size() == 2 for c in ctxs)
    for g, c in enumerate(ctxs):
        b, e = distributed.spp_shard(16, g, 2)
        c.film_clear()
        c.render(seed=5, spp_total=16, spp_begin=b, spp_end=e)
    api.film_allreduce_all(ctxs)
    films = [c.film_download() for c in ctxs]
    one = api.Context(0)
    one.upload_scene(scene)
    one.film_clear()
    one.render(seed=5, spp_total=16)
    whole = one.film_download()
    for f in films:
        assert np.allclose(f, whole, rtol=1e-4, atol=1e-5)
    for c in ctxs:
        c.close()


def test_film_comm_entry_points_without_a_gpu(built):
    """The collective entry points exist, NCCL is bound at run time, and nothing works without a device."""
    import torch
    v = api.nccl_version()
    assert v > 20000
    assert len(api.comm_unique_id()) == api.COMM_ID_BYTES
    if not torch.cuda.is_available():
        with pytest.raises(api.GoblinError):
            api.Context(0)
