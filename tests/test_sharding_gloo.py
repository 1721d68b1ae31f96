"""Multi-rank host logic on CPU: world_size-2 gloo run of the scene sharding + final metric gather."""
import os
import socket
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from infgen_b200.sharding import shard_scenes, gather_scene_metrics


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_scenes, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ids = shard_scenes(n_scenes, rank, world)
    vals = torch.tensor([[float(i) * 10.0, float(rank)] for i in ids]).reshape(len(ids), 2)
    full = gather_scene_metrics(ids, vals, n_scenes)
    q.put((rank, ids, full.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_covers_every_scene_once():
    for n, w in ((256, 8), (7, 2), (1, 4), (64, 8)):
        seen = sorted(i for r in range(w) for i in shard_scenes(n, r, w))
        assert seen == list(range(n))


def test_gather_world1():
    full = gather_scene_metrics([0, 1, 2], torch.tensor([[1.0], [2.0], [3.0]]), 3)
    assert full.flatten().tolist() == [1.0, 2.0, 3.0]


def test_gloo_world2_gather():
    world, n_scenes = 2, 7
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_scenes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[i * 10.0, float(i % world)] for i in range(n_scenes)]
    for rank, ids, full in res:
        assert ids == list(range(rank, n_scenes, world))
        assert full == want
