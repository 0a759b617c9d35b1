"""CPU (gloo, world_size 2): the page-sharding / host-gather logic used for N > 1 GPUs."""
import os
import socket
import zlib

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from onnxocr_b200 import multigpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _page_result(page):
    return {"crc": zlib.crc32(page.tobytes()), "shape": page.shape}


def _worker(rank, world, port, n_pages, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    pages = [rng.integers(0, 256, (8 + i, 9, 3), dtype=np.uint8) for i in range(n_pages)]
    out = multigpu.ocr_pages_sharded(pages, _page_result)
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pages", [7, 2, 1])
def test_sharded_equals_single(n_pages):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pages, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    pages = [rng.integers(0, 256, (8 + i, 9, 3), dtype=np.uint8) for i in range(n_pages)]
    assert got == [_page_result(p) for p in pages]


def test_shard_indices_partition():
    for n in (0, 1, 5, 16, 17):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in multigpu.shard_indices(n, r, world))
            assert seen == list(range(n))


def test_gather_single_process():
    out = multigpu.gather_by_page([(1, "b"), (0, "a")], 2)
    assert out == ["a", "b"]
