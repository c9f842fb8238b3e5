"""CPU tier, world_size 2 (and 3) over gloo: the N > 1 host logic -- row ownership, the framebuffer gather and
the max reduction -- without a GPU.  The device side of the same path (a row shard rendered by mirogpu_render is
bit-identical to those rows of the full frame) is covered in tests/test_gpu_render.py."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import importlib, os, sys
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    sh = importlib.import_module("cse168-raytracer_b200.sharding")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    H, W = 37, 5
    full = torch.arange(H * W * 3, dtype=torch.float32).reshape(H, W, 3)
    rows, n = sh.rows_of_rank(H, world, rank)
    assert rows == (0, H, world, rank) and n == len(range(rank, H, world))
    local = full[rank::world].clone()
    assert local.shape[0] == n
    got = sh.gather_rows(local, H, world, rank)
    assert torch.equal(got, full), "gathered frame differs"
    got8 = sh.gather_rows((local %% 251).to(torch.uint8), H, world, rank)
    assert torch.equal(got8, (full %% 251).to(torch.uint8))
    # the preallocated gather of bench.py's e2e path: the frame holds this rank's rows, junk elsewhere; both row-count cases
    for hh in (H, 6 * world):
        ref = (torch.arange(hh * W * 3, dtype=torch.int64).reshape(hh, W, 3) %% 251).to(torch.uint8)
        g = sh.RowGather(hh, W, 3, torch.uint8, "cpu", world, rank)
        for rep in range(2):
            frame = torch.full((hh, W, 3), 7, dtype=torch.uint8)
            frame[rank::world] = ref[rank::world]
            assert torch.equal(g(frame), ref), "RowGather frame differs"
    m = sh.reduce_max(torch.tensor([float(rank) + 0.5]))
    assert float(m) == world - 0.5
    # every row is owned by exactly one rank
    owner = torch.zeros(H, dtype=torch.int64); owner[rank::world] += 1
    dist.all_reduce(owner)
    assert bool((owner == 1).all())
    dist.barrier()
    if rank == 0:
        print("GLOO_OK", world)
''') % ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharding_and_gather_over_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert f"GLOO_OK {world}" in r.stdout


def test_rows_partition_the_frame():
    import importlib
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("cse168-raytracer_b200.sharding")
    for h in (1, 7, 1080):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                (rb, re_, rs, rp), n = sh.rows_of_rank(h, world, r)
                rows = list(range(rb + rp, re_, rs))
                assert len(rows) == n <= sh.max_rows(h, world)
                seen += rows
            assert sorted(seen) == list(range(h))
