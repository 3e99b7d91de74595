"""CPU tests of the multi-GPU host logic: slab partition arithmetic, split/merge of the state, and the rendezvous of
the NCCL id across world_size-2 processes (gloo)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nt,world", [(9, 2), (17, 3), (33, 8), (513, 8), (5, 4)])
def test_partition_covers_the_grid(nt, world):
    from dotsocp_b200 import slab
    p = slab.partition(nt, world)
    assert p[0][0] == 0 and p[-1][1] == nt - 1 and p[-1][3] == nt
    for a, b in zip(p[:-1], p[1:]):
        assert a[1] == b[0] and a[3] == b[2] and a[1] > a[0]
    assert sum(t[1] - t[0] for t in p) == nt - 1 and sum(t[3] - t[2] for t in p) == nt
    sizes = [t[1] - t[0] for t in p]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world", [1, 2, 3])
def test_split_merge_round_trip(world):
    from dotsocp_b200 import slab
    nt, nx, ny = 7, 5, 4
    rng = np.random.default_rng(0)
    P = nx * ny
    L = (nt - 1) * P
    Q = L + nt * (nx - 1) * ny + nt * nx * (ny - 1)
    phi, q, alpha, c = rng.standard_normal(nt * P), rng.standard_normal(Q), rng.standard_normal(Q), rng.standard_normal(nt * P)
    z, beta = np.asfortranarray(rng.standard_normal((L, 10))), np.asfortranarray(rng.standard_normal((L, 10)))
    parts = [slab.split_state(r, world, nt, nx, ny, phi, q, z, alpha, beta, c) for r in range(world)]
    for r in range(world):
        sz = slab.local_sizes(r, world, nt, nx, ny)
        assert parts[r][0].size == sz["N"] and parts[r][1].size == sz["Q"] and parts[r][2].shape == (sz["L"], 10)
    m = slab.merge_state(world, nt, nx, ny, parts)
    for a, b in zip(m, (phi, q, z, alpha, beta)):
        assert np.array_equal(a, b)


def test_two_process_rendezvous_gloo(built, tmp_path):
    """world_size 2 over gloo: rank 0's id reaches rank 1 unchanged and both agree on the partition and on a merged state."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    code = textwrap.dedent("""
        import os, sys, hashlib
        sys.path.insert(0, %r)
        import numpy as np
        import torch.distributed as dist
        from dotsocp_b200 import slab
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        obj = [os.urandom(128) if rank == 0 else None]       # stands in for ncclGetUniqueId (no GPU here)
        dist.broadcast_object_list(obj, src=0)
        ident = obj[0]
        nt, nx, ny = 9, 5, 4
        rng = np.random.default_rng(1)
        P = nx * ny; L = (nt - 1) * P; Q = L + nt * (nx - 1) * ny + nt * nx * (ny - 1)
        phi, q, alpha, c = (rng.standard_normal(n) for n in (nt * P, Q, Q, nt * P))
        z, beta = (np.asfortranarray(rng.standard_normal((L, 10))) for _ in range(2))
        mine = slab.split_state(rank, world, nt, nx, ny, phi, q, z, alpha, beta, c)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine[:5])
        merged = slab.merge_state(world, nt, nx, ny, gathered)
        ok = all(np.array_equal(a, b) for a, b in zip(merged, (phi, q, z, alpha, beta)))
        digests = [None] * world
        dist.all_gather_object(digests, hashlib.sha1(ident).hexdigest())
        assert ok and len(set(digests)) == 1 and len(ident) == 128
        print("rank", rank, "ok")
        dist.destroy_process_group()
    """ % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    script = tmp_path / "rendezvous.py"
    script.write_text(code)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
