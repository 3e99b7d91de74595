"""Staged host<->device transfers (csrc/hostcopy.cu): upload -> download must return the caller's arrays bit for bit, for
pieces that span several ring chunks, with chunk sizes that do not divide them, fresh and in-place destinations."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent("""
    import sys
    import numpy as np
    sys.path.insert(0, %r)
    import dotsocp_b200 as dp
    nt, nx, ny = 9, 513, 257
    rng = np.random.default_rng(11)
    N = nt * nx * ny
    L = (nt - 1) * nx * ny
    Q = L + nt * (nx - 1) * ny + nt * nx * (ny - 1)
    phi, q, alpha = rng.standard_normal(N), rng.standard_normal(Q), rng.standard_normal(Q)
    z = np.asfortranarray(rng.standard_normal((L, 10)))
    beta = np.asfortranarray(rng.standard_normal((L, 10)))
    c = np.zeros(N)
    c[:nx * ny] = rng.standard_normal(nx * ny)
    c[-nx * ny:] = rng.standard_normal(nx * ny)
    for world in (1, 3):
        with dp.Session("dot2d", nt, nx, ny, world=world) as s:
            s.upload(phi, q, z, alpha, beta, c)
            fresh = s.download()
            dst = (np.full(N, np.nan), np.full(Q, np.nan), np.full((L, 10), np.nan, order="F"), np.full(Q, np.nan),
                   np.full((L, 10), np.nan, order="F"))
            inplace = s.download(out=dst)
        for got in (fresh, inplace):
            for a, b, name in zip(got, (phi, q, z, alpha, beta), ("phi", "q", "z", "alpha", "beta")):
                assert np.array_equal(a, b), (world, name)
        assert all(x is y for x, y in zip(inplace, dst))
    # an interior non-zero of model.c is rejected (found by the pool's parallel scan)
    bad = c.copy()
    bad[4 * nx * ny + 5] = 1.0
    with dp.Session("dot2d", nt, nx, ny) as s:
        try:
            s.upload(phi, q, z, alpha, beta, bad)
        except dp.DotsocpError as e:
            assert "interior" in str(e)
        else:
            raise AssertionError("interior entry of c accepted")
    print("HOSTCOPY_OK")
""") % ROOT


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"DOTSOCP_COPY_CHUNK_MB": "1", "DOTSOCP_COPY_THREADS": "3"},
                                 {"DOTSOCP_COPY_CHUNK_MB": "32"},
                                 {"DOTSOCP_HOSTCOPY": "plain"}])
def test_upload_download_roundtrip(gpu, env, tmp_path):
    script = tmp_path / "roundtrip.py"
    script.write_text(SCRIPT)
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, str(script)], env=e, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "HOSTCOPY_OK" in r.stdout, r.stdout + r.stderr
