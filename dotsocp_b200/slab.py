"""Time-slab partition helpers for multi-GPU sessions (one process per GPU).

The C library cuts the (nt-1) cell layers into `world` contiguous slabs (solver.cu: dotsocp_create); slab r owns cell
layers [tc0, tc1) and node levels [tn0, tn1) (the last slab also owns level nt-1).  In a distributed session the host
arrays passed to upload/download hold only the slab's owned part:

    phi, c            owned node levels                                (tn1-tn0)*nx*ny
    q, alpha, weight  [ q0 owned cells | bx owned levels | by owned levels ]
    z, beta           ncol columns of (tc1-tc0)*nx*ny doubles, column-major
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib


def default_cuts(nt, world):
    """first cell layer of every slab (world + 1 entries) -- same arithmetic as dotsocp_create"""
    return [r * (nt - 1) // world for r in range(world)] + [nt - 1]


def partition(nt, world, cuts=None):
    """[(tc0, tc1, tn0, tn1)] for every slab.  cuts: explicit partition (default_cuts; a refined session -- Session.refined,
    dotsocp_create_refined -- uses the coarse cuts doubled so that slab r covers the same physical time on every level)."""
    cuts = default_cuts(nt, world) if cuts is None else list(cuts)
    assert len(cuts) == world + 1 and cuts[0] == 0 and cuts[-1] == nt - 1
    return [(cuts[r], cuts[r + 1], cuts[r], nt if r == world - 1 else cuts[r + 1]) for r in range(world)]


def split_state(rank, world, nt, nx, ny, phi, q, z, alpha, beta, c, weight=None, cuts=None):
    """slab-local copies of global arrays (MATLAB linear order); any of the arrays may be None"""
    tc0, tc1, tn0, tn1 = partition(nt, world, cuts)[rank]
    P, PBX, PBY = nx * ny, (nx - 1) * ny, nx * (ny - 1)
    L = (nt - 1) * P
    NBX = nt * PBX

    def node(a):
        return None if a is None else np.ascontiguousarray(a[tn0 * P: tn1 * P])

    def stag(a):
        if a is None:
            return None
        return np.concatenate([a[tc0 * P: tc1 * P], a[L + tn0 * PBX: L + tn1 * PBX], a[L + NBX + tn0 * PBY: L + NBX + tn1 * PBY]])

    def cols(a):
        return None if a is None else np.asfortranarray(a[tc0 * P: tc1 * P, :])
    return (node(phi), stag(q), cols(z), stag(alpha), cols(beta), node(c), stag(weight))


def merge_state(world, nt, nx, ny, parts, ncol=10, cuts=None):
    """inverse of split_state for (phi, q, z, alpha, beta): `parts[r]` is the tuple downloaded by slab r"""
    P, PBX, PBY = nx * ny, (nx - 1) * ny, nx * (ny - 1)
    L = (nt - 1) * P
    NBX = nt * PBX
    Q = L + NBX + nt * PBY
    phi, q, alpha = np.empty(nt * P), np.empty(Q), np.empty(Q)
    z, beta = np.empty((L, ncol), order="F"), np.empty((L, ncol), order="F")
    for r, (tc0, tc1, tn0, tn1) in enumerate(partition(nt, world, cuts)):
        p_phi, p_q, p_z, p_alpha, p_beta = parts[r][:5]
        phi[tn0 * P: tn1 * P] = p_phi
        n0, n1 = (tc1 - tc0) * P, (tn1 - tn0) * PBX
        for dst, src in ((q, p_q), (alpha, p_alpha)):
            dst[tc0 * P: tc1 * P] = src[:n0]
            dst[L + tn0 * PBX: L + tn1 * PBX] = src[n0:n0 + n1]
            dst[L + NBX + tn0 * PBY: L + NBX + tn1 * PBY] = src[n0 + n1:]
        z[tc0 * P: tc1 * P, :] = p_z
        beta[tc0 * P: tc1 * P, :] = p_beta
    return phi, q, z, alpha, beta


def local_sizes(rank, world, nt, nx, ny, cuts=None):
    tc0, tc1, tn0, tn1 = partition(nt, world, cuts)[rank]
    P, PBX, PBY = nx * ny, (nx - 1) * ny, nx * (ny - 1)
    return {"N": (tn1 - tn0) * P, "L": (tc1 - tc0) * P, "Q": (tc1 - tc0) * P + (tn1 - tn0) * (PBX + PBY)}


REUSE_COMM = b"\0" * 128   # pass as nccl_id to re-use the process-wide communicator of an earlier session


def nccl_unique_id():
    """128-byte ncclUniqueId (call on rank 0, broadcast to the others, pass to Session(..., nccl_id=...))."""
    buf = C.create_string_buffer(128)
    check(lib().dotsocp_nccl_unique_id(buf))
    return buf.raw


def broadcast_unique_id(dist, rank):
    """rank 0 creates the id, everybody receives it through the given torch.distributed module (any backend)."""
    obj = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]
