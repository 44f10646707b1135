"""World-size-2 CPU test (gloo) of the multi-GPU decomposition: the slab layout, the halo sets and
the reduction points the GPU library uses (csrc/ec3d_gpu.cu h_halo / h_allreduce), run here as a
numpy BiCGSTABwr over the oracle's CSR rows with torch.distributed doing the halo exchange and the
scalar all-reduces.  Result must match the single-process oracle."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eddy_currents_3d_b200 import plate
    from eddy_currents_3d_b200.slab import make_slabs
    from oracle import oracle
    p = plate(32, "B")
    run = oracle.OracleRun(p)
    run.step(solve=False)
    b = run.rhs.copy()
    A = run.A
    n = p.nCellsGlob
    slabs = make_slabs(p, world)
    me = slabs[rank]
    own = np.zeros(n, bool)
    for s in me.owned:
        own[s] = True
    visible = own.copy()
    for s in me.halo_lo + me.halo_hi:
        visible[s] = True
    rows = np.flatnonzero(own)
    irow0 = A.irow.astype(np.int64) - 1
    # halo depth check: every owned row only references owned or halo columns
    for r in rows[:: max(1, rows.size // 4000)]:
        cols = A.jcol[irow0[r]:irow0[r + 1]] - 1
        assert visible[cols].all(), (rank, r)
    import scipy.sparse as sp
    M = sp.csr_matrix((A.valA, A.jcol - 1, irow0), shape=(n, n))[rows]

    def halo(v):
        reqs = []
        if rank > 0:
            for s_out, s_in in zip(me.send_lo, me.halo_lo):
                reqs.append(dist.isend(torch.from_numpy(v[s_out].copy()), rank - 1))
                buf = torch.empty(s_in.stop - s_in.start, dtype=torch.float64)
                reqs.append((dist.irecv(buf, rank - 1), s_in, buf))
        if rank < world - 1:
            for s_out, s_in in zip(me.send_hi, me.halo_hi):
                reqs.append(dist.isend(torch.from_numpy(v[s_out].copy()), rank + 1))
                buf = torch.empty(s_in.stop - s_in.start, dtype=torch.float64)
                reqs.append((dist.irecv(buf, rank + 1), s_in, buf))
        for q in reqs:
            if isinstance(q, tuple):
                q[0].wait(); v[q[1]] = q[2].numpy()
            else:
                q.wait()

    def spmv(v):
        halo(v)
        out = np.zeros(n)
        out[rows] = M @ np.where(visible, v, np.nan)      # NaN would expose a missing halo entry
        return out

    def dots(*pairs):
        t = torch.tensor([float(np.dot(a[rows], c[rows])) for a, c in pairs], dtype=torch.float64)
        dist.all_reduce(t)
        return [float(x) for x in t]

    tol, itmax = p.tolerance, p.itmax
    x = np.zeros(n)
    R = np.zeros(n); R[rows] = b[rows] - spmv(x)[rows]
    R0, P = R.copy(), R.copy()
    bb, rr0 = dots((b, b), (R, R0))
    bnorm = np.sqrt(bb)
    it = 0
    while True:
        if it > itmax:
            break
        it += 1
        AP = spmv(P)
        (apr0,) = dots((AP, R0))
        alpha = rr0 / apr0
        S = R - alpha * AP
        (ss,) = dots((S, S))
        if np.sqrt(ss) / bnorm < tol:
            x = x + alpha * P
            break
        AS = spmv(S)
        ass, asas = dots((AS, S), (AS, AS))
        omega = ass / asas
        x = x + alpha * P + omega * S
        R = S - omega * AS
        rr, rr0n = dots((R, R), (R, R0))
        if np.sqrt(rr) / bnorm < tol:
            break
        beta = (alpha / omega) * rr0n / rr0
        P = R + beta * (P - omega * AP)
        rr0 = rr0n
        if abs(rr0n) / bnorm < tol:
            R0, P, rr0 = R.copy(), R.copy(), rr
    # gather the owned parts on rank 0
    xt = torch.from_numpy(np.where(own, x, 0.0))
    dist.all_reduce(xt)
    if rank == 0:
        x_ref = np.zeros(n)
        it_ref = oracle.bicgstabwr(A.valA, A.irow, A.jcol, b, x_ref, tol, itmax)
        err = float(np.linalg.norm(xt.numpy() - x_ref) / np.linalg.norm(x_ref))
        ret["it"], ret["it_ref"], ret["err"] = it, it_ref, err
        ret["n_owned"] = [s.n_owned for s in slabs]
    dist.destroy_process_group()


def test_two_rank_slab_bicgstab_matches_oracle():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["it"] == ret["it_ref"], dict(ret)
    assert ret["err"] < 1e-9, dict(ret)
    assert sum(ret["n_owned"]) == 3 * 32 ** 3 + 32 ** 3 // 8


def test_slab_layout_covers_everything():
    from eddy_currents_3d_b200 import plate
    from eddy_currents_3d_b200.slab import make_slabs
    p = plate(32, "A")
    for nr in (1, 2, 3, 4, 8):
        slabs = make_slabs(p, nr)
        cover = np.zeros(p.nCellsGlob, np.int32)
        for s in slabs:
            for sl in s.owned:
                cover[sl] += 1
        assert np.all(cover == 1)
        for a, b in zip(slabs[:-1], slabs[1:]):
            assert a.k1 == b.k0
            # what a sends up is what b receives from below, and vice versa
            assert [(s.start, s.stop) for s in a.send_hi] == [(s.start, s.stop) for s in b.halo_lo]
            assert [(s.start, s.stop) for s in b.send_lo] == [(s.start, s.stop) for s in a.halo_hi]
