"""N > 1 host logic on CPU (gloo, world_size 2 and 3): the z-slab bookkeeping and the neighbour halo
exchange of fluidnet_b200.slab reproduce the global field in every rank's padded slab."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, gnz, halo, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from fluidnet_b200.slab import SlabDecomposition
        rs = np.random.RandomState(3)
        U = torch.from_numpy(rs.rand(1, 3, gnz, 5, 7).astype(np.float32))
        rho = torch.from_numpy(rs.rand(1, 1, gnz, 5, 7).astype(np.float32))
        d = SlabDecomposition(gnz, rank, world, halo)
        assert d.z1 - d.z0 >= halo and d.zoff >= 0 and d.zoff + d.nz <= gnz
        lU, lr = d.scatter(U), d.scatter(rho)
        # owned planes partition the domain
        cover = torch.zeros(gnz)
        cover[d.z0:d.z1] = 1
        dist.all_reduce(cover)
        assert torch.all(cover == 1)
        for width in (halo, 1, 3 if halo >= 3 else 1):
            a, b = lU.clone(), lr.clone()
            a[:, :, :d.own_lo] = float("nan"); a[:, :, d.own_hi:] = float("nan")
            b[:, :, :d.own_lo] = float("nan"); b[:, :, d.own_hi:] = float("nan")
            d.exchange([a, b], width)
            lo = d.own_lo - (width if rank > 0 else 0)
            hi = d.own_hi + (width if rank < world - 1 else 0)
            assert torch.equal(a[:, :, lo:hi], lU[:, :, lo:hi])
            assert torch.equal(b[:, :, lo:hi], lr[:, :, lo:hi])
            if lo > 0:
                assert torch.isnan(a[:, :, :lo]).all()      # planes beyond the exchanged width untouched
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:       # pragma: no cover
        q.put((rank, "FAIL: %r" % (e,)))
        raise


@pytest.mark.parametrize("world,gnz,halo", [(2, 16, 6), (3, 20, 4), (2, 13, 6)])
def test_halo_exchange_gloo(world, gnz, halo):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, gnz, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(r[1] == "ok" for r in res), res


def test_decomposition_rejects_thin_slabs():
    from fluidnet_b200.slab import SlabDecomposition
    with pytest.raises(AssertionError):
        SlabDecomposition(16, 0, 8, halo=6)
