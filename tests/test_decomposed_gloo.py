"""world_size-2 test of the processor-patch migration protocol on CPU (gloo): two ranks each own one brick of a
periodic box (decomposed_box), run the oracle's move on their parcels and exchange the leavers exactly as
Cloud<T>::move does (BASIC/Cloud/Cloud.C:258-455: per-neighbour transfer lists, loop until no rank sent).
The union of the two clouds must equal the single-domain run: same parcels, same global cells, same positions,
and the per-neighbour migration counts must match those implied by the single-domain cell changes."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hystrath_b200 import capi, meshgen
from oracle.pyoracle import Oracle
from tests import helpers as H

N_LOCAL = (4, 4, 3)
L_LOCAL = (0.016, 0.016, 0.012)
PROCS = (2, 1, 1)
STEPS = 3


def _models(fnum):
    return capi.build_models("NoBinaryCollision", nEquivalentParticles=fnum, deltaT=6e-6, seed=77)


def _global_reference():
    """Single-domain run of the whole 8x4x3 box."""
    sp = [H.argon()]
    mesh = meshgen.box_mesh((8, 4, 3), (0.032, 0.016, 0.012))
    fnum = 1e20 * 0.032 * 0.016 * 0.012 / (96 * 40)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(_models(fnum))
    o.mesh_fill([0], [1e20], 300.0, velocity=(150.0, 0.0, 0.0))
    start = o.download_parcels()
    o.evolve(STEPS)
    return fnum, start, H.by_id(o.download_parcels())


def _worker(rank, world, port, q, renumber=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fnum, start, _ = _global_reference()          # deterministic: both ranks generate the same global cloud
        sp = [H.argon()]
        mesh = meshgen.decomposed_box(N_LOCAL, L_LOCAL, PROCS, rank)
        new_of_old = np.arange(mesh.n_cells, dtype=np.int32)
        if renumber:   # every rank relabels its own brick (renumberMesh per processor directory); patch faces keep their order
            mesh, new_of_old = meshgen.renumber_cells(mesh, meshgen.morton_order(mesh))
        old_of_new = np.argsort(new_of_old)
        o = Oracle()
        o.set_mesh(mesh); o.set_species(sp); o.set_models(_models(fnum))
        # my share of the global cloud: global cell (i,j,k) -> rank i // 4, local cell (i % 4, j, k)
        gi, gj, gk = start.cell % 8, (start.cell // 8) % 4, start.cell // 32
        mine = (gi // 4) == rank
        loc = (gi % 4 + 4 * (gj + 4 * gk)).astype(np.int32)
        p = capi.ParcelData(int(mine.sum()), 1, allocate=False, position=start.position[mine], U=start.U[mine], cell=new_of_old[loc[mine]],
                            typeId=start.typeId[mine], origId=start.origId[mine])
        o.upload_parcels(p)
        sent_total = np.zeros(world, np.int64)
        for _ in range(STEPS):
            o.evolve_begin()
            while True:
                d, i = o.outbox()
                out = {dst: (d[i[:, 0] == dst], i[i[:, 0] == dst]) for dst in range(world) if dst != rank}
                for dst, (dd, _) in out.items():
                    sent_total[dst] += len(dd)
                gathered = [None] * world
                dist.all_gather_object(gathered, out)           # pBufs.finishedSends + the transfers themselves
                if not any(len(v[0]) for g in gathered for v in g.values()):
                    break                                         # reduce(transfered, orOp<bool>()) == false
                for src in range(world):
                    if src != rank and rank in gathered[src] and len(gathered[src][rank][0]):
                        o.receive_and_move(src, *gathered[src][rank])
            o.evolve_end()
        res = o.download_parcels()
        lc = old_of_new[res.cell]
        li, lj, lk = lc % 4, (lc // 4) % 4, lc // 16
        gcell = (li + 4 * rank) + 8 * (lj + 4 * lk)
        q.put((rank, res.origId.copy(), res.position.copy(), gcell.astype(np.int32), sent_total))
    finally:
        dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("renumber", [False, True])
def test_two_rank_migration_matches_single_domain(renumber):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, renumber)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fnum, start, ref = _global_reference()
    ids = np.concatenate([r[1] for r in results])
    pos = np.concatenate([r[2] for r in results])
    cell = np.concatenate([r[3] for r in results])
    order = np.argsort(ids)
    assert np.array_equal(ids[order], ref["origId"])              # nothing lost or duplicated
    assert np.array_equal(cell[order], ref["cell"])               # cell indexing bit-exact across the partition
    assert np.allclose(pos[order], ref["position"], rtol=0, atol=1e-15)
    # every parcel that ends on the other rank was shipped at least once; the counts are symmetric-ish and non-zero
    sent = {r[0]: r[4] for r in results}
    assert sent[0][1] > 0 and sent[1][0] > 0
    owner_end = (ref["cell"] % 8) // 4
    start_sorted = H.by_id(start)
    owner_start = (start_sorted["cell"] % 8) // 4
    net01 = int(((owner_start == 0) & (owner_end == 1)).sum())
    net10 = int(((owner_start == 1) & (owner_end == 0)).sum())
    assert sent[0][1] >= net01 and sent[1][0] >= net10
    assert (sent[0][1] - sent[1][0]) == (net01 - net10)           # conservation of parcels per rank
