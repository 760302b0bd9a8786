"""Two-GPU test of the NCCL processor-patch migration (stage 1 across bricks): the union of the two ranks'
clouds after a few steps must equal the single-domain oracle run (same parcels, same global cells, positions)."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from hystrath_b200 import capi, meshgen
from tests import helpers as H
from tests.test_decomposed_gloo import L_LOCAL, N_LOCAL, PROCS, STEPS, _global_reference, _models

pytestmark = pytest.mark.gpu


def _worker(rank, world, q_id, q_out):
    try:
        torch.cuda.set_device(rank)
        fnum, start, _ = _global_reference()
        mesh = meshgen.decomposed_box(N_LOCAL, L_LOCAL, PROCS, rank)
        eng = capi.Engine(rank, rank, world)
        eng.set_mesh(mesh); eng.set_species([H.argon()]); eng.set_models(_models(fnum))
        if rank == 0:
            ident = capi.nccl_unique_id()
            for _ in range(world - 1):
                q_id.put(ident)
        else:
            ident = q_id.get(timeout=120)
        eng.init_comm(ident)
        gi, gj, gk = start.cell % 8, (start.cell // 8) % 4, start.cell // 32
        mine = (gi // 4) == rank
        loc = (gi % 4 + 4 * (gj + 4 * gk)).astype(np.int32)
        p = capi.ParcelData(int(mine.sum()), 1, allocate=False, position=start.position[mine], U=start.U[mine], cell=loc[mine],
                            typeId=start.typeId[mine], origId=start.origId[mine])
        eng.upload_parcels(p)          # tetFace/tetPt located by the library
        migrated = 0
        for _ in range(STEPS):
            eng.evolve(1)
            c = eng.counters()
            migrated += c.migratedOut
        res = eng.download_parcels()
        li, lj, lk = res.cell % 4, (res.cell // 4) % 4, res.cell // 16
        gcell = (li + 4 * rank) + 8 * (lj + 4 * lk)
        q_out.put((rank, res.origId.copy(), res.position.copy(), gcell.astype(np.int32), int(migrated)))
        eng.close()
    except Exception as e:  # surface the failure in the parent
        q_out.put((rank, repr(e)))


def test_two_gpu_migration_matches_single_domain_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, q_id, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q_out.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for r in results:
        assert len(r) == 5, r
    fnum, start, ref = _global_reference()
    ids = np.concatenate([r[1] for r in results])
    pos = np.concatenate([r[2] for r in results])
    cell = np.concatenate([r[3] for r in results])
    order = np.argsort(ids)
    assert np.array_equal(ids[order], ref["origId"])
    assert np.array_equal(cell[order], ref["cell"])
    assert np.allclose(pos[order], ref["position"], rtol=0, atol=1e-15)
    assert all(r[4] > 0 for r in results)
