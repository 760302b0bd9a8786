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


# ---- the same with walls: a channel (diffuse walls at y = 0 and y = Ly) cut in x between the two GPUs --------------------------------
def _channel_models(fnum, patch):
    pm = [dict(patch=patch, boundaryModel="dsmcDiffuseWallPatch", temperature=700.0, velocity=(120.0, 0.0, 0.0))]
    return capi.build_models("NoBinaryCollision", nEquivalentParticles=fnum, deltaT=6e-6, seed=78, patch_models=pm)


def _channel_reference():
    from oracle.pyoracle import Oracle

    sides = {"xmin": ("cyclic",), "xmax": ("cyclic",), "ymin": ("wall", "walls"), "ymax": ("wall", "walls"), "zmin": ("cyclic",), "zmax": ("cyclic",)}
    mesh = meshgen.box_mesh((8, 4, 3), (0.032, 0.016, 0.012), sides=sides)
    fnum = 1e20 * 0.032 * 0.016 * 0.012 / (96 * 40)
    o = Oracle()
    o.set_mesh(mesh); o.set_species([H.argon()]); o.set_models(_channel_models(fnum, mesh.patch_index("walls")))
    o.mesh_fill([0], [1e20], 300.0, velocity=(150.0, 0.0, 0.0))
    start = o.download_parcels()
    o.evolve(STEPS)
    return fnum, start, H.by_id(o.download_parcels()), o.wall_accumulators()


def _channel_worker(rank, world, q_id, q_out):
    try:
        torch.cuda.set_device(rank)
        fnum, start, _, _ = _channel_reference()
        mesh = meshgen.decomposed_box(N_LOCAL, L_LOCAL, PROCS, rank, outer=("cyclic", ("wall", "walls"), "cyclic"))
        eng = capi.Engine(rank, rank, world)
        eng.set_mesh(mesh); eng.set_species([H.argon()]); eng.set_models(_channel_models(fnum, mesh.patch_index("walls")))
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
        # (origProc, origId) is a parcel's identity and keys its wall-model random stream: the single-domain run created them all on proc 0
        p = capi.ParcelData(int(mine.sum()), 1, allocate=False, position=start.position[mine], U=start.U[mine], cell=loc[mine],
                            typeId=start.typeId[mine], origId=start.origId[mine], origProc=np.zeros(int(mine.sum()), np.int32))
        eng.upload_parcels(p)
        eng.evolve(STEPS)
        res = eng.download_parcels()
        li, lj, lk = res.cell % 4, (res.cell // 4) % 4, res.cell // 16
        gcell = (li + 4 * rank) + 8 * (lj + 4 * lk)
        w = eng.wall_accumulators()
        q_out.put((rank, res.origId.copy(), res.position.copy(), gcell.astype(np.int32), res.U.copy(), w.sum(axis=(0, 1))))
        eng.close()
    except Exception as e:
        q_out.put((rank, repr(e)))


def test_two_gpu_channel_with_diffuse_walls_matches_single_domain_oracle():
    """Walls and migration together: a parcel re-emitted by a wall on one GPU may cross to the other in the same step; wall draws are
    keyed by (origId, hit, step), so the union of the two clouds equals the single-domain run, wall-sampled sums included."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_channel_worker, args=(r, 2, q_id, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q_out.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for r in results:
        assert len(r) == 6, r
    fnum, start, ref, wref = _channel_reference()
    ids = np.concatenate([r[1] for r in results])
    pos = np.concatenate([r[2] for r in results])
    cell = np.concatenate([r[3] for r in results])
    U = np.concatenate([r[4] for r in results])
    order = np.argsort(ids)
    assert np.array_equal(ids[order], ref["origId"])
    assert np.array_equal(cell[order], ref["cell"])
    assert np.allclose(U[order], ref["U"], rtol=0, atol=1e-9)
    assert np.allclose(pos[order], ref["position"], rtol=0, atol=1e-12)
    hit = (ref["U"] != H.by_id(start)["U"]).any(1)
    assert hit.sum() > 200
    wsum = results[0][5] + results[1][5]
    wr = wref.sum(axis=(0, 1))
    assert np.abs(wsum - wr).max() / np.abs(wr).max() < 1e-9 and np.abs(wr).max() > 0


@pytest.mark.parametrize("time_step_model", ["constant", "variable"])
def test_two_gpu_driver_runs_decomposed_case(tmp_path, time_step_model):
    """With timeStepModel variable the reference cell is the smallest of the WHOLE mesh (dsmcVariableTimeStepModel::findRefCell reduces the
    minimum volume over the ranks: dsmcb200_allreduce_min); on this uniform mesh every cell then keeps nEquivalentParticles and deltaT.
    dsmcb200_run -parallel on a decomposePar-style case: processor0/ and processor1/ each hold their brick (polyMesh with processor and
    processorCyclic patches, start-time cloud), the dictionaries sit at the case root; one process per GPU, ncclUniqueId handed over through
    the case directory.  A closed channel keeps its parcels; both ranks write their time directories; the log carries the global counts."""
    import subprocess

    from hystrath_b200 import case as casew
    from oracle.pyoracle import Oracle

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    case = str(tmp_path)
    run = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hystrath_b200", "dsmcb200_run")
    fnum = 1e20 * 0.032 * 0.016 * 0.012 / (96 * 60)
    total = 0
    for rank in range(2):
        mesh = meshgen.decomposed_box(N_LOCAL, L_LOCAL, PROCS, rank, outer=("cyclic", ("wall", "walls"), "cyclic"))
        root = os.path.join(case, f"processor{rank}")
        casew.write_poly_mesh(root, mesh)
        pm = [dict(patch=mesh.patch_index("walls"), boundaryModel="dsmcDiffuseWallPatch", temperature=700.0, velocity=(120.0, 0.0, 0.0))]
        o = Oracle()
        o.set_mesh(mesh); o.set_species([H.argon()])
        o.set_models(capi.build_models("VariableHardSphere", nEquivalentParticles=fnum, deltaT=6e-6, seed=90 + rank, patch_models=pm))
        o.mesh_fill([0], [1e20], 300.0, velocity=(150.0, 0.0, 0.0))
        p = o.download_parcels()
        sig, _ = o.download_cellstate()
        casew.write_cloud(root, "0", p, sig, mesh)
        total += p.n
    casew.write_dict(os.path.join(case, "constant", "dsmcProperties"), "constant", "dsmcProperties", """
nEquivalentParticles            %.10g;
seedNumber                      11;
timeStepModel                   TIME_STEP_MODEL;
BinaryCollisionModel            VariableHardSphere;
collisionPartnerSelectionModel  noTimeCounter;
typeIdList                      (Ar);
moleculeProperties
{
    Ar { mass 66.3e-27; diameter 4.17e-10; omega 0.81; alpha 1.0; }
}
""".replace("TIME_STEP_MODEL", time_step_model) % fnum)
    casew.write_dict(os.path.join(case, "system", "controlDict"), "system", "controlDict", """
application dsmcFoam+; nTerminalOutputs 5; startFrom latestTime; startTime 0; stopAt endTime; endTime 6e-5; deltaT 6e-6;
writeControl timeStep; writeInterval 10; writeFormat ascii; writePrecision 10; timeFormat general; timePrecision 10;
""")
    casew.write_dict(os.path.join(case, "system", "boundariesDict"), "system", "boundariesDict", """
dsmcPatchBoundaries
(
    boundary
    {
        patchBoundaryProperties { patchName walls; }
        boundaryModel   dsmcDiffuseWallPatch;
        dsmcDiffuseWallPatchProperties { temperature 700; velocity (120 0 0); }
    }
);
dsmcCyclicBoundaries ( );
dsmcGeneralBoundaries ( );
""")
    casew.write_dict(os.path.join(case, "system", "fieldPropertiesDict"), "system", "fieldPropertiesDict", """
dsmcFields
(
    field
    {
        fieldModel dsmcVolFields;
        timeProperties { timeOption write; resetAtOutput on; }
        dsmcVolFieldsProperties { fieldName Ar; typeIds (Ar); }
    }
);
""")
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([run, "-case", case, "-parallel"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se + so
    log = outs[0][0]
    assert f"Number of DSMC particles        = {total}" in log and "End stage 0" in log
    assert ("Variable time-step model:" in log) == (time_step_model == "variable")
    n_end = 0
    for rank in range(2):
        tdir = os.path.join(case, f"processor{rank}", "6e-05")
        assert os.path.isdir(tdir), os.listdir(os.path.join(case, f"processor{rank}"))
        _, cell = ff_read_positions(os.path.join(tdir, "lagrangian", "dsmc", "positions"))
        n_end += len(cell)
        assert os.path.exists(os.path.join(tdir, "rhoN_Ar")) and os.path.exists(os.path.join(tdir, "dsmcSigmaTcRMax"))
    assert n_end == total                                      # diffuse walls re-emit, processor / processorCyclic patches hand over
    if time_step_model == "variable":                          # the model's fields: uniform on a uniform mesh, the same on both ranks
        from hystrath_b200 import foamfile as ff
        for rank in range(2):
            tdir = os.path.join(case, f"processor{rank}", "6e-05")
            assert np.allclose(ff.read_internal_field(os.path.join(tdir, "nParticles")), fnum, rtol=1e-9)
            assert np.allclose(ff.read_internal_field(os.path.join(tdir, "deltaT")), 6e-6, rtol=1e-9)


def ff_read_positions(path):
    from hystrath_b200 import foamfile as ff

    return ff.read_positions(path)


# ---- collisions on: the decomposed run is reproducible parcel by parcel -----------------------------------------------------------------
# Cloud<T>::move ships leavers in cloud-list order and appends arrivals neighbour by neighbour (BASIC/Cloud/Cloud.C:283-306,364-397), so
# the in-cell order that buildCellOccupancy produces -- and with it every NTC pair -- is a function of the decomposition alone.  The
# engine restores that order (orderMigrants), hence the 2-GPU run equals the 2-rank oracle run: same list order, same collisions.
LB_STEPS = 4


def _lb_models(fnum):
    return capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=fnum, deltaT=6e-6, seed=79)   # variable Zv table


def _lb_start():
    from oracle.pyoracle import Oracle

    sp = H.air5()[:2]
    mesh = meshgen.box_mesh((8, 4, 3), (0.032, 0.016, 0.012))
    fnum = 2e21 * 0.032 * 0.016 * 0.012 / (96 * 40)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(_lb_models(fnum))
    o.mesh_fill([0, 1], [1.5e21, 0.5e21], 4000.0, 4000.0, 4000.0, velocity=(300.0, 0.0, 0.0))
    return fnum, o.download_parcels(), o.download_cellstate()[0]


def _lb_share(start, sigma, rank):
    gi, gj, gk = start.cell % 8, (start.cell // 8) % 4, start.cell // 32
    mine = (gi // 4) == rank
    loc = (gi % 4 + 4 * (gj + 4 * gk)).astype(np.int32)
    p = capi.ParcelData(int(mine.sum()), 1, allocate=False, position=start.position[mine], U=start.U[mine], ERot=start.ERot[mine],
                        cell=loc[mine], typeId=start.typeId[mine], vibLevel=start.vibLevel[mine], origId=start.origId[mine])
    c = np.arange(48)
    gcell = (c % 4 + 4 * rank) + 8 * ((c // 4) % 4 + 4 * (c // 16))
    return p, sigma[gcell].copy()


def _lb_oracle_two_ranks():
    """The protocol of tests/test_decomposed_gloo.py with both ranks in this process."""
    from oracle.pyoracle import Oracle

    fnum, start, sigma = _lb_start()
    ranks = []
    for r in range(2):
        o = Oracle()
        o.set_mesh(meshgen.decomposed_box(N_LOCAL, L_LOCAL, PROCS, r)); o.set_species(H.air5()[:2]); o.set_models(_lb_models(fnum))
        p, sig = _lb_share(start, sigma, r)
        o.upload_parcels(p)
        o.upload_cellstate(sig, None)
        ranks.append(o)
    sent = np.zeros((LB_STEPS, 2), np.int64)     # [step][rank]: parcels handed to the other rank, summed over the rounds of Cloud<T>::move
    rounds = np.zeros(LB_STEPS, np.int64)
    for step in range(LB_STEPS):
        for o in ranks:
            o.evolve_begin()
        while True:
            boxes = [o.outbox() for o in ranks]
            if not any(len(d) for d, _ in boxes):
                break
            rounds[step] += 1
            for r, o in enumerate(ranks):
                sent[step, r] += int((boxes[r][1][:, 0] == 1 - r).sum())
            for r, o in enumerate(ranks):
                d, i = boxes[1 - r]
                sel = i[:, 0] == r
                if sel.any():
                    o.receive_and_move(1 - r, d[sel], i[sel])
        for o in ranks:
            o.evolve_end()
    return [(o.download_parcels(), o.counters()) for o in ranks], sent, rounds


def _lb_worker(rank, world, q_id, q_out):
    try:
        torch.cuda.set_device(rank)
        fnum, start, sigma = _lb_start()
        eng = capi.Engine(rank, rank, world)
        eng.set_mesh(meshgen.decomposed_box(N_LOCAL, L_LOCAL, PROCS, rank)); eng.set_species(H.air5()[:2]); eng.set_models(_lb_models(fnum))
        if rank == 0:
            ident = capi.nccl_unique_id()
            for _ in range(world - 1):
                q_id.put(ident)
        else:
            ident = q_id.get(timeout=120)
        eng.init_comm(ident)
        p, sig = _lb_share(start, sigma, rank)
        eng.upload_parcels(p)
        eng.upload_cellstate(sig, None)
        collisions = 0
        sent, recv, rounds = [], [], []
        for _ in range(LB_STEPS):
            eng.evolve(1)
            c = eng.counters()
            collisions += c.collisions
            assert c.nNeighbours == 1 and c.neighbourProc[0] == 1 - rank
            sent.append(int(c.migratedTo[0])); recv.append(int(c.migratedFrom[0])); rounds.append(int(c.migrationRounds))
            assert c.migratedOut == c.migratedTo[0] and c.migratedIn == c.migratedFrom[0]
        res = eng.download_parcels()
        q_out.put((rank, res.origId.copy(), res.cell.copy(), res.U.copy(), res.ERot.copy(), res.vibLevel.copy(), int(collisions), sent, recv, rounds))
        eng.close()
    except Exception as e:
        q_out.put((rank, repr(e)))


def test_two_gpu_run_with_collisions_equals_the_two_rank_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_lb_worker, args=(r, 2, q_id, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    ref, ref_sent, ref_rounds = _lb_oracle_two_ranks()
    results = sorted([q_out.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    for r in results:
        assert len(r) == 10, r
    for rank, (_, ids, cell, U, erot, vib, ncoll, sent, recv, rounds) in enumerate(results):
        o, oc = ref[rank]
        # migration counts per neighbour and per step, and the number of transfer rounds, are those of the reference's algorithm
        assert sent == ref_sent[:, rank].tolist() and recv == ref_sent[:, 1 - rank].tolist() and sum(sent) > 0
        assert rounds == ref_rounds.tolist()
        assert np.array_equal(ids, o.origId)              # the cloud in the same list order: arrivals included
        assert np.array_equal(cell, o.cell)
        assert ncoll == oc["collisions"] and ncoll > 100  # the same NTC pairs were selected and accepted
        assert np.array_equal(vib, o.vibLevel)            # variable-Zv vibrational exchange (host-tabulated 1/Zv) included
        assert np.allclose(U, o.U, rtol=0, atol=1e-8) and np.allclose(erot, o.ERot, rtol=1e-9, atol=1e-30)


# ---- BASELINE configs[3] on two GPUs: the capsule forebody cut into two bricks along x (the upstream half and the half with the shield) ----
CAP_N, CAP_STEPS = (7, 10, 10), 5


def _capsule_rank(r):
    from hystrath_b200 import cases
    return cases.capsule_forebody(CAP_N, ppc=14, density_scale=40.0, procs=(2, 1, 1), rank=r, seed=0xD5C00004 + 7183 * r)


def _capsule_oracle(r):
    from oracle.pyoracle import Oracle
    mesh, sp, md, fill = _capsule_rank(r)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md); o.set_rank(r)
    o.mesh_fill(fill["type_ids"], fill["number_densities"], fill["Ttra"], fill["Trot"], fill["Tvib"], 0.0, fill["velocity"])
    return o


def _capsule_oracle_two_ranks():
    ranks = [_capsule_oracle(r) for r in range(2)]
    sent = np.zeros((CAP_STEPS, 2), np.int64)
    for step in range(CAP_STEPS):
        for o in ranks:
            o.evolve_begin()
        while True:
            boxes = [o.outbox() for o in ranks]
            if not any(len(d) for d, _ in boxes):
                break
            for r in range(2):
                sent[step, r] += int((boxes[r][1][:, 0] == 1 - r).sum())
            for r, o in enumerate(ranks):
                d, i = boxes[1 - r]
                sel = i[:, 0] == r
                if sel.any():
                    o.receive_and_move(1 - r, d[sel], i[sel])
        for o in ranks:
            o.evolve_end()
    return [(o.download_parcels(), o.counters(), o.wall_accumulators()) for o in ranks], sent


def _capsule_worker(rank, world, q_id, q_out):
    try:
        torch.cuda.set_device(rank)
        mesh, sp, md, fill = _capsule_rank(rank)
        eng = capi.Engine(rank, rank, world)
        eng.set_mesh(mesh); eng.set_species(sp); eng.set_models(md)
        if rank == 0:
            ident = capi.nccl_unique_id()
            for _ in range(world - 1):
                q_id.put(ident)
        else:
            ident = q_id.get(timeout=120)
        eng.init_comm(ident)
        o = _capsule_oracle(rank)                      # the rank's own dsmcMeshFill, as decomposed dsmcInitialise+ does it
        eng.upload_parcels(o.download_parcels())
        eng.upload_cellstate(*o.download_cellstate())
        sent, collisions, inserted, deleted = [], 0, 0, 0
        for _ in range(CAP_STEPS):
            eng.evolve(1)
            c = eng.counters()
            sent.append(int(c.migratedTo[0])); collisions += c.collisions; inserted += c.inserted; deleted += c.deleted
        res = eng.download_parcels()
        q_out.put((rank, res.origId.copy(), res.origProc.copy(), res.cell.copy(), res.typeId.copy(), res.vibLevel.copy(), sent, int(collisions),
                   int(inserted), int(deleted), eng.wall_accumulators()))
        eng.close()
    except Exception as e:
        q_out.put((rank, repr(e)))


def test_two_gpu_capsule_forebody_equals_the_two_rank_oracle():
    """The capsule forebody (BASELINE configs[3]) at test size on two GPUs: inflow, wall hits on the warped shield faces, deletion,
    Larsen-Borgnakke collisions and NCCL migration across the processor patch; per rank the cloud (identity and order of every parcel,
    cells, species, vibrational levels), the parcels sent per step, insertions, deletions and collision counts equal the two-rank oracle."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_capsule_worker, args=(r, 2, q_id, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    ref, ref_sent = _capsule_oracle_two_ranks()
    results = sorted([q_out.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    for r in results:
        assert len(r) == 11, r
    for rank, (_, ids, procs_, cell, typ, vib, sent, ncoll, nins, ndel, wall) in enumerate(results):
        o, oc, ow = ref[rank]
        assert sent == ref_sent[:, rank].tolist()
        assert np.array_equal(ids, o.origId) and np.array_equal(procs_, o.origProc)
        assert np.array_equal(cell, o.cell) and np.array_equal(typ, o.typeId) and np.array_equal(vib, o.vibLevel)
        assert ncoll == oc["collisions"] and nins == oc["inserted"] and ndel == oc["deleted"]
        if rank == 1:
            assert np.abs(ow).max() > 0 and np.abs(wall - ow).max() <= 1e-9 * np.abs(ow).max()      # the shield is on rank 1
    assert ref_sent[:, 0].sum() > 100 and ref[0][1]["inserted"] > 0 and ref[1][1]["deleted"] > 0


# ---- dsmcAxisymmetric on two GPUs: a weighted box cut in two along the polar axis; parcels take their radial weight across the processor patch ----
AXI_STEPS = 5


def _axi_rank(r):
    outer = ("cyclic", (("symmetryPlane", "axis"), ("wall", "top")), ("symmetryPlane", "sides"))
    mesh = meshgen.decomposed_box((4, 3, 1), (0.04, 0.03, 0.01), (1, 2, 1), r, outer=outer)
    rev, pol, ang = capi.axisymmetric_axes()
    md = capi.build_models("VariableHardSphere", nEquivalentParticles=1e9, deltaT=4e-6, seed=5 + 7183 * r, coordinateSystem="dsmcAxisymmetric",
                           angularCoordinate=ang, patch_models=[dict(patch=mesh.patch_index("top"), boundaryModel="dsmcDiffuseWallPatch",
                                                                      temperature=400.0, velocity=(0, 0, 0))] if r == 1 else [])
    return mesh, md, pol


def _axi_oracle(r):
    from oracle.pyoracle import Oracle
    mesh, md, pol = _axi_rank(r)
    o = Oracle()
    o.set_mesh(mesh); o.set_species([H.argon()]); o.set_models(md); o.set_rank(r)
    cc, cv, fc, *_ = o.geometry()
    rwf, _ = capi.axisymmetric_rwf(cc, fc, pol, 100.0, radial_extent=0.06)
    o.set_cell_fields(RWF=rwf)
    o.mesh_fill([0], [4e18], 300.0)
    return o, rwf


def _axi_oracle_two_ranks():
    ranks = [_axi_oracle(r)[0] for r in range(2)]
    sent = np.zeros((AXI_STEPS, 2), np.int64)
    for step in range(AXI_STEPS):
        for o in ranks:
            o.evolve_begin()
        while True:
            boxes = [o.outbox() for o in ranks]
            if not any(len(d) for d, _ in boxes):
                break
            for r in range(2):
                sent[step, r] += int((boxes[r][1][:, 0] == 1 - r).sum())
            for r, o in enumerate(ranks):
                d, i = boxes[1 - r]
                sel = i[:, 0] == r
                if sel.any():
                    o.receive_and_move(1 - r, d[sel], i[sel])
        for o in ranks:
            o.evolve_end()
    return [(o.download_parcels(), o.weighting_counts()) for o in ranks], sent


def _axi_worker(rank, world, q_id, q_out):
    try:
        torch.cuda.set_device(rank)
        mesh, md, pol = _axi_rank(rank)
        eng = capi.Engine(rank, rank, world)
        eng.set_mesh(mesh); eng.set_species([H.argon()]); eng.set_models(md)
        if rank == 0:
            ident = capi.nccl_unique_id()
            for _ in range(world - 1):
                q_id.put(ident)
        else:
            ident = q_id.get(timeout=120)
        eng.init_comm(ident)
        o, rwf = _axi_oracle(rank)
        eng.set_cell_fields(RWF=rwf)
        eng.upload_parcels(o.download_parcels())
        eng.upload_cellstate(*o.download_cellstate())
        sent = []
        for _ in range(AXI_STEPS):
            eng.evolve(1)
            sent.append(int(eng.counters().migratedTo[0]))
        res = eng.download_parcels()
        q_out.put((rank, res.origId.copy(), res.origProc.copy(), res.cell.copy(), res.radialWeight.copy(), res.U.copy(), sent))
        eng.close()
    except Exception as e:
        q_out.put((rank, repr(e)))


def test_two_gpu_axisymmetric_run_equals_the_two_rank_oracle():
    """Radial weighting on a decomposed mesh: a parcel that crosses the processor patch arrives with the weight of the cell it started the
    step in (it travels next to the 96-byte record) and is cloned / deleted against the weight of the cell it ends in, on the receiving
    rank.  Per rank the cloud (identity, order, cells, weights, velocities of the clones included) equals the two-rank oracle."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_axi_worker, args=(r, 2, q_id, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    ref, ref_sent = _axi_oracle_two_ranks()
    results = sorted([q_out.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    for r in results:
        assert len(r) == 7, r
    for rank, (_, ids, procs_, cell, rwf, U, sent) in enumerate(results):
        o, (cloned, deleted) = ref[rank]
        assert sent == ref_sent[:, rank].tolist() and sum(sent) > 20
        assert cloned > 0 and deleted > 0
        assert np.array_equal(ids, o.origId) and np.array_equal(procs_, o.origProc) and np.array_equal(cell, o.cell)
        assert np.array_equal(rwf, o.radialWeight)
        assert np.allclose(U, o.U, rtol=1e-9, atol=1e-7)
