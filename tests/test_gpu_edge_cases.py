"""GPU: degenerate inputs through the C ABI -- empty cloud, a single parcel, a cloud that is deleted completely, invalid uploads.
The reference loops over an empty IDLList / empty cellOccupancy lists without special cases (Cloud.C:204-312, noTimeCounter.C:96-155:
nC <= 1 cells select nothing); the engine must do the same and keep its counters and accumulators consistent."""
import numpy as np
import pytest

from hystrath_b200 import capi, meshgen
from oracle.pyoracle import Oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _box(model="VariableHardSphere", sides=None, patch_models=None):
    mesh = meshgen.box_mesh((4, 4, 4), (0.016,) * 3, sides=sides)
    md = capi.build_models(model, nEquivalentParticles=1e20 * 0.016 ** 3 / (64 * 30), deltaT=5e-6, seed=9, patch_models=patch_models or [])
    return mesh, [H.argon()], md


def test_empty_cloud_evolves():
    mesh, sp, md = _box()
    eng = capi.Engine(0)
    eng.set_mesh(mesh); eng.set_species(sp); eng.set_models(md)
    eng.upload_parcels(capi.ParcelData(0, 1))
    eng.evolve(3)
    assert eng.num_parcels() == 0
    assert np.array_equal(eng.occupancy(), np.zeros(65, dtype=np.int32))
    acc, coll, nt = eng.accumulators()
    assert nt == 3 and not acc.any() and not coll.any()
    c = eng.counters()
    assert c.collisions == 0 and c.collisionCandidates == 0
    eng.close()


def test_single_parcel_flies_and_never_collides():
    mesh, sp, md = _box()
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    p = capi.ParcelData(1, 1)
    p.position[:] = [[0.0021, 0.0093, 0.0157]]
    p.U[:] = [[412.0, -233.0, 157.0]]
    p.cell[:] = 0 + 4 * (2 + 4 * 3)
    p.tetFace = p.tetPt = None  # tet not known: located on upload like Cloud<T>::initCloud does after reading `positions`
    p.origId[:] = 7
    for x in (eng, ora):
        x.upload_parcels(p)
        x.evolve(40)
    g, o = eng.download_parcels(), ora.download_parcels()
    assert g.n == o.n == 1
    assert np.array_equal(g.position, o.position) and np.array_equal(g.cell, o.cell) and np.array_equal(g.U, p.U)
    free = np.mod(p.position + 40 * 5e-6 * p.U, 0.016)
    assert np.allclose(g.position, free, atol=1e-12)
    assert eng.counters().collisions == 0
    occ = eng.occupancy()
    assert occ[-1] == 1 and (np.diff(occ) == 1).sum() == 1
    eng.close()


def test_cloud_deleted_completely():
    sides = {s: ("patch", "outlet") for s in meshgen.SIDES}
    mesh, sp, md = _box(sides=sides, patch_models=[dict(patch=0, boundaryModel="dsmcDeletionPatch")])
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, [0], [1e20], 300.0, velocity=(3000.0, 0.0, 0.0))
    n0 = eng.num_parcels()
    assert n0 > 1000
    for x in (eng, ora):
        x.evolve(4)             # 3000 m/s * 20 us = 0.06 m >> 0.016 m: everything has left through the deletion patch
    assert eng.num_parcels() == ora.num_parcels() == 0
    eng.evolve(2)               # and an empty cloud keeps stepping
    assert eng.num_parcels() == 0 and eng.occupancy()[-1] == 0
    eng.close()


def test_invalid_uploads_fail_loudly():
    mesh, sp, md = _box()
    eng = capi.Engine(0)
    eng.set_mesh(mesh); eng.set_species(sp); eng.set_models(md)
    p = capi.ParcelData(4, 1)
    p.position[:] = 0.001
    p.cell[:] = [0, 1, 64, 2]     # cell 64 does not exist
    p.tetFace[:] = 0
    p.tetPt[:] = 1
    with pytest.raises(capi.Dsmcb200Error):
        eng.upload_parcels(p)
    p.cell[:] = [0, 0, 0, 0]
    p.typeId[:] = [0, 0, 3, 0]    # typeId 3 is not in typeIdList
    with pytest.raises(capi.Dsmcb200Error):
        eng.upload_parcels(p)
    eng.close()


def test_cloud_without_tet_indices_is_located_on_the_device():
    """Cloud<T>::initCloud -> particle::initCellFacePtOrDeleteLostParticle (BASIC/particle/particleI.H:851-996): the `positions` file holds
    only "(x y z) cell"; tetFace / tetPt are re-derived per parcel (first tet of the cell with tetrahedron::inside); a parcel that is not in
    the cell its label names is looked for in the cells around it and then in the whole mesh (findCellFacePt) and takes that cell; a parcel just outside the mesh
    (rounding, or by 5 % of a cell) is moved towards the cell centre in steps of 1e-5 of the distance until a tet of the cell claims it, and
    keeps that position (particleI.H:927-976); parcels outside the 10 %-inflated cell bounding box that no cell claims are deleted."""
    mesh, sp, md = _box()
    h = 0.004
    rng = np.random.default_rng(5)
    n_in = 2000
    cell = rng.integers(0, 64, n_in).astype(np.int32)
    ijk = np.stack([cell % 4, (cell // 4) % 4, cell // 16], 1)
    pos = (ijk + rng.random((n_in, 3))) * h
    # special points of cell 21 = (1, 1, 1): centre, a face centre, a vertex (all "inside" by the > SMALL rule)
    special = np.array([[1.5, 1.5, 1.5], [1.0, 1.5, 1.5], [1.0, 1.0, 1.0]]) * h
    # labelled 21 but lying in other cells: across the x-min face by rounding and by 5 % (cell 20), two cells further (cell 23: the rings
    # around the label), at the other end of the mesh (cell 63: the mesh-wide search)
    elsewhere = np.array([[1.0 - 1e-9, 1.5, 1.5], [0.95, 1.5, 1.5], [3.5, 1.5, 1.5], [3.25, 3.5, 3.75]]) * h
    # labelled 20 = (0, 1, 1) on the wall: outside the mesh by rounding and by 5 % (in the inflated box: both found by the walk), by half a cell (lost)
    outside = np.array([[-1e-9, 1.5, 1.5], [-0.05, 1.5, 1.5], [-0.5, 1.5, 1.5]]) * h
    position = np.concatenate([pos, special, elsewhere, outside])
    cells = np.concatenate([cell, np.full(7, 21, np.int32), np.full(3, 20, np.int32)])
    expect_cell = np.concatenate([cell, [21, 21, 21, 20, 20, 23, 63, 20, 20]]).astype(np.int32)
    n = len(position)
    p = capi.ParcelData(n, 1, allocate=False, position=position, U=np.zeros((n, 3)), cell=cells, typeId=np.zeros(n, np.int32),
                        origId=np.arange(n, dtype=np.int32))
    eng = capi.Engine(0)
    eng.set_mesh(mesh); eng.set_species(sp); eng.set_models(md)
    eng.upload_parcels(p)
    assert eng.num_parcels() == n - 1                       # the lost parcel is deleted
    g = H.by_id(eng.download_parcels())
    assert np.array_equal(g["origId"], np.arange(n - 1).astype(np.int32))
    assert np.array_equal(g["position"][: n - 3], position[: n - 3])   # the search for another cell does not move a parcel ...
    assert np.array_equal(g["cell"], expect_cell)
    # ... the walk does: the first multiple of 1e-5 (cC - position) that lies in the cell, one step for the rounding case, the
    # ~9100th for the parcel 5 % of a cell outside (0.05 h of the 0.55 h to the centre)
    cC = np.array([0.5, 1.5, 1.5]) * h
    for j, (kmin, kmax) in ((n - 3, (1, 1)), (n - 2, (9000, 9200))):
        k = (g["position"][j] - position[j])[0] / (1e-5 * (cC - position[j])[0])
        assert kmin - 1e-6 <= k <= kmax + 1e-6 and abs(k - round(k)) < 1e-6, k
        assert np.allclose(g["position"][j], position[j] + round(k) * 1e-5 * (cC - position[j]), rtol=0, atol=1e-16)
        assert 0 <= g["position"][j][0] < 1e-3 * h
    # the oracle's own findTetFacePt on the parcels that are inside the cell they end up with
    ora = Oracle()
    ora.set_mesh(mesh); ora.set_species(sp); ora.set_models(md)
    m = n_in + 7
    q = capi.ParcelData(m, 1, allocate=False, position=position[:m], U=np.zeros((m, 3)), cell=expect_cell[:m],
                        typeId=np.zeros(m, np.int32), origId=np.arange(m, dtype=np.int32))
    ora.upload_parcels(q)
    o = H.by_id(ora.download_parcels())
    assert np.array_equal(g["tetFace"][:m], o["tetFace"]) and np.array_equal(g["tetPt"][:m], o["tetPt"])
    # the parcel found by the walk sits in a tet of the x-min wall face of cell 20 and flies on normally
    own = np.asarray(mesh.owner)
    f = g["tetFace"][m]
    assert own[f] == 20 and f >= mesh.n_internal
    eng.evolve(2)
    assert eng.num_parcels() == n - 1
    eng.close()


def test_cells_with_thousands_of_parcels_keep_list_order_and_collide_like_the_oracle():
    """Cells far beyond the in-warp ordering pass (> 1024 parcels: the heatBath tutorials put 1e5 in one cell): the block-level sort
    restores cloud-list order, so occupancy, NTC pairs and collision counts are those of the oracle, run after run."""
    mesh = meshgen.box_mesh((2, 1, 1), (0.02, 0.01, 0.01))
    sp = [H.argon()]
    fnum = 1e20 * 0.02 * 0.01 * 0.01 / 5200.0
    md = capi.build_models("VariableHardSphere", nEquivalentParticles=fnum, deltaT=4e-6, seed=321)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    ora.set_reorder(True)
    start = H.same_start(eng, ora, [0], [1e20], 300.0)
    assert start.n > 4500 and np.bincount(start.cell).min() > 2000
    seen = 0
    for _ in range(3):
        eng.evolve(1)
        ora.evolve(1)
        tot = ora.counters()["collisions"]
        c = eng.counters()
        assert c.unsortedLargeCells == 0
        assert c.collisions == tot - seen > 100
        seen = tot
    g, o = eng.download_parcels(), ora.download_parcels()
    assert np.array_equal(g.origId, o.origId)            # cell-major, list order inside the cell
    assert np.array_equal(g.cell, o.cell) and np.array_equal(eng.occupancy(), ora.occupancy())
    assert np.allclose(g.U, o.U, rtol=0, atol=1e-9)
    eng.close()
