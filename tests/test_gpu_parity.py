"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Integer/index results must be bit-exact; floating point within the tolerance stated."""
import numpy as np
import pytest

from hystrath_b200 import capi, meshgen
from oracle.pyoracle import Oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def periodic_case(n=(8, 8, 8), L=0.032, ppc=24, species=None, model="VariableHardSphere", dens=1e20, dt=5e-6, **kw):
    mesh = meshgen.box_mesh(n, (L, L * n[1] / n[0], L * n[2] / n[0]))
    species = species or [H.argon()]
    vol = L * (L * n[1] / n[0]) * (L * n[2] / n[0])
    fnum = dens * vol / (np.prod(n) * ppc)
    models = capi.build_models(model, nEquivalentParticles=fnum, deltaT=dt, seed=0xD5C00001, **kw)
    return mesh, species, models


def test_move_periodic_box_bit_exact():
    mesh, sp, md = periodic_case()
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, [0], [1e20], 300.0)
    for x in (eng, ora):
        x.stage(capi.STAGE_MOVE)
        x.stage(capi.STAGE_SORT)
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert np.array_equal(g["origId"], o["origId"])
    assert np.array_equal(g["cell"], o["cell"])
    assert np.array_equal(g["tetFace"], o["tetFace"])
    assert np.array_equal(g["tetPt"], o["tetPt"])
    assert np.array_equal(g["position"], o["position"])  # same FP64 operations, no FMA contraction: bit-exact
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    assert (g["cell"] != H.by_id(ora.download_parcels())["cell"]).sum() == 0
    eng.close()


def test_move_on_renumbered_mesh_bit_exact():
    """A z-order renumbered mesh (meshgen.renumber_cells: flipped internal faces, re-sorted face list) through move + sort + collide."""
    base, sp, md = periodic_case((8, 6, 5))
    mesh, _ = meshgen.renumber_cells(base, meshgen.morton_order(base))
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, [0], [1e20], 300.0)
    for x in (eng, ora):
        for _ in range(3):          # three free flights (no collisions in between, so positions stay bit-comparable)
            x.stage(capi.STAGE_MOVE)
            x.stage(capi.STAGE_SORT)
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert np.array_equal(g["cell"], o["cell"])
    assert np.array_equal(g["tetFace"], o["tetFace"]) and np.array_equal(g["tetPt"], o["tetPt"])
    assert np.array_equal(g["position"], o["position"])
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    for x in (eng, ora):
        x.evolve(2)                 # and the full step, collisions included
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert np.array_equal(g["cell"], o["cell"])
    assert np.allclose(g["U"], o["U"], rtol=0, atol=1e-9)
    eng.close()


@pytest.mark.parametrize("kind", ["prism", "tet"])
def test_move_on_non_hex_cells_bit_exact(kind):
    """Prism / tetrahedral cells with triangular faces inside a diffuse-wall box: one free flight bit-exact, then two full steps."""
    mesh, locate = meshgen.split_box_mesh((4, 3, 3), (0.016, 0.012, 0.012), kind)
    sp = [H.argon()]
    md = capi.build_models("VariableHardSphere", nEquivalentParticles=1e20 * 0.016 * 0.012 * 0.012 / (mesh.n_cells * 30), deltaT=5e-6,
                           seed=11, patch_models=[dict(patch=0, boundaryModel="dsmcDiffuseWallPatch", temperature=400.0, velocity=(0, 0, 0))])
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    start = H.by_id(H.same_start(eng, ora, [0], [1e20], 300.0))   # the oracle's fill, uploaded to the engine (tets located on upload)
    for x in (eng, ora):
        x.stage(capi.STAGE_MOVE)
        x.stage(capi.STAGE_SORT)
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert np.array_equal(g["cell"], o["cell"])
    assert np.array_equal(g["tetFace"], o["tetFace"]) and np.array_equal(g["tetPt"], o["tetPt"])
    # a diffuse reflection draws through log / sin / cos (libm vs CUDA ulps): bit-exact for the parcels that met no wall
    hit = (o["U"] != start["U"]).any(1)
    assert 0 < hit.sum() < len(hit) // 2
    assert np.array_equal(g["position"][~hit], o["position"][~hit])
    assert np.array_equal(g["U"][~hit], o["U"][~hit])
    assert np.allclose(g["position"], o["position"], rtol=0, atol=1e-12) and np.allclose(g["U"], o["U"], rtol=0, atol=1e-9)
    assert np.array_equal(locate(g["position"]), g["cell"])
    for x in (eng, ora):
        x.evolve(2)
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert np.array_equal(g["cell"], o["cell"])
    assert np.allclose(g["U"], o["U"], rtol=0, atol=1e-9)
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    eng.close()


def test_move_across_refinement_interface_bit_exact():
    """Pentagonal faces with a hanging node (2:1 refinement interface): base points, tet links and tracking against the oracle."""
    mesh, locate = meshgen.refined_interface_mesh()
    sp = [H.argon()]
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e20 * 3 * 0.004 ** 3 / (6 * 400), deltaT=2e-6, seed=3,
                           patch_models=[dict(patch=0, boundaryModel="dsmcSpecularWallPatch")])
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, [0], [1e20], 300.0)
    for x in (eng, ora):
        x.evolve(6)
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert len(g["cell"]) == len(o["cell"])
    assert np.array_equal(g["cell"], o["cell"])
    assert np.array_equal(g["tetFace"], o["tetFace"]) and np.array_equal(g["tetPt"], o["tetPt"])
    assert np.array_equal(g["position"], o["position"]) and np.array_equal(g["U"], o["U"])   # specular walls: no libm in the loop
    assert np.array_equal(locate(g["position"]), g["cell"])
    eng.close()


def test_sort_is_stable_and_matches_oracle_order():
    mesh, sp, md = periodic_case((6, 5, 4))
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, [0], [1e20], 300.0)
    for x in (eng, ora):
        x.stage(capi.STAGE_MOVE)
        x.stage(capi.STAGE_SORT)
    g, o = eng.download_parcels(), ora.download_parcels()
    assert np.array_equal(g.origId, o.origId)  # identical physical order: cell-major, list order inside a cell
    assert np.all(np.diff(g.cell) >= 0)
    eng.close()


def test_collide_vhs_matches_oracle_and_conserves():
    mesh, sp, md = periodic_case(ppc=40)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    start = H.same_start(eng, ora, [0], [1e20], 300.0)
    for x in (eng, ora):
        x.stage(capi.STAGE_COLLIDE)
    g, o = eng.download_parcels(), ora.download_parcels()
    assert np.array_equal(g.origId, o.origId)
    c = eng.counters()
    oc = ora.counters()
    assert c.collisionCandidates == oc["collisionCandidates"]
    assert c.collisions == oc["collisions"] and c.collisions > 0
    # libm differences (pow, sin, cos) only: 1e-12 relative to the thermal speed
    assert np.allclose(g.U, o.U, rtol=0, atol=1e-9)
    gs, gr = eng.download_cellstate()
    os_, or_ = ora.download_cellstate()
    assert np.allclose(gs, os_, rtol=1e-13)
    assert np.array_equal(gr, or_)
    # momentum and kinetic energy of every cell are conserved to 1e-12 relative
    off = eng.occupancy()
    m = sp[0].mass
    for arr0, arr1 in ((start.U, g.U),):
        p0 = np.add.reduceat(arr0, off[:-1], axis=0) * m
        p1 = np.add.reduceat(arr1, off[:-1], axis=0) * m
        e0 = np.add.reduceat((arr0 ** 2).sum(1), off[:-1]) * 0.5 * m
        e1 = np.add.reduceat((arr1 ** 2).sum(1), off[:-1]) * 0.5 * m
        scale = np.sqrt((arr0 ** 2).sum(1)).mean() * m * 40
        assert np.abs(p1 - p0).max() / scale < 1e-12
        assert np.abs(e1 / e0 - 1).max() < 1e-12
    eng.close()


def test_collide_larsen_borgnakke_air5_matches_oracle():
    sp = H.air5()
    mesh, _, md = periodic_case((6, 6, 6), ppc=60, species=sp, model="LarsenBorgnakkeVariableHardSphere", dens=1e21, dt=2e-6,
                                rotationalRelaxationCollisionNumber=5.0)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    dens = [0.6e21, 0.2e21, 0.05e21, 0.1e21, 0.05e21]
    start = H.same_start(eng, ora, [0, 1, 2, 3, 4], dens, 5000.0, 5000.0, 5000.0)
    for x in (eng, ora):
        x.stage(capi.STAGE_COLLIDE)
    g, o = eng.download_parcels(), ora.download_parcels()
    assert eng.counters().collisions == ora.counters()["collisions"] > 0
    assert np.array_equal(g.vibLevel, o.vibLevel)
    assert np.array_equal(g.ELevel, o.ELevel)
    assert np.allclose(g.U, o.U, rtol=0, atol=1e-8)
    assert np.allclose(g.ERot, o.ERot, rtol=1e-10, atol=1e-30)
    # total energy (translational + rotational + vibrational) per cell conserved to 1e-12
    mass = np.array([s.mass for s in sp])
    thv = np.array([s.thetaV[0] for s in sp])

    def energy(p):
        return 0.5 * mass[p.typeId] * (p.U ** 2).sum(1) + p.ERot + p.vibLevel[:, 0] * H.KB * thv[p.typeId]

    off = eng.occupancy()
    e0 = np.add.reduceat(energy(start), off[:-1])
    e1 = np.add.reduceat(energy(g), off[:-1])
    assert np.abs(e1 / e0 - 1).max() < 1e-12
    eng.close()


def test_collide_inverse_zv_2008_with_macroscopic_temperature_matches_oracle():
    """inverseZvFormulation "2008": Zv at fields().overallT(cell) where it exists (> SMALL), the collision temperature elsewhere
    (dsmcCloud.C:1441-1456); the temperature field comes in through dsmcb200_upload_overall_temperature."""
    sp = H.air5()
    mesh, _, md = periodic_case((6, 5, 4), ppc=50, species=sp, model="LarsenBorgnakkeVariableHardSphere", dens=1e21, dt=2e-6,
                                inverseZvFormulation="2008")
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    start = H.same_start(eng, ora, [0, 1, 2, 3, 4], [0.6e21, 0.2e21, 0.05e21, 0.1e21, 0.05e21], 5000.0, 5000.0, 5000.0)
    rng = np.random.default_rng(3)
    Tov = rng.uniform(2000.0, 60000.0, mesh.n_cells)
    Tov[::7] = 0.0                                   # cells without a macroscopic temperature yet: fallback
    for x in (eng, ora):
        x.upload_overall_temperature(Tov)
        x.stage(capi.STAGE_COLLIDE)
    g, o = eng.download_parcels(), ora.download_parcels()
    assert eng.counters().collisions == ora.counters()["collisions"] > 1000
    assert np.array_equal(g.vibLevel, o.vibLevel) and (o.vibLevel != start.vibLevel).sum() > 50
    assert np.allclose(g.U, o.U, rtol=0, atol=1e-9) and np.allclose(g.ERot, o.ERot, rtol=1e-10, atol=1e-32)
    eng.close()


def test_collide_soft_sphere_models_match_oracle():
    """VariableSoftSphere and LarsenBorgnakkeVariableSoftSphere (Bird eq. 2.22 scattering with the species' alpha)."""
    sp = H.air5()
    dens = [0.6e21, 0.2e21, 0.05e21, 0.1e21, 0.05e21]
    for model in ("VariableSoftSphere", "LarsenBorgnakkeVariableSoftSphere"):
        mesh, _, md = periodic_case((5, 5, 5), ppc=60, species=sp, model=model, dens=1e21, dt=2e-6)
        eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
        lb = model.startswith("Larsen")
        start = H.same_start(eng, ora, [0, 1, 2, 3, 4], dens, 5000.0, 5000.0 if lb else 0.0, 5000.0 if lb else 0.0)
        for x in (eng, ora):
            x.stage(capi.STAGE_COLLIDE)
        g, o = eng.download_parcels(), ora.download_parcels()
        assert eng.counters().collisions == ora.counters()["collisions"] > 0
        assert np.allclose(g.U, o.U, rtol=0, atol=1e-8)
        mass = np.array([s.mass for s in sp])
        off = eng.occupancy()
        p0 = np.add.reduceat(start.U * mass[start.typeId][:, None], off[:-1], axis=0)
        p1 = np.add.reduceat(g.U * mass[g.typeId][:, None], off[:-1], axis=0)
        assert np.abs(p1 - p0).max() / (mass.max() * 3000 * 60) < 1e-12
        if lb:
            assert np.array_equal(g.vibLevel, o.vibLevel)
            assert np.allclose(g.ERot, o.ERot, rtol=1e-10, atol=1e-30)
        eng.close()


def test_collide_cells_above_and_below_the_lane_kernel_limit():
    """Mean occupancy 250: about half of the cells exceed the 255-parcel limit of collideLaneKernel and are handed to
    collideBigCellsKernel through the big-cell list; both must reproduce the serial oracle."""
    sp = H.air5()
    mesh, _, md = periodic_case((4, 4, 3), ppc=250, species=sp, model="LarsenBorgnakkeVariableHardSphere", dens=1e21, dt=2e-6,
                                rotationalRelaxationCollisionNumber=5.0)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    dens = [0.6e21, 0.2e21, 0.05e21, 0.1e21, 0.05e21]
    H.same_start(eng, ora, [0, 1, 2, 3, 4], dens, 5000.0, 5000.0, 5000.0)
    occ = np.diff(eng.occupancy())
    assert (occ > 255).any() and (occ <= 255).any()
    for x in (eng, ora):
        x.stage(capi.STAGE_COLLIDE)
    g, o = eng.download_parcels(), ora.download_parcels()
    c, oc = eng.counters(), ora.counters()
    assert c.collisionCandidates == oc["collisionCandidates"]
    assert c.collisions == oc["collisions"] > 0
    assert np.array_equal(g.vibLevel, o.vibLevel)
    assert np.array_equal(g.ELevel, o.ELevel)
    assert np.allclose(g.U, o.U, rtol=0, atol=1e-8)
    assert np.allclose(g.ERot, o.ERot, rtol=1e-10, atol=1e-30)
    gs, gr = eng.download_cellstate()
    os_, or_ = ora.download_cellstate()
    assert np.allclose(gs, os_, rtol=1e-13)
    assert np.array_equal(gr, or_)
    eng.close()


def test_sample_matches_oracle():
    sp = H.air5()
    mesh, _, md = periodic_case((5, 5, 5), ppc=50, species=sp, model="LarsenBorgnakkeVariableHardSphere", dens=1e21, dt=2e-6,
                                measureHeatFluxShearStress=True)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, [0, 1, 2, 3, 4], [0.6e21, 0.2e21, 0.05e21, 0.1e21, 0.05e21], 3000.0, 3000.0, 3000.0, (100.0, 20.0, -5.0))
    for x in (eng, ora):
        x.stage(capi.STAGE_COLLIDE)
        x.stage(capi.STAGE_SAMPLE)
        x.stage(capi.STAGE_SAMPLE)
    ga, gc, gn = eng.accumulators()
    oa, oc, on = ora.accumulators()
    assert gn == on == 2
    assert np.array_equal(ga[:, :, 0], oa[:, :, 0])  # parcel counts: exact
    scale = np.abs(oa).max(axis=(0, 1), keepdims=True) + 1e-300
    assert (np.abs(ga - oa) / scale).max() < 1e-12  # FP64 sums in a different order
    assert np.allclose(gc, oc, rtol=1e-12)
    eng.close()


def test_evolve_multi_step_tracks_oracle():
    mesh, sp, md = periodic_case((8, 8, 8), ppc=30)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, [0], [1e20], 300.0)
    seen = 0
    for _ in range(5):
        eng.evolve(1)
        ora.evolve(1)
        # the engine reports the step, the oracle the run: collision and candidate counts step by step
        total = ora.counters()
        c = eng.counters()
        assert c.collisions == total["collisions"] - seen > 0
        seen = total["collisions"]
    g, o = eng.download_parcels(), ora.download_parcels()
    assert g.n == o.n
    assert np.array_equal(g.origId, o.origId)
    assert np.array_equal(g.cell, o.cell)           # cell indexing bit-exact over 5 full steps
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-12)
    eng.close()


def wall_case(model="LarsenBorgnakkeVariableHardSphere"):
    # couette-like: x cyclic, y diffuse walls (moving upper wall), z empty (2-D)
    sides = {"xmin": ("cyclic",), "xmax": ("cyclic",), "ymin": ("wall", "lowerWall"), "ymax": ("wall", "upperWall"),
             "zmin": ("empty", "frontAndBack"), "zmax": ("empty", "frontAndBack")}
    mesh = meshgen.box_mesh((5, 20, 1), (0.05, 0.2, 0.01), sides=sides)
    sp = H.air5()[:2]
    pm = [dict(patch=mesh.patch_index("lowerWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=2000.0, velocity=(0, 0, 0)),
          dict(patch=mesh.patch_index("upperWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=3000.0, velocity=(300.0, 0, 0))]
    vol = 0.05 * 0.2 * 0.01
    md = capi.build_models(model, nEquivalentParticles=1e20 * vol / (100 * 60), deltaT=4e-6, seed=7, patch_models=pm,
                           inverseZvFormulation="pre-2008")
    return mesh, sp, md


def test_move_with_diffuse_walls_and_empty_patches():
    mesh, sp, md = wall_case()
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, [0, 1], [0.8e20, 0.2e20], 2500.0, 2500.0, 2500.0)
    for x in (eng, ora):
        x.stage(capi.STAGE_MOVE)
        x.stage(capi.STAGE_SORT)
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert np.array_equal(g["origId"], o["origId"])
    assert np.array_equal(g["cell"], o["cell"])
    assert np.all(g["position"][:, 2] == 0.005)  # constrainToMeshCentre
    hit = np.any(g["U"] != H.by_id(ora.download_parcels())["U"], axis=1)
    assert np.allclose(g["U"], o["U"], rtol=1e-12, atol=1e-9)
    assert np.allclose(g["position"], o["position"], rtol=0, atol=1e-13)
    assert np.array_equal(g["vibLevel"], o["vibLevel"])
    gw, ow = eng.wall_accumulators(), ora.wall_accumulators()
    assert gw.shape == ow.shape and gw.shape[0] == 10
    assert np.abs(ow).sum() > 0
    scale = np.abs(ow).max(axis=(0, 1), keepdims=True) + 1e-300
    assert (np.abs(gw - ow) / scale).max() < 1e-10
    eng.close()


def test_sample_interval_matches_oracle():
    """sampleInterval 2 over 5 steps: cells and wall faces are measured on steps 2 and 4 only (dsmcVolFields.C:1073-1081,1362;
    boundaryMeas_ / cellMeas_ cleaned every step, dsmcCloud.C:923-925)."""
    mesh, sp, _ = wall_case()
    pm = [dict(patch=mesh.patch_index("lowerWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=2000.0, velocity=(0, 0, 0)),
          dict(patch=mesh.patch_index("upperWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=3000.0, velocity=(300.0, 0, 0))]
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=1e20 * 0.05 * 0.2 * 0.01 / (100 * 60), deltaT=4e-6, seed=7,
                           patch_models=pm, inverseZvFormulation="pre-2008", sampleInterval=2)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, [0, 1], [0.8e20, 0.2e20], 2500.0, 2500.0, 2500.0)
    eng.evolve(5)
    ora.evolve(5)
    ga, gc, gn = eng.accumulators()
    oa, oc, on = ora.accumulators()
    assert gn == on == 2
    assert np.array_equal(ga[:, :, 0], oa[:, :, 0])
    scale = np.abs(oa).max(axis=(0, 1), keepdims=True) + 1e-300
    assert (np.abs(ga - oa) / scale).max() < 1e-10
    assert np.allclose(gc, oc, rtol=1e-10)
    gw, ow = eng.wall_accumulators(), ora.wall_accumulators()
    assert np.abs(ow).sum() > 0
    scale = np.abs(ow).max(axis=(0, 1), keepdims=True) + 1e-300
    assert (np.abs(gw - ow) / scale).max() < 1e-9
    eng.close()


def test_sample_sets_with_their_own_intervals_match_oracle_runs():
    """field{} entries with different sampleIntervals (dsmcField.C:113-152): one set of sums per interval (dsmcb200_set_sample_sets).  The
    cloud does not know about sampling, so set k of ONE engine run equals the accumulators of an oracle run made with that interval --
    cell sums, collision sums, wall measurements (only the steps the set samples leave any) and nTimeSteps; resetting one set leaves the
    others alone."""
    intervals = [1, 3, 2]
    mesh, sp, md = wall_case()
    eng = capi.Engine(0)
    eng.set_sample_sets(intervals)
    eng.set_mesh(mesh); eng.set_species(sp); eng.set_models(md)
    oras = []
    for iv in intervals:
        md_k = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=md.nEquivalentParticles, deltaT=4e-6, seed=7,
                                 patch_models=[dict(patch=mesh.patch_index("lowerWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=2000.0, velocity=(0, 0, 0)),
                                               dict(patch=mesh.patch_index("upperWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=3000.0, velocity=(300.0, 0, 0))],
                                 inverseZvFormulation="pre-2008", sampleInterval=iv)
        o = Oracle()
        o.set_mesh(mesh); o.set_species(sp); o.set_models(md_k)
        oras.append(o)
    H.same_start(eng, oras[0], [0, 1], [0.8e20, 0.2e20], 2500.0, 2500.0, 2500.0)
    start = oras[0].download_parcels()
    sig, rem = oras[0].download_cellstate()
    for o in oras[1:]:
        o.upload_parcels(start)
        o.upload_cellstate(sig, rem)
    eng.evolve(7)
    for o in oras:
        o.evolve(7)

    def check(k, o):
        eng.select_sample_set(k)
        ga, gc, gn = eng.accumulators()
        oa, oc, on = o.accumulators()
        assert gn == on == 7 // intervals[k]
        scale = np.abs(oa).max(axis=(0, 1), keepdims=True) + 1e-300
        assert np.abs(oa).sum() > 0 and (np.abs(ga - oa) / scale).max() < 1e-9
        assert np.allclose(gc, oc, rtol=1e-9, atol=0)
        gw, ow = eng.wall_accumulators(), o.wall_accumulators()
        ws = np.abs(ow).max(axis=(0, 1), keepdims=True) + 1e-300
        assert np.abs(ow).sum() > 0 and (np.abs(gw - ow) / ws).max() < 1e-9
        return ga, gw

    sums = [check(k, o) for k, o in enumerate(oras)]
    assert np.abs(sums[0][0]).sum() > np.abs(sums[2][0]).sum() > np.abs(sums[1][0]).sum()     # 7, 3 and 2 samples
    eng.select_sample_set(1)
    eng.reset_accumulators()
    a1, _, n1 = eng.accumulators()
    assert n1 == 0 and not a1.any() and not eng.wall_accumulators().any()
    eng.select_sample_set(2)
    a2, _, n2 = eng.accumulators()
    assert n2 == 3 and np.array_equal(a2, sums[2][0]) and np.array_equal(eng.wall_accumulators(), sums[2][1])
    with pytest.raises(capi.Dsmcb200Error, match="no such set"):
        eng.select_sample_set(3)
    eng.close()


def test_diffuse_wall_with_linear_temperature_matches_oracle():
    """dsmcDiffuseWallPatch::getLocalTemperature (dsmcDiffuseWallPatch.C:141-148): groundLevelTemperature / formationLevelTemperature /
    depthAxis -- the wall temperature is a linear function of the hit position along the depth axis of the mesh bounds."""
    sides = {"xmin": ("wall", "walls"), "xmax": ("wall", "walls"), "ymin": ("symmetryPlane", "ends"), "ymax": ("symmetryPlane", "ends"),
             "zmin": ("cyclic",), "zmax": ("cyclic",)}
    mesh = meshgen.box_mesh((3, 12, 2), (0.003, 0.06, 0.002), sides=sides)
    sp = H.air5()[:2]
    pm = [dict(patch=mesh.patch_index("walls"), boundaryModel="dsmcDiffuseWallPatch", temperature=3000.0, formationLevelTemperature=500.0,
               depthAxis="y", velocity=(0.0, 50.0, 0.0))]
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=1e20 * 0.003 * 0.06 * 0.002 / (72 * 80), deltaT=2e-6, seed=13,
                           patch_models=pm)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    start = H.by_id(H.same_start(eng, ora, [0, 1], [0.8e20, 0.2e20], 1000.0, 1000.0, 1000.0))
    for x in (eng, ora):
        x.stage(capi.STAGE_MOVE)
        x.stage(capi.STAGE_SORT)
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert np.array_equal(g["cell"], o["cell"])
    assert np.allclose(g["U"], o["U"], rtol=1e-12, atol=1e-9) and np.allclose(g["position"], o["position"], rtol=0, atol=1e-13)
    assert np.array_equal(g["vibLevel"], o["vibLevel"])
    hit = np.abs((o["U"] ** 2).sum(1) / (start["U"] ** 2).sum(1) - 1) > 1e-9
    assert hit.sum() > 300
    # hotter at the ground level (y_max) than at the formation level (y_min)
    y, e = o["position"][hit, 1], (o["U"][hit] ** 2).sum(1)
    assert e[y > 0.045].mean() > 2.5 * e[y < 0.015].mean()
    gw, ow = eng.wall_accumulators(), ora.wall_accumulators()
    scale = np.abs(ow).max(axis=(0, 1), keepdims=True) + 1e-300
    assert (np.abs(gw - ow) / scale).max() < 1e-10
    eng.close()


def test_diffuse_specular_wall_matches_oracle():
    """dsmcDiffuseSpecularWallPatch: the diffuse / specular draw comes first in the hit's Philox stream on both sides."""
    sides = {"xmin": ("cyclic",), "xmax": ("cyclic",), "ymin": ("wall", "lowerWall"), "ymax": ("wall", "upperWall"),
             "zmin": ("empty", "frontAndBack"), "zmax": ("empty", "frontAndBack")}
    mesh = meshgen.box_mesh((5, 20, 1), (0.05, 0.2, 0.01), sides=sides)
    sp = H.air5()[:2]
    pm = [dict(patch=mesh.patch_index("lowerWall"), boundaryModel="dsmcDiffuseSpecularWallPatch", temperature=2000.0, velocity=(0, 0, 0),
               diffuseFraction=0.4),
          dict(patch=mesh.patch_index("upperWall"), boundaryModel="dsmcDiffuseSpecularWallPatch", temperature=3000.0, velocity=(300.0, 0, 0),
               diffuseFraction=0.85)]
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=1e20 * 0.05 * 0.2 * 0.01 / (100 * 60), deltaT=4e-6, seed=7,
                           patch_models=pm, inverseZvFormulation="pre-2008")
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    start = H.by_id(H.same_start(eng, ora, [0, 1], [0.8e20, 0.2e20], 2500.0, 2500.0, 2500.0))
    for x in (eng, ora):
        x.stage(capi.STAGE_MOVE)
        x.stage(capi.STAGE_SORT)
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert np.array_equal(g["cell"], o["cell"])
    assert np.allclose(g["U"], o["U"], rtol=1e-12, atol=1e-9) and np.allclose(g["position"], o["position"], rtol=0, atol=1e-13)
    assert np.array_equal(g["vibLevel"], o["vibLevel"])
    hit = (o["U"] != start["U"]).any(1)
    specular = hit & (np.abs((o["U"] ** 2).sum(1) / (start["U"] ** 2).sum(1) - 1) < 1e-12)
    assert 0 < specular.sum() < hit.sum()                        # both branches taken
    assert np.array_equal(g["U"][specular], o["U"][specular])    # no libm in a specular reflection: bit-exact
    gw, ow = eng.wall_accumulators(), ora.wall_accumulators()
    scale = np.abs(ow).max(axis=(0, 1), keepdims=True) + 1e-300
    assert (np.abs(gw - ow) / scale).max() < 1e-10
    eng.close()


def test_cll_wall_matches_oracle():
    """dsmcCLLWallPatch (Cercignani-Lampis-Lord kernel with Lord's rotational extension, dsmcCLLWallPatch.C:100-300): same draws in the same
    order from the hit's Philox stream on both sides; a moving wall, a diatomic and a monatomic species; vibration / electronic levels untouched.
    The second wall has both coefficients zero: specular, and it measures nothing (dsmcCLLWallPatch.C:82-89)."""
    sides = {"xmin": ("cyclic",), "xmax": ("cyclic",), "ymin": ("wall", "lowerWall"), "ymax": ("wall", "upperWall"),
             "zmin": ("cyclic",), "zmax": ("cyclic",)}
    mesh = meshgen.box_mesh((5, 12, 3), (0.05, 0.12, 0.03), sides=sides)
    a5 = H.air5()
    sp = [a5[0], a5[3]]   # N2, N
    pm = [dict(patch=mesh.patch_index("lowerWall"), boundaryModel="dsmcCLLWallPatch", temperature=1200.0, velocity=(250.0, 0, -40.0),
               normalAccommodationCoefficient=0.7, tangentialAccommodationCoefficient=0.45, rotationalEnergyAccommodationCoefficient=0.6),
          dict(patch=mesh.patch_index("upperWall"), boundaryModel="dsmcCLLWallPatch", temperature=500.0, velocity=(0, 0, 0),
               normalAccommodationCoefficient=0.0, tangentialAccommodationCoefficient=0.0, rotationalEnergyAccommodationCoefficient=0.0)]
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=1e20 * 0.05 * 0.12 * 0.03 / (180 * 60), deltaT=6e-6, seed=19,
                           patch_models=pm, inverseZvFormulation="pre-2008")
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    start = H.by_id(H.same_start(eng, ora, [0, 1], [0.7e20, 0.3e20], 2000.0, 2000.0, 2000.0))
    for x in (eng, ora):
        x.stage(capi.STAGE_MOVE)
        x.stage(capi.STAGE_SORT)
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert np.array_equal(g["cell"], o["cell"])
    assert np.allclose(g["U"], o["U"], rtol=1e-11, atol=1e-8) and np.allclose(g["position"], o["position"], rtol=0, atol=1e-13)
    assert np.allclose(g["ERot"], o["ERot"], rtol=1e-10, atol=1e-30)
    assert np.array_equal(g["vibLevel"], o["vibLevel"]) and np.array_equal(g["vibLevel"], start["vibLevel"])
    assert np.array_equal(g["ELevel"], start["ELevel"])
    hit = (o["U"] != start["U"]).any(1)
    upper = hit & (o["position"][:, 1] > 0.06)
    lower = hit & ~upper
    assert upper.sum() > 100 and lower.sum() > 100
    # specular limit of the kernel: speed kept
    assert np.allclose((o["U"][upper] ** 2).sum(1), (start["U"][upper] ** 2).sum(1), rtol=1e-10)
    # the rotational energy of a hit N2 changed on the accommodating wall, an atom's stays zero
    n2 = o["typeId"] == 0
    assert (o["ERot"][lower & n2] != start["ERot"][lower & n2]).mean() > 0.99 and np.all(o["ERot"][~n2] == 0)
    gw, ow = eng.wall_accumulators(), ora.wall_accumulators()
    scale = np.abs(ow).max(axis=(0, 1), keepdims=True) + 1e-300
    assert (np.abs(gw - ow) / scale).max() < 1e-10
    # faces of the zero-coefficient wall carry nothing (measurePropertiesAtWall_ = false)
    per_face = np.abs(ow).reshape(ow.shape[0], -1).sum(1)   # measurement faces in dsmcPatchBoundaries order: 15 lower, 15 upper
    assert per_face.shape == (30,) and np.all(per_face[:15] > 0) and np.all(per_face[15:] == 0)
    assert np.all(np.abs(gw).reshape(30, -1).sum(1)[15:] == 0)
    eng.evolve(3)
    ora.evolve(3)
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert np.array_equal(g["cell"], o["cell"]) and np.allclose(g["U"], o["U"], rtol=1e-9, atol=1e-7)
    eng.close()


def test_inflow_deletion_specular_counts_match_oracle():
    sides = {"xmin": ("patch", "inlet"), "xmax": ("patch", "outlet"), "ymin": ("wall", "plate"), "ymax": ("symmetryPlane", "top"),
             "zmin": ("cyclic",), "zmax": ("cyclic",)}
    mesh = meshgen.box_mesh((10, 6, 4), (0.1, 0.06, 0.04), sides=sides)
    sp = [H.argon()]
    pm = [dict(patch=mesh.patch_index("inlet"), boundaryModel="dsmcDeletionPatch"),
          dict(patch=mesh.patch_index("outlet"), boundaryModel="dsmcDeletionPatch"),
          dict(patch=mesh.patch_index("plate"), boundaryModel="dsmcSpecularWallPatch")]
    inflow = [dict(patch=mesh.patch_index("inlet"), typeIds=[0], numberDensities=[1e20], velocity=(1936.0, 0, 0), translationalTemperature=300.0)]
    vol = 0.1 * 0.06 * 0.04
    md = capi.build_models("VariableHardSphere", nEquivalentParticles=1e20 * vol / (240 * 30), deltaT=2e-6, seed=11, patch_models=pm, inflows=inflow)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, [0], [1e20], 300.0, velocity=(1936.0, 0, 0))
    n0 = ora.num_parcels()
    eng.evolve(4)
    ora.evolve(4)
    g, o = eng.download_parcels(), ora.download_parcels()
    assert g.n == o.n and g.n != n0
    assert np.array_equal(g.origId, o.origId)
    assert np.array_equal(g.cell, o.cell)
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-12)
    assert np.allclose(g.U, o.U, rtol=1e-12, atol=1e-9)
    eng.close()


def test_face_tracker_fluxes_match_oracle():
    """dsmcFaceTracker (DSMC/faceTracker/dsmcFaceTracker.C:124-198) on every face kind of one case: internal faces, cyclic (credited to
    the coupled face), inflow insertions (dsmcCloud.C:429-437), deletion patches, a specular wall and a symmetry plane.  Parcel counts
    per (species, face) are integers: exact."""
    sides = {"xmin": ("patch", "inlet"), "xmax": ("patch", "outlet"), "ymin": ("wall", "plate"), "ymax": ("symmetryPlane", "top"),
             "zmin": ("cyclic",), "zmax": ("cyclic",)}
    mesh = meshgen.box_mesh((10, 6, 4), (0.1, 0.06, 0.04), sides=sides)
    sp = H.air5()[:2]
    pm = [dict(patch=mesh.patch_index("inlet"), boundaryModel="dsmcDeletionPatch"),
          dict(patch=mesh.patch_index("outlet"), boundaryModel="dsmcDeletionPatch"),
          dict(patch=mesh.patch_index("plate"), boundaryModel="dsmcSpecularWallPatch")]
    inflow = [dict(patch=mesh.patch_index("inlet"), typeIds=[0, 1], numberDensities=[0.8e20, 0.2e20], velocity=(1500.0, 0, 0),
                   translationalTemperature=300.0, rotationalTemperature=300.0, vibrationalTemperature=300.0)]
    vol = 0.1 * 0.06 * 0.04
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=1e20 * vol / (240 * 30), deltaT=2e-6, seed=11,
                           patch_models=pm, inflows=inflow, trackFaceFluxes=True)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, [0, 1], [0.8e20, 0.2e20], 300.0, 300.0, 300.0, velocity=(1500.0, 0, 0))
    nI = mesh.n_internal
    for _ in range(3):
        eng.evolve(1)
        ora.evolve(1)
        gp, gm = eng.face_fluxes()
        op, om = ora.face_fluxes()
        assert np.array_equal(gp, op)
        for k, p in enumerate(mesh.patches):
            assert np.abs(op[:, p["start"]:p["start"] + p["size"]]).sum() > 0, p["name"]   # every patch kind saw crossings
        assert np.abs(op[:, :nI]).sum() > 1000
        scale = np.abs(om).max() + 1e-300
        assert np.abs(gm - om).max() / scale < 1e-12
    assert eng.num_parcels() == ora.num_parcels()
    eng.close()


def test_equilibrium_collision_rate_within_one_percent():
    # VHS equilibrium collision rate, single species, periodic box
    mesh, sp, md = periodic_case((16, 16, 16), L=0.064, ppc=32)
    eng = capi.Engine(0)
    eng.set_mesh(mesh); eng.set_species(sp); eng.set_models(md)
    eng.mesh_fill([0], [1e20], 300.0)
    n = eng.num_parcels()
    eng.evolve(20)   # let sigmaTcRMax settle
    total = 0
    steps = 60
    for _ in range(steps):
        eng.evolve(1)
        total += eng.counters().collisions
    vol = 0.064 ** 3
    expected = H.vhs_equilibrium_collision_rate(1e20, 300.0, sp[0]) * vol * md.deltaT * steps / md.nEquivalentParticles
    assert abs(total / expected - 1) < 0.01, (total, expected, n)
    eng.close()


def test_mesh_fill_on_device_matches_oracle():
    sp = H.air5()
    mesh, _, md = periodic_case((4, 4, 4), ppc=40, species=sp, model="LarsenBorgnakkeVariableHardSphere", dens=1e21)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    dens = [0.6e21, 0.2e21, 0.05e21, 0.1e21, 0.05e21]
    eng.mesh_fill([0, 1, 2, 3, 4], dens, 4000.0, 4000.0, 4000.0)
    ora.mesh_fill([0, 1, 2, 3, 4], dens, 4000.0, 4000.0, 4000.0)
    g, o = eng.download_parcels(), ora.download_parcels()
    assert g.n == o.n
    assert np.array_equal(g.cell, o.cell) and np.array_equal(g.typeId, o.typeId)
    assert np.array_equal(g.tetFace, o.tetFace) and np.array_equal(g.tetPt, o.tetPt)
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-15)
    assert np.allclose(g.U, o.U, rtol=1e-12, atol=1e-9)
    assert np.array_equal(g.vibLevel, o.vibLevel)
    eng.close()


def test_zone_fill_on_device_matches_oracle():
    """dsmcZoneFill (initialiseDsmcParcels/derived/dsmcZoneFill/dsmcZoneFill.C:71-272): two configurations of one dsmcInitialiseDict, a
    driver and a driven section at different states, the second zone's cells in descending order; the cloud is the oracle's parcel for
    parcel (identified by origId: zone order, then tet, species and insertion order), sigmaTcRMax is set zone by zone, and three steps
    from there stay on the oracle's."""
    sp = H.air5()[:2]
    mesh, _, md = periodic_case((6, 4, 4), ppc=40, species=sp, model="LarsenBorgnakkeVariableHardSphere", dens=1e21)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    x = np.arange(mesh.n_cells) % 6
    left, right = np.flatnonzero(x < 2).astype(np.int32), np.flatnonzero(x >= 2)[::-1].astype(np.int32)
    for e in (eng, ora):
        e.upload_parcels(capi.ParcelData(0, 1))
        e.zone_fill(left, [0, 1], [2.4e21, 0.6e21], 3000.0, 3000.0, 3000.0, velocity=(400.0, 0, 0))
        e.zone_fill(right, [0], [0.4e21], 300.0, 300.0, 300.0)
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert len(g["cell"]) == len(o["cell"]) > 3000
    assert np.array_equal(g["origId"], o["origId"]) and len(np.unique(g["origId"])) == len(g["origId"])
    assert np.array_equal(g["cell"], o["cell"]) and np.array_equal(g["typeId"], o["typeId"])
    assert np.array_equal(g["tetFace"], o["tetFace"]) and np.array_equal(g["tetPt"], o["tetPt"])
    assert np.allclose(g["position"], o["position"], rtol=0, atol=1e-15)
    assert np.allclose(g["U"], o["U"], rtol=1e-12, atol=1e-9) and np.allclose(g["ERot"], o["ERot"], rtol=1e-12, atol=1e-30)
    assert np.array_equal(g["vibLevel"], o["vibLevel"])
    in_left = np.isin(o["cell"], left)
    assert in_left.sum() > 3 * (~in_left).sum() and np.all(o["typeId"][~in_left] == 0)
    gs, _ = eng.download_cellstate()
    os_, _ = ora.download_cellstate()
    assert np.allclose(gs, os_, rtol=1e-14) and len(np.unique(np.round(os_ / os_.max(), 12))) == 2
    seen = ora.counters()["collisions"]
    for _ in range(3):
        eng.evolve(1)
        ora.evolve(1)
        total = ora.counters()["collisions"]     # the oracle's counter runs on, the engine's is the last step's
        assert eng.counters().collisions == total - seen > 0
        seen = total
    g, o = H.by_id(eng.download_parcels()), H.by_id(ora.download_parcels())
    assert np.array_equal(g["cell"], o["cell"]) and np.allclose(g["U"], o["U"], rtol=1e-9, atol=1e-7)
    eng.close()


def test_cylinder_ogrid_inflow_wall_matches_oracle():
    """BASELINE configs[1] topology at test size: body-fitted O-grid (non-axis-aligned hexahedra, 2-D with empty
    patches), hypersonic free stream, diffuse cylinder wall, inflow + deletion on the outer boundary."""
    from hystrath_b200 import cases

    mesh, sp, md, fill = cases.lofthouse_cylinder(nr=24, ntheta=64, ppc=20, r_out=0.4)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    H.same_start(eng, ora, fill["type_ids"], fill["number_densities"], fill["Ttra"], velocity=fill["velocity"])
    n0 = ora.num_parcels()
    eng.evolve(6)
    ora.evolve(6)
    g, o = eng.download_parcels(), ora.download_parcels()
    oc = ora.counters()
    assert oc["inserted"] > 0 and oc["deleted"] > 0
    assert g.n == o.n
    assert np.array_equal(g.origId, o.origId)
    assert np.array_equal(g.cell, o.cell)
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    assert np.array_equal(g.tetFace, o.tetFace) and np.array_equal(g.tetPt, o.tetPt)
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-12)
    assert np.allclose(g.U, o.U, rtol=1e-12, atol=1e-8)
    gw, ow = eng.wall_accumulators(), ora.wall_accumulators()
    assert np.abs(ow).sum() > 0
    scale = np.abs(ow).max(axis=(0, 1), keepdims=True) + 1e-300
    assert (np.abs(gw - ow) / scale).max() < 1e-9
    # every parcel stays between the cylinder and the outer boundary
    r = np.hypot(g.position[:, 0], g.position[:, 1])
    assert r.min() >= 0.1524 * np.cos(np.pi / 64) - 1e-12 and r.max() <= 0.4 + 1e-12
    eng.close()


def test_air_wedge_small_scale_matches_oracle():
    """BASELINE configs[2] at test size (hystrath_b200.cases.air_wedge: sheared hexahedra, diffuse 1000 K wedge, free-stream inflow +
    deletion on inlet / top / outlet, symmetry planes in z, 5-species Larsen-Borgnakke with variable Zv): insertions, deletions, cells,
    list order and collision counts are the oracle's step by step; velocities to libm ulps; wall accumulators to 1e-10."""
    from hystrath_b200 import cases

    mesh, sp, md, fill = cases.air_wedge(40, 20, 2, ppc=20, density_scale=40.0)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    ora.set_reorder(True)
    H.same_start(eng, ora, fill["type_ids"], fill["number_densities"], fill["Ttra"], fill["Trot"], fill["Tvib"], fill["velocity"])
    seen = dict(collisions=0, inserted=0, deleted=0)
    for _ in range(6):
        eng.evolve(1)
        ora.evolve(1)
        tot, c = ora.counters(), eng.counters()
        assert c.inserted == tot["inserted"] - seen["inserted"] > 0
        assert c.deleted == tot["deleted"] - seen["deleted"] > 0
        assert c.collisions == tot["collisions"] - seen["collisions"] > 50
        seen = tot
    g, o = eng.download_parcels(), ora.download_parcels()
    assert g.n == o.n
    assert np.array_equal(g.origId, o.origId) and np.array_equal(g.typeId, o.typeId)
    assert np.array_equal(g.cell, o.cell) and np.array_equal(eng.occupancy(), ora.occupancy())
    assert np.array_equal(g.vibLevel, o.vibLevel)
    assert len(np.unique(g.typeId)) == 5                      # every species path ran
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-9) and np.allclose(g.U, o.U, rtol=1e-11, atol=1e-7)
    gw, ow = eng.wall_accumulators(), ora.wall_accumulators()
    assert np.abs(ow).max() > 0 and np.abs(gw - ow).max() <= 1e-9 * np.abs(ow).max()
    eng.close()


def test_capsule_forebody_small_scale_matches_oracle():
    """BASELINE configs[3] at test size (hystrath_b200.cases.capsule_forebody: hexahedra compressed towards a spherical-segment heat
    shield, so most faces are warped and every track runs through the tet decomposition; diffuse 1000 K shield, free-stream inflow and
    deletion on the inlet and the four lateral planes, outflow around the shoulder; 5-species Larsen-Borgnakke with variable Zv):
    insertions, deletions, cells, list order and collision counts are the oracle's step by step; wall accumulators to 1e-9."""
    from hystrath_b200 import cases

    mesh, sp, md, fill = cases.capsule_forebody((14, 12, 12), ppc=16, density_scale=40.0)
    assert [p["name"] for p in mesh.patches] == ["flow", "capsule", "outflow"] and mesh.patches[1]["size"] > 30
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    ora.set_reorder(True)
    H.same_start(eng, ora, fill["type_ids"], fill["number_densities"], fill["Ttra"], fill["Trot"], fill["Tvib"], fill["velocity"])
    seen = dict(collisions=0, inserted=0, deleted=0)
    for _ in range(6):
        eng.evolve(1)
        ora.evolve(1)
        tot, c = ora.counters(), eng.counters()
        assert c.inserted == tot["inserted"] - seen["inserted"] > 0
        assert c.deleted == tot["deleted"] - seen["deleted"] > 0
        assert c.collisions == tot["collisions"] - seen["collisions"] > 50
        seen = tot
    g, o = eng.download_parcels(), ora.download_parcels()
    assert g.n == o.n
    assert np.array_equal(g.origId, o.origId) and np.array_equal(g.typeId, o.typeId)
    assert np.array_equal(g.cell, o.cell) and np.array_equal(eng.occupancy(), ora.occupancy())
    assert np.array_equal(g.vibLevel, o.vibLevel)
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-9) and np.allclose(g.U, o.U, rtol=1e-11, atol=1e-7)
    gw, ow = eng.wall_accumulators(), ora.wall_accumulators()
    assert np.abs(ow).max() > 0 and np.abs(gw - ow).max() <= 1e-9 * np.abs(ow).max()
    eng.close()


def test_one_million_parcels_track_the_oracle():
    """BASELINE configs[0] at its own size (32^3 cells, ~1.05 M argon parcels, the C1 case of bench.py): three full steps, collision counts
    step by step, cloud order, cells and occupancy exact, positions to rounding -- the parity statement of the small cases at the size of a
    bench workload (work list with hundreds of windows per block, 32 768 cells through both collide kernels' bookkeeping)."""
    mesh, sp, md = periodic_case((32, 32, 32), L=0.128, ppc=32, dens=1e20, dt=5e-6)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    start = H.same_start(eng, ora, [0], [1e20], 300.0)
    assert abs(start.n - 1048576) < 5000
    seen = 0
    for _ in range(3):
        eng.evolve(1)
        ora.evolve(1)
        total = ora.counters()
        assert eng.counters().collisions == total["collisions"] - seen > 50000
        seen = total["collisions"]
    g, o = eng.download_parcels(), ora.download_parcels()
    assert g.n == o.n == start.n
    assert np.array_equal(g.origId, o.origId) and np.array_equal(g.cell, o.cell)
    assert np.array_equal(g.tetFace, o.tetFace) and np.array_equal(g.tetPt, o.tetPt)
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-12) and np.allclose(g.U, o.U, rtol=1e-9, atol=1e-7)
    ga, gc, _ = eng.accumulators()
    oa, oc, _ = ora.accumulators()
    assert np.array_equal(ga[:, :, 0], oa[:, :, 0]) and np.allclose(gc, oc, rtol=1e-12)
    eng.close()
