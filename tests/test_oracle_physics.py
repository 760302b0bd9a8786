"""CPU tests of the oracle's physics: the identities SURVEY.md 8c lists as the parity anchors where no
reference-side vector exists (conservation per collision, NTC candidate counts, equilibrium collision rate,
wall / inflow fluxes)."""
import numpy as np
import pytest

from hystrath_b200 import capi, meshgen
from oracle.pyoracle import Oracle
from tests import helpers as H


def box(n, L, species, model="VariableHardSphere", ppc=30, dens=1e20, dt=5e-6, sides=None, **kw):
    mesh = meshgen.box_mesh(n, L, sides=sides)
    vol = np.prod(L)
    fnum = dens * vol / (np.prod(n) * ppc)
    md = capi.build_models(model, nEquivalentParticles=fnum, deltaT=dt, seed=42, **kw)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(species); o.set_models(md)
    return mesh, md, o


def test_vhs_collisions_conserve_momentum_and_energy_per_cell():
    sp = [H.argon()]
    mesh, md, o = box((4, 4, 4), (0.016,) * 3, sp, ppc=50)
    o.mesh_fill([0], [1e20], 300.0)
    a = o.download_parcels()
    o.stage(capi.STAGE_COLLIDE)
    b = o.download_parcels()
    assert o.counters()["collisions"] > 50
    off = o.occupancy()
    m = sp[0].mass
    p0, p1 = np.add.reduceat(a.U, off[:-1], axis=0) * m, np.add.reduceat(b.U, off[:-1], axis=0) * m
    e0, e1 = np.add.reduceat((a.U ** 2).sum(1), off[:-1]), np.add.reduceat((b.U ** 2).sum(1), off[:-1])
    assert np.abs(p1 - p0).max() / (m * 400 * 50) < 1e-12
    assert np.abs(e1 / e0 - 1).max() < 1e-12
    assert np.array_equal(a.position, b.position)


def test_larsen_borgnakke_conserves_total_energy():
    sp = H.air5()
    mesh, md, o = box((3, 3, 3), (0.012,) * 3, sp, "LarsenBorgnakkeVariableHardSphere", ppc=80, dens=1e21, dt=2e-6,
                      rotationalRelaxationCollisionNumber=1.0, vibrationalRelaxationCollisionNumber=1.0)
    o.mesh_fill([0, 1, 2, 3, 4], [0.5e21, 0.2e21, 0.1e21, 0.1e21, 0.1e21], 8000.0, 2000.0, 1000.0)
    mass = np.array([s.mass for s in sp]); thv = np.array([s.thetaV[0] for s in sp])

    def energy(p):
        return 0.5 * mass[p.typeId] * (p.U ** 2).sum(1) + p.ERot + p.vibLevel[:, 0] * H.KB * thv[p.typeId]

    a = o.download_parcels()
    o.stage(capi.STAGE_COLLIDE)
    b = o.download_parcels()
    off = o.occupancy()
    e0, e1 = np.add.reduceat(energy(a), off[:-1]), np.add.reduceat(energy(b), off[:-1])
    assert np.abs(e1 / e0 - 1).max() < 1e-12
    assert (a.vibLevel != b.vibLevel).sum() > 0 and (a.ERot != b.ERot).sum() > 0
    # atoms carry no internal energy
    assert np.all(b.ERot[b.typeId >= 3] == 0) and np.all(b.vibLevel[b.typeId >= 3] == 0)


def test_variable_soft_sphere_conserves_momentum_and_energy_and_scatters_forward():
    """VariableSoftSphere (collisions/derived/VariableSoftSphere/VariableSoftSphere.C:195-262): Bird eq. 2.22 keeps |c_r|, so
    momentum and kinetic energy per cell are conserved; with alpha > 1 the deflection cosine 2 R^(1/alpha) - 1 is biased
    forward: <cos chi> = (alpha - 1)/(alpha + 1)."""
    ar = H.argon()
    ar.alpha = 1.66
    sp = [ar]
    mesh, md, o = box((4, 4, 4), (0.016,) * 3, sp, "VariableSoftSphere", ppc=400, dens=3e20)
    o.mesh_fill([0], [3e20], 300.0)
    a = o.download_parcels()
    o.stage(capi.STAGE_COLLIDE)
    b = o.download_parcels()
    assert o.counters()["collisions"] > 2000
    off = o.occupancy()
    m = sp[0].mass
    p0, p1 = np.add.reduceat(a.U, off[:-1], axis=0) * m, np.add.reduceat(b.U, off[:-1], axis=0) * m
    e0, e1 = np.add.reduceat((a.U ** 2).sum(1), off[:-1]), np.add.reduceat((b.U ** 2).sum(1), off[:-1])
    assert np.abs(p1 - p0).max() / (m * 400 * 400) < 1e-12
    assert np.abs(e1 / e0 - 1).max() < 1e-12
    # parcels that collided exactly once: their velocity change gives the pair's deflection cosine
    changed = np.nonzero((a.U != b.U).any(1))[0]
    dU = b.U[changed] - a.U[changed]
    # pairs share |dU|: match partners inside each cell by equal and opposite momentum change
    cos = []
    cell = a.cell[changed]
    for c in np.unique(cell):
        idx = changed[cell == c]
        d = b.U[idx] - a.U[idx]
        for i in range(len(idx)):
            j = np.nonzero(np.abs(d + d[i]).max(1) < 1e-9 * np.abs(d[i]).max())[0]
            if len(j) == 1 and j[0] > i:
                cr0 = a.U[idx[i]] - a.U[idx[j[0]]]
                cr1 = b.U[idx[i]] - b.U[idx[j[0]]]
                assert abs(np.linalg.norm(cr1) / np.linalg.norm(cr0) - 1) < 1e-11
                cos.append(cr0 @ cr1 / (cr0 @ cr0))
    cos = np.array(cos)
    assert len(cos) > 1000
    expect = (1.66 - 1) / (1.66 + 1)
    assert abs(cos.mean() - expect) < 4 * cos.std() / np.sqrt(len(cos))


def test_larsen_borgnakke_soft_sphere_follows_the_reference_formula():
    """LarsenBorgnakkeVariableSoftSphere::collide rescales c_r after the energy exchange and hands it to Bird's eq. 2.22 together with
    the PRE-exchange relative velocity components (LarsenBorgnakkeVariableSoftSphere.C:110-147, VariableSoftSphere.C:195-262).
    Eq. 2.22 is a rotation only when c_r = |c_r components|, so the reference conserves total energy exactly for pairs that exchanged
    nothing and only approximately otherwise; momentum is conserved always.  The restatement keeps that behaviour (SURVEY quirk list)."""
    sp = H.air5()
    mass = np.array([s.mass for s in sp]); thv = np.array([s.thetaV[0] for s in sp])

    def energy(p):
        return 0.5 * mass[p.typeId] * (p.U ** 2).sum(1) + p.ERot + p.vibLevel[:, 0] * H.KB * thv[p.typeId]

    for zrot, exact in ((1e30, True), (1.0, False)):
        mesh, md, o = box((3, 3, 3), (0.012,) * 3, sp, "LarsenBorgnakkeVariableSoftSphere", ppc=80, dens=1e21, dt=2e-6,
                          rotationalRelaxationCollisionNumber=zrot, vibrationalRelaxationCollisionNumber=zrot,
                          electronicRelaxationCollisionNumber=1e30)
        o.mesh_fill([0, 1, 2, 3, 4], [0.5e21, 0.2e21, 0.1e21, 0.1e21, 0.1e21], 8000.0, 2000.0, 1000.0)
        a = o.download_parcels()
        o.stage(capi.STAGE_COLLIDE)
        b = o.download_parcels()
        off = o.occupancy()
        assert o.counters()["collisions"] > 50
        p0 = np.add.reduceat(a.U * mass[a.typeId][:, None], off[:-1], axis=0)
        p1 = np.add.reduceat(b.U * mass[b.typeId][:, None], off[:-1], axis=0)
        assert np.abs(p1 - p0).max() / (mass.max() * 5000 * 80) < 1e-12
        e0, e1 = np.add.reduceat(energy(a), off[:-1]), np.add.reduceat(energy(b), off[:-1])
        if exact:
            assert np.abs(e1 / e0 - 1).max() < 1e-12
            assert np.array_equal(a.ERot, b.ERot)
        else:
            assert (a.ERot != b.ERot).sum() > 0 and np.abs(e1 / e0 - 1).max() > 1e-6


def test_ntc_candidate_count_and_remainder():
    sp = [H.argon()]
    mesh, md, o = box((4, 4, 4), (0.016,) * 3, sp, ppc=40)
    o.mesh_fill([0], [1e20], 300.0)
    sig, rem = o.download_cellstate()
    nC = np.diff(o.occupancy()).astype(float)
    o.stage(capi.STAGE_COLLIDE)
    sig1, rem1 = o.download_cellstate()
    V = (0.004) ** 3
    selected = rem + 0.5 * nC * (nC - 1) * md.nEquivalentParticles * sig * md.deltaT / V
    assert o.counters()["collisionCandidates"] == int(np.floor(selected).sum())
    assert np.allclose(rem1, selected - np.floor(selected), atol=1e-12)
    assert np.all(sig1 >= sig)


def test_equilibrium_collision_rate_within_one_percent_cpu():
    sp = [H.argon()]
    mesh, md, o = box((10, 10, 10), (0.04,) * 3, sp, ppc=40)
    o.mesh_fill([0], [1e20], 300.0)
    o.evolve(15)
    c0 = o.counters()["collisions"]
    steps = 40
    o.evolve(steps)
    total = o.counters()["collisions"] - c0
    expected = H.vhs_equilibrium_collision_rate(1e20, 300.0, sp[0]) * 0.04 ** 3 * md.deltaT * steps / md.nEquivalentParticles
    assert abs(total / expected - 1) < 0.01, (total, expected)


def test_heat_bath_relaxes_towards_equipartition():
    sp = H.air5()[:1]
    mesh, md, o = box((2, 2, 2), (0.008,) * 3, sp, "LarsenBorgnakkeVariableHardSphere", ppc=2000, dens=1e21, dt=2e-6,
                      rotationalRelaxationCollisionNumber=1.0)
    o.mesh_fill([0], [1e21], 3000.0, 300.0, 0.0)
    a = o.download_parcels()
    o.evolve(40)
    b = o.download_parcels()
    m = sp[0].mass
    Ttr = lambda p: m * (p.U ** 2).sum(1).mean() / (3 * H.KB)
    Trot = lambda p: p.ERot.mean() / H.KB
    assert Trot(a) < 350 and Ttr(a) > 2900
    assert abs(Ttr(b) - Trot(b)) < 0.08 * Ttr(b)          # rotation equilibrated with translation
    assert Ttr(b) < Ttr(a)


def test_diffuse_wall_reemits_at_wall_temperature_and_measures_fluxes():
    sides = {"xmin": ("cyclic",), "xmax": ("cyclic",), "ymin": ("wall", "cold"), "ymax": ("wall", "hot"), "zmin": ("cyclic",), "zmax": ("cyclic",)}
    mesh = meshgen.box_mesh((2, 6, 2), (0.01, 0.03, 0.01), sides=sides)
    sp = [H.argon()]
    pm = [dict(patch=mesh.patch_index("cold"), boundaryModel="dsmcDiffuseWallPatch", temperature=300.0),
          dict(patch=mesh.patch_index("hot"), boundaryModel="dsmcDiffuseWallPatch", temperature=300.0)]
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e20 * 3e-6 / (24 * 400), deltaT=5e-6, seed=3, patch_models=pm)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
    o.mesh_fill([0], [1e20], 300.0)
    n0 = o.num_parcels()
    o.evolve(30)
    assert o.num_parcels() == n0                 # walls never delete
    w = o.wall_accumulators()
    assert w.shape == (8, 1, 17)
    # equilibrium gas against an isothermal wall at the same temperature: no net heat flux within scatter, pressure = n k T
    q = w[:, 0, 13].sum() / 30 / 8
    fD = w[:, 0, 14:17] / 30
    p_wall = np.abs(fD[:, 1]).mean()
    assert abs(p_wall / (1e20 * H.KB * 300.0) - 1) < 0.05
    incident_energy_flux = 1e20 * np.sqrt(8 * H.KB * 300 / (np.pi * sp[0].mass)) / 4 * 2 * H.KB * 300
    assert abs(q) < 0.05 * incident_energy_flux
    b = o.download_parcels()
    assert b.position[:, 1].min() >= 0 and b.position[:, 1].max() <= 0.03


def test_free_stream_inflow_flux_matches_bird_4_22():
    sides = {"xmin": ("patch", "inlet"), "xmax": ("patch", "outlet"), "ymin": ("cyclic",), "ymax": ("cyclic",), "zmin": ("cyclic",), "zmax": ("cyclic",)}
    mesh = meshgen.box_mesh((8, 3, 3), (0.08, 0.03, 0.03), sides=sides)
    sp = [H.argon()]
    pm = [dict(patch=mesh.patch_index("inlet"), boundaryModel="dsmcDeletionPatch"), dict(patch=mesh.patch_index("outlet"), boundaryModel="dsmcDeletionPatch")]
    U = 500.0
    inflow = [dict(patch=mesh.patch_index("inlet"), typeIds=[0], numberDensities=[1e20], velocity=(U, 0, 0), translationalTemperature=300.0)]
    fnum = 1e20 * 0.08 * 0.03 * 0.03 / (72 * 200)
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=fnum, deltaT=2e-6, seed=9, patch_models=pm, inflows=inflow)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
    p = capi.ParcelData(0, 1)
    o.upload_parcels(p)
    steps = 200
    o.evolve(steps)
    from math import erf, exp, pi, sqrt
    cmp_ = sqrt(2 * H.KB * 300 / sp[0].mass)
    s = U / cmp_
    flux = 1e20 * cmp_ * (exp(-s * s) + sqrt(pi) * s * (1 + erf(s))) / (2 * sqrt(pi))      # Bird eq. 4.22
    expected = flux * 0.03 * 0.03 * md.deltaT * steps / fnum
    ins = o.counters()["inserted"]
    assert abs(ins / expected - 1) < 3.5 / np.sqrt(expected) + 1e-3
    q = o.download_parcels()
    assert q.n > 0 and abs(q.U[:, 0].mean() / U - 1) < 0.1


def test_cyclic_box_keeps_parcels_and_cells_consistent():
    sp = [H.argon()]
    mesh, md, o = box((5, 4, 3), (0.02, 0.016, 0.012), sp, "NoBinaryCollision", ppc=60)
    o.mesh_fill([0], [1e20], 300.0, velocity=(300.0, -200.0, 100.0))
    a = H.by_id(o.download_parcels())
    o.evolve(6)
    b = H.by_id(o.download_parcels())
    assert len(b["origId"]) == len(a["origId"])
    L = np.array([0.02, 0.016, 0.012])
    free = a["position"] + 6 * md.deltaT * a["U"]
    assert np.allclose(b["position"], np.mod(free, L), atol=1e-12)      # ballistic flight with periodic wrap
    ijk = np.floor(b["position"] / 0.004).astype(int)
    assert np.array_equal(ijk[:, 0] + 5 * (ijk[:, 1] + 4 * ijk[:, 2]), b["cell"])
    assert np.array_equal(a["U"], b["U"])


def test_renumbered_mesh_tracks_like_the_original():
    """meshgen.renumber_cells (a renumberMesh equivalent: cells relabelled along a z-order curve, internal faces flipped / re-sorted
    into upper-triangular order) yields a valid polyMesh: ballistic flight with periodic wrap ends in the relabelled cell of the
    direct point location."""
    sp = [H.argon()]
    base = meshgen.box_mesh((6, 5, 4), (0.024, 0.02, 0.016))
    new_of_old = meshgen.morton_order(base)
    mesh, _ = meshgen.renumber_cells(base, new_of_old)
    assert sorted(new_of_old) == list(range(120)) and not np.array_equal(new_of_old, np.arange(120))
    assert np.all(mesh.owner[: mesh.n_internal] < mesh.neighbour)
    key = mesh.owner[: mesh.n_internal].astype(np.int64) * mesh.n_cells + mesh.neighbour
    assert np.all(np.diff(key) > 0)                                     # upper-triangular face order
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e20 * 0.024 * 0.02 * 0.016 / (120 * 50), deltaT=5e-6, seed=42)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
    o.mesh_fill([0], [1e20], 300.0, velocity=(300.0, -200.0, 100.0))
    a = H.by_id(o.download_parcels())
    ijk = np.floor(a["position"] / 0.004).astype(int)
    assert np.array_equal(new_of_old[ijk[:, 0] + 6 * (ijk[:, 1] + 5 * ijk[:, 2])], a["cell"])
    o.evolve(6)
    b = H.by_id(o.download_parcels())
    assert len(b["origId"]) == len(a["origId"])
    free = a["position"] + 6 * md.deltaT * a["U"]
    assert np.allclose(b["position"], np.mod(free, [0.024, 0.02, 0.016]), atol=1e-12)
    ijk = np.floor(b["position"] / 0.004).astype(int)
    assert np.array_equal(new_of_old[ijk[:, 0] + 6 * (ijk[:, 1] + 5 * ijk[:, 2])], b["cell"])


@pytest.mark.parametrize("kind", ["prism", "tet"])
def test_tracking_on_non_hex_cells(kind):
    """Triangular prisms / tetrahedra (triangular faces: one tet per face triangle, base points from polyMeshTetDecomposition) inside a
    specular box: after 8 steps every parcel sits in the cell that direct point location gives, none is lost, |U| is unchanged."""
    mesh, locate = meshgen.split_box_mesh((3, 3, 2), (0.012, 0.012, 0.008), kind)
    assert set(np.diff(mesh.face_offsets)) == ({3, 4} if kind == "prism" else {3})
    sp = [H.argon()]
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e20 * 0.012 * 0.012 * 0.008 / (mesh.n_cells * 40), deltaT=5e-6, seed=3,
                           patch_models=[dict(patch=0, boundaryModel="dsmcSpecularWallPatch")])
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
    o.mesh_fill([0], [1e20], 300.0)
    a = H.by_id(o.download_parcels())
    assert np.array_equal(locate(a["position"]), a["cell"])
    o.evolve(8)
    b = H.by_id(o.download_parcels())
    assert len(b["cell"]) == len(a["cell"]) > 1000
    assert np.array_equal(locate(b["position"]), b["cell"])
    assert np.allclose((a["U"] ** 2).sum(1), (b["U"] ** 2).sum(1), rtol=1e-12)
    assert (a["cell"] != b["cell"]).mean() > 0.5


def test_tracking_across_a_refinement_interface_with_five_point_faces():
    """A 2:1 refinement interface: the coarse cell has nine faces, four of them pentagons with a hanging node on a straight edge (one
    fan triangle of the face is degenerate for most base points: polyMeshTetDecomposition::findBasePoint).  Cell volumes are exact,
    every parcel stays locatable through 12 steps in a specular box."""
    mesh, locate = meshgen.refined_interface_mesh()
    assert sorted(set(np.diff(mesh.face_offsets))) == [4, 5] and mesh.n_cells == 6
    sp = [H.argon()]
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e20 * 3 * 0.004 ** 3 / (6 * 400), deltaT=2e-6, seed=3,
                           patch_models=[dict(patch=0, boundaryModel="dsmcSpecularWallPatch")])
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
    _, cv, *_ = o.geometry()
    assert np.allclose(np.asarray(cv) / 0.004 ** 3, [1, 1, 0.25, 0.25, 0.25, 0.25], rtol=1e-12)
    o.mesh_fill([0], [1e20], 300.0)
    a = H.by_id(o.download_parcels())
    assert np.array_equal(locate(a["position"]), a["cell"])
    o.evolve(12)
    b = H.by_id(o.download_parcels())
    assert len(b["cell"]) == len(a["cell"]) > 2000
    assert np.array_equal(locate(b["position"]), b["cell"])
    assert (a["cell"] != b["cell"]).mean() > 0.5 and o.counters()["trackingRescues"] == 0


def test_diffuse_specular_wall_splits_by_diffuse_fraction():
    """dsmcDiffuseSpecularWallPatch (Maxwell model, patchBoundaries/mixed/dsmcDiffuseSpecularWallPatch/dsmcDiffuseSpecularWallPatch.C:97-115):
    a wall hit is diffuse with probability diffuseFraction, specular otherwise.  A specular hit keeps |U|; a diffuse one resamples it."""
    sp = [H.argon()]
    sides = {s: ("wall", "walls") for s in meshgen.SIDES}
    frac = 0.3
    mesh, md, o = box((4, 4, 4), (0.016,) * 3, sp, "NoBinaryCollision", ppc=400, sides=sides, dt=2e-6,
                      patch_models=[dict(patch=0, boundaryModel="dsmcDiffuseSpecularWallPatch", temperature=900.0, velocity=(0, 0, 0),
                                         diffuseFraction=frac)])
    o.mesh_fill([0], [1e20], 300.0)
    a = H.by_id(o.download_parcels())
    o.evolve(1)
    b = H.by_id(o.download_parcels())
    hit = (a["U"] != b["U"]).any(1)
    assert hit.sum() > 1500
    keeps_speed = np.abs((b["U"][hit] ** 2).sum(1) / (a["U"][hit] ** 2).sum(1) - 1) < 1e-12
    n = hit.sum()
    # corner double hits are rare at this time step: the specular share is 1 - diffuseFraction within 4 sigma (+1 % slack)
    assert abs(keeps_speed.mean() - (1 - frac)) < 4 * np.sqrt(frac * (1 - frac) / n) + 0.01
    # the diffuse ones come off at the wall temperature: mean kinetic energy of a half-range Maxwellian flux = 2 k T_w
    ek = 0.5 * sp[0].mass * (b["U"][hit][~keeps_speed] ** 2).sum(1)
    assert abs(ek.mean() / (2 * H.KB * 900.0) - 1) < 0.1


def test_sample_interval_skips_cell_and_wall_measurements():
    """dsmcVolFields::calculateField samples only when sampleInterval_ <= ++sampleCounter_ (dsmcVolFields.C:1073-1081,1362);
    boundaryMeas_ / cellMeas_ are cleaned every step (dsmcCloud.C:923-925), so un-sampled steps leave no trace.  The same seeded
    run with sampleInterval 1 sees identical parcels (sampling draws nothing), hence interval 3 == steps 3 and 6 of it."""
    sides = {"xmin": ("cyclic",), "xmax": ("cyclic",), "ymin": ("wall", "cold"), "ymax": ("wall", "hot"), "zmin": ("cyclic",), "zmax": ("cyclic",)}
    mesh = meshgen.box_mesh((2, 6, 2), (0.01, 0.03, 0.01), sides=sides)
    sp = [H.argon()]
    pm = [dict(patch=mesh.patch_index("cold"), boundaryModel="dsmcDiffuseWallPatch", temperature=300.0),
          dict(patch=mesh.patch_index("hot"), boundaryModel="dsmcDiffuseWallPatch", temperature=500.0)]

    def run(interval, stops):
        md = capi.build_models("VariableHardSphere", nEquivalentParticles=1e20 * 3e-6 / (24 * 100), deltaT=5e-6, seed=3, patch_models=pm,
                               sampleInterval=interval)
        o = Oracle()
        o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
        o.mesh_fill([0], [1e20], 300.0)
        out, done = [], 0
        for s in stops:
            o.evolve(s - done)
            done = s
            out.append((o.accumulators(), o.wall_accumulators().copy()))
        return out

    every = run(1, [2, 3, 5, 6, 7])
    third = run(3, [7])[0]
    (a3, c3, n3), w3 = third
    assert n3 == 2
    (a_2, c_2, _), w_2 = every[0]
    (a_3, c_3, _), w_3 = every[1]
    (a_5, c_5, _), w_5 = every[2]
    (a_6, c_6, _), w_6 = every[3]
    assert np.allclose(a3, (a_3 - a_2) + (a_6 - a_5), rtol=1e-12, atol=0)
    assert np.allclose(c3, (c_3 - c_2) + (c_6 - c_5), rtol=1e-12, atol=0)
    assert np.abs(w3).sum() > 0
    scale = np.abs(w3).max(axis=(0, 1), keepdims=True) + 1e-300
    assert (np.abs(w3 - ((w_3 - w_2) + (w_6 - w_5))) / scale).max() < 1e-9


def test_face_tracker_fluxes_balance_the_cell_occupancy():
    """dsmcFaceTracker::trackFaceTransition (DSMC/faceTracker/dsmcFaceTracker.C:124-198): per step and species, the signed parcel
    flux through the internal faces of a cell (S_f points owner -> neighbour) is its change of occupancy; a specular wall hit is
    tracked with the reflected velocity (-1 on the wall face) and massIdFlux = mass x parcelIdFlux."""
    sides = {k: ("wall", "walls") for k in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")}
    mesh = meshgen.box_mesh((4, 3, 5), (0.02, 0.015, 0.025), sides=sides)
    sp = H.air5()[:2]
    pm = [dict(patch=mesh.patch_index("walls"), boundaryModel="dsmcSpecularWallPatch")]
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=1e20 * 0.02 * 0.015 * 0.025 / (60 * 40), deltaT=5e-6, seed=9,
                           patch_models=pm, trackFaceFluxes=True)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
    o.mesh_fill([0, 1], [0.7e20, 0.3e20], 400.0, 400.0, 400.0)
    nI = mesh.n_internal
    owner, neigh = np.asarray(mesh.owner), np.asarray(mesh.neighbour)
    for _ in range(3):
        before = o.download_parcels()
        o.evolve(1)
        after = o.download_parcels()
        pf, mf = o.face_fluxes()
        assert np.abs(pf[:, :nI]).sum() > 100
        for s in range(2):
            dn = np.bincount(after.cell[after.typeId == s], minlength=mesh.n_cells) - np.bincount(before.cell[before.typeId == s], minlength=mesh.n_cells)
            net = np.zeros(mesh.n_cells)
            np.add.at(net, owner[:nI], -pf[s, :nI])
            np.add.at(net, neigh[:nI], pf[s, :nI])
            assert np.array_equal(net, dn)
            assert np.allclose(mf[s], sp[s].mass * pf[s], rtol=1e-12, atol=1e-12 * sp[s].mass)   # +m -m leaves rounding residue
        assert (pf[:, nI:] <= 0).all() and pf[:, nI:].sum() < -10     # every wall hit leaves with U . S_f < 0


def test_inverse_zv_formulation_2008_uses_the_macroscopic_temperature():
    """inverseZvFormulation "2008" (dsmcCloud.C:1441-1456): Zv is evaluated at fields().overallT(cell); while that is <= SMALL the
    quantised collision temperature is used, i.e. the run equals "pre-2008".  Zv(T = theta_d) = 1 (every eligible collision
    exchanges vibrational energy), Zv(T -> 0) -> infinity (none does)."""
    sp = H.air5()[:1]   # N2: theta_v 3371 K, theta_d 113500 K

    def run(formulation, Tov=None):
        mesh, md, o = box((4, 4, 4), (0.016,) * 3, sp, model="LarsenBorgnakkeVariableHardSphere", ppc=60, dens=1e21, dt=2e-6,
                          inverseZvFormulation=formulation)
        o.mesh_fill([0], [1e21], 6000.0, 6000.0, 6000.0)
        if Tov is not None:
            o.upload_overall_temperature(np.full(mesh.n_cells, Tov))
        before = o.download_parcels().vibLevel.copy()
        o.stage(capi.STAGE_COLLIDE)
        after = o.download_parcels()
        return int((after.vibLevel != before).sum()), after, o.counters()["collisions"]

    n_pre, a_pre, c_pre = run("pre-2008")
    n_fb, a_fb, c_fb = run("2008")                      # no Tov yet: the reference's fallback
    assert c_pre == c_fb > 500 and n_pre == n_fb > 0
    assert np.array_equal(a_pre.vibLevel, a_fb.vibLevel) and np.array_equal(a_pre.U, a_fb.U)
    n_cold, _, c_cold = run("2008", Tov=1.0)
    assert c_cold > 500 and n_cold == 0                 # 1/Zv(1 K) = 0: no vibrational exchange at all
    n_hot, _, _ = run("2008", Tov=113500.0)
    assert n_hot > 5 * n_pre                            # 1/Zv(theta_d) = 1


def test_diffuse_wall_linear_temperature_along_the_depth_axis():
    """dsmcDiffuseWallPatch::getLocalTemperature (dsmcDiffuseWallPatch.C:141-148): T(y) = groundLevelTemperature + (y - y_max) *
    (groundLevelTemperature - formationLevelTemperature) / (y_max - y_min) with the bounds of the mesh; a diffusely re-emitted parcel
    carries <m U^2 / 2> = 2 k T(y_hit)."""
    sides = {"xmin": ("wall", "walls"), "xmax": ("wall", "walls"), "ymin": ("symmetryPlane", "ends"), "ymax": ("symmetryPlane", "ends"),
             "zmin": ("cyclic",), "zmax": ("cyclic",)}
    mesh = meshgen.box_mesh((2, 10, 1), (0.002, 0.1, 0.002), sides=sides)
    sp = [H.argon()]
    Tg, Tf = 1000.0, 200.0
    pm = [dict(patch=mesh.patch_index("walls"), boundaryModel="dsmcDiffuseWallPatch", temperature=Tg, formationLevelTemperature=Tf, depthAxis="y")]
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e20 * 0.002 * 0.1 * 0.002 / (20 * 3000), deltaT=2e-6, seed=21, patch_models=pm)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
    o.mesh_fill([0], [1e20], 300.0)
    before = H.by_id(o.download_parcels())
    o.evolve(1)
    after = H.by_id(o.download_parcels())
    e0, e1 = (before["U"] ** 2).sum(1), (after["U"] ** 2).sum(1)
    hit = np.abs(e1 / e0 - 1) > 1e-9                    # a symmetry-plane reflection keeps |U|, a diffuse wall does not
    assert hit.sum() > 8000
    y = after["position"][hit, 1]                       # within |U_y| dt ~ 1 mm of the hit position
    T_est = sp[0].mass * (after["U"][hit] ** 2).sum(1) / (4 * H.KB)
    for lo in (0.0, 0.02, 0.04, 0.06, 0.08):
        sel = (y >= lo) & (y < lo + 0.02)
        T_expected = Tg + ((lo + 0.01) - 0.1) * (Tg - Tf) / 0.1
        assert sel.sum() > 1000
        assert abs(T_est[sel].mean() / T_expected - 1) < 0.06, (lo, T_est[sel].mean(), T_expected)


def _weighted_box():
    """4 x 6 x 1 box whose rows of cells carry radial weights growing with y, as an axisymmetric mesh's do"""
    sides = {"xmin": ("cyclic",), "xmax": ("cyclic",), "ymin": ("symmetryPlane", "axis"), "ymax": ("wall", "top"),
             "zmin": ("symmetryPlane", "front"), "zmax": ("symmetryPlane", "back")}
    mesh = meshgen.box_mesh((4, 6, 1), (0.04, 0.06, 0.01), sides=sides)
    rev, pol, ang = capi.axisymmetric_axes()
    md = capi.build_models("VariableHardSphere", nEquivalentParticles=1e9, deltaT=2e-6, seed=5, coordinateSystem="dsmcAxisymmetric",
                           angularCoordinate=ang, patch_models=[dict(patch=mesh.patch_index("top"), boundaryModel="dsmcSpecularWallPatch")])
    return mesh, md, pol, ang


def test_axisymmetric_axes_follow_the_dictionary_keywords():
    """dsmcAxisymmetric::checkCoordinateSystemInputs (dsmcAxisymmetric.C:337-420)"""
    assert capi.axisymmetric_axes() == (0, 1, 2)
    assert capi.axisymmetric_axes("z") == (2, 0, 1) and capi.axisymmetric_axes("z", "y") == (2, 1, 0)
    assert capi.axisymmetric_axes("y") == (1, 2, 0) and capi.axisymmetric_axes("y", "x") == (1, 0, 2)
    assert capi.axisymmetric_axes("x", "z") == (0, 2, 1)
    with pytest.raises(capi.Dsmcb200Error, match="badly defined"):
        capi.axisymmetric_axes("z", "z")


def test_radial_weighting_clones_and_deletes_with_the_weight_ratio():
    """dsmcAxisymmetric::axisymmetricWeighting (dsmcAxisymmetric.C:50-209): a parcel that arrives in a cell with a smaller RWF is cloned
    floor(old/new - 1) times plus once with the remaining probability (the clone mirrors the angular velocity component, keeps
    everything else and takes the new weight); in a cell with a larger RWF it survives with probability old/new.  So the number of
    molecules a cell's parcels stand for, sum of RWF, is conserved in expectation."""
    mesh, md, pol, ang = _weighted_box()
    o = Oracle()
    o.set_mesh(mesh); o.set_species([H.argon()]); o.set_models(md)
    cc, cv, fc, *_ = o.geometry()
    rwf, ext = capi.axisymmetric_rwf(cc, fc, pol, 100.0)
    assert abs(ext - 0.06) < 1e-15 and abs(rwf[0] - (1 + 99 * 0.005 / 0.06)) < 1e-12
    o.set_cell_fields(RWF=rwf)
    o.mesh_fill([0], [4e18], 300.0)     # n V / (F_N RWF) parcels per cell: fewer, heavier parcels away from the axis
    p = o.download_parcels()
    assert np.array_equal(p.radialWeight, rwf[p.cell])
    n_cell = np.bincount(p.cell, minlength=mesh.n_cells)
    expect = 4e18 * cv / (1e9 * rwf)
    assert np.abs(n_cell - expect).max() < 5 * np.sqrt(expect.max())
    # give every parcel the weight 20: cells with RWF < 20 clone, cells with RWF > 20 delete
    p.radialWeight[:] = 20.0
    o.upload_parcels(p)
    o.stage(capi.STAGE_SORT)
    q = o.download_parcels()
    cloned, deleted = o.weighting_counts()
    assert q.n == p.n + cloned - deleted and cloned > 0 and deleted > 0
    assert np.array_equal(q.radialWeight, rwf[q.cell])
    new = q.origId >= p.n
    assert new.sum() == cloned
    # every clone has a parent with the same position and the mirrored angular velocity component
    parent = {tuple(x): k for k, x in enumerate(p.position)}
    for k in np.nonzero(new)[0][:200]:
        j = parent[tuple(q.position[k])]
        u = p.U[j].copy(); u[ang] *= -1.0
        assert np.array_equal(q.U[k], u) and q.cell[k] == p.cell[j]
    # per row of cells: sum of weights before and after agree within the binomial scatter
    row = (np.arange(mesh.n_cells) // 4) % 6
    for r in range(6):
        w0 = 20.0 * np.isin(p.cell, np.nonzero(row == r)[0]).sum()
        w1 = q.radialWeight[np.isin(q.cell, np.nonzero(row == r)[0])].sum()
        n_r = np.isin(p.cell, np.nonzero(row == r)[0]).sum()
        assert abs(w1 - w0) < 5 * max(rwf[row == r][0], 20.0) * np.sqrt(n_r), (r, w0, w1)
    # ratio 20 / RWF > 2 in the first row: at least one clone per parcel there
    first = np.nonzero(row == 0)[0]
    assert 20.0 / rwf[first[0]] - 1 > 1.0
    assert np.isin(q.cell, first).sum() >= 2 * np.isin(p.cell, first).sum()


def _spherical_box():
    """4 x 4 x 4 box around the origin of a spherical coordinate system in its corner: weights grow with the square of the radius"""
    sides = {s: ("wall", "walls") for s in meshgen.SIDES}
    mesh = meshgen.box_mesh((4, 4, 4), (0.04, 0.04, 0.04), sides=sides)
    md = capi.build_models("VariableHardSphere", nEquivalentParticles=1e9, deltaT=2e-6, seed=5, coordinateSystem="dsmcSpherical",
                           patch_models=[dict(patch=mesh.patch_index("walls"), boundaryModel="dsmcSpecularWallPatch")])
    return mesh, md


def test_spherical_weighting_clones_without_mirroring():
    """dsmcSpherical (spherical/dsmcSpherical.C:50-275): RWF = 1 + (maxRWF - 1) (r / radialExtent)^2 about the origin, the same clone / delete
    rule as the axisymmetric system, and a clone that keeps its parent's velocity."""
    mesh, md = _spherical_box()
    o = Oracle()
    o.set_mesh(mesh); o.set_species([H.argon()]); o.set_models(md)
    cc, cv, fc, *_ = o.geometry()
    rwf, ext = capi.spherical_rwf(cc, fc, (0, 0, 0), 100.0)
    assert abs(ext - np.sqrt(0.04 ** 2 + 0.035 ** 2 + 0.035 ** 2)) < 1e-15    # the farthest face centre: an outer face of the corner cell
    assert abs(rwf[0] - (1 + 99 * (3 * 0.005 ** 2) / ext ** 2)) < 1e-12 and rwf.max() < 100
    o.set_cell_fields(RWF=rwf)
    o.mesh_fill([0], [4e18], 300.0)
    p = o.download_parcels()
    assert np.array_equal(p.radialWeight, rwf[p.cell])
    p.radialWeight[:] = 20.0
    o.upload_parcels(p)
    o.stage(capi.STAGE_SORT)
    q = o.download_parcels()
    cloned, deleted = o.weighting_counts()
    assert q.n == p.n + cloned - deleted and cloned > 50 and deleted > 50
    assert np.array_equal(q.radialWeight, rwf[q.cell])
    new = np.nonzero(q.origId >= p.n)[0]
    parent = {tuple(x): k for k, x in enumerate(p.position)}
    for k in new[:200]:
        j = parent[tuple(q.position[k])]
        assert np.array_equal(q.U[k], p.U[j]) and q.cell[k] == p.cell[j]          # not mirrored
    # the same parcels of a 2-D case are NOT pulled to the mesh centre under a non-Cartesian system (dsmcParcel.C:76)
    o.evolve(2)
    assert o.num_parcels() > 0


def test_variable_time_step_scales_weights_and_steps_with_the_cell_volume():
    """dsmcVariableTimeStepModel (dsmcVariableTimeStepModel.C:48-100): nParticles and deltaT of a cell grow with its volume, their ratio is
    uniform.  A parcel moves by U deltaT(its cell) (dsmcParcel.C:62-63) and the candidate pairs of a cell use its own weight and step
    (noTimeCounter.C:101,148); dsmcMeshFill puts n V / nParticles(cell) parcels into a cell: the same number in every cell."""
    sides = {s: ("cyclic",) for s in meshgen.SIDES}
    mesh = meshgen.box_mesh((6, 2, 2), (0.06, 0.02, 0.02), sides=sides)
    x = mesh.points[:, 0].copy()
    mesh.points[:, 0] = 0.06 * (x / 0.06) ** 2     # cells grow along x
    md = capi.build_models("VariableHardSphere", nEquivalentParticles=2e8, deltaT=1e-7, seed=3)
    o = Oracle()
    o.set_mesh(mesh); o.set_species([H.argon()]); o.set_models(md)
    cc, cv, *_ = o.geometry()
    n, dt = capi.variable_time_step(cv, 2e8, 1e-7)
    assert np.allclose(n / dt, 2e8 / 1e-7, rtol=1e-14) and np.isclose(n.min(), 2e8, rtol=1e-14) and np.isclose(n.max() / n.min(), cv.max() / cv.min())
    o.set_cell_fields(nParticles=n, deltaT=dt)
    o.mesh_fill([0], [1e20], 300.0)
    p = o.download_parcels()
    per_cell = np.bincount(p.cell, minlength=mesh.n_cells)
    expect = 1e20 * cv.min() / 2e8
    assert np.abs(per_cell - expect).max() < 5 * np.sqrt(expect)          # the same number of parcels in every cell
    o.stage(capi.STAGE_MOVE)
    q = o.download_parcels()
    o2 = np.argsort(q.origId)
    moved = q.position[o2] - p.position
    moved[:, 0] -= 0.06 * np.round(moved[:, 0] / 0.06); moved[:, 1] -= 0.02 * np.round(moved[:, 1] / 0.02); moved[:, 2] -= 0.02 * np.round(moved[:, 2] / 0.02)
    assert np.allclose(moved, p.U * dt[p.cell][:, None], rtol=1e-9, atol=1e-12)
    # candidate pairs: 0.5 N (N - 1) nParticles sigmaTcRMax deltaT / V per cell
    o.upload_parcels(p)
    sig, rem = o.download_cellstate()
    o.upload_cellstate(sig, np.zeros_like(rem))
    o.stage(capi.STAGE_COLLIDE)
    N = per_cell.astype(float)
    cand = np.floor(0.5 * N * (N - 1) * n * sig * dt / cv)
    assert o.counters()["collisionCandidates"] == int(cand.sum())


def test_capsule_mesh_bricks_tile_the_single_domain_mesh():
    """meshgen.capsule_mesh (BASELINE configs[3]): the eight bricks of a 2 x 2 x 2 decomposition have the patch list of the whole mesh
    (empty where a brick does not touch a boundary), their cells tile the single-domain volume, the heat shield's area is the sum of
    the bricks' shares, and processor patches pair up face by face (same centres on both sides)."""
    whole = meshgen.capsule_mesh((12, 10, 10), 0.15)
    o = Oracle(); o.set_mesh(whole)
    _, cv, fc, fa, *_ = o.geometry()
    cap = whole.patches[whole.patch_index("capsule")]
    area = np.linalg.norm(fa[cap["start"]:cap["start"] + cap["size"]], axis=1).sum()
    h = whole.capsule["height"]
    # volume removed by the body = volume of the spherical segment (to the accuracy of the faceted surface)
    seg = np.pi * h * h * (3 * whole.capsule["Rs"] - h) / 3.0
    assert abs((1.8 * 1.5 * 1.5 - cv.sum()) / seg - 1) < 0.15
    vol, a_sum, proc_faces = 0.0, 0.0, {}
    for r in range(8):
        m = meshgen.capsule_mesh((6, 5, 5), 0.15, (2, 2, 2), r)
        assert [p["name"] for p in m.patches[:3]] == ["flow", "capsule", "outflow"]
        b = Oracle(); b.set_mesh(m)
        _, v, c, a, *_ = b.geometry()
        vol += v.sum()
        p = m.patches[1]
        a_sum += np.linalg.norm(a[p["start"]:p["start"] + p["size"]], axis=1).sum()
        for q in m.patches[3:]:
            assert q["type"] == "processor"
            proc_faces[(r, q["neighbProcNo"])] = c[q["start"]:q["start"] + q["size"]]
    assert abs(vol / cv.sum() - 1) < 1e-12 and abs(a_sum / area - 1) < 1e-12
    for (r, nb), c in proc_faces.items():
        assert np.allclose(c, proc_faces[(nb, r)], atol=1e-12)      # the two sides list the shared faces in the same order


def _cll_box(aN, sigT, aR, Tw=900.0, species=None, Trot=0.0, wall_velocity=(0, 0, 0)):
    sides = {"xmin": ("cyclic",), "xmax": ("cyclic",), "ymin": ("wall", "walls"), "ymax": ("wall", "walls"), "zmin": ("cyclic",), "zmax": ("cyclic",)}
    sp = species or [H.argon()]
    wall = meshgen.box_mesh((4, 4, 4), (0.016,) * 3, sides=sides).patch_index("walls")
    pm = [dict(patch=wall, boundaryModel="dsmcCLLWallPatch", temperature=Tw, velocity=wall_velocity, normalAccommodationCoefficient=aN,
               tangentialAccommodationCoefficient=sigT, rotationalEnergyAccommodationCoefficient=aR)]
    mesh, md, o = box((4, 4, 4), (0.016,) * 3, sp, "NoBinaryCollision", ppc=1500, sides=sides, dt=5e-6, patch_models=pm)
    o.mesh_fill([0], [1e20], 300.0, Trot)
    a = H.by_id(o.download_parcels())
    o.evolve(1)
    b = H.by_id(o.download_parcels())
    hit = (a["U"] != b["U"]).any(1)
    assert hit.sum() > 4000
    return sp, a, b, hit, o


def test_cll_wall_with_zero_coefficients_is_specular_and_measures_nothing():
    """dsmcCLLWallPatch.C:82-89: alphaN = alphaT = 0 reduces the kernel to a specular wall and switches the wall measurements off."""
    sp, a, b, hit, o = _cll_box(0.0, 0.0, 0.0)
    ua, ub = a["U"][hit], b["U"][hit]
    assert np.allclose(ub[:, 1], -ua[:, 1], rtol=1e-12)
    assert np.allclose(ub[:, [0, 2]], ua[:, [0, 2]], rtol=1e-10, atol=1e-9)
    assert np.abs(o.wall_accumulators()).sum() == 0


def test_cll_wall_moments_follow_the_accommodation_coefficients():
    """dsmcCLLWallPatch::controlParticle (dsmcCLLWallPatch.C:100-300).  In units of the wall's most probable speed the reflected normal
    component has <vn'^2> = alphaN + (1 - alphaN) vn^2, and the component along the incident tangential direction has
    <vt1'> = sqrt(1 - alphaT) |vt| with alphaT = sigma (2 - sigma), i.e. (1 - sigma) |vt|: the definitions of the two coefficients."""
    aN, sigT, Tw = 0.6, 0.35, 900.0
    sp, a, b, hit, o = _cll_box(aN, sigT, 1.0, Tw)
    vmp = np.sqrt(2 * H.KB * Tw / sp[0].mass)
    ua, ub = a["U"][hit] / vmp, b["U"][hit] / vmp
    n = hit.sum()
    resid = ub[:, 1] ** 2 - (1 - aN) * ua[:, 1] ** 2
    assert abs(resid.mean() - aN) < 4 * resid.std() / np.sqrt(n)
    ta = ua[:, [0, 2]]
    t1 = ta / np.linalg.norm(ta, axis=1, keepdims=True)
    vt1 = (ub[:, [0, 2]] * t1).sum(1)
    resid = vt1 - (1 - sigT) * np.linalg.norm(ta, axis=1)
    assert abs(resid.mean()) < 4 * resid.std() / np.sqrt(n)
    # the component normal to both has no memory of the incident one: variance alphaT / 2
    vt2 = ub[:, 0] * t1[:, 1] - ub[:, 2] * t1[:, 0]
    alphaT = sigT * (2 - sigT)
    assert abs(vt2.mean()) < 4 * vt2.std() / np.sqrt(n) and abs(vt2.var() / (alphaT / 2) - 1) < 0.06
    # reflected parcels leave the wall
    y = b["position"][hit, 1]
    assert np.all(np.where(y < 0.008, ub[:, 1] > 0, ub[:, 1] < 0))
    assert np.abs(o.wall_accumulators()).sum() > 0


def test_cll_wall_full_accommodation_is_diffuse_and_lord_rotation():
    """alphaN = sigma_t = 1 is the diffuse wall at T_w (flux mean kinetic energy 2 k T_w, plus the wall's velocity tangentially); Lord's
    rotational extension for a diatomic: <ERot'> = alphaR k T_w + (1 - alphaR) ERot (dsmcCLLWallPatch.C:247-253)."""
    Tw, aR = 900.0, 0.4
    n2 = H.air5()[:1]
    sp, a, b, hit, o = _cll_box(1.0, 1.0, aR, Tw, species=n2, Trot=300.0, wall_velocity=(150.0, 0, 0))
    ub = b["U"][hit] - np.array([150.0, 0, 0])
    ek = 0.5 * sp[0].mass * (ub ** 2).sum(1)
    assert abs(ek.mean() / (2 * H.KB * Tw) - 1) < 4 * ek.std() / ek.mean() / np.sqrt(hit.sum())
    assert abs(ub[:, 0].mean()) < 4 * ub[:, 0].std() / np.sqrt(hit.sum())
    resid = b["ERot"][hit] - (1 - aR) * a["ERot"][hit]
    assert abs(resid.mean() / (aR * H.KB * Tw) - 1) < 4 * resid.std() / resid.mean() / np.sqrt(hit.sum())
    assert np.array_equal(a["vibLevel"], b["vibLevel"]) and np.array_equal(a["ELevel"], b["ELevel"])   # untouched (commented out upstream)


def test_zone_fill_is_the_mesh_fill_of_the_zone_cells():
    """dsmcZoneFill (initialiseDsmcParcels/derived/dsmcZoneFill/dsmcZoneFill.C:71-272): one zone holding every cell in ascending order gives the
    cloud of dsmcMeshFill; two zones give each its own state, only their own cells, and sigmaTcRMax of their own most abundant species."""
    sp = H.air5()[:2]
    mesh, md, o = box((6, 4, 4), (0.024, 0.016, 0.016), sp, "LarsenBorgnakkeVariableHardSphere", ppc=40, inverseZvFormulation="pre-2008")
    o.mesh_fill([0, 1], [0.8e20, 0.2e20], 500.0, 500.0, 500.0)
    a = o.download_parcels()
    o.upload_parcels(capi.ParcelData(0, 1))
    o.zone_fill(np.arange(mesh.n_cells), [0, 1], [0.8e20, 0.2e20], 500.0, 500.0, 500.0)
    b = o.download_parcels()
    assert a.n == b.n > 3000
    for k in ("position", "U", "ERot", "cell", "tetFace", "tetPt", "typeId", "vibLevel", "origId"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k

    x = np.arange(mesh.n_cells) % 6
    left, right = np.flatnonzero(x < 2), np.flatnonzero(x >= 2)[::-1]
    o.upload_parcels(capi.ParcelData(0, 1))
    o.zone_fill(left, [0, 1], [2.4e20, 0.6e20], 3000.0, 3000.0, 3000.0, velocity=(400.0, 0, 0))
    n_left = o.num_parcels()
    o.zone_fill(right, [1], [0.5e20], 300.0, 300.0, 300.0)
    n_all = o.num_parcels()
    c = H.by_id(o.download_parcels())                                     # the cloud is kept in cell order; origId is the insertion order
    assert np.array_equal(c["origId"], np.arange(n_all))                  # appended, ids continue
    assert np.all(np.isin(c["cell"][:n_left], left)) and np.all(np.isin(c["cell"][n_left:], right)) and np.all(c["typeId"][n_left:] == 1)
    assert c["cell"][n_left] == right[0] == mesh.n_cells - 1              # the zone's order, not the mesh's
    per_cell = 40 / 1e20
    assert abs(n_left / (len(left) * 3.0e20 * per_cell) - 1) < 0.03 and abs((n_all - n_left) / (len(right) * 0.5e20 * per_cell) - 1) < 0.03
    U = c["U"]
    assert abs(U[:n_left, 0].mean() - 400.0) < 25.0 and abs(U[n_left:, 0].mean()) < 10.0
    ek = lambda U, m: 0.5 * m * ((U - U.mean(0)) ** 2).sum(1).mean() / (1.5 * H.KB)
    assert abs(ek(U[n_left:], sp[1].mass) / 300.0 - 1) < 0.05
    s, _ = o.download_cellstate()
    vmp = lambda T, m: np.sqrt(2 * H.KB * T / m)
    assert np.allclose(s[left], np.pi * sp[0].diameter ** 2 * vmp(3000.0, sp[0].mass), rtol=1e-12)
    assert np.allclose(s[right], np.pi * sp[1].diameter ** 2 * vmp(300.0, sp[1].mass), rtol=1e-12)
