"""GPU tests of the coordinate-system options next to the path: dsmcAxisymmetric radial weighting
(DSMC/coordinateSystem/derived/axisymmetric/dsmcAxisymmetric.C) and dsmcVariableTimeStepModel
(DSMC/coordinateSystem/timeStepModel/derived/variableTimeStepModel), through the per-cell fields of the C ABI -- against the CPU oracle
on the same seeded input and against the fields the reference ships for its axisymmetric tutorial."""
import numpy as np
import pytest

from hystrath_b200 import capi, meshgen
from oracle import fields_ref
from oracle.pyoracle import Oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_radial_weighting_stage_matches_oracle_parcel_by_parcel():
    """axisymmetricWeighting on a cloud whose parcels all carry the weight 20 in cells with weights from 9 to 91: which parcels are cloned
    (how often) and deleted, the order of the clones in the cloud, their mirrored velocity and the rebuilt occupancy are the oracle's."""
    from tests.test_oracle_physics import _weighted_box
    mesh, md, pol, ang = _weighted_box()
    eng, ora = H.setup_pair(mesh, [H.argon()], md, capi.Engine, Oracle)
    cc, cv, fc, *_ = ora.geometry()
    rwf, _ = capi.axisymmetric_rwf(cc, fc, pol, 100.0)
    for x in (eng, ora):
        x.set_cell_fields(RWF=rwf)
    ora.mesh_fill([0], [4e18], 300.0)
    p = ora.download_parcels()
    p.radialWeight[:] = 20.0
    for x in (eng, ora):
        x.upload_parcels(p)
        x.stage(capi.STAGE_SORT)
    g, o = eng.download_parcels(), ora.download_parcels()
    cloned, deleted = ora.weighting_counts()
    assert cloned > 100 and deleted > 100
    assert g.n == o.n == p.n + cloned - deleted
    for k in ("origId", "cell", "tetFace", "tetPt", "typeId", "position", "U", "radialWeight"):
        assert np.array_equal(getattr(g, k), getattr(o, k)), k
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    eng.close()


def test_spherical_weighting_and_steps_match_oracle():
    """dsmcSpherical (spherical/dsmcSpherical.C): the weighting stage with weights growing as r^2 about the origin -- clones with the
    parent's velocity -- parcel by parcel, then three full steps (specular walls, VHS collisions weighted per cell, weighting after every
    move) against the oracle."""
    from tests.test_oracle_physics import _spherical_box
    mesh, md = _spherical_box()
    eng, ora = H.setup_pair(mesh, [H.argon()], md, capi.Engine, Oracle)
    cc, cv, fc, *_ = ora.geometry()
    rwf, _ = capi.spherical_rwf(cc, fc, (0, 0, 0), 100.0)
    for x in (eng, ora):
        x.set_cell_fields(RWF=rwf)
    ora.mesh_fill([0], [4e18], 300.0)
    p = ora.download_parcels()
    sig, rem = ora.download_cellstate()
    eng.upload_cellstate(sig, rem)
    p.radialWeight[:] = 20.0
    for x in (eng, ora):
        x.upload_parcels(p)
        x.stage(capi.STAGE_SORT)
    g, o = eng.download_parcels(), ora.download_parcels()
    cloned, deleted = ora.weighting_counts()
    assert cloned > 50 and deleted > 50 and g.n == o.n == p.n + cloned - deleted
    for k in ("origId", "cell", "tetFace", "tetPt", "typeId", "position", "U", "radialWeight"):
        assert np.array_equal(getattr(g, k), getattr(o, k)), k
    seen = 0
    for _ in range(3):
        eng.evolve(1)
        ora.evolve(1)
        assert eng.num_parcels() == ora.num_parcels()
        tot = ora.counters()["collisions"]
        assert eng.counters().collisions == tot - seen > 20      # cells of 9600 parcels (the centre) down to 60 (the far corner)
        seen = tot
    g, o = eng.download_parcels(), ora.download_parcels()
    assert np.array_equal(g.origId, o.origId) and np.array_equal(g.cell, o.cell) and np.array_equal(g.radialWeight, o.radialWeight)
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-13) and np.allclose(g.U, o.U, rtol=1e-9, atol=1e-7)
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    eng.close()


def test_axisymmetric_steps_match_oracle():
    """Five full steps of the reference's axisymmetric tutorial (wedge mesh with prisms on the axis, free-stream inflow weighted by the
    face cell's RWF, deletion, diffuse wall, symmetry planes, cloning / deletion after every move): the cloud is the oracle's parcel by
    parcel; wall accumulators (weighted by the parcels' RWF) agree to rounding."""
    gold = H.axisym_gold()
    eng, ora = capi.Engine(0), Oracle()
    mesh, spd, npc, cv = H.axisym_setup(ora, gold, ora.geometry)
    H.axisym_setup(eng, gold, eng.geometry)
    ora.mesh_fill([0], [float(gold["numberDensity"])], float(gold["temperature"]), 0, 0, 0, tuple(gold["velocity"]))
    start = ora.download_parcels()
    sig, rem = ora.download_cellstate()
    eng.upload_parcels(start)
    eng.upload_cellstate(sig, rem)
    for _ in range(5):
        eng.evolve(1)
        ora.evolve(1)
        assert eng.num_parcels() == ora.num_parcels()
    cloned, deleted = ora.weighting_counts()
    assert cloned > 100 and deleted > 100
    g, o = eng.download_parcels(), ora.download_parcels()
    assert np.array_equal(g.origId, o.origId) and np.array_equal(g.cell, o.cell)
    assert np.array_equal(g.radialWeight, o.radialWeight)
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-14)
    assert np.allclose(g.U, o.U, rtol=1e-9, atol=1e-7)
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    gw, ow = eng.wall_accumulators(), ora.wall_accumulators()
    assert np.abs(ow).max() > 0
    scale = np.abs(ow).max(axis=(0, 1), keepdims=True) + 1e-300
    assert (np.abs(gw - ow) / scale).max() < 1e-9
    ga, _, _ = eng.accumulators()
    oa, _, _ = ora.accumulators()
    assert np.array_equal(ga[:, :, 0], oa[:, :, 0])
    eng.close()


def test_variable_time_step_steps_match_oracle():
    """dsmcVariableTimeStepModel on a box whose cells grow along x: per-cell nParticles and deltaT through dsmcb200_set_cell_fields;
    three full steps equal the oracle's (cells, occupancy, collision counts; positions to rounding)."""
    mesh = meshgen.box_mesh((8, 4, 4), (0.08, 0.04, 0.04))
    x = mesh.points[:, 0].copy()
    mesh.points[:, 0] = 0.08 * (0.5 * (x / 0.08) + 0.5 * (x / 0.08) ** 2)
    sp = [H.argon()]
    md = capi.build_models("VariableHardSphere", nEquivalentParticles=4e9, deltaT=2e-6, seed=21)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    _, cv, *_ = ora.geometry()
    n, dt = capi.variable_time_step(cv, 4e9, 2e-6)
    for e in (eng, ora):
        e.set_cell_fields(nParticles=n, deltaT=dt)
    H.same_start(eng, ora, [0], [1e20], 300.0)
    assert np.allclose(eng.cell_fields()[1], dt)
    seen = 0
    for _ in range(3):
        eng.evolve(1)
        ora.evolve(1)
        total = ora.counters()
        assert eng.counters().collisions == total["collisions"] - seen > 0
        seen = total["collisions"]
    g, o = eng.download_parcels(), ora.download_parcels()
    assert np.array_equal(g.origId, o.origId) and np.array_equal(g.cell, o.cell)
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-13)
    eng.close()


def test_axisymmetric_tutorial_follows_the_shipped_fields():
    """The reference's axisymmetricFlatnosedCylinder tutorial on the GPU at its own size (4000 cells, ~184 000 parcels): 3000 steps to
    the steady state, 3000 sampled steps, then number density, temperature, velocity and parcels per cell against the fields the
    reference ships (averaged over 40 000 steps).  Over the field: no bias (mean deviation < 0.5 %, rms < 2.5 %).  Per cell: within the
    range the shipped field spans over the cell and its neighbours (the bow shock is a jump over two cells), widened by 4.5 sigma of
    the sampled parcel count (sqrt(1 / (N nSteps)), a factor 3 for the correlation of successive steps) plus 3 %."""
    gold = H.axisym_gold()
    eng = capi.Engine(0)
    mesh, spd, npc, cv = H.axisym_setup(eng, gold, eng.geometry)
    eng.mesh_fill([0], [float(gold["numberDensity"])], float(gold["temperature"]), 0, 0, 0, tuple(gold["velocity"]))
    eng.evolve(3000)
    eng.reset_accumulators()
    n_s = 3000
    eng.evolve(n_s)
    assert abs(eng.num_parcels() / gold["dsmcNMean_Ar"].sum() - 1) < 0.01          # 183 564 parcels in the shipped steady state
    acc, coll, nt = eng.accumulators()
    f = fields_ref.derive(acc, coll, nt, spd, [0], npc, cv, has_internal=False)
    N = gold["dsmcNMean_Ar"]
    sig = 3.0 / np.sqrt(N * n_s)
    for k, g in (("dsmcNMean", N), ("rhoN", gold["rhoN_Ar"]), ("Ttra", gold["Ttra_Ar"])):
        r = f[k] / g - 1
        assert abs(r.mean()) < 0.005 and np.sqrt((r ** 2).mean()) < 0.025, (k, r.mean(), np.sqrt((r ** 2).mean()))   # no bias over the field
        lo, hi = H.axisym_envelope(g)
        tol = (4.5 * sig + 0.03) * g
        if k == "Ttra":
            # ahead of the shock the temperature of a cell is carried by the few fast molecules scattered back from the shock layer:
            # a fraction (T - T_inf) / Theta of the parcels, Theta = m U_inf^2 / (3 k) = 1600 K, whose count is what scatters
            theta = float(gold["mass"]) * 1000.0 ** 2 / (3 * H.KB)
            tol = 4.5 * 3.0 * np.sqrt(np.maximum(g - float(gold["temperature"]), 10.0) * theta / (N * n_s)) + 0.03 * g
        excess = np.maximum(lo - tol - f[k], f[k] - hi - tol) / tol
        assert excess.max() < 0.0, (k, excess.max(), int(excess.argmax()))
    cbar = np.sqrt(2 * H.KB * gold["Ttra_Ar"] / float(gold["mass"]))
    for d in range(2):
        g = gold["U_Ar"][:, d]
        lo, hi = H.axisym_envelope(g)
        tol = 4.5 * sig * cbar + 0.03 * 1000.0
        excess = np.maximum(lo - tol - f["UMean"][:, d], f["UMean"][:, d] - hi - tol) / tol
        assert excess.max() < 0.0, ("U", d, excess.max(), int(excess.argmax()))
    # the stagnation region in front of the flat face: density rise and temperature of the shock layer
    assert 6.0 < f["rhoN"].max() / 1e21 < 1.1 * gold["rhoN_Ar"].max() / 1e21
    eng.close()
