"""GPU: cell labels inside the engine (dsmcb200_set_cell_order, the in-memory counterpart of OpenFOAM's renumberMesh).  The engine that
relabels the caller's mesh itself must give, label for label, what the oracle gives on the mesh relabelled with the same table: the whole
translation sits in the entry points (parcels, cell state, accumulators, occupancy, geometry), the kernels never see the caller's labels."""
import copy

import numpy as np
import pytest

from hystrath_b200 import capi, meshgen
from oracle.pyoracle import Oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _case():
    # x cyclic, y diffuse walls (the upper one moving), z cyclic: walls, coupled patches and Larsen-Borgnakke collisions of N2 / O2
    sides = {"xmin": ("cyclic",), "xmax": ("cyclic",), "ymin": ("wall", "lowerWall"), "ymax": ("wall", "upperWall"), "zmin": ("cyclic",), "zmax": ("cyclic",)}
    mesh = meshgen.box_mesh((6, 7, 5), (0.06, 0.07, 0.05), sides=sides)
    sp = H.air5()[:2]
    pm = [dict(patch=mesh.patch_index("lowerWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=2000.0, velocity=(0, 0, 0)),
          dict(patch=mesh.patch_index("upperWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=3000.0, velocity=(300.0, 0, 0))]
    vol = 0.06 * 0.07 * 0.05
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=1e20 * vol / (210 * 40), deltaT=4e-6, seed=99, patch_models=pm,
                           inverseZvFormulation="pre-2008")
    return mesh, sp, md


def _relabelled(p, table):
    q = copy.copy(p)
    q.cell = table[p.cell].astype(np.int32)
    return q


@pytest.mark.parametrize("mode", ["z-curve", "given"])
def test_engine_with_its_own_cell_labels_equals_the_oracle_on_the_relabelled_mesh(mode):
    mesh, sp, md = _case()
    eng = capi.Engine(0)
    if mode == "given":
        eng.set_cell_order(new_of_old=np.random.default_rng(5).permutation(mesh.n_cells))
    else:
        eng.set_cell_order("z-curve")
    eng.set_mesh(mesh); eng.set_species(sp); eng.set_models(md)
    t = eng.cell_order()
    assert sorted(t.tolist()) == list(range(mesh.n_cells)) and not np.array_equal(t, np.arange(mesh.n_cells))
    # the start: a cloud and a cell state in the CALLER's labels (made by an oracle on the caller's mesh)
    src = Oracle()
    src.set_mesh(mesh); src.set_species(sp); src.set_models(md)
    src.mesh_fill([0, 1], [0.8e20, 0.2e20], 2500.0, 2500.0, 2500.0, 0.0, (0, 0, 0))
    start = src.download_parcels()
    sig, _ = src.download_cellstate()
    sig = sig * (1.0 + 0.5 * np.random.default_rng(1).random(mesh.n_cells))     # per-cell values, so a wrong row shows
    rem = np.random.default_rng(2).random(mesh.n_cells)
    eng.upload_parcels(start)
    eng.upload_cellstate(sig, rem)
    s2, r2 = eng.download_cellstate()
    assert np.array_equal(s2, sig) and np.array_equal(r2, rem)                  # round trip in the caller's labels

    # the oracle on the relabelled mesh, fed the same cloud in the engine's labels
    ora = Oracle()
    ora.set_mesh(meshgen.relabel_cells(mesh, t)); ora.set_species(sp); ora.set_models(md)
    ora.set_reorder(True)
    ora.upload_parcels(_relabelled(start, t))
    inv = np.empty_like(t); inv[t] = np.arange(len(t), dtype=np.int32)
    ora.upload_cellstate(sig[inv], rem[inv])

    collisions = 0
    for _ in range(3):
        eng.evolve(1)
        ora.evolve(1)
        collisions += eng.counters().collisions      # the engine counts per step, the oracle from the start
    g, o = eng.download_parcels(), ora.download_parcels()
    assert g.n == o.n == start.n
    assert np.array_equal(g.origId, o.origId)                   # the same cloud order: cell-major in the engine's labels
    assert np.array_equal(t[g.cell], o.cell)                    # ... and the caller sees its own labels
    assert np.array_equal(g.tetFace, o.tetFace) and np.array_equal(g.tetPt, o.tetPt)
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-13)      # velocities differ by the rounding of pow() on the two sides
    assert np.allclose(g.U, o.U, rtol=1e-12, atol=1e-9) and np.array_equal(g.vibLevel, o.vibLevel)
    assert collisions == ora.counters()["collisions"] > 50
    # per-cell rows come back in the caller's labels
    ga, gc, _ = eng.accumulators()
    oa, oc, _ = ora.accumulators()
    assert np.abs(oa).sum() > 0 and np.allclose(ga, oa[t], rtol=1e-12, atol=0) and np.allclose(gc, oc[t], rtol=1e-12, atol=0)
    gs, gr = eng.download_cellstate()
    os_, or_ = ora.download_cellstate()
    assert np.allclose(gs, os_[t], rtol=1e-12) and np.allclose(gr, or_[t], rtol=0, atol=1e-9)
    # occupancy in the caller's labels = that of the downloaded cloud ordered by the caller's cells
    off = eng.occupancy()
    assert np.array_equal(np.diff(off), np.bincount(g.cell, minlength=mesh.n_cells))
    # geometry in the caller's labels
    cc, cv, *_ = eng.geometry()
    occ, ocv, *_ = ora.geometry()
    assert np.allclose(cc, occ[t], rtol=0, atol=1e-15) and np.allclose(cv, ocv[t], rtol=1e-14)
    eng.close()


def test_mesh_fill_and_uploads_without_tet_indices_follow_the_cell_order():
    """dsmcMeshFill on the device with the engine's labels = the oracle's fill on the relabelled mesh; a cloud uploaded without tet
    indices is located in the caller's cells."""
    mesh, sp, md = _case()
    eng = capi.Engine(0)
    eng.set_cell_order("z-curve")
    eng.set_mesh(mesh); eng.set_species(sp); eng.set_models(md)
    t = eng.cell_order()
    ora = Oracle()
    ora.set_mesh(meshgen.relabel_cells(mesh, t)); ora.set_species(sp); ora.set_models(md)
    for x in (eng, ora):
        x.mesh_fill([0, 1], [0.8e20, 0.2e20], 2500.0, 2500.0, 2500.0, 0.0, (0, 0, 0))
    g, o = eng.download_parcels(), ora.download_parcels()
    assert g.n == o.n > 5000 and np.array_equal(t[g.cell], o.cell) and np.array_equal(g.position, o.position) and np.array_equal(g.typeId, o.typeId)
    # the same cloud again, this time without tet indices: every parcel is found in the cell the caller names
    p = copy.copy(g)
    p.tetFace = None; p.tetPt = None
    eng.upload_parcels(p)
    h = eng.download_parcels()
    assert h.n == g.n
    a, b = H.by_id(h), H.by_id(g)
    assert np.array_equal(a["cell"], b["cell"]) and np.array_equal(a["position"], b["position"])
    eng.close()


def test_axisymmetric_case_with_cell_fields_under_a_cell_order():
    """Per-cell nParticles / RWF (dsmcb200_set_cell_fields) and the parcels' radial weights under the engine's own labels: three steps of the
    axisymmetric tutorial (prisms on the axis, weighted inflow, cloning / deletion, diffuse wall) equal the oracle on the relabelled mesh."""
    gold = H.axisym_gold()
    base = meshgen.axisymmetric_cylinder_mesh()
    eng = capi.Engine(0)
    eng.set_cell_order("z-curve")
    _, _, npc, _ = H.axisym_setup(eng, gold, eng.geometry, mesh=base)
    t = eng.cell_order()
    assert not np.array_equal(t, np.arange(base.n_cells))
    ora = Oracle()
    _, _, npc_o, _ = H.axisym_setup(ora, gold, ora.geometry, mesh=meshgen.relabel_cells(base, t))
    assert np.allclose(npc, npc_o[t], rtol=1e-14)
    _, _, rwf = eng.cell_fields()
    assert np.allclose(rwf * float(gold["nEquivalentParticles"]), npc, rtol=1e-14)     # what was set comes back in the caller's labels
    ora.mesh_fill([0], [float(gold["numberDensity"])], float(gold["temperature"]), 0, 0, 0, tuple(gold["velocity"]))
    start = ora.download_parcels()
    sig, rem = ora.download_cellstate()
    inv = np.empty_like(t); inv[t] = np.arange(len(t), dtype=np.int32)
    eng.upload_parcels(_relabelled(start, inv))
    eng.upload_cellstate(sig[t], rem[t])
    for _ in range(3):
        eng.evolve(1)
        ora.evolve(1)
        assert eng.num_parcels() == ora.num_parcels()
    cloned, deleted = ora.weighting_counts()
    assert cloned > 50 and deleted > 50
    g, o = eng.download_parcels(), ora.download_parcels()
    assert np.array_equal(g.origId, o.origId) and np.array_equal(t[g.cell], o.cell)
    assert np.array_equal(g.radialWeight, o.radialWeight)
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-14) and np.allclose(g.U, o.U, rtol=1e-9, atol=1e-7)
    gw, ow = eng.wall_accumulators(), ora.wall_accumulators()
    scale = np.abs(ow).max(axis=(0, 1), keepdims=True) + 1e-300
    assert np.abs(ow).max() > 0 and (np.abs(gw - ow) / scale).max() < 1e-9
    ga, _, _ = eng.accumulators()
    oa, _, _ = ora.accumulators()
    assert np.array_equal(ga[:, :, 0], oa[t][:, :, 0])
    eng.close()


def test_cell_order_is_checked():
    mesh, sp, md = _case()
    eng = capi.Engine(0)
    bad = np.arange(mesh.n_cells, dtype=np.int32); bad[3] = 4          # not a permutation
    eng.set_cell_order(new_of_old=bad)
    with pytest.raises(capi.Dsmcb200Error, match="not a permutation"):
        eng.set_mesh(mesh)
    eng.set_cell_order(new_of_old=np.arange(mesh.n_cells - 1, dtype=np.int32))
    with pytest.raises(capi.Dsmcb200Error, match="number of cells"):
        eng.set_mesh(mesh)
    eng.set_cell_order("as-given")
    eng.set_mesh(mesh)
    assert np.array_equal(eng.cell_order(), np.arange(mesh.n_cells))
    with pytest.raises(capi.Dsmcb200Error, match="before set_mesh"):
        eng.set_cell_order("z-curve")
    eng.close()
