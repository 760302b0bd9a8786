"""CPU pin of the move kernel's per-visit code: hystrath_b200/csrc/move_core.h on the tables baked by host_mesh.cpp (through the harness
tests/native/libmovecheck.so) against the oracle's restatement of particle::trackToFace -- bit for bit, on hex, renumbered, prism,
tetrahedral and 2:1-refined meshes, with the fast (division-free plane test) and the verbatim path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from hystrath_b200 import capi, meshgen
from oracle.pyoracle import Oracle
from tests import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "native")])
    lib = C.CDLL(os.path.join(HERE, "native", "libmovecheck.so"))
    lib.movecheck_run.restype = C.c_int
    lib.movecheck_run.argtypes = [C.c_void_p, C.c_double, C.c_int64] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
    lib.movecheck_tets.restype = C.c_int
    lib.movecheck_tets.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def run_harness(lib, mesh, dt, start, steps=1, force_slow=False):
    st = mesh.as_struct()
    pos = np.ascontiguousarray(start.position, np.float64).copy()
    U = np.ascontiguousarray(start.U, np.float64).copy()
    cell, tf, tp = (np.ascontiguousarray(a, np.int32).copy() for a in (start.cell, start.tetFace, start.tetPt))
    stats = np.zeros(4, np.int64)
    tot = np.zeros(4, np.int64)
    for _ in range(steps):
        rc = lib.movecheck_run(C.addressof(st), dt, len(cell), _ptr(pos), _ptr(U), _ptr(cell), _ptr(tf), _ptr(tp), int(force_slow), _ptr(stats))
        assert rc == 0, rc
        tot[:3] += stats[:3]
    return dict(position=pos, U=U, cell=cell, tetFace=tf, tetPt=tp, rescues=int(tot[0]), visits=int(tot[1]), slow=int(tot[2]))


def oracle_free_flight(mesh, sp, md, steps, dens=1e20, T=300.0, velocity=(0, 0, 0)):
    ora = Oracle()
    ora.set_mesh(mesh); ora.set_species(sp); ora.set_models(md)
    ora.mesh_fill([0], [dens], T, 0.0, 0.0, 0.0, velocity)
    start = ora.download_parcels()
    for _ in range(steps):
        ora.stage(capi.STAGE_MOVE)
    return start, ora.download_parcels(), ora.counters()


def check(h, end_o, start):
    """harness output (in start order) against the oracle's cloud (compacted: deleted parcels removed, order kept)"""
    keep = h["cell"] >= 0
    assert keep.sum() == end_o.n
    ids = start.origId[keep]
    assert np.array_equal(ids, end_o.origId)
    assert np.array_equal(h["cell"][keep], end_o.cell)
    assert np.array_equal(h["tetFace"][keep], end_o.tetFace) and np.array_equal(h["tetPt"][keep], end_o.tetPt)
    assert np.array_equal(h["position"][keep], end_o.position)     # identical FP64 operations: bit for bit
    assert np.array_equal(h["U"][keep], end_o.U)


@pytest.mark.parametrize("force_slow", [False, True])
def test_periodic_hex_box(harness, force_slow):
    mesh = meshgen.box_mesh((8, 6, 5), (0.032, 0.024, 0.020))
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e20 * 0.032 * 0.024 * 0.020 / (240 * 30), deltaT=5e-6, seed=5)
    start, end, cnt = oracle_free_flight(mesh, [H.argon()], md, 3)
    h = run_harness(harness, mesh, 5e-6, start, 3, force_slow)
    check(h, end, start)
    assert h["rescues"] == cnt["trackingRescues"]
    assert h["visits"] > 2 * start.n
    if not force_slow:
        assert h["slow"] < 0.01 * h["visits"]      # the tolerance band is the exception


def test_renumbered_mesh(harness):
    base = meshgen.box_mesh((8, 6, 5), (0.032, 0.024, 0.020))
    mesh, _ = meshgen.renumber_cells(base, meshgen.morton_order(base))
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e20 * 0.032 * 0.024 * 0.020 / (240 * 30), deltaT=5e-6, seed=5)
    start, end, _ = oracle_free_flight(mesh, [H.argon()], md, 3)
    check(run_harness(harness, mesh, 5e-6, start, 3), end, start)


@pytest.mark.parametrize("kind", ["prism", "tet"])
def test_non_hex_cells_with_specular_walls(harness, kind):
    mesh, locate = meshgen.split_box_mesh((4, 3, 3), (0.016, 0.012, 0.012), kind)
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e20 * 0.016 * 0.012 * 0.012 / (mesh.n_cells * 30), deltaT=5e-6, seed=11,
                           patch_models=[dict(patch=0, boundaryModel="dsmcSpecularWallPatch")])
    start, end, _ = oracle_free_flight(mesh, [H.argon()], md, 4)
    h = run_harness(harness, mesh, 5e-6, start, 4)
    check(h, end, start)
    assert np.array_equal(locate(h["position"]), h["cell"])


def test_refinement_interface(harness):
    mesh, locate = meshgen.refined_interface_mesh()
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e20 * 3 * 0.004 ** 3 / (6 * 400), deltaT=2e-6, seed=3,
                           patch_models=[dict(patch=0, boundaryModel="dsmcSpecularWallPatch")])
    start, end, _ = oracle_free_flight(mesh, [H.argon()], md, 6)
    h = run_harness(harness, mesh, 2e-6, start, 6)
    check(h, end, start)
    assert np.array_equal(locate(h["position"]), h["cell"])


def test_two_dimensional_cases_with_empty_patches(harness):
    """One cell thick with empty front and back: the constrained track runs parallel to the planes of the empty faces, whose
    denominators are exactly zero -- they are not crossed and must not push the visit off the fast path.  A box with specular
    walls and the cylinder O-grid of BASELINE configs[1] (skewed hexahedra, specular cylinder, deletion on the outer patch)."""
    mesh = meshgen.box_mesh((10, 8, 1), (0.02, 0.016, 0.002), sides={"xmin": ("cyclic",), "xmax": ("cyclic",), "ymin": ("wall", "walls"),
                                                                       "ymax": ("wall", "walls"), "zmin": ("empty", "frontAndBack"), "zmax": ("empty", "frontAndBack")})
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e20 * 0.02 * 0.016 * 0.002 / (80 * 40), deltaT=4e-6, seed=9,
                           patch_models=[dict(patch=mesh.patch_index("walls"), boundaryModel="dsmcSpecularWallPatch")])
    start, end, _ = oracle_free_flight(mesh, [H.argon()], md, 4, velocity=(200.0, 50.0, 0.0))
    h = run_harness(harness, mesh, 4e-6, start, 4)
    check(h, end, start)
    assert h["slow"] == h["rescues"]       # only the rescue corrections (the fill's parcels are snapped to the mid-plane) leave the fast path

    mesh = meshgen.cylinder_ogrid(12, 40, 0.05, 0.2, grading=3.0)
    vol = np.pi * (0.2 ** 2 - 0.05 ** 2) * float(np.ptp(mesh.points[:, 2]))
    md = capi.build_models("NoBinaryCollision", nEquivalentParticles=1e19 * vol / (mesh.n_cells * 20), deltaT=2e-5, seed=9,
                           patch_models=[dict(patch=mesh.patch_index("cylinder"), boundaryModel="dsmcSpecularWallPatch"),
                                         dict(patch=mesh.patch_index("outer"), boundaryModel="dsmcDeletionPatch")])
    start, end, cnt = oracle_free_flight(mesh, [H.argon()], md, 5, dens=1e19, velocity=(400.0, 0.0, 0.0))
    h = run_harness(harness, mesh, 2e-5, start, 5)
    assert cnt["deleted"] > 0
    check(h, end, start)


def test_tet_table_invariants(harness):
    """cell-major numbering: in-cell links stay inside the cell, `across` links are mutual and keep (face, tetPt)."""
    for mesh in (meshgen.box_mesh((4, 3, 2), (0.4, 0.3, 0.2)), meshgen.split_box_mesh((3, 2, 2), (0.3, 0.2, 0.2), "tet")[0],
                 meshgen.refined_interface_mesh()[0]):
        st = mesh.as_struct()
        cap = 200000
        cell, face, tetPt, across = (np.zeros(cap, np.int32) for _ in range(4))
        nbr = np.zeros((cap, 3), np.int32)
        start = np.zeros(mesh.n_cells + 1, np.int32)
        n = harness.movecheck_tets(C.addressof(st), cap, _ptr(cell), _ptr(face), _ptr(tetPt), _ptr(across), _ptr(nbr), _ptr(start))
        assert 0 < n <= cap
        cell, face, tetPt, across, nbr = cell[:n], face[:n], tetPt[:n], across[:n], nbr[:n]
        assert start[0] == 0 and start[-1] == n and np.all(np.diff(start) > 0)
        assert np.array_equal(cell, np.repeat(np.arange(mesh.n_cells), np.diff(start)))
        assert np.all(cell[nbr] == cell[:, None])                       # tris 1-3 lead to tets of the same cell
        assert np.all((nbr >= start[cell][:, None]) & (nbr < start[cell + 1][:, None]))
        internal = across >= 0
        assert np.array_equal(across[across[internal]], np.nonzero(internal)[0])
        assert np.array_equal(face[across[internal]], face[internal]) and np.array_equal(tetPt[across[internal]], tetPt[internal])
        assert np.all(cell[across[internal]] != cell[internal])
        assert np.all(face[~internal] - mesh.n_internal == -1 - across[~internal])
