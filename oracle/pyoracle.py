"""ctypes face of oracle/liboracle.so -- the CPU restatement of the reference path.

TEST INFRASTRUCTURE ONLY (see oracle.cpp header): imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs, never by the hystrath_b200 package.
The interface structs are the PODs of include/dsmcb200.h (via hystrath_b200.capi).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from hystrath_b200 import capi
from hystrath_b200.capi import AccumInfo, MeshData, Models, ParcelData, Species, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    P = C.c_void_p
    lib.oracle_create.restype = P
    lib.oracle_last_error.restype = C.c_char_p
    lib.oracle_num_parcels.restype = C.c_int64
    lib.oracle_outbox_size.restype = C.c_int64
    for name in ("oracle_destroy", "oracle_last_error", "oracle_num_parcels", "oracle_outbox_size", "oracle_evolve_begin", "oracle_evolve_end"):
        getattr(lib, name).argtypes = [P]
    lib.oracle_set_mesh.argtypes = [P, C.POINTER(capi.Mesh)]
    lib.oracle_set_species.argtypes = [P, C.c_int, C.POINTER(Species)]
    lib.oracle_set_models.argtypes = [P, C.POINTER(Models)]
    lib.oracle_set_reorder.argtypes = [P, C.c_int]
    lib.oracle_set_step.argtypes = [P, C.c_uint32]
    lib.oracle_set_rank.argtypes = [P, C.c_int]
    lib.oracle_set_reactions.argtypes = [P, C.c_int, C.c_void_p]
    lib.oracle_reaction_counts.argtypes = [P, C.c_void_p]
    lib.oracle_set_threads.argtypes = [C.c_int]
    lib.oracle_get_threads.restype = C.c_int
    lib.oracle_upload_parcels.argtypes = [P, C.c_int64, C.POINTER(capi.ParcelsSoA)]
    lib.oracle_download_parcels.argtypes = [P, C.POINTER(capi.ParcelsSoA)]
    lib.oracle_upload_cellstate.argtypes = [P, C.c_void_p, C.c_void_p]
    lib.oracle_set_cell_fields.argtypes = [P, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_weighting_counts.argtypes = [P, C.c_void_p, C.c_void_p]
    lib.oracle_download_cellstate.argtypes = [P, C.c_void_p, C.c_void_p]
    lib.oracle_mesh_fill.argtypes = [P, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.oracle_zone_fill.argtypes = [P, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.oracle_stage.argtypes = [P, C.c_int]
    lib.oracle_evolve.argtypes = [P, C.c_int]
    lib.oracle_outbox_get.argtypes = [P, C.c_void_p, C.c_void_p]
    lib.oracle_receive_and_move.argtypes = [P, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    lib.oracle_download_occupancy.argtypes = [P, C.c_void_p]
    lib.oracle_accum_info.argtypes = [P, C.POINTER(AccumInfo)]
    lib.oracle_download_accumulators.argtypes = [P, C.c_void_p, C.c_void_p]
    lib.oracle_wall_info.argtypes = [P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.oracle_download_wall_accumulators.argtypes = [P, C.c_void_p]
    lib.oracle_download_face_fluxes.argtypes = [P, C.c_void_p, C.c_void_p]
    lib.oracle_upload_overall_temperature.argtypes = [P, C.c_void_p]
    lib.oracle_get_counters.argtypes = [P, C.c_void_p]
    lib.oracle_download_geometry.argtypes = [P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _LIB = lib
    return lib


def set_threads(n):
    """OpenMP threads of the oracle in this process (overrides OMP_NUM_THREADS); returns the team size actually obtained."""
    lib = load()
    lib.oracle_set_threads(int(n))
    return int(lib.oracle_get_threads())


class Oracle:
    """Same surface as hystrath_b200.capi.Engine, computed on the CPU by the restatement."""

    def __init__(self):
        self.lib = load()
        self.h = C.c_void_p(self.lib.oracle_create())
        self.max_modes = 1

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.lib.oracle_last_error(self.h).decode())

    def set_mesh(self, mesh: MeshData):
        self._mesh = mesh
        st = mesh.as_struct()
        self._ck(self.lib.oracle_set_mesh(self.h, C.byref(st)))

    def set_species(self, species):
        arr = (Species * len(species))(*species)
        self.n_species = len(species)
        self.max_modes = max(1, max(s.nVibrationalModes for s in species))
        self._ck(self.lib.oracle_set_species(self.h, len(species), arr))

    def set_models(self, models):
        self._models = models
        self._ck(self.lib.oracle_set_models(self.h, C.byref(models)))

    def set_reorder(self, on):
        self.lib.oracle_set_reorder(self.h, int(on))

    def set_step(self, step):
        self.lib.oracle_set_step(self.h, step)

    def set_cell_fields(self, nParticles=None, deltaT=None, RWF=None):
        a = [None if x is None else np.ascontiguousarray(x, np.float64) for x in (nParticles, deltaT, RWF)]
        self._ck(self.lib.oracle_set_cell_fields(self.h, *[None if x is None else _ptr(x) for x in a]))

    def weighting_counts(self):
        """parcels cloned / deleted by dsmcAxisymmetric::axisymmetricWeighting since creation"""
        a, b = C.c_int64(), C.c_int64()
        self.lib.oracle_weighting_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def set_reactions(self, reactions):
        self._reactions = reactions
        self._ck(self.lib.oracle_set_reactions(self.h, reactions._n, C.cast(reactions, C.c_void_p)))

    def reaction_counts(self):
        n = self._reactions._n
        out = np.zeros((max(n, 1), 3), np.int64)
        self.lib.oracle_reaction_counts(self.h, _ptr(out))
        return out[:n]

    def set_rank(self, rank):
        """Pstream::myProcNo() of this instance: origProc of the parcels it creates."""
        self.lib.oracle_set_rank(self.h, rank)

    def upload_parcels(self, p: ParcelData):
        st = p.as_struct()
        self._ck(self.lib.oracle_upload_parcels(self.h, p.n, C.byref(st)))

    def num_parcels(self):
        return self.lib.oracle_num_parcels(self.h)

    def download_parcels(self):
        out = ParcelData(self.num_parcels(), self.max_modes)
        st = out.as_struct()
        self._ck(self.lib.oracle_download_parcels(self.h, C.byref(st)))
        return out

    def upload_cellstate(self, sigma=None, remainder=None):
        sigma = None if sigma is None else np.ascontiguousarray(sigma, np.float64)
        remainder = None if remainder is None else np.ascontiguousarray(remainder, np.float64)
        self._ck(self.lib.oracle_upload_cellstate(self.h, _ptr(sigma), _ptr(remainder)))

    def download_cellstate(self):
        n = self._mesh.n_cells
        s, r = np.zeros(n), np.zeros(n)
        self._ck(self.lib.oracle_download_cellstate(self.h, _ptr(s), _ptr(r)))
        return s, r

    def mesh_fill(self, type_ids, number_densities, Ttra, Trot=0.0, Tvib=0.0, Telec=0.0, velocity=(0.0, 0.0, 0.0)):
        t = np.ascontiguousarray(type_ids, np.int32)
        nd = np.ascontiguousarray(number_densities, np.float64)
        v = np.ascontiguousarray(velocity, np.float64)
        self._ck(self.lib.oracle_mesh_fill(self.h, len(t), _ptr(t), _ptr(nd), Ttra, Trot, Tvib, Telec, _ptr(v)))

    def zone_fill(self, zone_cells, type_ids, number_densities, Ttra, Trot=0.0, Tvib=0.0, Telec=0.0, velocity=(0.0, 0.0, 0.0)):
        z = np.ascontiguousarray(zone_cells, np.int32)
        t = np.ascontiguousarray(type_ids, np.int32)
        nd = np.ascontiguousarray(number_densities, np.float64)
        v = np.ascontiguousarray(velocity, np.float64)
        self._ck(self.lib.oracle_zone_fill(self.h, C.c_int64(len(z)), _ptr(z), len(t), _ptr(t), _ptr(nd), Ttra, Trot, Tvib, Telec, _ptr(v)))

    def evolve(self, n=1):
        self._ck(self.lib.oracle_evolve(self.h, n))

    def stage(self, s):
        self._ck(self.lib.oracle_stage(self.h, s))

    def occupancy(self):
        off = np.zeros(self._mesh.n_cells + 1, np.int32)
        self._ck(self.lib.oracle_download_occupancy(self.h, _ptr(off)))
        return off

    def accumulators(self):
        i = AccumInfo()
        self._ck(self.lib.oracle_accum_info(self.h, C.byref(i)))
        acc = np.zeros((i.nCells, i.nSpecies, i.nQuantities))
        coll = np.zeros((i.nCells, 2))
        self._ck(self.lib.oracle_download_accumulators(self.h, _ptr(acc), _ptr(coll)))
        return acc, coll, i.nTimeSteps

    def wall_accumulators(self):
        nf, nq = C.c_int32(), C.c_int32()
        self._ck(self.lib.oracle_wall_info(self.h, C.byref(nf), C.byref(nq)))
        w = np.zeros((nf.value, self.n_species, nq.value))
        if nf.value:
            self._ck(self.lib.oracle_download_wall_accumulators(self.h, _ptr(w)))
        return w

    def upload_overall_temperature(self, Tov):
        t = np.ascontiguousarray(Tov, np.float64)
        self._ck(self.lib.oracle_upload_overall_temperature(self.h, _ptr(t)))

    def face_fluxes(self):
        """dsmcFaceTracker parcelIdFlux / massIdFlux of the last step, each [nSpecies][nFaces]."""
        pf, mf = np.zeros((self.n_species, self._mesh.n_faces)), np.zeros((self.n_species, self._mesh.n_faces))
        self._ck(self.lib.oracle_download_face_fluxes(self.h, _ptr(pf), _ptr(mf)))
        return pf, mf

    def counters(self):
        c = np.zeros(6, np.int64)
        self._ck(self.lib.oracle_get_counters(self.h, _ptr(c)))
        return dict(nParcels=int(c[0]), collisions=int(c[1]), collisionCandidates=int(c[2]), trackingRescues=int(c[3]),
                    deleted=int(c[4]), inserted=int(c[5]))

    def geometry(self):
        m = self._mesh
        cc, cv = np.zeros((m.n_cells, 3)), np.zeros(m.n_cells)
        fc, fa = np.zeros((m.n_faces, 3)), np.zeros((m.n_faces, 3))
        tb = np.zeros(m.n_faces, np.int32)
        self._ck(self.lib.oracle_download_geometry(self.h, _ptr(cc), _ptr(cv), _ptr(fc), _ptr(fa), _ptr(tb)))
        return cc, cv, fc, fa, tb

    # ---- decomposed runs (the MPI block of Cloud<T>::move is driven by the caller) ----
    def evolve_begin(self):
        self._ck(self.lib.oracle_evolve_begin(self.h))

    def outbox(self):
        n = self.lib.oracle_outbox_size(self.h)
        d, i = np.zeros((n, 9)), np.zeros((n, 12), np.int32)
        self._ck(self.lib.oracle_outbox_get(self.h, _ptr(d), _ptr(i)))
        return d, i

    def receive_and_move(self, from_proc, d, i):
        d = np.ascontiguousarray(d, np.float64)
        i = np.ascontiguousarray(i, np.int32)
        self._ck(self.lib.oracle_receive_and_move(self.h, from_proc, len(d), _ptr(d), _ptr(i)))

    def evolve_end(self):
        self._ck(self.lib.oracle_evolve_end(self.h))
