"""ctypes binding of include/dsmcb200.h (libdsmcb200.so).

This is the thin Python face of the C ABI used by tests/ and bench.py.  The product path is
the shared library itself: if it has not been built, or no CUDA device is present, calls fail
loudly -- there is no CPU fallback (north_star).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MAX_SPECIES = 8
MAX_VIB_MODES = 3
MAX_ELEC_LEVELS = 16
NAME_LEN = 64
MAX_NEIGHBOURS = 16

# dsmcb200_patch_type
PATCH_WALL, PATCH_PATCH, PATCH_CYCLIC, PATCH_PROCESSOR, PATCH_EMPTY = 0, 1, 2, 3, 4
PATCH_SYMMETRYPLANE, PATCH_SYMMETRY, PATCH_WEDGE, PATCH_PROCESSORCYCLIC = 5, 6, 7, 8
PATCH_TYPE_NAMES = {
    "wall": PATCH_WALL, "patch": PATCH_PATCH, "cyclic": PATCH_CYCLIC, "processor": PATCH_PROCESSOR,
    "empty": PATCH_EMPTY, "symmetryPlane": PATCH_SYMMETRYPLANE, "symmetry": PATCH_SYMMETRY,
    "wedge": PATCH_WEDGE, "processorCyclic": PATCH_PROCESSORCYCLIC,
}
COLL_NONE, COLL_VHS, COLL_LB_VHS = 0, 1, 2
BND_NONE, BND_DIFFUSE_WALL, BND_SPECULAR_WALL, BND_DELETION = 0, 1, 2, 3
STAGE_INFLOW, STAGE_MOVE, STAGE_SORT, STAGE_COLLIDE, STAGE_SAMPLE = 0, 1, 2, 3, 4

# run-time selection names of the reference -> ABI enums
COLLISION_MODEL_NAMES = {
    "NoBinaryCollision": COLL_NONE,
    "VariableHardSphere": COLL_VHS,
    "LarsenBorgnakkeVariableHardSphere": COLL_LB_VHS,
    "VariableSoftSphere": 3,
    "LarsenBorgnakkeVariableSoftSphere": 4,
}
PATCH_MODEL_NAMES = {
    "dsmcDiffuseWallPatch": BND_DIFFUSE_WALL,
    "dsmcSpecularWallPatch": BND_SPECULAR_WALL,
    "dsmcDeletionPatch": BND_DELETION,
    "dsmcDiffuseSpecularWallPatch": 4,
    "dsmcCLLWallPatch": 5,
}


class Patch(C.Structure):
    _fields_ = [("name", C.c_char * NAME_LEN), ("type", C.c_int32), ("start", C.c_int32), ("size", C.c_int32),
                ("neighbPatch", C.c_int32), ("myProcNo", C.c_int32), ("neighbProcNo", C.c_int32),
                ("referPatch", C.c_int32), ("hasSeparation", C.c_int32), ("separation", C.c_double * 3)]


class Mesh(C.Structure):
    _fields_ = [("nPoints", C.c_int32), ("nFaces", C.c_int32), ("nInternalFaces", C.c_int32), ("nCells", C.c_int32),
                ("nPatches", C.c_int32),
                ("points", C.c_void_p), ("faceOffsets", C.c_void_p), ("facePoints", C.c_void_p), ("owner", C.c_void_p),
                ("neighbour", C.c_void_p), ("patches", C.POINTER(Patch)),
                ("cellCentres", C.c_void_p), ("cellVolumes", C.c_void_p), ("faceCentres", C.c_void_p),
                ("faceAreas", C.c_void_p), ("tetBasePtIs", C.c_void_p)]


class Species(C.Structure):
    _fields_ = [("name", C.c_char * NAME_LEN), ("mass", C.c_double), ("diameter", C.c_double), ("omega", C.c_double),
                ("alpha", C.c_double), ("rotationalDegreesOfFreedom", C.c_double), ("nVibrationalModes", C.c_int32),
                ("charge", C.c_int32), ("thetaV", C.c_double * MAX_VIB_MODES), ("Zref", C.c_double * MAX_VIB_MODES),
                ("TrefZv", C.c_double * MAX_VIB_MODES), ("thetaD", C.c_double), ("nElectronicLevels", C.c_int32),
                ("pad_", C.c_int32), ("electronicEnergyList", C.c_double * MAX_ELEC_LEVELS),
                ("electronicDegeneracyList", C.c_int32 * MAX_ELEC_LEVELS)]


class PatchModel(C.Structure):
    _fields_ = [("patch", C.c_int32), ("model", C.c_int32), ("temperature", C.c_double), ("velocity", C.c_double * 3),
                ("diffuseFraction", C.c_double), ("linearTemperature", C.c_int32), ("depthAxis", C.c_int32),
                ("formationLevelTemperature", C.c_double), ("normalAccommodationCoefficient", C.c_double),
                ("tangentialAccommodationCoefficient", C.c_double), ("rotationalEnergyAccommodationCoefficient", C.c_double)]


class Inflow(C.Structure):
    _fields_ = [("patch", C.c_int32), ("nTypes", C.c_int32), ("typeIds", C.c_int32 * MAX_SPECIES),
                ("numberDensities", C.c_double * MAX_SPECIES), ("velocity", C.c_double * 3),
                ("translationalTemperature", C.c_double), ("rotationalTemperature", C.c_double),
                ("vibrationalTemperature", C.c_double), ("electronicTemperature", C.c_double)]


class Models(C.Structure):
    _fields_ = [("collisionModel", C.c_int32), ("invZvFormulation", C.c_int32), ("Tref", C.c_double),
                ("rotationalRelaxationCollisionNumber", C.c_double), ("vibrationalRelaxationCollisionNumber", C.c_double),
                ("electronicRelaxationCollisionNumber", C.c_double), ("nEquivalentParticles", C.c_double),
                ("deltaT", C.c_double), ("seed", C.c_uint64), ("kB", C.c_double), ("nPatchModels", C.c_int32),
                ("nInflows", C.c_int32), ("patchModels", C.POINTER(PatchModel)), ("inflows", C.POINTER(Inflow)),
                ("measureHeatFluxShearStress", C.c_int32), ("measureClassifications", C.c_int32),
                ("trackFaceFluxes", C.c_int32), ("coordinateSystem", C.c_int32), ("sampleInterval", C.c_int32), ("angularCoordinate", C.c_int32)]


class Reaction(C.Structure):
    _fields_ = [("model", C.c_int32), ("reactants", C.c_int32 * 2), ("allowSplitting", C.c_int32),
                ("dissociationProducts", (C.c_int32 * 2) * 2), ("exchangeProducts", C.c_int32 * 2),
                ("heatOfReactionExchange", C.c_double), ("aCoeff", C.c_double), ("bCoeff", C.c_double)]


REACTION_MODEL_NAMES = {"dissociationQK": 1, "exchangeQK": 2, "dissociationExchangeQK": 3}


def build_reactions(type_id_list, reactions):
    """system/chemReactDict `reactions ( ... )` -> array of dsmcb200_reaction.  `reactions`: list of dicts with the dictionary's keys
    (reactionModel, reactants, allowSplitting, dissociationProducts, exchangeProducts, heatOfReactionExchange, aCoeff, bCoeff);
    species by name.  Unknown model names fail as dsmcReaction::New does (dsmcReaction.C:140-158)."""
    ids = {n: i for i, n in enumerate(type_id_list)}

    def tid(name):
        if name not in ids:
            raise Dsmcb200Error(f"Cannot find type id: {name}")
        return ids[name]

    arr = (Reaction * max(1, len(reactions)))()
    for k, d in enumerate(reactions):
        name = d["reactionModel"]
        if name not in REACTION_MODEL_NAMES:
            raise Dsmcb200Error(f"dsmcReaction::New(const dictionary&) : \n    unknown dsmc reaction model type {name}, constructor not in hash "
                                f"table\n\n    Valid reaction types are :\n{sorted(REACTION_MODEL_NAMES)}")
        r = arr[k]
        r.model = REACTION_MODEL_NAMES[name]
        if len(d["reactants"]) != 2:
            raise Dsmcb200Error("There should be two reactants")
        r.reactants[0], r.reactants[1] = tid(d["reactants"][0]), tid(d["reactants"][1])
        r.allowSplitting = 1 if d.get("allowSplitting", True) else 0
        for i in range(2):
            r.dissociationProducts[i][0] = r.dissociationProducts[i][1] = -1
        r.exchangeProducts[0] = r.exchangeProducts[1] = -1
        if r.model != 2:
            prods = d["dissociationProducts"]
            if len(prods) != 2:
                raise Dsmcb200Error(f"There should be two lists of products, instead of {len(prods)}")
            for i, lst in enumerate(prods):
                if len(lst) not in (0, 2):
                    raise Dsmcb200Error(f"There should be 2 dissociation products instead of {len(lst)}")
                for j, nm in enumerate(lst):
                    r.dissociationProducts[i][j] = tid(nm)
        if r.model != 1:
            ex = d["exchangeProducts"]
            if len(ex) != 2:
                raise Dsmcb200Error(f"There should be two products, instead of {len(ex)}")
            r.exchangeProducts[0], r.exchangeProducts[1] = tid(ex[0]), tid(ex[1])
            r.heatOfReactionExchange = float(d["heatOfReactionExchange"])
            r.aCoeff, r.bCoeff = float(d["aCoeff"]), float(d["bCoeff"])
    arr._n = len(reactions)
    return arr


class ParcelsSoA(C.Structure):
    _fields_ = [("position", C.c_void_p), ("U", C.c_void_p), ("ERot", C.c_void_p), ("cell", C.c_void_p),
                ("tetFace", C.c_void_p), ("tetPt", C.c_void_p), ("typeId", C.c_void_p), ("vibLevel", C.c_void_p),
                ("ELevel", C.c_void_p), ("newParcel", C.c_void_p), ("classification", C.c_void_p), ("origId", C.c_void_p),
                ("maxModes", C.c_int32), ("pad_", C.c_int32), ("origProc", C.c_void_p), ("radialWeight", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [("nParcels", C.c_int64), ("collisions", C.c_int64), ("collisionCandidates", C.c_int64),
                ("trackingRescues", C.c_int64), ("deleted", C.c_int64), ("inserted", C.c_int64),
                ("migratedOut", C.c_int64), ("migratedIn", C.c_int64), ("unsortedLargeCells", C.c_int64),
                ("mass", C.c_double), ("linearKineticEnergy", C.c_double), ("rotationalEnergy", C.c_double),
                ("vibrationalEnergy", C.c_double), ("electronicEnergy", C.c_double), ("stageMs", C.c_double * 8),
                ("nNeighbours", C.c_int32), ("migrationRounds", C.c_int32), ("neighbourProc", C.c_int32 * MAX_NEIGHBOURS),
                ("migratedTo", C.c_int64 * MAX_NEIGHBOURS), ("migratedFrom", C.c_int64 * MAX_NEIGHBOURS),
                ("nMolecules", C.c_double), ("cloned", C.c_int64)]


class AccumInfo(C.Structure):
    _fields_ = [("nCells", C.c_int32), ("nSpecies", C.c_int32), ("nQuantities", C.c_int32), ("nModes", C.c_int32),
                ("nTimeSteps", C.c_double)]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class MeshData:
    """polyMesh arrays (numpy) + patch list; keeps the ctypes struct alive."""

    def __init__(self, points, face_offsets, face_points, owner, neighbour, patches):
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        self.face_offsets = np.ascontiguousarray(face_offsets, dtype=np.int32)
        self.face_points = np.ascontiguousarray(face_points, dtype=np.int32)
        self.owner = np.ascontiguousarray(owner, dtype=np.int32)
        self.neighbour = np.ascontiguousarray(neighbour, dtype=np.int32)
        self.patches = list(patches)  # dicts: name,type,start,size,[neighbPatch,myProcNo,neighbProcNo,separation]
        self.n_cells = int(self.owner.max()) + 1 if len(self.owner) else 0
        self.n_faces = len(self.owner)
        self.n_internal = len(self.neighbour)

    def patch_index(self, name):
        for i, p in enumerate(self.patches):
            if p["name"] == name:
                return i
        raise KeyError(name)

    def as_struct(self):
        arr = (Patch * max(1, len(self.patches)))()
        for i, p in enumerate(self.patches):
            arr[i].name = p["name"].encode()[: NAME_LEN - 1]
            t = p["type"]
            arr[i].type = PATCH_TYPE_NAMES[t] if isinstance(t, str) else int(t)
            arr[i].start = int(p["start"])
            arr[i].size = int(p["size"])
            arr[i].neighbPatch = int(p.get("neighbPatch", -1))
            arr[i].myProcNo = int(p.get("myProcNo", -1))
            arr[i].neighbProcNo = int(p.get("neighbProcNo", -1))
            arr[i].referPatch = int(p.get("referPatch", -1))
            sep = p.get("separation")
            arr[i].hasSeparation = 0 if sep is None else 1
            if sep is not None:
                for d in range(3):
                    arr[i].separation[d] = float(sep[d])
        m = Mesh()
        m.nPoints = len(self.points)
        m.nFaces = self.n_faces
        m.nInternalFaces = self.n_internal
        m.nCells = self.n_cells
        m.nPatches = len(self.patches)
        m.points = _ptr(self.points)
        m.faceOffsets = _ptr(self.face_offsets)
        m.facePoints = _ptr(self.face_points)
        m.owner = _ptr(self.owner)
        m.neighbour = _ptr(self.neighbour)
        m.patches = arr
        self._keep = arr
        return m


def make_species(name, mass, diameter, omega, alpha=1.0, rotationalDegreesOfFreedom=0.0, thetaV=(), Zref=(), TrefZv=(),
                 thetaD=0.0, charge=0, electronicEnergyList=(0.0,), electronicDegeneracyList=(1,)):
    """dsmcParcel::constantProperties from the moleculeProperties keywords (dsmcParcelI.H:37-200)."""
    s = Species()
    s.name = name.encode()[: NAME_LEN - 1]
    s.mass, s.diameter, s.omega, s.alpha = mass, diameter, omega, alpha
    s.rotationalDegreesOfFreedom = rotationalDegreesOfFreedom
    s.nVibrationalModes = len(thetaV)
    if not (len(Zref) == len(thetaV) == len(TrefZv)):
        raise ValueError("Number of characteristic vibrational temperatures / Zref / referenceTempForZref differ")
    for i, v in enumerate(thetaV):
        s.thetaV[i] = v
        s.Zref[i] = Zref[i]
        s.TrefZv[i] = TrefZv[i]
    s.thetaD = thetaD
    s.charge = charge
    s.nElectronicLevels = len(electronicEnergyList)
    for i, v in enumerate(electronicEnergyList):
        s.electronicEnergyList[i] = v
        s.electronicDegeneracyList[i] = electronicDegeneracyList[i]
    return s


class ParcelData:
    """Host SoA of a cloud in the layout of dsmcb200_parcels_soa."""

    FIELDS = [("position", np.float64, 3), ("U", np.float64, 3), ("ERot", np.float64, 1), ("cell", np.int32, 1),
              ("tetFace", np.int32, 1), ("tetPt", np.int32, 1), ("typeId", np.int32, 1), ("vibLevel", np.int32, 0),
              ("ELevel", np.int32, 1), ("newParcel", np.int32, 1), ("classification", np.int32, 1), ("origId", np.int32, 1),
              ("origProc", np.int32, 1), ("radialWeight", np.float64, 1)]

    def __init__(self, n=0, max_modes=1, allocate=True, **arrays):
        self.n = n
        self.max_modes = max_modes
        for name, dt, w in self.FIELDS:
            width = max_modes if name == "vibLevel" else w
            if name in arrays and arrays[name] is not None:
                a = np.ascontiguousarray(arrays[name], dtype=dt)
                setattr(self, name, a)
            elif allocate:
                shape = (n, width) if width > 1 or name == "vibLevel" else (n,)
                setattr(self, name, np.zeros(shape, dtype=dt))
            else:
                setattr(self, name, None)

    def as_struct(self):
        s = ParcelsSoA()
        for name, _, _ in self.FIELDS:
            setattr(s, name, _ptr(getattr(self, name)))
        s.maxModes = self.max_modes
        return s

    def truncated(self, n):
        out = ParcelData(0, self.max_modes, allocate=False)
        out.n = n
        for name, _, _ in self.FIELDS:
            a = getattr(self, name)
            setattr(out, name, None if a is None else a[:n])
        return out


_LIB = None


def lib_path():
    # DSMCB200_LIB selects an alternative build of the same library (kernel tuning experiments)
    return os.environ.get("DSMCB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdsmcb200.so")


def load_library():
    """dlopen libdsmcb200.so; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the dsmcb200 engine)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    P = C.c_void_p
    sigs = {
        "dsmcb200_abi_version": ([], C.c_int),
        "dsmcb200_create": ([C.POINTER(P), C.c_int, C.c_int, C.c_int], C.c_int),
        "dsmcb200_destroy": ([P], None),
        "dsmcb200_last_error": ([P], C.c_char_p),
        "dsmcb200_nccl_unique_id": ([C.c_void_p], C.c_int),
        "dsmcb200_init_comm": ([P, C.c_void_p], C.c_int),
        "dsmcb200_set_mesh": ([P, C.POINTER(Mesh)], C.c_int),
        "dsmcb200_set_species": ([P, C.c_int, C.POINTER(Species)], C.c_int),
        "dsmcb200_set_models": ([P, C.POINTER(Models)], C.c_int),
        "dsmcb200_set_cell_fields": ([P, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "dsmcb200_download_cell_fields": ([P, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "dsmcb200_set_reactions": ([P, C.c_int, C.POINTER(Reaction)], C.c_int),
        "dsmcb200_reaction_counts": ([P, C.c_int, C.c_void_p], C.c_int),
        "dsmcb200_reserve": ([P, C.c_int64], C.c_int),
        "dsmcb200_upload_parcels": ([P, C.c_int64, C.POINTER(ParcelsSoA)], C.c_int),
        "dsmcb200_download_parcels": ([P, C.c_int64, C.POINTER(C.c_int64), C.POINTER(ParcelsSoA)], C.c_int),
        "dsmcb200_upload_cellstate": ([P, C.c_void_p, C.c_void_p], C.c_int),
        "dsmcb200_download_cellstate": ([P, C.c_void_p, C.c_void_p], C.c_int),
        "dsmcb200_mesh_fill": ([P, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p], C.c_int),
        "dsmcb200_zone_fill": ([P, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p], C.c_int),
        "dsmcb200_evolve": ([P, C.c_int], C.c_int),
        "dsmcb200_stage": ([P, C.c_int], C.c_int),
        "dsmcb200_set_step": ([P, C.c_uint32], C.c_int),
        "dsmcb200_download_occupancy": ([P, C.c_void_p], C.c_int),
        "dsmcb200_set_cell_order": ([P, C.c_int, C.c_void_p, C.c_int32], C.c_int),
        "dsmcb200_download_cell_order": ([P, C.c_void_p], C.c_int),
        "dsmcb200_accum_info_get": ([P, C.POINTER(AccumInfo)], C.c_int),
        "dsmcb200_set_sample_sets": ([P, C.c_int, C.c_void_p], C.c_int),
        "dsmcb200_select_sample_set": ([P, C.c_int], C.c_int),
        "dsmcb200_download_accumulators": ([P, C.c_void_p, C.c_void_p], C.c_int),
        "dsmcb200_upload_accumulators": ([P, C.c_void_p, C.c_void_p, C.c_double], C.c_int),
        "dsmcb200_reset_accumulators": ([P], C.c_int),
        "dsmcb200_wall_info": ([P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)], C.c_int),
        "dsmcb200_download_wall_accumulators": ([P, C.c_void_p], C.c_int),
        "dsmcb200_upload_wall_accumulators": ([P, C.c_void_p], C.c_int),
        "dsmcb200_download_face_fluxes": ([P, C.c_void_p, C.c_void_p], C.c_int),
        "dsmcb200_upload_overall_temperature": ([P, C.c_void_p], C.c_int),
        "dsmcb200_get_counters": ([P, C.POINTER(Counters)], C.c_int),
        "dsmcb200_kernel_times": ([P, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "dsmcb200_download_geometry": ([P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p], C.c_int),
        "dsmcb200_allreduce_sum": ([P, C.c_void_p, C.c_int], C.c_int),
        "dsmcb200_allreduce_min": ([P, C.c_void_p, C.c_int], C.c_int),
        "dsmcb200_timer_start": ([P], C.c_int),
        "dsmcb200_timer_stop": ([P, C.POINTER(C.c_float)], C.c_int),
    }
    for name, (args, res) in sigs.items():
        if os.environ.get("DSMCB200_LIB") and not hasattr(lib, name):
            continue   # an older build selected for an A/B run: it simply lacks the newer entry points
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.argtypes = args
        fn.restype = res
    _LIB = lib
    return lib


EXPORTED_SYMBOLS = [
    "dsmcb200_abi_version", "dsmcb200_create", "dsmcb200_destroy", "dsmcb200_last_error", "dsmcb200_nccl_unique_id",
    "dsmcb200_init_comm", "dsmcb200_set_mesh", "dsmcb200_set_species", "dsmcb200_set_models", "dsmcb200_set_reactions",
    "dsmcb200_reaction_counts", "dsmcb200_set_cell_fields", "dsmcb200_download_cell_fields", "dsmcb200_reserve",
    "dsmcb200_upload_parcels", "dsmcb200_download_parcels", "dsmcb200_upload_cellstate", "dsmcb200_download_cellstate",
    "dsmcb200_mesh_fill", "dsmcb200_zone_fill", "dsmcb200_evolve", "dsmcb200_stage", "dsmcb200_set_step", "dsmcb200_download_occupancy",
    "dsmcb200_set_cell_order", "dsmcb200_download_cell_order", "dsmcb200_set_sample_sets", "dsmcb200_select_sample_set",
    "dsmcb200_accum_info_get", "dsmcb200_download_accumulators", "dsmcb200_upload_accumulators",
    "dsmcb200_reset_accumulators", "dsmcb200_wall_info", "dsmcb200_download_wall_accumulators",
    "dsmcb200_upload_wall_accumulators", "dsmcb200_download_face_fluxes", "dsmcb200_upload_overall_temperature", "dsmcb200_get_counters",
    "dsmcb200_kernel_times", "dsmcb200_download_geometry", "dsmcb200_timer_start", "dsmcb200_timer_stop", "dsmcb200_allreduce_sum", "dsmcb200_allreduce_min",
]


class Dsmcb200Error(RuntimeError):
    pass


def build_models(collisionModel="VariableHardSphere", nEquivalentParticles=1.0, deltaT=1e-6, seed=1, Tref=273.0,
                 rotationalRelaxationCollisionNumber=5.0, vibrationalRelaxationCollisionNumber=0.0,
                 electronicRelaxationCollisionNumber=500.0, inverseZvFormulation="", kB=0.0, patch_models=(), inflows=(),
                 measureHeatFluxShearStress=False, measureClassifications=False, sampleInterval=1, trackFaceFluxes=False,
                 coordinateSystem="dsmcCartesian", angularCoordinate=2):
    """POD form of constant/dsmcProperties + boundariesDict.  Unknown model names raise with the
    reference's 'Valid ... types are' message shape (BinaryCollisionModel.C:70-85)."""
    if collisionModel not in COLLISION_MODEL_NAMES:
        raise Dsmcb200Error(f"BinaryCollisionModel::New(const dictionary&, CloudType&) : \n    unknown BinaryCollisionModelType type "
                            f"{collisionModel}, constructor not in hash table\n\n    Valid BinaryCollisionModel types are :\n"
                            f"{sorted(COLLISION_MODEL_NAMES)}")
    if coordinateSystem not in COORDINATE_SYSTEM_NAMES:
        raise Dsmcb200Error(f"dsmcCoordinateSystem::New(const dictionary&) : \n    unknown dsmcCoordinateSystem type {coordinateSystem}, "
                            f"constructor not in hash table\n\n    Valid coordinate system types are :\n{sorted(COORDINATE_SYSTEM_NAMES)}")
    m = Models()
    m.coordinateSystem = COORDINATE_SYSTEM_NAMES[coordinateSystem]
    m.angularCoordinate = int(angularCoordinate)
    m.collisionModel = COLLISION_MODEL_NAMES[collisionModel]
    m.invZvFormulation = {"pre-2008": 0, "2008": 1}.get(inverseZvFormulation, 2)
    m.Tref = Tref
    m.rotationalRelaxationCollisionNumber = rotationalRelaxationCollisionNumber
    m.vibrationalRelaxationCollisionNumber = vibrationalRelaxationCollisionNumber
    m.electronicRelaxationCollisionNumber = electronicRelaxationCollisionNumber
    m.nEquivalentParticles = nEquivalentParticles
    m.deltaT = deltaT
    m.seed = seed
    m.kB = kB
    pm = (PatchModel * max(1, len(patch_models)))()
    for i, d in enumerate(patch_models):
        name = d["boundaryModel"]
        if name not in PATCH_MODEL_NAMES:
            raise Dsmcb200Error(f"dsmcPatchBoundary::New(const dictionary&) : \n    unknown dsmcPatchBoundary type {name}, "
                                f"constructor not in hash table\n\n    Valid patch boundary types are :\n{sorted(PATCH_MODEL_NAMES)}")
        pm[i].patch = d["patch"]
        pm[i].model = PATCH_MODEL_NAMES[name]
        pm[i].temperature = d.get("temperature", 0.0)
        pm[i].diffuseFraction = d.get("diffuseFraction", 0.0)
        for key in ("normalAccommodationCoefficient", "tangentialAccommodationCoefficient", "rotationalEnergyAccommodationCoefficient"):
            setattr(pm[i], key, d.get(key, 0.0))       # dsmcCLLWallPatchProperties
        if "formationLevelTemperature" in d:   # linear T(depth), dsmcDiffuseWallPatch.C:141-148
            pm[i].linearTemperature = 1
            pm[i].formationLevelTemperature = d["formationLevelTemperature"]
            pm[i].depthAxis = {"x": 0, "y": 1, "z": 2}[d.get("depthAxis", "y")]
        for k in range(3):
            pm[i].velocity[k] = d.get("velocity", (0.0, 0.0, 0.0))[k]
    inf = (Inflow * max(1, len(inflows)))()
    for i, d in enumerate(inflows):
        inf[i].patch = d["patch"]
        inf[i].nTypes = len(d["typeIds"])
        for k, t in enumerate(d["typeIds"]):
            inf[i].typeIds[k] = t
            inf[i].numberDensities[k] = d["numberDensities"][k]
        for k in range(3):
            inf[i].velocity[k] = d["velocity"][k]
        inf[i].translationalTemperature = d["translationalTemperature"]
        inf[i].rotationalTemperature = d.get("rotationalTemperature", 0.0)
        inf[i].vibrationalTemperature = d.get("vibrationalTemperature", 0.0)
        inf[i].electronicTemperature = d.get("electronicTemperature", 0.0)
    m.sampleInterval = int(sampleInterval)
    m.trackFaceFluxes = int(bool(trackFaceFluxes))
    m.nPatchModels = len(patch_models)
    m.nInflows = len(inflows)
    m.patchModels = pm
    m.inflows = inf
    m.measureHeatFluxShearStress = int(measureHeatFluxShearStress)
    m.measureClassifications = int(measureClassifications)
    m._keep = (pm, inf)
    return m


COORDINATE_SYSTEM_NAMES = {"dsmcCartesian": 0, "dsmcAxisymmetric": 1, "dsmcSpherical": 2}


def axisymmetric_axes(revolutionAxis="", polarAxis=""):
    """(revolutionAxis, polarAxis, angularCoordinate) as component labels from the axisymmetricProperties keywords
    (dsmcAxisymmetric::checkCoordinateSystemInputs, dsmcAxisymmetric.C:337-420; defaults x, y, z)."""
    rev, pol, ang = 0, 1, 2
    bad = Dsmcb200Error("Revolution and polar axes are badly defined in constant/dsmcProperties axisymmetricProperties{}")
    if revolutionAxis == "z":
        rev = 2
        if polarAxis == "":
            pol, ang = 0, 1
        elif polarAxis == "y":
            pol, ang = 1, 0
        elif polarAxis == "x":
            pol, ang = 0, 1
        else:
            raise bad
    elif revolutionAxis == "y":
        rev = 1
        if polarAxis == "":
            pol, ang = 2, 0
        elif polarAxis == "x":
            pol, ang = 0, 2
        elif polarAxis == "z":
            pol, ang = 2, 0
        else:
            raise bad
    elif revolutionAxis == "x":
        if polarAxis == "z":
            pol, ang = 2, 1
        elif polarAxis != "y":
            raise bad
    return rev, pol, ang


def axisymmetric_rwf(cell_centres, face_centres, polar_axis, max_rwf, radial_extent=None):
    """dsmcAxisymmetric::recalculateRWF, radial weighting method "cell" (dsmcAxisymmetric.C:236-275): RWF = 1 + (maxRWF - 1) r / radialExtent
    with r = |cell centre . polar axis| and radialExtent = gMax of the face centres' polar component (:447-456) -- a global maximum:
    on a decomposed mesh pass the extent of the whole domain."""
    fc = np.asarray(face_centres)[:, polar_axis]
    if radial_extent is None:
        radial_extent = fc.max()
        if not radial_extent > 0:
            radial_extent = -fc.min()
    rwf = np.ones(len(cell_centres))
    rwf += (max_rwf - 1.0) * np.abs(np.asarray(cell_centres)[:, polar_axis]) / radial_extent
    return rwf, radial_extent


def spherical_rwf(cell_centres, face_centres, origin, max_rwf, radial_extent=None):
    """dsmcSpherical::recalculateRWF, radial weighting method "cell" (dsmcSpherical.C:232-275): RWF = 1 + (maxRWF - 1) (r / radialExtent)^2
    with r the distance of the cell centre from `origin` and radialExtent = gMax of the face centres' distances (:355-366)."""
    o = np.asarray(origin, float)
    d = np.asarray(face_centres) - o
    if radial_extent is None:
        radial_extent = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2 + d[:, 2] ** 2).max()
    c = np.asarray(cell_centres) - o
    radius = np.sqrt(c[:, 0] ** 2 + c[:, 1] ** 2 + c[:, 2] ** 2)
    rwf = np.ones(len(c))
    rwf += (max_rwf - 1.0) * (radius / radial_extent) ** 2
    return rwf, radial_extent


def variable_time_step(cell_volumes, n_equivalent_particles, delta_t):
    """dsmcVariableTimeStepModel (variableTimeStepModel/dsmcVariableTimeStepModel.C:48-100): the smallest cell keeps nEquivalentParticles
    and deltaT, every other cell scales both with its volume, so that nParticles / deltaT is uniform and fluxes are conserved."""
    V = np.asarray(cell_volumes, float)
    ref = int(np.nonzero(np.abs(V - V.min()) < 1e-15)[0][0])
    n = n_equivalent_particles * V / V[ref]
    ratio = n[ref] / delta_t
    return n, n / ratio


class Engine:
    """One dsmcb200 context (one GPU / rank)."""

    def __init__(self, device=0, rank=0, n_ranks=1):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.dsmcb200_create(C.byref(h), device, rank, n_ranks)
        if rc != 0:
            raise Dsmcb200Error(f"dsmcb200_create failed ({rc}): a CUDA device is required; there is no CPU fallback")
        self.h = h
        self.max_modes = 1
        self._mesh = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.dsmcb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise Dsmcb200Error(f"dsmcb200 error {rc}: {self.lib.dsmcb200_last_error(self.h).decode()}")

    CELL_ORDER = {"as-given": 0, "z-curve": 1, "given": 2}

    def set_cell_order(self, mode="z-curve", new_of_old=None):
        """Cell labels inside the engine (dsmcb200_set_cell_order; before set_mesh): "as-given", "z-curve" or "given" with a table."""
        if new_of_old is not None:
            t = np.ascontiguousarray(new_of_old, dtype=np.int32)
            self._ck(self.lib.dsmcb200_set_cell_order(self.h, self.CELL_ORDER["given"], _ptr(t), len(t)))
        else:
            self._ck(self.lib.dsmcb200_set_cell_order(self.h, self.CELL_ORDER[mode], None, 0))

    def set_sample_sets(self, intervals):
        """One set of sums per distinct sampleInterval of the case's field{} entries (dsmcb200_set_sample_sets; before the engine is finalised)."""
        t = np.ascontiguousarray(intervals, dtype=np.int32)
        self._ck(self.lib.dsmcb200_set_sample_sets(self.h, len(t), _ptr(t)))

    def select_sample_set(self, k):
        """The set accumulators() / wall_accumulators() / reset_accumulators() / upload_accumulators() act on."""
        self._ck(self.lib.dsmcb200_select_sample_set(self.h, int(k)))

    def cell_order(self):
        """new_of_old: the engine's label of the caller's cell k."""
        t = np.zeros(self._mesh.n_cells, np.int32)
        self._ck(self.lib.dsmcb200_download_cell_order(self.h, _ptr(t)))
        return t

    def set_mesh(self, mesh: MeshData):
        self._mesh = mesh
        st = mesh.as_struct()
        self._ck(self.lib.dsmcb200_set_mesh(self.h, C.byref(st)))

    def set_species(self, species):
        arr = (Species * len(species))(*species)
        self.n_species = len(species)
        self.max_modes = max(1, max(s.nVibrationalModes for s in species))
        self._ck(self.lib.dsmcb200_set_species(self.h, len(species), arr))

    def set_models(self, models: Models):
        self._models = models
        self._ck(self.lib.dsmcb200_set_models(self.h, C.byref(models)))

    def set_cell_fields(self, nParticles=None, deltaT=None, RWF=None):
        """per-cell nParticles (time-step model), deltaT and radial weighting factor; None = uniform"""
        a = [None if x is None else np.ascontiguousarray(x, np.float64) for x in (nParticles, deltaT, RWF)]
        self._ck(self.lib.dsmcb200_set_cell_fields(self.h, *[None if x is None else _ptr(x) for x in a]))

    def cell_fields(self):
        n = self.accum_info().nCells
        out = [np.zeros(n) for _ in range(3)]
        self._ck(self.lib.dsmcb200_download_cell_fields(self.h, *[_ptr(x) for x in out]))
        return out

    def set_reactions(self, reactions):
        """reactions: the array of build_reactions()"""
        self._reactions = reactions
        self._ck(self.lib.dsmcb200_set_reactions(self.h, reactions._n, reactions))

    def reaction_counts(self):
        """[nReactions][3]: dissociations of reactant 0, of reactant 1, exchanges -- totals since set_reactions"""
        n = self._reactions._n
        out = np.zeros((max(n, 1), 3), np.int64)
        self._ck(self.lib.dsmcb200_reaction_counts(self.h, n, _ptr(out)))
        return out[:n]

    def init_comm(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.dsmcb200_init_comm(self.h, buf))

    def reserve(self, n):
        self._ck(self.lib.dsmcb200_reserve(self.h, int(n)))

    def upload_parcels(self, p: ParcelData):
        st = p.as_struct()
        self._ck(self.lib.dsmcb200_upload_parcels(self.h, p.n, C.byref(st)))

    def num_parcels(self):
        n = C.c_int64()
        self._ck(self.lib.dsmcb200_download_parcels(self.h, 0, C.byref(n), None))
        return n.value

    def download_parcels(self, out: ParcelData | None = None):
        n = self.num_parcels()
        if out is None:
            out = ParcelData(n, self.max_modes)
        st = out.as_struct()
        nn = C.c_int64()
        self._ck(self.lib.dsmcb200_download_parcels(self.h, out.n if out.position is None else len(out.position), C.byref(nn), C.byref(st)))
        return out.truncated(nn.value) if nn.value != out.n else out

    def upload_cellstate(self, sigma=None, remainder=None):
        sigma = None if sigma is None else np.ascontiguousarray(sigma, np.float64)
        remainder = None if remainder is None else np.ascontiguousarray(remainder, np.float64)
        self._ck(self.lib.dsmcb200_upload_cellstate(self.h, _ptr(sigma), _ptr(remainder)))

    def download_cellstate(self):
        n = self._mesh.n_cells
        s, r = np.zeros(n), np.zeros(n)
        self._ck(self.lib.dsmcb200_download_cellstate(self.h, _ptr(s), _ptr(r)))
        return s, r

    def mesh_fill(self, type_ids, number_densities, Ttra, Trot=0.0, Tvib=0.0, Telec=0.0, velocity=(0.0, 0.0, 0.0)):
        t = np.ascontiguousarray(type_ids, np.int32)
        nd = np.ascontiguousarray(number_densities, np.float64)
        v = np.ascontiguousarray(velocity, np.float64)
        self._ck(self.lib.dsmcb200_mesh_fill(self.h, len(t), _ptr(t), _ptr(nd), Ttra, Trot, Tvib, Telec, _ptr(v)))

    def zone_fill(self, zone_cells, type_ids, number_densities, Ttra, Trot=0.0, Tvib=0.0, Telec=0.0, velocity=(0.0, 0.0, 0.0)):
        """dsmcZoneFill: the mesh fill for the cells of one cellZone, appended to the cloud."""
        z = np.ascontiguousarray(zone_cells, np.int32)
        t = np.ascontiguousarray(type_ids, np.int32)
        nd = np.ascontiguousarray(number_densities, np.float64)
        v = np.ascontiguousarray(velocity, np.float64)
        self._ck(self.lib.dsmcb200_zone_fill(self.h, C.c_int64(len(z)), _ptr(z), len(t), _ptr(t), _ptr(nd), Ttra, Trot, Tvib, Telec, _ptr(v)))

    def evolve(self, n_steps=1):
        self._ck(self.lib.dsmcb200_evolve(self.h, n_steps))

    def stage(self, stage):
        self._ck(self.lib.dsmcb200_stage(self.h, stage))

    def set_step(self, step):
        self._ck(self.lib.dsmcb200_set_step(self.h, step))

    def occupancy(self):
        off = np.zeros(self._mesh.n_cells + 1, np.int32)
        self._ck(self.lib.dsmcb200_download_occupancy(self.h, _ptr(off)))
        return off

    def accum_info(self):
        i = AccumInfo()
        self._ck(self.lib.dsmcb200_accum_info_get(self.h, C.byref(i)))
        return i

    def accumulators(self, acc=None, coll=None):
        """dsmcb200_download_accumulators; acc / coll may be caller-owned (e.g. pinned) float64 arrays of the right size."""
        i = self.accum_info()
        if acc is None:
            acc = np.zeros((i.nCells, i.nSpecies, i.nQuantities))
        if coll is None:
            coll = np.zeros((i.nCells, 2))
        assert acc.dtype == np.float64 and acc.size == i.nCells * i.nSpecies * i.nQuantities and acc.flags.c_contiguous
        assert coll.dtype == np.float64 and coll.size == i.nCells * 2 and coll.flags.c_contiguous
        self._ck(self.lib.dsmcb200_download_accumulators(self.h, _ptr(acc), _ptr(coll)))
        return acc, coll, i.nTimeSteps

    def upload_accumulators(self, acc, coll, n_time_steps):
        acc = np.ascontiguousarray(acc, np.float64)
        coll = np.ascontiguousarray(coll, np.float64)
        self._ck(self.lib.dsmcb200_upload_accumulators(self.h, _ptr(acc), _ptr(coll), float(n_time_steps)))

    def reset_accumulators(self):
        self._ck(self.lib.dsmcb200_reset_accumulators(self.h))

    def wall_accumulators(self):
        nf, nq = C.c_int32(), C.c_int32()
        self._ck(self.lib.dsmcb200_wall_info(self.h, C.byref(nf), C.byref(nq)))
        w = np.zeros((nf.value, self.n_species, nq.value))
        if nf.value:
            self._ck(self.lib.dsmcb200_download_wall_accumulators(self.h, _ptr(w)))
        return w

    def upload_overall_temperature(self, Tov):
        """fields().overallT(cell) for inverseZvFormulation "2008" (dsmcCloud.C:1441-1456)."""
        t = np.ascontiguousarray(Tov, np.float64)
        assert t.shape == (self._mesh.n_cells,)
        self._ck(self.lib.dsmcb200_upload_overall_temperature(self.h, _ptr(t)))

    def face_fluxes(self):
        """dsmcFaceTracker parcelIdFlux / massIdFlux of the last step, each [nSpecies][nFaces] (models.trackFaceFluxes)."""
        pf, mf = np.zeros((self.n_species, self._mesh.n_faces)), np.zeros((self.n_species, self._mesh.n_faces))
        self._ck(self.lib.dsmcb200_download_face_fluxes(self.h, _ptr(pf), _ptr(mf)))
        return pf, mf

    def counters(self):
        c = Counters()
        self._ck(self.lib.dsmcb200_get_counters(self.h, C.byref(c)))
        return c

    def kernel_times(self, reset=False):
        cap = 64
        names = C.create_string_buffer(cap * NAME_LEN)
        ms = (C.c_float * cap)()
        launches = (C.c_int64 * cap)()
        n = C.c_int()
        self._ck(self.lib.dsmcb200_kernel_times(self.h, cap, C.byref(n), names, ms, launches))
        out = {}
        for k in range(n.value):
            nm = names.raw[k * NAME_LEN:(k + 1) * NAME_LEN].split(b"\0")[0].decode()
            out[nm] = (float(ms[k]), int(launches[k]))
        if reset:
            self._ck(self.lib.dsmcb200_kernel_times(self.h, 0, C.byref(n), None, None, None))
        return out

    def timer_start(self):
        self._ck(self.lib.dsmcb200_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.lib.dsmcb200_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def geometry(self):
        m = self._mesh
        cc, cv = np.zeros((m.n_cells, 3)), np.zeros(m.n_cells)
        fc, fa = np.zeros((m.n_faces, 3)), np.zeros((m.n_faces, 3))
        tb = np.zeros(m.n_faces, np.int32)
        self._ck(self.lib.dsmcb200_download_geometry(self.h, _ptr(cc), _ptr(cv), _ptr(fc), _ptr(fa), _ptr(tb)))
        return cc, cv, fc, fa, tb


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.dsmcb200_nccl_unique_id(buf)
    if rc != 0:
        raise Dsmcb200Error("dsmcb200_nccl_unique_id failed (libnccl.so.2 not loadable?)")
    return buf.raw
