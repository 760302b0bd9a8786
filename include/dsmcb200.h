/*
 * dsmcb200.h -- C ABI of libdsmcb200.so, the B200-native engine behind
 * dsmcFoam+'s per-timestep particle loop (dsmcCloud::evolve()).
 *
 * The reference (hyStrath, OpenFOAM v1706 add-on) has no FFI of its own: its
 * plugin surface is OpenFOAM run-time selection.  Each
 * entry point below therefore cites the reference member function whose work
 * it takes over; the OpenFOAM-side shim in INTEGRATION.md is the binding a
 * maintainer adds to call them.  Paths are relative to the reference root;
 * DSMC/ = src/lagrangian/dsmc/, BASIC/ = src/lagrangian/basic/.
 *
 * Conventions
 *  - plain C, POD structs, raw pointers + sizes; no C++/torch types.
 *  - caller owns every host buffer, the library copies; the library owns all
 *    device memory; download buffers are caller-allocated with capacity
 *    passed in.
 *  - every call returns 0 on success or a negative dsmcb200_status; the text
 *    is available through dsmcb200_last_error().  Nothing throws or exits
 *    across this boundary (the shim turns errors into FatalErrorIn, cf.
 *    DSMC/clouds/dsmcCloudI.H:141-147).
 *  - one ctx per GPU / rank; a ctx is not re-entrant; calls are synchronous
 *    on return unless stated otherwise.
 *  - label -> int32_t (reference build is Int32), scalar -> double, SI units.
 *  - there is NO CPU fallback: dsmcb200_create fails if no CUDA device.
 */
#ifndef DSMCB200_H
#define DSMCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSMCB200_ABI_VERSION 6   /* 6 = 5 + dsmcb200_zone_fill (dsmcZoneFill);
                                    5 = 4 + the three accommodation coefficients at the end of dsmcb200_patch_model (dsmcCLLWallPatch), COORD_SPHERICAL;
                                    4 = 3 + set_cell_order / download_cell_order, set_sample_sets / select_sample_set, allreduce_min */
#define DSMCB200_MAX_NEIGHBOURS 16
#define DSMCB200_MAX_SPECIES 8
#define DSMCB200_MAX_VIB_MODES 3
#define DSMCB200_MAX_ELEC_LEVELS 16
#define DSMCB200_NAME_LEN 64

typedef struct dsmcb200_ctx dsmcb200_ctx;

typedef enum {
    DSMCB200_OK = 0,
    DSMCB200_ERR_INVALID = -1,   /* bad argument / inconsistent input      */
    DSMCB200_ERR_CUDA = -2,      /* CUDA runtime failure                   */
    DSMCB200_ERR_STATE = -3,     /* call order violated (e.g. no mesh yet) */
    DSMCB200_ERR_CAPACITY = -4,  /* a fixed-capacity buffer overflowed     */
    DSMCB200_ERR_UNSUPPORTED = -5, /* model name outside the scoped path   */
    DSMCB200_ERR_NCCL = -6
} dsmcb200_status;

/* polyPatch types dispatched by particle::trackToFace
 * (BASIC/particle/particleTemplates.C:1110-1164). */
typedef enum {
    DSMCB200_PATCH_WALL = 0,
    DSMCB200_PATCH_PATCH = 1,
    DSMCB200_PATCH_CYCLIC = 2,
    DSMCB200_PATCH_PROCESSOR = 3,
    DSMCB200_PATCH_EMPTY = 4,
    DSMCB200_PATCH_SYMMETRYPLANE = 5,
    DSMCB200_PATCH_SYMMETRY = 6,
    DSMCB200_PATCH_WEDGE = 7,
    DSMCB200_PATCH_PROCESSORCYCLIC = 8
} dsmcb200_patch_type;

typedef struct {
    char name[DSMCB200_NAME_LEN];
    int32_t type;        /* dsmcb200_patch_type                                  */
    int32_t start;       /* first face (global face index)                       */
    int32_t size;        /* nFaces                                               */
    int32_t neighbPatch; /* cyclic: index of neighbourPatch; else -1             */
    int32_t myProcNo;    /* processor patches; else -1                           */
    int32_t neighbProcNo;
    int32_t referPatch;  /* processorCyclic: the cyclic patch it derives from    */
    int32_t hasSeparation; /* 1: separation[] below is authoritative; 0: cyclic
                              patches get (nf & (Cr - Cf))*nf from the geometry   */
    double separation[3];  /* coupledPolyPatch::separation() of THIS patch: a
                              parcel arriving on it has position -= separation    */
} dsmcb200_patch;

/* polyMesh as read from constant/polyMesh/{points,faces,owner,neighbour,boundary}. */
typedef struct {
    int32_t nPoints, nFaces, nInternalFaces, nCells, nPatches;
    const double* points;        /* [nPoints*3]                    */
    const int32_t* faceOffsets;  /* [nFaces+1] CSR into facePoints */
    const int32_t* facePoints;
    const int32_t* owner;        /* [nFaces]                       */
    const int32_t* neighbour;    /* [nInternalFaces]               */
    const dsmcb200_patch* patches;
    /* Optional precomputed geometry (NULL -> computed by the library with the
     * primitiveMesh algorithms).  The OpenFOAM shim passes
     * OpenFOAM's own arrays so geometry is identical by construction. */
    const double* cellCentres;   /* [nCells*3] */
    const double* cellVolumes;   /* [nCells]   */
    const double* faceCentres;   /* [nFaces*3] */
    const double* faceAreas;     /* [nFaces*3] */
    const int32_t* tetBasePtIs;  /* [nFaces]   */
} dsmcb200_mesh;

/* dsmcParcel::constantProperties (DSMC/parcels/dsmcParcel.H:72-236,
 * parsed at DSMC/parcels/dsmcParcelI.H:37-200). */
typedef struct {
    char name[DSMCB200_NAME_LEN];
    double mass, diameter, omega, alpha;
    double rotationalDegreesOfFreedom;
    int32_t nVibrationalModes;
    int32_t charge;
    double thetaV[DSMCB200_MAX_VIB_MODES];   /* characteristicVibrationalTemperature */
    double Zref[DSMCB200_MAX_VIB_MODES];
    double TrefZv[DSMCB200_MAX_VIB_MODES];   /* referenceTempForZref                 */
    double thetaD;                            /* dissociationTemperature              */
    int32_t nElectronicLevels;
    int32_t pad_;
    double electronicEnergyList[DSMCB200_MAX_ELEC_LEVELS];
    int32_t electronicDegeneracyList[DSMCB200_MAX_ELEC_LEVELS];
} dsmcb200_species;

typedef enum {
    DSMCB200_COLL_NONE = 0,                /* NoBinaryCollision                  */
    DSMCB200_COLL_VHS = 1,                 /* VariableHardSphere                 */
    DSMCB200_COLL_LB_VHS = 2,              /* LarsenBorgnakkeVariableHardSphere  */
    DSMCB200_COLL_VSS = 3,                 /* VariableSoftSphere (collisions/derived/VariableSoftSphere/VariableSoftSphere.C:78-262) */
    DSMCB200_COLL_LB_VSS = 4               /* LarsenBorgnakkeVariableSoftSphere  */
} dsmcb200_collision_model;

typedef enum {
    DSMCB200_BND_NONE = 0,
    DSMCB200_BND_DIFFUSE_WALL = 1,   /* dsmcDiffuseWallPatch  */
    DSMCB200_BND_SPECULAR_WALL = 2,  /* dsmcSpecularWallPatch */
    DSMCB200_BND_DELETION = 3,       /* dsmcDeletionPatch     */
    DSMCB200_BND_DIFFUSE_SPECULAR_WALL = 4, /* mixed/dsmcDiffuseSpecularWallPatch (Maxwell model: diffuse with probability diffuseFraction) */
    DSMCB200_BND_CLL_WALL = 5               /* dsmcCLLWallPatch (Cercignani-Lampis-Lord kernel, Lord's extension to rotation) */
} dsmcb200_patch_model_kind;

/* One entry of system/boundariesDict dsmcPatchBoundaries
 * (DSMC/boundaries/basic/dsmcPatchBoundary/dsmcPatchBoundary.C:56-125). */
typedef struct {
    int32_t patch;          /* index into mesh patches */
    int32_t model;          /* dsmcb200_patch_model_kind */
    double temperature;     /* dsmcDiffuseWallPatchProperties.temperature */
    double velocity[3];
    double diffuseFraction; /* dsmcDiffuseSpecularWallPatchProperties.diffuseFraction (dsmcDiffuseSpecularWallPatch.C:67) */
    /* dsmcDiffuseWallPatch::getLocalTemperature (dsmcDiffuseWallPatch.C:141-148): T(x) = temperature + (x[depthAxis] - maxDepth) *
     * (temperature - formationLevelTemperature) / lengthPatch with maxDepth / lengthPatch from the (rank-local) mesh bounds (:181-184).
     * `temperature` above is groundLevelTemperature when that key is given (:53-60).  linearTemperature = 0: the uniform wall
     * (formationLevelTemperature defaults to temperature, :62-63). */
    int32_t linearTemperature;
    int32_t depthAxis;      /* 0 x, 1 y (default), 2 z (:169-179) */
    double formationLevelTemperature;
    /* dsmcCLLWallPatchProperties (dsmcCLLWallPatch.C:57-62): normal / tangential / rotational-energy accommodation coefficients */
    double normalAccommodationCoefficient, tangentialAccommodationCoefficient, rotationalEnergyAccommodationCoefficient;
} dsmcb200_patch_model;

/* One dsmcFreeStreamInflowPatch of dsmcGeneralBoundaries
 * (DSMC/boundaries/derived/generalBoundaries/dsmcFreeStreamInflowPatch/
 *  dsmcFreeStreamInflowPatch.C:393-470). */
typedef struct {
    int32_t patch;
    int32_t nTypes;
    int32_t typeIds[DSMCB200_MAX_SPECIES];
    double numberDensities[DSMCB200_MAX_SPECIES];
    double velocity[3];
    double translationalTemperature, rotationalTemperature;
    double vibrationalTemperature, electronicTemperature;
} dsmcb200_inflow;

/* constant/dsmcProperties + system/controlDict + system/boundariesDict,
 * reduced to POD (DSMC/clouds/dsmcCloud.C:586-691). */
typedef struct {
    int32_t collisionModel;        /* dsmcb200_collision_model                    */
    int32_t invZvFormulation;      /* "pre-2008"->0, "2008"->1, default 2         */
    double Tref;                   /* VariableHardSphereCoeffs.Tref (273)         */
    double rotationalRelaxationCollisionNumber;   /* 5   */
    double vibrationalRelaxationCollisionNumber;  /* 0 -> variable Zv */
    double electronicRelaxationCollisionNumber;   /* 500 */
    double nEquivalentParticles;
    double deltaT;
    uint64_t seed;                 /* Philox key (dsmcProperties seedNumber)      */
    double kB;                     /* physicoChemical::k; 0 -> 1.38065e-23 (v1706)*/
    int32_t nPatchModels;
    int32_t nInflows;
    const dsmcb200_patch_model* patchModels;
    const dsmcb200_inflow* inflows;
    int32_t measureHeatFluxShearStress; /* sample the optional 2nd-moment set     */
    int32_t measureClassifications;
    int32_t trackFaceFluxes;       /* dsmcFaceTracker counters (off by default), see dsmcb200_download_face_fluxes */
    int32_t coordinateSystem;      /* dsmcb200_coordinate_system: dsmcProperties `coordinateSystem` (dsmcCoordinateSystem.C:81) */
    int32_t sampleInterval;        /* dsmcVolFieldsProperties.sampleInterval: stage 5 runs every n-th step (0, 1: every step;
                                      dsmcVolFields.C:1073-1081,1362)             */
    int32_t angularCoordinate;     /* dsmcAxisymmetric: the component of U a cloned parcel gets mirrored in (0, 1, 2;
                                      dsmcAxisymmetric.C:82, 337-420)             */
} dsmcb200_models;

/* dsmcProperties `coordinateSystem`:
 *   dsmcCartesian      DSMC/coordinateSystem/derived/Cartesian/dsmcCartesian.C        uniform weights, 2-D tracks constrained to the mesh centre
 *   dsmcAxisymmetric   DSMC/coordinateSystem/derived/axisymmetric/dsmcAxisymmetric.C  radial weighting factors: parcels carry the RWF of the
 *                      cell they started the step in; after the move a parcel is cloned / deleted with the ratio of old and new RWF
 *                      (axisymmetricWeighting, :50-209) and the occupancy is rebuilt.  The per-cell RWF comes from
 *                      dsmcb200_set_cell_fields (the caller evaluates recalculateRWF, :236-275). */
typedef enum {
    DSMCB200_COORD_CARTESIAN = 0,
    DSMCB200_COORD_AXISYMMETRIC = 1,
    DSMCB200_COORD_SPHERICAL = 2    /* dsmcSpherical: the same weighting stage, clones keep their velocity (dsmcSpherical.C:50-216) */
} dsmcb200_coordinate_system;

/* One entry of system/chemReactDict `reactions ( name { reactionModel M; reactants (A B); allowSplitting yes; ... } )`
 * (DSMC/reactions/basic/dsmcReaction/dsmcReaction.C:79-120).  Quantum-kinetic models:
 *   dissociationQK          DSMC/reactions/derived/dissociationQK/dissociationQK.C         AB + M -> A + B + M
 *   exchangeQK              DSMC/reactions/derived/exchangeQK/exchangeQK.C                 AB + C -> A + BC
 *   dissociationExchangeQK  DSMC/reactions/derived/mixed/dissociationExchangeQK/...C       both, competing */
typedef enum {
    DSMCB200_REACT_DISSOCIATION_QK = 1,
    DSMCB200_REACT_EXCHANGE_QK = 2,
    DSMCB200_REACT_DISSOCIATION_EXCHANGE_QK = 3
} dsmcb200_reaction_model;

typedef struct {
    int32_t model;                       /* dsmcb200_reaction_model                                          */
    int32_t reactants[2];                /* typeIds, in dictionary order                                     */
    int32_t allowSplitting;              /* dsmcReaction.C:96 (default yes); no: reactions are only counted  */
    int32_t dissociationProducts[2][2];  /* dissociationQKProperties.dissociationProducts ((a b) (c d)); -1 -1 for an empty list */
    int32_t exchangeProducts[2];         /* exchangeQKProperties.exchangeProducts, as listed                 */
    double heatOfReactionExchange;       /* Kelvin, > 0 exothermic (exchangeQK.C:363-367)                    */
    double aCoeff, bCoeff;               /* activation-energy coefficients (exchangeQK.C:196-206)            */
} dsmcb200_reaction;

/* Host-side SoA view of the cloud.  Layout of vectors is OpenFOAM's
 * (x y z) interleaved.  Optional arrays may be NULL on upload (defaults:
 * ERot 0, vibLevel 0, ELevel 0, newParcel -1, classification 0, origId = i,
 * stepFraction 0). */
typedef struct {
    double* position;      /* [3n] */
    double* U;             /* [3n] */
    double* ERot;          /* [n]  */
    int32_t* cell;         /* [n]  */
    int32_t* tetFace;      /* [n]  */
    int32_t* tetPt;        /* [n]  */
    int32_t* typeId;       /* [n]  */
    int32_t* vibLevel;     /* [n*maxModes], parcel-major */
    int32_t* ELevel;       /* [n]  */
    int32_t* newParcel;    /* [n]  */
    int32_t* classification; /* [n] */
    int32_t* origId;       /* [n]  */
    int32_t maxModes;      /* stride of vibLevel */
    int32_t pad_;
    int32_t* origProc;     /* [n] particle::origProc_ (BASIC/particle/particle.H:134): with origId the identity of a parcel; NULL on
                              upload = this rank.  Kept on the device only when nRanks > 1. */
    double* radialWeight;  /* [n] dsmcParcel::RWF_ (the lagrangian field `radialWeight`, dsmcParcelIO.C): NULL on upload = the RWF of
                              the parcel's cell; written on download only with dsmcAxisymmetric */
} dsmcb200_parcels_soa;

/* Counters of one evolve() (noTimeCounter.C:312-337, dsmcCloud.C:935-985,
 * Cloud.H:193-204). */
typedef struct {
    int64_t nParcels;
    int64_t collisions;
    int64_t collisionCandidates;
    int64_t trackingRescues;
    int64_t deleted;
    int64_t inserted;
    int64_t migratedOut;
    int64_t migratedIn;
    int64_t unsortedLargeCells;  /* always 0: every cell is put into cloud-list order (kept for ABI stability) */
    double mass, linearKineticEnergy, rotationalEnergy, vibrationalEnergy, electronicEnergy;
    double stageMs[8];  /* last step: inflow, move, migrate, sort, collide, sample, info, total */
    /* processor-patch transfers of the last step, per neighbour processor, summed over the rounds of Cloud<T>::move
     * (the sizes of particleTransferLists / the received lists, BASIC/Cloud/Cloud.C:283-306,356-404) */
    int32_t nNeighbours;
    int32_t migrationRounds;
    int32_t neighbourProc[DSMCB200_MAX_NEIGHBOURS];
    int64_t migratedTo[DSMCB200_MAX_NEIGHBOURS];
    int64_t migratedFrom[DSMCB200_MAX_NEIGHBOURS];
    double nMolecules;  /* sum of nParticles(cell) over the parcels: infoMeasurements[6] (dsmcCloudI.H:268-297); the energies above carry
                           the same per-parcel weight */
    int64_t cloned;     /* parcels added by the radial weighting of the last step (dsmcAxisymmetric.C:64-166); its deletions are in `deleted` */
} dsmcb200_counters;

/* Sampled per-cell accumulators of stage 5: the per-species moment sums from
 * which every dsmcVolFields instance (any typeIds subset) is derived
 * (DSMC/macroscopicProperties/derived/combined/dsmcVolFields/dsmcVolFields.C:1115-1237). */
typedef struct {
    int32_t nCells, nSpecies, nQuantities, nModes;
    double nTimeSteps;
} dsmcb200_accum_info;

/* Index of quantity q inside one species block of an accumulator row. */
enum {
    DSMCB200_Q_N = 0,   /* dsmcNSpeciesCum             */
    DSMCB200_Q_PX = 1,  /* sum U (momentum / mass)     */
    DSMCB200_Q_PY = 2,
    DSMCB200_Q_PZ = 3,
    DSMCB200_Q_CC = 4,  /* sum U.U                     */
    DSMCB200_Q_EROT = 5,
    DSMCB200_Q_EELEC = 6,
    DSMCB200_Q_EVIB0 = 7 /* + mode                      */
};

/* ---- life cycle ------------------------------------------------------- */
/* replaces: dsmcCloud ctor, DSMC/clouds/dsmcCloud.C:586-691 (device side) */
int dsmcb200_create(dsmcb200_ctx** out, int device, int rank, int nRanks);
void dsmcb200_destroy(dsmcb200_ctx*);
const char* dsmcb200_last_error(const dsmcb200_ctx*);
int dsmcb200_abi_version(void);

/* NCCL plumbing for processor-patch migration (replaces PstreamBuffers,
 * BASIC/Cloud/Cloud.C:255,324-454).  id is ncclUniqueId (128 bytes). */
int dsmcb200_nccl_unique_id(void* id128);
int dsmcb200_init_comm(dsmcb200_ctx*, const void* id128);

/* ---- configuration ---------------------------------------------------- */
/* replaces: polyMesh addressing + tetBasePtIs used by particle::trackToFace,
 * BASIC/particle/particleTemplates.C:741-743,830-861 */
int dsmcb200_set_mesh(dsmcb200_ctx*, const dsmcb200_mesh*);
/* Cell labels inside the engine (what OpenFOAM's renumberMesh does to the case files, done in memory instead): the engine works on the
 * same polyMesh with its cells relabelled so that neighbours in space are neighbours in the sorted cloud and in the tet table, and every
 * entry point that takes or returns cell labels or per-cell rows (parcels, cell state, cell fields, accumulators, geometry, occupancy,
 * overall temperature) translates.  Only owner / neighbour entries change: faces, points, patches and the face list of every cell stay as
 * they are, so a run equals, label for label, the run on the mesh relabelled with the same table.  Call before set_mesh.
 * mode AS_GIVEN: the caller's labels (default); Z_CURVE: along a z-order curve through the cell centres; GIVEN: newOfOld[nCells]. */
typedef enum {
    DSMCB200_CELL_ORDER_AS_GIVEN = 0,
    DSMCB200_CELL_ORDER_Z_CURVE = 1,
    DSMCB200_CELL_ORDER_GIVEN = 2
} dsmcb200_cell_order;
int dsmcb200_set_cell_order(dsmcb200_ctx*, int mode, const int32_t* newOfOldOrNull, int32_t nCells);
/* the table in use after set_mesh: the engine's label of the caller's cell k (identity for AS_GIVEN) */
int dsmcb200_download_cell_order(dsmcb200_ctx*, int32_t* newOfOld);
/* replaces: dsmcCloud::buildConstProps, DSMC/clouds/dsmcCloud.C:40-59 */
int dsmcb200_set_species(dsmcb200_ctx*, int nSpecies, const dsmcb200_species*);
/* replaces: BinaryCollisionModel::New / collisionPartnerSelection::New /
 * dsmcBoundaries ctor, DSMC/clouds/dsmcCloud.C:653-683 */
/* replaces: dsmcReactions ctor + initialConfiguration (DSMC/reactions/basic/dsmcReactions/dsmcReactions.C:69-165): the reaction
 * list and the typeId-pair addressing (more than one model for a pair is an error).  Call after set_species, before the first
 * upload / evolve; n = 0 clears.  The reference's validity checks (products of a diatomic must be atoms, ...) are applied. */
int dsmcb200_set_reactions(dsmcb200_ctx*, int n, const dsmcb200_reaction* reactions);
/* replaces: nTotDissociationReactions_ / nTotExchangeReactions_ of the reaction models (dissociationQK.C:276-277, exchangeQK.C:278-279):
 * counts3n[3 r + {0, 1, 2}] = dissociations of reactant 0, of reactant 1, exchanges of reaction r since set_reactions (this rank) */
int dsmcb200_reaction_counts(dsmcb200_ctx*, int n, int64_t* counts3n);
int dsmcb200_set_models(dsmcb200_ctx*, const dsmcb200_models*);
/* replaces: the volScalarFields behind dsmcCloud::nParticles(cell) and deltaTValue(cell) (DSMC/clouds/dsmcCloudI.H:70-100):
 *   nParticles[nCells]  dsmcTimeStepModel::nParticles_ (timeStepModel/basic/dsmcTimeStepModelI.H:72-90); NULL = models.nEquivalentParticles
 *   deltaT[nCells]      dsmcVariableTimeStepModel::deltaT_ (variableTimeStepModel/dsmcVariableTimeStepModel.C:91-100); NULL = models.deltaT
 *   RWF[nCells]         dsmcAxisymmetric::RWF_ (axisymmetric/dsmcAxisymmetricI.H:55-73); NULL = 1
 * A cell's parcels stand for nParticles[c] * RWF[c] molecules and advance by deltaT[c] per step: candidate pairs (noTimeCounter.C:101,148),
 * inflow (dsmcFreeStreamInflowPatch.C:109-140), dsmcMeshFill (dsmcMeshFill.C:146), wall measurements (dsmcPatchBoundary.C:275-289,467),
 * dsmcParcel::move (dsmcParcel.C:62-97).  The values of a boundary face are those of its cell, as the reference sets them
 * (dsmcVariableTimeStepModel.C:83-93, dsmcAxisymmetric.C:219-230).  Call after set_mesh, at any time; the arrays are copied. */
int dsmcb200_set_cell_fields(dsmcb200_ctx*, const double* nParticles, const double* deltaT, const double* RWF);
/* The three fields as the engine uses them (uniform values expanded); any pointer may be NULL. */
int dsmcb200_download_cell_fields(dsmcb200_ctx*, double* nParticles, double* deltaT, double* RWF);
/* Reserve device storage for up to maxParcels (0 -> grow on demand). */
int dsmcb200_reserve(dsmcb200_ctx*, int64_t maxParcels);

/* ---- cloud state ------------------------------------------------------ */
/* replaces: Cloud<T>::initCloud + dsmcParcel::readFields,
 * BASIC/Cloud/CloudIO.C:111-165, DSMC/parcels/dsmcParcelIO.C:133-335.
 * If tetFace/tetPt are NULL they are located on the device like
 * particle::initCellFacePtOrDeleteLostParticle (BASIC/particle/particleI.H:851-996): first tet of the
 * given cell with tetrahedron::inside, else a walk of 1e-5 steps towards the cell centre; parcels outside
 * the 10 %-inflated cell bounding box (or not locatable) are deleted and counted in counters.deleted. */
int dsmcb200_upload_parcels(dsmcb200_ctx*, int64_t n, const dsmcb200_parcels_soa*);
int dsmcb200_download_parcels(dsmcb200_ctx*, int64_t capacity, int64_t* n, dsmcb200_parcels_soa*);
/* replaces: read of <time>/dsmcSigmaTcRMax and
 * buildCollisionSelectionRemainderFromScratch, DSMC/clouds/dsmcCloud.C:105-116,625-636.
 * remainder NULL -> uniform Philox randoms. */
int dsmcb200_upload_cellstate(dsmcb200_ctx*, const double* sigmaTcRMax, const double* remainderOrNull);
int dsmcb200_download_cellstate(dsmcb200_ctx*, double* sigmaTcRMax, double* remainder);
/* replaces: dsmcMeshFill::setInitialConfiguration,
 * DSMC/initialiseDsmcParcels/derived/dsmcMeshFill/dsmcMeshFill.C:70-240
 * (synthetic loads are generated on the device). */
int dsmcb200_mesh_fill(dsmcb200_ctx*, int nTypes, const int32_t* typeIds, const double* numberDensities,
                       double translationalT, double rotationalT, double vibrationalT, double electronicT,
                       const double velocity[3]);
/* replaces: dsmcZoneFill::setInitialConfiguration,
 * DSMC/initialiseDsmcParcels/derived/dsmcZoneFill/dsmcZoneFill.C:71-272: the insertion of dsmcMeshFill for the cells of one
 * cellZone (zoneCells: the zone's cell labels in the zone's order), APPENDED to the cloud as dsmcAllConfigurations runs the
 * configurations of system/dsmcInitialiseDict one after the other; sigmaTcRMax is set for the zone's cells only (:258-268).
 * An empty cloud to start from: dsmcb200_upload_parcels with n = 0. */
int dsmcb200_zone_fill(dsmcb200_ctx*, int64_t nZoneCells, const int32_t* zoneCells, int nTypes, const int32_t* typeIds,
                       const double* numberDensities, double translationalT, double rotationalT, double vibrationalT,
                       double electronicT, const double velocity[3]);

/* ---- the hot path ----------------------------------------------------- */
/* replaces: dsmcCloud::evolve(), DSMC/clouds/dsmcCloud.C:819-926 */
int dsmcb200_evolve(dsmcb200_ctx*, int nSteps);

typedef enum {
    DSMCB200_STAGE_INFLOW = 0,  /* dsmcFreeStreamInflowPatch::controlParcelsBeforeMove */
    DSMCB200_STAGE_MOVE = 1,    /* Cloud<dsmcParcel>::move, BASIC/Cloud/Cloud.C:204     */
    DSMCB200_STAGE_SORT = 2,    /* dsmcCloud::buildCellOccupancy, dsmcCloud.C:63-74     */
    DSMCB200_STAGE_COLLIDE = 3, /* noTimeCounter::collide, noTimeCounter.C:81           */
    DSMCB200_STAGE_SAMPLE = 4   /* dsmcVolFields::calculateField, dsmcVolFields.C:1071  */
} dsmcb200_stage_id;
/* Single-stage entry points for parity tests and ncu. */
int dsmcb200_stage(dsmcb200_ctx*, int stage);
/* Time-step index used as the Philox step word (evolve increments it). */
int dsmcb200_set_step(dsmcb200_ctx*, uint32_t step);

/* ---- results ---------------------------------------------------------- */
/* cellOccupancy as CSR offsets [nCells+1] (valid after STAGE_SORT). */
int dsmcb200_download_occupancy(dsmcb200_ctx*, int32_t* cellOffsets);
/* replaces: the accumulators of dsmcVolFields (calculateField, :1115-1237) and
 * resumeSampling_<name> (writeOut/readIn, :647-835).
 * acc layout: [nCells][nSpecies][nQuantities]; coll: [nCells][2] = (nCollsCum, collisionSeparationCum). */
/* Sample sets: every field{} of fieldPropertiesDict has its own sampleInterval (dsmcField.C:113-152, dsmcVolFields.C:1073-1081); fields
 * that share one share a set of sums.  set_sample_sets (before the first call that finalises the engine) declares one set per distinct
 * interval -- set 0 replaces models.sampleInterval --, each with its own cell, collision and wall accumulators and nTimeSteps;
 * select_sample_set chooses the set accum_info_get, download / upload / reset_accumulators and download / upload_wall_accumulators act on
 * (default 0). */
int dsmcb200_set_sample_sets(dsmcb200_ctx*, int nSets, const int32_t* sampleIntervals);
int dsmcb200_select_sample_set(dsmcb200_ctx*, int set);
int dsmcb200_accum_info_get(dsmcb200_ctx*, dsmcb200_accum_info*);
int dsmcb200_download_accumulators(dsmcb200_ctx*, double* acc, double* coll);
int dsmcb200_upload_accumulators(dsmcb200_ctx*, const double* acc, const double* coll, double nTimeSteps);
int dsmcb200_reset_accumulators(dsmcb200_ctx*);      /* dsmcVolFields resetField / resetAtOutput */
/* Wall-patch accumulators of boundaryMeasurements
 * (DSMC/boundaryMeasurements/boundaryMeasurements.C:364-401):
 * layout [nBoundaryFaces][nSpecies][nWallQuantities]. */
int dsmcb200_wall_info(dsmcb200_ctx*, int32_t* nBoundaryFaces, int32_t* nWallQuantities);
int dsmcb200_download_wall_accumulators(dsmcb200_ctx*, double* wall);
/* restart of the wall sampling: the *BF_ arrays dsmcVolFields::readIn restores (dsmcVolFields.C:723-738) */
int dsmcb200_upload_wall_accumulators(dsmcb200_ctx*, const double* wall);
/* dsmcFaceTracker (DSMC/faceTracker/dsmcFaceTracker.C:124-198; models.trackFaceFluxes = 1): parcelIdFlux_ and massIdFlux_ of the
 * last step, each [nSpecies][nFaces] (all faces, internal first).  A face crossing adds sign(U . S_f) (x mass), U after the
 * boundary interaction; a cyclic hit is credited, unsigned, to the coupled face; a parcel inserted on an inflow face counts as
 * a crossing of it (dsmcCloud.C:429-437).  The arrays are cleared at the start of every step (trackingInfo_.clean(),
 * dsmcCloud.C:923).  Replaces cloud.tracker().parcelIdFlux() / massIdFlux() (dsmcFaceTracker.H). */
int dsmcb200_download_face_fluxes(dsmcb200_ctx*, double* parcelIdFlux, double* massIdFlux);
/* inverseZvFormulation "2008" (LarsenBorgnakkeVariableHardSphereCoeffs; invZvFormulation = 1): the vibrational collision number is
 * evaluated at the macroscopic overall temperature of the cell, fields().overallT(cellI) = Tov_ of the first dsmcVolFields entry as of
 * its last write (dsmcCloud.C:1441-1456, dsmcFieldProperties.C:235-240, dsmcVolFields.H:244-247).  The caller computes Tov when it
 * writes the fields and hands it over; cells with Tov <= SMALL (and every cell before the first upload) use the quantised
 * collision temperature, as the reference does.  Tov: [nCells]. */
int dsmcb200_upload_overall_temperature(dsmcb200_ctx*, const double* Tov);
int dsmcb200_get_counters(dsmcb200_ctx*, dsmcb200_counters*);
/* Per-kernel device time of the last step, for bench.py: names[i] is filled with up to
 * DSMCB200_NAME_LEN chars; returns the count through *n (capacity in). */
int dsmcb200_kernel_times(dsmcb200_ctx*, int capacity, int* n, char* names, float* ms, int64_t* launches);
/* Sum of vals[0..n) over all ranks (replaces the reduce(..., sumOp) calls of dsmcCloud::info and
 * noTimeCounter::collide, DSMC/clouds/dsmcCloud.C:938-958); a no-op on one rank. */
int dsmcb200_allreduce_sum(dsmcb200_ctx*, double* vals, int n);
/* reduce(x, minOp<scalar>()) over the ranks, e.g. gMin(mesh.V()) of dsmcVariableTimeStepModel::findRefCell (n <= 8) */
int dsmcb200_allreduce_min(dsmcb200_ctx*, double* vals, int n);
/* CUDA-event stopwatch on the context's launching stream (bench.py times the K steps with it). */
int dsmcb200_timer_start(dsmcb200_ctx*);
int dsmcb200_timer_stop(dsmcb200_ctx*, float* ms);
/* Derived geometry back to the caller (tests / writers). */
int dsmcb200_download_geometry(dsmcb200_ctx*, double* cellCentres, double* cellVolumes,
                               double* faceCentres, double* faceAreas, int32_t* tetBasePtIs);

#ifdef __cplusplus
}
#endif
#endif /* DSMCB200_H */
