// engine.h -- device-side data layout of libdsmcb200 and the launchers of its kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "host_mesh.h"
#include "philox.h"
#include "vec3.h"

namespace dsmc {

constexpr int MAX_SPECIES = DSMCB200_MAX_SPECIES;
constexpr int MAX_MODES = DSMCB200_MAX_VIB_MODES;
constexpr int MAX_ELEC = DSMCB200_MAX_ELEC_LEVELS;
constexpr int MAX_PATCHES = 64;
constexpr int MAX_NEIGHBOURS = DSMCB200_MAX_NEIGHBOURS;
constexpr int MAX_INFLOWS = 8;
constexpr int ZV_TABLE = 128;  // tabulated iMax range of the variable vibrational collision number
constexpr int MAX_REACTIONS = 32;   // entries of system/chemReactDict

struct DevSpecies {
    double mass, d, omega, alpha, rotDof, thetaD;
    double thetaV[MAX_MODES], Zref[MAX_MODES], TrefZv[MAX_MODES];
    double eElec[MAX_ELEC];
    int32_t gElec[MAX_ELEC];
    int32_t nVib, charge, nElec, type;
};

struct DevPatch {
    int32_t type;        // dsmcb200_patch_type
    int32_t model;       // dsmcb200_patch_model_kind (wall / patch types)
    int32_t start, size;
    int32_t nbrPatch;    // cyclic: neighbourPatch
    int32_t pad_;
    int32_t nbrSlot;     // processor patches: slot in the neighbour list
    int32_t nbrOrdinal;  // processor patches: ordinal of this patch among those facing that neighbour
    double T;            // wall temperature
    double vel[3];       // wall velocity
    double sep[3];       // cyclic separation (receiving side subtracts)
    double diffuseFraction;  // dsmcDiffuseSpecularWallPatch
    // dsmcDiffuseWallPatch::getLocalTemperature: T + (x[depthAxis] - maxDepth) * (T - Tformation) / lengthPatch
    int32_t linearT, depthAxis;
    double Tformation, maxDepth, lengthPatch;
    double alphaN, alphaT, alphaR;   // dsmcCLLWallPatch: normal, tangential*(2 - tangential), rotational accommodation
};

// one quantum-kinetic reaction model of system/chemReactDict (dsmcb200_reaction after dsmcReaction::setProperties)
struct DevReaction {
    int32_t model;              // dsmcb200_reaction_model
    int32_t reactants[2];
    int32_t allowSplitting;
    int32_t dissProd[2][2];     // products of the dissociation of reactant r
    int32_t exchProd[2];        // [0] the molecule, [1] the atom (exchangeQK.C:141-170)
    int32_t posMolReactant;     // which reactant is the molecule of the exchange
    int32_t pad_;
    double heatDissJ[2];        // kB * thetaD of reactant r (dissociationQK.C:120-140)
    double heatExchJ;           // heatOfReactionExchange * kB
    double aDash;               // aCoeff * chiB^b * Gamma(chiB) / Gamma(chiB + b), chiB = 2.5 - omegaPQ of the reactants (exchangeQK.C:196-200)
    double bCoeff;
};

struct DevParams {
    int32_t nSpecies, nPatches, collisionModel, invZvFormulation;
    int32_t coordinateSystem, angularCoordinate;   // dsmcb200_coordinate_system; dsmcAxisymmetric: the mirrored velocity component of a clone
    int32_t nModes;  // max vibrational modes over species (stride of vib arrays)
    int32_t hasInternalEnergy, measureFlux, measureClass;
    int32_t solutionD[3];
    int32_t nInternalFaces;
    double centre[3];  // bounds mid-point for constrainToMeshCentre
    double nParticles, deltaT, kB, Tref;
    double invZrot, Zvib, invZelec;
    const double* invZvTab;  // [species][partner][mode][ZV_TABLE]: 1/Zv (or 1/(5 Zv)) of the quantised collision temperature

    uint64_t seed;
    DevSpecies sp[MAX_SPECIES];
    DevPatch patch[MAX_PATCHES];
    // per species pair: pi*dPQ^2, exp(lgamma(2.5-omegaPQ)), omegaPQ, reduced mass
    double vhsA[MAX_SPECIES][MAX_SPECIES], vhsG[MAX_SPECIES][MAX_SPECIES];
    double omegaPQ[MAX_SPECIES][MAX_SPECIES], mR[MAX_SPECIES][MAX_SPECIES];
    // dsmcReactions: the reaction model of a typeId pair (-1: none), dsmcReactions.C:137-165
    int32_t nReactions;
    int8_t pairReaction[MAX_SPECIES][MAX_SPECIES];
    DevReaction reactions[MAX_REACTIONS];
};

// Structure-of-arrays cloud: one array per component so every stage streams coalesced.
struct ParcelArrays {
    double *px, *py, *pz, *ux, *uy, *uz, *erot;
    int32_t *cell, *tet, *origId;
    int32_t* vib[MAX_MODES];
    uint8_t *typeId, *elevel, *cls;
    uint8_t* origProc;   // particle::origProc_: with origId the unique identity of a parcel (keys its wall-model random stream)
    double* rwf;         // dsmcParcel::RWF_ (dsmcAxisymmetric only, else nullptr): the radial weight of the cell the parcel started the step in
};

// dsmcCloud::nParticles(cell) = nPts[cell] * rwf[cell] and deltaTValue(cell) (dsmcCloudI.H:70-100); nullptr = the uniform value of DevParams
struct CellFields {
    const double *nPts, *dt, *rwf;
    __host__ __device__ double nParticlesTs(const double uniform, int32_t c) const { return nPts ? nPts[c] : uniform; }
    __host__ __device__ double nParticles(const double uniform, int32_t c) const { const double n = nPts ? nPts[c] : uniform; return rwf ? n * rwf[c] : n; }
    __host__ __device__ double deltaT(const double uniform, int32_t c) const { return dt ? dt[c] : uniform; }
    __host__ __device__ double RWF(int32_t c) const { return rwf ? rwf[c] : 1.0; }
};

// Packed record shipped across a processor patch (BASIC/particle/particleIO.C:121-132 +
// DSMC/parcels/dsmcParcelIO.C:491-501 carry the same information).
struct alignas(16) MigRec {
    double pos[3], U[3], erot, stepFraction;
    int32_t patchOrdinal, patchFace, tetLocal, origId;
    int32_t vib[MAX_MODES];
    uint8_t typeId, elevel, cls, origProc;
};
static_assert(sizeof(MigRec) == 96, "MigRec must be 96 bytes");

struct DevCounters {
    unsigned long long collisions, candidates, rescues, deleted, migratedOut, unsortedLargeCells, overflow;
    unsigned long long trackingFailures;   // parcels dropped by the move kernel's iteration guard (a corrupt tet table): an error
    int32_t nMig[MAX_NEIGHBOURS];
    int32_t nInserted;
    int32_t bigCells;  // cells handed from collideLaneKernel to collideBigCellsKernel this step
    int32_t bigSortCells;  // cells handed from segmentSortKernel to bigSegmentSortKernel
    int32_t giantSortCells;  // ... and from there to giantSortKernel (more than GIANT_SORT parcels: the one-cell heat baths)
    int32_t nBorn;         // parcels created by dissociations this step (dsmcCloud::addNewParcel)
    int32_t weightDeleted; // parcels deleted by the radial weighting this step
    int32_t nPendingLocate; // parcels an upload without tet indices sends on to the mesh-wide search
    int32_t pad_;
    unsigned long long nReact[MAX_REACTIONS][3];   // this step: dissociations of reactant 0, of reactant 1, exchanges
};

// the second product of a dissociation (dissociationQK.C:355-370) until it joins the cloud in (cell, candidate) order
struct BornRec {
    unsigned long long key;     // cell << 32 | candidate index: the order in which the reference's serial loop creates them
    double pos[3], U[3], rwf;
    int32_t cell, tet;
    uint8_t typeId, cls, pad_[6];
};

// wall accumulator quantities per (measured face, species)
enum WallQ { WQ_RHON = 0, WQ_RHON_INT, WQ_RHON_ELEC, WQ_RHOM, WQ_LINKE, WQ_MCC, WQ_MOMX, WQ_MOMY, WQ_MOMZ,
             WQ_EROT, WQ_ZETAROT, WQ_EVIB, WQ_EELEC, WQ_Q, WQ_FDX, WQ_FDY, WQ_FDZ, WQ_EVIBMOD0, WQ_BASE = WQ_EVIBMOD0 };

constexpr int MOVE_PMAX = 1024;   // parcels per work-list entry of the move kernel (upper bound)
constexpr int MOVE_TAIL = 512;    // parcels per entry of the unsorted tail
#ifndef MOVE_NBUF_SZ
#define MOVE_NBUF_SZ 4
#endif
constexpr int MOVE_NBUF = MOVE_NBUF_SZ;   // entries in flight per block (ring of shared-memory windows)

struct MoveArgs {
    ParcelArrays p;
    CellFields cf;
    int32_t weighted;             // coordinate system other than dsmcCartesian: selects the kernel instance that looks at cf / p.rwf
    // work list (launchMovePlan): plan[0 .. *planTotal) = {parcelBeg, parcelEnd, tetBeg, nTets}
    const int4* plan;
    const int32_t* planTotal;
    int32_t gridBlocks;           // persistent blocks (<= SMs)
    int32_t stageTets;            // capacity of one shared-memory window in tet records
    int32_t tailStart;            // parcels >= tailStart carry a step fraction in sfTail[i - tailStart]
    const double* sfTail;
    const TetRec* tets;
    const BFaceRec* bfaces;
    const double* bfaceArea;      // [nBFaces*3] face area vectors
    const DevParams* P;
    double* wallAcc;              // [nMeasFaces][nSpecies][nWallQ]
    int32_t nWallQ;
    int32_t wallsDue;             // 0: this step is not sampled (sampleInterval), wall hits leave no measurement
    double* faceFlux;             // dsmcFaceTracker: [2][nSpecies][nFacesAll] (parcelIdFlux, massIdFlux) or nullptr
    const double* faceAreas;      // [nFacesAll*3], read only by the face tracker
    int32_t nFacesAll;
    MigRec* migBuf;               // [MAX_NEIGHBOURS][migCapacity]
    double* migRwf;               // [MAX_NEIGHBOURS][migCapacity] the leavers' radial weights (dsmcAxisymmetric), next to the 96-byte records; or nullptr
    int32_t* migKey;              // [MAX_NEIGHBOURS][migCapacity] cloud index of the packed parcel: the sender's list order
    int32_t migCapacity;
    int32_t* cellCount;           // histogram for the sort (stage 2), fused here
    DevCounters* counters;
    uint32_t step;
    int32_t cllWalls;             // some patch carries a dsmcCLLWallPatch: selects the kernel instance that contains the CLL scattering kernel
};

struct CollideArgs {
    ParcelArrays p;
    CellFields cf;
    const int32_t* cellOffset;    // [nCells+1]
    int32_t nCells;
    const double* cellCentres;    // [nCells*3]
    const double* cellVolumes;
    double* sigmaTcRMax;
    double* remainder;
    double* nCollsStep;           // cellMeasurements: nColls_ of this step
    double* collSepStep;          // cellMeasurements: collisionSeparation_ of this step
    const double* overallT;       // [nCells] fields().overallT(cell) for inverseZvFormulation "2008", or nullptr
    int32_t nModes;               // vibrational modes stored per parcel (selects the kernel instance)
    int32_t* bigScratch;          // [nParcels] sub-cell index lists of cells too large for shared memory
    int32_t* bigList;             // [nCells] ids of those cells (written by the lane kernel)
    const uint8_t* octKey;        // [nParcels] sub-cell (octant) of every sorted parcel, written by the sort's gather
    const int32_t* giantList;     // [nGiant] cells of more than GIANT_SORT parcels (written by the sort): one block each
    int32_t nGiant;
    BornRec* born;                // [bornCapacity] parcels created by reactions (nullptr without chemistry)
    int32_t bornCapacity;
    const DevParams* P;
    DevCounters* counters;
    uint32_t step;
};

struct SampleArgs {
    ParcelArrays p;
    const int32_t* cellOffset;
    int32_t nCells;
    double* acc;                  // [nCells][nSpecies][nQ]
    int32_t nQ, nSpecies;
    int32_t nParcels;             // sorted parcels (= cellOffset[nCells])
    int32_t nCloud;               // all parcels: [nParcels, nCloud) were created by this step's reactions and are in no cell list
    double* collCum;              // [nCells][2]
    const double* nCollsStep;
    const double* collSepStep;
    const DevParams* P;
};

struct KernelTimer;  // engine.cu

// per-step work list of the move kernel over the cell-sorted part of the cloud
struct MovePlanArgs {
    const int32_t* groupCell;     // [nGroups+1] runs of cells whose tet records fit the window (HostMesh::stageGroupCell)
    int32_t nGroups;
    const int32_t* cellOffset;    // occupancy CSR of the sorted cloud
    const int32_t* cellTetStart;  // [nCells+1]
    int32_t* nSub;                // [nGroups+1] scratch: blocks per run
    int32_t* subBase;             // [nGroups+1] exclusive scan; subBase[nGroups] = entries of the sorted part
    int32_t maxTets;
    int32_t tailBeg, tailEnd;     // unsorted parcels appended to the list in pieces of MOVE_TAIL
    int4* plan;                   // [nGroups + N/MOVE_PMAX + tail pieces + 1]
    int32_t* planTotal;           // out: entries in the list
};
size_t moveSharedBytes(int32_t stageTets);
int32_t moveMaxStageTets();

// ---- launchers (each returns cudaGetLastError()) ----
cudaError_t launchMove(const MoveArgs& a, cudaStream_t s);
cudaError_t launchMovePlan(const MovePlanArgs& m, int32_t* scanScratch, cudaStream_t s);
cudaError_t launchExclusiveScan(const int32_t* in, int32_t* out, int32_t* out2, int32_t n, int32_t* blockSums, cudaStream_t s);
int32_t scanScratchInts(int32_t n);
cudaError_t launchScatterIndex(const int32_t* cell, int32_t n, int32_t* cursor, int32_t* perm, cudaStream_t s);
cudaError_t launchSegmentSort(const int32_t* cellOffset, int32_t nCells, int32_t* perm, DevCounters* c, int32_t* bigList, int32_t* giantList, cudaStream_t s);
constexpr int32_t GIANT_LIST = 32768;   // entries of giantList: 2^31 parcels / GIANT_SORT
// cells of more than GIANT_SORT parcels (listed in giantList by the block-level sort): ordered through a bitmap of the cell's index range, one block per
// cell at a time; bitmap: blocks * wordsPerBlock words of scratch, wordsPerBlock >= nIn / 32 + 2
constexpr int32_t GIANT_SORT = 65536;
constexpr int GIANT_SORT_BLOCKS = 4;
cudaError_t launchGiantSort(const int32_t* cellOffset, int32_t* perm, const int32_t* giantList, int32_t nGiant, uint32_t* bitmap,
                            int64_t wordsPerBlock, cudaStream_t s);
cudaError_t launchGather(const ParcelArrays& src, const ParcelArrays& dst, const int32_t* perm, const double* cellCentres,
                         uint8_t* octKey, int32_t nOut, int32_t nModes, bool hasInternal, cudaStream_t s);
cudaError_t launchHistogram(const int32_t* cell, int32_t n, int32_t* cellCount, cudaStream_t s);
cudaError_t launchCollide(const CollideArgs& a, cudaStream_t s);
cudaError_t launchSample(const SampleArgs& a, cudaStream_t s);
// the parcels reactions created join the cloud at [base, base + n) in key order; work: 2 n keys + 2 n ints
size_t orderBornTempBytes(int32_t capacity);
cudaError_t launchAppendBorn(const ParcelArrays& p, const BornRec* born, int32_t n, int32_t base, int32_t origIdBase, int32_t origProc, int32_t nModes,
                             unsigned long long* keyWork, int32_t* idxWork, void* temp, size_t tempBytes, cudaStream_t s);
cudaError_t launchInfo(const ParcelArrays& p, const CellFields& cf, int32_t n, const DevParams* P, double* out5, double* scratch, cudaStream_t s);

// dsmcAxisymmetric::axisymmetricWeighting (dsmcAxisymmetric.C:50-209) over the sorted cloud [0, n): pass 0 gives every parcel its cell's RWF,
// marks the parcels to delete (cell = -1) and counts the clones of each parcel; pass 1 (counts scanned) writes the clones behind the cloud
struct WeightArgs {
    ParcelArrays p;
    CellFields cf;
    int32_t n, base, capacity;
    int32_t angularCoordinate;
    int32_t* counts;              // [n + 1] clones per parcel -> offsets after the scan
    int32_t origIdBase, origProc, nModes;
    const DevParams* P;
    DevCounters* counters;
    uint32_t step;
};
cudaError_t launchWeighting(const WeightArgs& a, int pass, cudaStream_t s);
int32_t infoScratchDoubles();

struct FillArgs {
    ParcelArrays p;
    CellFields cf;
    int32_t nCells;
    const int32_t *cellFaceOffsets, *cellFaces, *faceOffsets, *facePoints, *owner, *tetBasePtIs, *cellTetStart;
    const double *points, *cellCentres;
    const DevParams* P;
    int32_t nTypes;
    int32_t typeIds[MAX_SPECIES];
    double numberDensities[MAX_SPECIES];
    double Ttra, Trot, Tvib, Telec;
    double velocity[3];
    int32_t* cellCount;   // pass 0 output / pass 1 input: offsets, one per entry of the fill
    int32_t origIdBase;
    int32_t origProc;     // this rank (particle::origProc_ of the parcels it creates)
    int32_t nFill;                // entries: nCells (dsmcMeshFill) or the size of the zone
    const int32_t* cellList;      // dsmcZoneFill: the zone's cells in the zone's order; nullptr: entry t is cell t
    int32_t slotBase;             // first slot of the fill in the parcel arrays (the cloud's size when a zone fill appends)
    uint32_t fillIndex;           // the k-th fill since the cloud was emptied draws from streams of its own
};
cudaError_t launchFill(const FillArgs& a, int pass, cudaStream_t s);

// Cloud<T>::move appends leavers to particleTransferLists[neighbour] in cloud-list order (BASIC/Cloud/Cloud.C:283-306); moveKernel
// packs them in atomic order, so each neighbour's records are put back into list order (ascending cloud index) before they are sent.
// scratch: n records; work: 3*n ints; temp: orderMigrantsTempBytes(capacity) bytes.
size_t orderMigrantsTempBytes(int32_t capacity);
cudaError_t orderMigrants(MigRec* records, MigRec* scratch, const int32_t* keys, int32_t* work, void* temp, size_t tempBytes, int32_t n,
                          cudaStream_t s, double* rwf = nullptr, double* rwfScratch = nullptr);

// particle::initCellFacePtOrDeleteLostParticle for a cloud that arrives without tetFace / tetPt (the `positions` file holds only
// "(x y z) cell", particleIO.C:51-58)
struct LocateArgs {
    double *px, *py, *pz;         // in; a parcel found by the walk towards its cell centre gets the new position (particleI.H:976)
    int32_t* cell;                // in: host cell label; out: -1 for a lost parcel
    int32_t* tet;                 // out: tet id cellTetStart[cell] + index of (tetFace, tetPt) among the cell's tets
    int32_t n, nCells;
    const int32_t *cellFaceOffsets, *cellFaces, *faceOffsets, *facePoints, *owner, *tetBasePtIs, *cellTetStart;
    const int32_t* neighbour;     // [nInternalFaces]: the search of polyMesh::findCellFacePt looks at the cells around the given one
    int32_t* pending;             // [n] parcels that go on to the mesh-wide search (nullptr: no such search)
    int32_t* nPending;
    double searchRadius2;         // square of the mesh's largest cell-centre-to-vertex distance
    int32_t nInternalFaces;
    const double *points, *cellCentres;
    unsigned long long* lost;     // parcels deleted (outside the inflated cell bounding box or not locatable)
};
cudaError_t launchLocate(const LocateArgs& a, cudaStream_t s);
cudaError_t launchLocateGlobal(const LocateArgs& a, int32_t nPending, cudaStream_t s);   // the parcels launchLocate listed in a.pending
constexpr int32_t LOCATE_GLOBAL_MAX = 65536;   // parcels one upload may send through the mesh-wide search

struct InflowArgs {
    ParcelArrays p;
    CellFields cf;
    int32_t nFaces;               // faces of the inflow patch
    int32_t patch, patchStart;
    const int32_t *faceOffsets, *facePoints, *owner, *tetBasePtIs;
    const BFaceRec* bfaces;       // tet0 of every boundary face
    int32_t nInternalFaces;
    const double *points, *faceCentres, *faceAreas;
    const DevParams* P;
    int32_t nTypes;
    int32_t typeIds[MAX_SPECIES];
    double numberDensities[MAX_SPECIES];
    double velocity[3];
    double Ttra, Trot, Tvib, Telec;
    double* accumulator;          // [nTypes][nFaces] accumulatedParcelsToInsert_
    int32_t* counts;              // [nTypes*nFaces] -> offsets after scan
    int32_t base;                 // first free parcel slot
    int32_t capacity;
    double* sfTail;               // step fractions of the new parcels, indexed slot - tailStart
    int32_t tailStart;
    int32_t origIdBase;
    int32_t origProc;
    double* faceFlux;             // dsmcFaceTracker arrays (see MoveArgs) or nullptr
    int32_t nFacesAll;
    DevCounters* counters;
    uint32_t step;
};
cudaError_t launchInflow(const InflowArgs& a, int pass, cudaStream_t s);

struct UnpackArgs {
    ParcelArrays p;
    const MigRec* recv;
    const double* recvRwf;        // radial weights of the arrivals (dsmcAxisymmetric) or nullptr
    int32_t nRecv, base;
    double* sfTail;
    int32_t tailStart;
    const int32_t* ordinalToPatch;  // [MAX_PATCHES] for the sending neighbour
    const BFaceRec* bfaces;
    const DevParams* P;
};
cudaError_t launchUnpack(const UnpackArgs& a, cudaStream_t s);

}  // namespace dsmc
