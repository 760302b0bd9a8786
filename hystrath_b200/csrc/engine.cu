// engine.cu -- the C ABI of libdsmcb200 (include/dsmcb200.h): context, device memory, stage
// sequencing of dsmcCloud::evolve() (DSMC/clouds/dsmcCloud.C:819-926) and NCCL migration
// (the MPI block of Cloud<T>::move, BASIC/Cloud/Cloud.C:258-455).
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "engine.h"

using namespace dsmc;

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time so single-GPU use does not need the library at all
// ------------------------------------------------------------------------------------------------
namespace {
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& err) {
        if (h) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) { err = "cannot dlopen libnccl.so.2"; return false; }
#define LD(f, s) f = reinterpret_cast<decltype(f)>(dlsym(h, s)); if (!f) { err = std::string("nccl symbol missing: ") + s; return false; }
        LD(GetUniqueId, "ncclGetUniqueId") LD(CommInitRank, "ncclCommInitRank") LD(CommDestroy, "ncclCommDestroy")
        LD(AllGather, "ncclAllGather") LD(AllReduce, "ncclAllReduce") LD(Send, "ncclSend") LD(Recv, "ncclRecv") LD(GroupStart, "ncclGroupStart")
        LD(GroupEnd, "ncclGroupEnd") LD(GetErrorString, "ncclGetErrorString")
#undef LD
        return true;
    }
};
NcclApi g_nccl;
constexpr int NCCL_CHAR = 0, NCCL_INT32 = 2;
}  // namespace

// ------------------------------------------------------------------------------------------------
// small conversion kernels (host AoS layouts <-> device SoA)
// ------------------------------------------------------------------------------------------------
__global__ void deinterleave3(const double* __restrict__ in, double* x, double* y, double* z, int32_t n) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { x[i] = in[3 * i]; y[i] = in[3 * i + 1]; z[i] = in[3 * i + 2]; }
}
__global__ void interleave3(const double* x, const double* y, const double* z, double* __restrict__ out, int32_t n) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[3 * i] = x[i]; out[3 * i + 1] = y[i]; out[3 * i + 2] = z[i]; }
}
__global__ void toTetId(const int32_t* cell, const int32_t* tetFace, const int32_t* tetPt, const int32_t* faceTet0, const int32_t* faceOffsets,
                        const int32_t* owner, const TetRec* tets, int32_t* tet, int32_t n, int32_t nFaces, int32_t nCells, int* bad) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t f = tetFace[i];
    if (cell[i] < 0) { tet[i] = 0; return; }   // a lost parcel (deleted by the sort), see upload_parcels
    if (f < 0 || f >= nFaces || cell[i] >= nCells) { *bad = 1; tet[i] = 0; return; }
    const int32_t nT = faceOffsets[f + 1] - faceOffsets[f] - 2;
    const int32_t tp = tetPt[i];
    const int32_t t0 = faceTet0[2 * size_t(f) + (owner[f] != cell[i] ? 1 : 0)];
    if (tp < 1 || tp > nT || t0 < 0 || tets[t0].cell != cell[i]) { *bad = 1; tet[i] = 0; return; }   // tetFace is not a face of the cell
    tet[i] = t0 + tp - 1;
}
__global__ void fromTetId(const int32_t* tet, const TetRec* tets, int32_t* tetFace, int32_t* tetPt, int32_t n) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const TetRec& r = tets[tet[i]];
    tetFace[i] = r.face;
    tetPt[i] = r.tetPt;
}
__global__ void i32ToU8(const int32_t* in, uint8_t* out, int32_t n) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = uint8_t(in[i]);
}
// the same with a range check: values outside [0, limit) raise *bad (typeId against typeIdList, dsmcParcelI.H constProps lookup)
__global__ void addDoubles(double* dst, const double* __restrict__ src, int32_t n) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}
// cell labels through a table (dsmcb200_set_cell_order); labels outside the mesh stay as they are for the checks that follow
__global__ void mapLabels(const int32_t* in, int32_t* out, const int32_t* __restrict__ table, int32_t n, int32_t nCells) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t v = in[i];
    out[i] = (v >= 0 && v < nCells) ? table[v] : v;
}
__global__ void i32ToU8Checked(const int32_t* in, uint8_t* out, int32_t n, int32_t limit, int* bad) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t v = in[i];
    if (v < 0 || v >= limit) { *bad = 2; out[i] = 0; return; }
    out[i] = uint8_t(v);
}
__global__ void u8ToI32(const uint8_t* in, int32_t* out, int32_t n) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
__global__ void stridedToMode(const int32_t* in, int32_t* out, int32_t n, int32_t stride, int32_t m) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[size_t(i) * stride + m];
}
__global__ void modeToStrided(const int32_t* in, int32_t* out, int32_t n, int32_t stride, int32_t m) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[size_t(i) * stride + m] = in[i];
}
__global__ void rwfOfCell(const int32_t* cell, const double* rwfCell, double* rwf, int32_t n) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rwf[i] = (rwfCell && cell[i] >= 0) ? rwfCell[cell[i]] : 1.0;
}
__global__ void iotaKernel(int32_t* out, int32_t n, int32_t base) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = base + i;
}
__global__ void fillRemainder(double* rem, int32_t nCells, uint64_t seed) {
    int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    Rng r;
    r.init(seed, uint32_t(c), 0u, 0u, STREAM_REMAINDER);
    rem[c] = r.sample01();  // dsmcCloud::buildCollisionSelectionRemainderFromScratch
}
__global__ void fillDoubleAt(double* p, const int32_t* at, int32_t n, double v) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[at[i]] = v;
}
__global__ void fillDouble(double* p, int64_t n, double v) {
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void fillCellIds(const int32_t* cellOffset, int32_t nCells, int32_t* cellOut) {
    const int lane = threadIdx.x & 31;
    const int32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int32_t nWarps = (gridDim.x * blockDim.x) >> 5;
    for (int32_t c = warp; c < nCells; c += nWarps)
        for (int32_t k = cellOffset[c] + lane; k < cellOffset[c + 1]; k += 32) cellOut[k] = c;
}

#define GRID(n) (int(((n) + 255) / 256)), 256

// ------------------------------------------------------------------------------------------------
struct ParcelBuffer {
    double* dslab = nullptr;
    int32_t* islab = nullptr;
    uint8_t* bslab = nullptr;
    ParcelArrays a{};
};

struct StageEvents { cudaEvent_t ev[8]; };

struct dsmcb200_ctx {
    int device = 0, rank = 0, nRanks = 1;
    std::string err;
    cudaStream_t stream = nullptr;
    HostMesh mesh;
    bool haveMesh = false, haveSpecies = false, haveModels = false, ready = false;
    std::vector<dsmcb200_species> species;
    dsmcb200_models models{};
    int sampleCounter = 0;   // steps since stage 5 last ran (sampleInterval)
    // dsmcb200_set_sample_sets: field{} entries with their own sampleInterval (dsmcField.C:113-152) sample into their own sums.  Set 0 is
    // dAcc / dCollCum / dWallAcc / nTimeSteps / sampleCounter with models.sampleInterval; the others live here.
    struct SampleSet { int interval = 1, counter = 0; double *dAcc = nullptr, *dCollCum = nullptr, *dWallAcc = nullptr; double nTimeSteps = 0; };
    std::vector<SampleSet> extraSets;
    std::vector<int32_t> setIntervals;
    int selectedSet = 0;
    double* dWallStep = nullptr;   // with more than one set: the wall measurements of the step, added to every set that samples it
    double* dOverallT = nullptr;   // fields().overallT(cell) uploaded by the caller at write times (inverseZvFormulation "2008")
    double* dFaceFlux = nullptr;   // dsmcFaceTracker: [2][nSpecies][nFaces] of the current step (models.trackFaceFluxes)
    std::vector<dsmcb200_patch_model> patchModels;
    std::vector<dsmcb200_inflow> inflows;
    // quantum-kinetic chemistry (dsmcb200_set_reactions): second products of dissociations are staged in dBorn during the collide stage
    std::vector<dsmcb200_reaction> reactions;
    std::vector<int64_t> reactionTotals;   // [nReactions][3] since set_reactions
    BornRec* dBorn = nullptr;
    unsigned long long* dBornKeys = nullptr;
    int32_t* dBornIdx = nullptr;
    void* dBornTemp = nullptr;
    size_t bornTempBytes = 0;
    int32_t bornCap = 0;
    // per-cell nParticles (time-step model), deltaT and radial weighting factor (dsmcb200_set_cell_fields); empty / nullptr = uniform
    std::vector<double> hNPts, hDt, hRWF;
    // dsmcb200_set_cell_order: the engine's label of the caller's cell k is newOfOld[k] (empty: the caller's labels are used as they are)
    int cellOrderMode = DSMCB200_CELL_ORDER_AS_GIVEN;
    std::vector<int32_t> newOfOld, oldOfNew, userOrder;
    int32_t *dNewOfOld = nullptr, *dOldOfNew = nullptr;
    double *dNPts = nullptr, *dDt = nullptr, *dRWF = nullptr;
    bool useRwf = false;           // dsmcAxisymmetric: parcels carry a radial weight
    int fillsDone = 0;             // fills since the cloud was emptied (stream key of dsmcb200_mesh_fill / zone_fill)
    bool cllWalls = false;         // a dsmcCLLWallPatch among the patch models (move kernel instance)
    int32_t* dWeightCounts = nullptr; int64_t weightCountsCap = 0;
    uint32_t* dGiantBitmap = nullptr; int64_t giantWords = 0;   // scratch of giantSortKernel
    int32_t nGiant = 0;   // cells of more than GIANT_SORT parcels found by the last sort
    double cellRadius2Max = 0.0;   // square of the largest cell-centre-to-vertex distance (mesh-wide parcel search)
    int32_t* dGiantList = nullptr;
    int64_t cloned = 0, weightDeleted = 0, weightDeletedStep = 0;
    DevParams hP{};
    DevParams* dP = nullptr;
    // device mesh
    TetRec* dTets = nullptr;
    BFaceRec* dBFaces = nullptr;
    double *dBFaceArea = nullptr, *dPoints = nullptr, *dCellCentres = nullptr, *dCellVolumes = nullptr, *dFaceCentres = nullptr,
           *dFaceAreas = nullptr;
    int32_t *dFaceOffsets = nullptr, *dFacePoints = nullptr, *dOwner = nullptr, *dNeighbour = nullptr, *dTetBasePtIs = nullptr, *dFaceTet0 = nullptr, *dCellTetStart = nullptr, *dGroupCell = nullptr,
            *dCellFaceOffsets = nullptr, *dCellFaces = nullptr;
    // work list of the move kernel (launchMovePlan)
    int32_t *dPlanSub = nullptr, *dPlanBase = nullptr;
    int4* dPlan = nullptr;
    int64_t planCap = 0;
    int32_t stageTets = 0, nGroups = 0, moveBlocks = 0;
    // cloud
    ParcelBuffer buf[2];
    int cur = 0;
    int64_t capacity = 0, N = 0;
    int nModes = 0;
    bool internal = false, useCls = false;
    int32_t *dCellCount = nullptr, *dCellOffset = nullptr, *dCursor = nullptr, *dPerm = nullptr, *dScanScratch = nullptr;
    uint8_t* dOctKey = nullptr;  // sub-cell of every sorted parcel (sort -> collide)
    double *dSigma = nullptr, *dRem = nullptr, *dNColls = nullptr, *dCollSep = nullptr, *dAcc = nullptr, *dCollCum = nullptr,
           *dWallAcc = nullptr, *dSfTail = nullptr, *dInfo = nullptr, *dInfoScratch = nullptr;
    int64_t sfCap = 0;
    DevCounters* dCounters = nullptr;
    DevCounters hCounters{};
    int* dBad = nullptr;
    double* dZvTab = nullptr;
    MigRec *dMigSend = nullptr, *dMigRecv = nullptr;
    double *dMigRwfSend = nullptr, *dMigRwfRecv = nullptr;   // the leavers' / arrivals' radial weights (dsmcAxisymmetric), [MAX_NEIGHBOURS][migCapacity]
    int32_t *dMigKey = nullptr, *dMigWork = nullptr;   // cloud index of each packed leaver; 3*migCapacity ints of sort workspace
    void* dMigTemp = nullptr; size_t migTempBytes = 0;
    int32_t migCapacity = 0;
    std::vector<double*> dInflowAcc;
    std::vector<int32_t*> dInflowCounts;
    int32_t* dInflowScan = nullptr;
    int nQ = 0, nWallQ = 0, nMeasFaces = 0;
    double nTimeSteps = 0;
    uint32_t step = 0;
    int64_t nextOrigId = 0;   // ids handed out by this rank; stored modulo 2^31 (with origProc the parcel's identity)
    bool occupancyValid = false;   // cellOffset AND the sub-cell keys describe the cloud (collide / sample may run)
    bool csrValid = false;         // parcels [0, sortedN) are in the cell order of cellOffset (move may use its work list)
    int64_t sortedN = 0;
    // counters of the last evolve
    dsmcb200_counters last{};
    // neighbours
    std::vector<int> nbrProcs;                       // slot -> proc
    std::vector<std::vector<int32_t>> ordinalToPatch;  // slot -> ordinal -> patch
    int32_t* dOrdinalToPatch = nullptr;              // [MAX_NEIGHBOURS][MAX_PATCHES]
    int32_t* dCountsMatrix = nullptr;                // [nRanks*nRanks]
    NcclComm comm = nullptr;
    // timing
    std::map<std::string, std::pair<double, int64_t>> ktimes;
    std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
    std::vector<cudaEvent_t> evPool;
    size_t evUsed = 0;
    bool timeKernels = true;
    cudaEvent_t tmr0 = nullptr, tmr1 = nullptr;
};

namespace {

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"; \
            return DSMCB200_ERR_CUDA;                                                                     \
        }                                                                                                 \
    } while (0)

int fail(dsmcb200_ctx* c, int code, const std::string& msg) { c->err = msg; return code; }

// host threads for the O(n) loops of upload / download (launchers such as torchrun export OMP_NUM_THREADS=1, so the team is sized here:
// the box's cores shared among the ranks of this node, at most 16)
int hostThreads(const dsmcb200_ctx* c) {
    if (const char* e = std::getenv("DSMCB200_HOST_THREADS")) return std::max(1, std::atoi(e));
    const unsigned hw = std::thread::hardware_concurrency();
    return int(std::max(1u, std::min(16u, (hw ? hw : 1u) / unsigned(std::max(1, c->nRanks)))));
}

template <class T>
cudaError_t devAlloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T)); }
template <class T>
void devFree(T*& p) { if (p) cudaFree(p); p = nullptr; }

template <class T>
cudaError_t upload(T** d, const std::vector<T>& h) {
    cudaError_t e = devAlloc(d, h.size());
    if (e != cudaSuccess) return e;
    if (!h.empty()) e = cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

cudaEvent_t getEvent(dsmcb200_ctx* c) {
    if (c->evUsed == c->evPool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->evPool.push_back(e);
    }
    return c->evPool[c->evUsed++];
}
struct KT {  // scoped kernel timer: events on the launching stream
    dsmcb200_ctx* c; const char* name; cudaEvent_t a, b;
    KT(dsmcb200_ctx* c_, const char* n) : c(c_), name(n) {
        if (c->timeKernels) { a = getEvent(c); b = getEvent(c); cudaEventRecord(a, c->stream); }
    }
    ~KT() {
        if (c->timeKernels) { cudaEventRecord(b, c->stream); c->pending.push_back({name, {a, b}}); }
    }
};
void resolveTimers(dsmcb200_ctx* c) {
    for (auto& p : c->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.second.first, p.second.second) == cudaSuccess) {
            auto& e = c->ktimes[p.first];
            e.first += ms; e.second += 1;
        }
    }
    c->pending.clear();
    c->evUsed = 0;
}

CellFields cellFields(const dsmcb200_ctx* c) { return CellFields{c->dNPts, c->dDt, c->dRWF}; }

void setArrays(ParcelBuffer& b, int64_t cap, int nModes, bool internal, bool useCls, bool useOrigProc, bool useRwf) {
    ParcelArrays& a = b.a;
    a.px = b.dslab; a.py = a.px + cap; a.pz = a.py + cap; a.ux = a.pz + cap; a.uy = a.ux + cap; a.uz = a.uy + cap;
    a.erot = internal ? a.uz + cap : nullptr;
    a.cell = b.islab; a.tet = a.cell + cap; a.origId = a.tet + cap;
    for (int m = 0; m < MAX_MODES; ++m) a.vib[m] = (internal && m < nModes) ? a.origId + cap * (1 + m) : nullptr;
    a.typeId = b.bslab; a.elevel = internal ? b.bslab + cap : nullptr; a.cls = useCls ? b.bslab + 2 * cap : nullptr;
    a.origProc = useOrigProc ? b.bslab + 3 * cap : nullptr;
    a.rwf = useRwf ? b.dslab + 7 * cap : nullptr;
}

int allocBuffer(dsmcb200_ctx* c, ParcelBuffer& b, int64_t cap) {
    // one extra double row and one extra int row so that every buffer can stage host AoS arrays
    CK(devAlloc(&b.dslab, size_t(cap) * (c->useRwf ? 8 : 7)));
    CK(devAlloc(&b.islab, size_t(cap) * (3 + MAX_MODES)));
    CK(devAlloc(&b.bslab, size_t(cap) * 4));
    setArrays(b, cap, c->nModes, c->internal, c->useCls, c->nRanks > 1, c->useRwf);
    return 0;
}

int ensureCapacity(dsmcb200_ctx* c, int64_t n) {
    if (n <= c->capacity) return 0;
    if (n >= (int64_t(1) << 31) - 1024) return fail(c, DSMCB200_ERR_CAPACITY, "more than 2^31 parcels on one GPU");
    int64_t cap = std::max<int64_t>(n, c->capacity + c->capacity / 4 + 4096);
    ParcelBuffer nb[2];
    for (int k = 0; k < 2; ++k) { int r = allocBuffer(c, nb[k], cap); if (r) return r; }
    if (c->N > 0) {
        const ParcelArrays& o = c->buf[c->cur].a;
        const ParcelArrays& d = nb[c->cur].a;
        const size_t nb8 = size_t(c->N) * 8, nb4 = size_t(c->N) * 4, nb1 = size_t(c->N);
        double* const od[8] = {o.px, o.py, o.pz, o.ux, o.uy, o.uz, o.erot, o.rwf};
        double* const dd[8] = {d.px, d.py, d.pz, d.ux, d.uy, d.uz, d.erot, d.rwf};
        for (int k = 0; k < 8; ++k) if (od[k]) CK(cudaMemcpyAsync(dd[k], od[k], nb8, cudaMemcpyDeviceToDevice, c->stream));
        int32_t* const oi[3] = {o.cell, o.tet, o.origId};
        int32_t* const di[3] = {d.cell, d.tet, d.origId};
        for (int k = 0; k < 3; ++k) CK(cudaMemcpyAsync(di[k], oi[k], nb4, cudaMemcpyDeviceToDevice, c->stream));
        for (int m = 0; m < MAX_MODES; ++m) if (o.vib[m]) CK(cudaMemcpyAsync(d.vib[m], o.vib[m], nb4, cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaMemcpyAsync(d.typeId, o.typeId, nb1, cudaMemcpyDeviceToDevice, c->stream));
        if (o.elevel) CK(cudaMemcpyAsync(d.elevel, o.elevel, nb1, cudaMemcpyDeviceToDevice, c->stream));
        if (o.cls) CK(cudaMemcpyAsync(d.cls, o.cls, nb1, cudaMemcpyDeviceToDevice, c->stream));
        if (o.origProc) CK(cudaMemcpyAsync(d.origProc, o.origProc, nb1, cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    for (int k = 0; k < 2; ++k) {
        devFree(c->buf[k].dslab); devFree(c->buf[k].islab); devFree(c->buf[k].bslab);
        c->buf[k] = nb[k];
    }
    devFree(c->dPerm);
    CK(devAlloc(&c->dPerm, size_t(cap)));
    devFree(c->dOctKey);
    CK(devAlloc(&c->dOctKey, size_t(cap)));
    c->occupancyValid = false;   // the sub-cell keys of the sorted cloud went with the old buffer
    c->capacity = cap;
    return 0;
}

int ensureSfTail(dsmcb200_ctx* c, int64_t n) {
    if (n <= c->sfCap) return 0;
    int64_t cap = std::max<int64_t>(n, c->sfCap * 2 + 4096);
    double* nd = nullptr;
    CK(devAlloc(&nd, size_t(cap)));
    if (c->dSfTail && c->sfCap) CK(cudaMemcpy(nd, c->dSfTail, size_t(c->sfCap) * 8, cudaMemcpyDeviceToDevice));
    devFree(c->dSfTail);
    c->dSfTail = nd;
    c->sfCap = cap;
    return 0;
}

// Per-cell rows between the caller's cell labels and the engine's (dsmcb200_set_cell_order).  Identity order: a plain copy.
cudaError_t cellRowsToDevice(dsmcb200_ctx* c, void* dev, const void* user, size_t rowBytes) {
    const size_t nC = size_t(c->mesh.nCells);
    if (c->newOfOld.empty()) return cudaMemcpy(dev, user, nC * rowBytes, cudaMemcpyHostToDevice);
    std::vector<char> tmp(nC * rowBytes);
    const char* u = static_cast<const char*>(user);
#pragma omp parallel for schedule(static) num_threads(hostThreads(c))
    for (int64_t k = 0; k < int64_t(nC); ++k) std::memcpy(&tmp[size_t(c->newOfOld[k]) * rowBytes], u + size_t(k) * rowBytes, rowBytes);
    return cudaMemcpy(dev, tmp.data(), nC * rowBytes, cudaMemcpyHostToDevice);
}
cudaError_t cellRowsToHost(dsmcb200_ctx* c, void* user, const void* dev, size_t rowBytes) {
    const size_t nC = size_t(c->mesh.nCells);
    if (c->newOfOld.empty()) return cudaMemcpy(user, dev, nC * rowBytes, cudaMemcpyDeviceToHost);
    std::vector<char> tmp(nC * rowBytes);
    cudaError_t e = cudaMemcpy(tmp.data(), dev, nC * rowBytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return e;
    char* u = static_cast<char*>(user);
#pragma omp parallel for schedule(static) num_threads(hostThreads(c))
    for (int64_t k = 0; k < int64_t(nC); ++k) std::memcpy(u + size_t(k) * rowBytes, &tmp[size_t(c->newOfOld[k]) * rowBytes], rowBytes);
    return cudaSuccess;
}

// the sums of the sample set the accumulator entry points act on (dsmcb200_select_sample_set)
double*& selAcc(dsmcb200_ctx* c) { return c->selectedSet == 0 ? c->dAcc : c->extraSets[c->selectedSet - 1].dAcc; }
double*& selColl(dsmcb200_ctx* c) { return c->selectedSet == 0 ? c->dCollCum : c->extraSets[c->selectedSet - 1].dCollCum; }
double*& selWall(dsmcb200_ctx* c) { return c->selectedSet == 0 ? c->dWallAcc : c->extraSets[c->selectedSet - 1].dWallAcc; }
double& selSteps(dsmcb200_ctx* c) { return c->selectedSet == 0 ? c->nTimeSteps : c->extraSets[c->selectedSet - 1].nTimeSteps; }

int uploadCellFields(dsmcb200_ctx* c) {
    const size_t nC = size_t(c->mesh.nCells);
    struct { std::vector<double>* h; double** d; } f[3] = {{&c->hNPts, &c->dNPts}, {&c->hDt, &c->dDt}, {&c->hRWF, &c->dRWF}};
    for (auto& x : f) {
        devFree(*x.d);
        if (x.h->empty()) continue;
        if (x.h->size() != nC) return fail(c, DSMCB200_ERR_INVALID, "set_cell_fields: array size differs from the number of cells");
        if (c->newOfOld.empty()) { CK(upload(x.d, *x.h)); }
        else { CK(devAlloc(x.d, nC)); CK(cellRowsToDevice(c, *x.d, x.h->data(), 8)); }
    }
    return 0;
}

// Everything that depends on mesh + species + models: device parameter block, tet table, accumulators.
int finalize(dsmcb200_ctx* c) {
    if (c->ready) return 0;
    if (!c->haveMesh || !c->haveSpecies || !c->haveModels) return fail(c, DSMCB200_ERR_STATE, "set_mesh, set_species and set_models must all be called first");
    HostMesh& M = c->mesh;
    DevParams& P = c->hP;
    std::memset(&P, 0, sizeof(P));
    const dsmcb200_models& md = c->models;
    P.nSpecies = int(c->species.size());
    P.nPatches = int(M.patches.size());
    if (P.nPatches > MAX_PATCHES) return fail(c, DSMCB200_ERR_CAPACITY, "more than 64 patches");
    P.collisionModel = md.collisionModel;
    P.invZvFormulation = md.invZvFormulation;
    if (md.coordinateSystem != DSMCB200_COORD_CARTESIAN && md.coordinateSystem != DSMCB200_COORD_AXISYMMETRIC && md.coordinateSystem != DSMCB200_COORD_SPHERICAL)
        return fail(c, DSMCB200_ERR_UNSUPPORTED, "dsmcCoordinateSystem::New(const dictionary&) : \n    unknown dsmcCoordinateSystem type " + std::to_string(md.coordinateSystem) +
                    ", constructor not in hash table\n\n    Valid coordinate system types are :\n3(dsmcAxisymmetric dsmcCartesian dsmcSpherical)");
    P.coordinateSystem = md.coordinateSystem;
    P.angularCoordinate = md.angularCoordinate;
    c->useRwf = md.coordinateSystem == DSMCB200_COORD_AXISYMMETRIC || md.coordinateSystem == DSMCB200_COORD_SPHERICAL;
    if (md.coordinateSystem == DSMCB200_COORD_AXISYMMETRIC) {
        if (md.angularCoordinate < 0 || md.angularCoordinate > 2) return fail(c, DSMCB200_ERR_INVALID, "dsmcAxisymmetric: angularCoordinate must be 0, 1 or 2");
    }
    P.kB = md.kB > 0 ? md.kB : 1.38065e-23;  // OpenFOAM v1706 physicoChemical::k (pinned by shipped couette fields)
    P.Tref = md.Tref > 0 ? md.Tref : 273.0;
    P.nParticles = md.nEquivalentParticles;
    P.deltaT = md.deltaT;
    P.seed = md.seed;
    P.invZrot = 1.0 / (md.rotationalRelaxationCollisionNumber > 0 ? md.rotationalRelaxationCollisionNumber : 5.0);
    P.Zvib = md.vibrationalRelaxationCollisionNumber;
    P.invZelec = 1.0 / (md.electronicRelaxationCollisionNumber > 0 ? md.electronicRelaxationCollisionNumber : 500.0);
    P.measureFlux = md.measureHeatFluxShearStress ? 1 : 0;
    P.measureClass = md.measureClassifications ? 1 : 0;
    P.nInternalFaces = M.nInternalFaces;
    for (int d = 0; d < 3; ++d) {
        P.solutionD[d] = M.solutionD[d];
        P.centre[d] = 0.5 * (comp(M.boundsMin, d) + comp(M.boundsMax, d));  // meshTools::constrainToMeshCentre
    }
    int nModes = 0;
    bool internal = md.collisionModel == DSMCB200_COLL_LB_VHS || md.collisionModel == DSMCB200_COLL_LB_VSS;
    for (int s = 0; s < P.nSpecies; ++s) {
        const dsmcb200_species& h = c->species[s];
        DevSpecies& d = P.sp[s];
        d.mass = h.mass; d.d = h.diameter; d.omega = h.omega; d.alpha = h.alpha; d.rotDof = h.rotationalDegreesOfFreedom;
        d.thetaD = h.thetaD; d.nVib = h.nVibrationalModes; d.charge = h.charge; d.nElec = std::max(1, h.nElectronicLevels);
        for (int m = 0; m < MAX_MODES; ++m) { d.thetaV[m] = h.thetaV[m]; d.Zref[m] = h.Zref[m]; d.TrefZv[m] = h.TrefZv[m]; }
        for (int l = 0; l < MAX_ELEC; ++l) { d.eElec[l] = l < d.nElec ? h.electronicEnergyList[l] : 0.0; d.gElec[l] = l < d.nElec ? h.electronicDegeneracyList[l] : 0; }
        if (h.nElectronicLevels <= 1) { d.eElec[0] = h.nElectronicLevels == 1 ? h.electronicEnergyList[0] : 0.0; d.gElec[0] = h.nElectronicLevels == 1 ? h.electronicDegeneracyList[0] : 1; }
        // dsmcParcel::constantProperties type_ (dsmcParcelI.H:158-199)
        if (d.charge == -1) d.type = 0;
        else d.type = (d.nVib == 0 ? 10 : (d.nVib == 1 ? 20 : 30)) + (d.charge == 1 ? 1 : 0);
        nModes = std::max(nModes, d.nVib);
        if (d.rotDof > 0 || d.nVib > 0 || d.nElec > 1) internal = true;
    }
    for (int p = 0; p < P.nSpecies; ++p)
        for (int q = 0; q < P.nSpecies; ++q) {
            const double dPQ = 0.5 * (P.sp[p].d + P.sp[q].d);
            const double omegaPQ = 0.5 * (P.sp[p].omega + P.sp[q].omega);
            const double mP = P.sp[p].mass, mQ = P.sp[q].mass;
            P.vhsA[p][q] = PI * (dPQ * dPQ);
            P.vhsG[p][q] = std::exp(std::lgamma(2.5 - omegaPQ));
            P.omegaPQ[p][q] = omegaPQ;
            P.mR[p][q] = mP * mQ / (mP + mQ);
        }
    P.nModes = nModes;
    P.hasInternalEnergy = internal ? 1 : 0;
    c->nModes = nModes; c->internal = internal; c->useCls = md.measureClassifications != 0;

    // patches, neighbours
    c->nbrProcs.clear(); c->ordinalToPatch.clear();
    for (int p = 0; p < P.nPatches; ++p) {
        const PatchInfo& pi = M.patches[p];
        DevPatch& d = P.patch[p];
        d.type = pi.type; d.model = DSMCB200_BND_NONE; d.start = pi.start; d.size = pi.size; d.nbrPatch = pi.neighbPatch;
        d.nbrSlot = -1; d.nbrOrdinal = -1; d.T = 0;
        d.sep[0] = pi.separation.x; d.sep[1] = pi.separation.y; d.sep[2] = pi.separation.z;
        if (pi.type == DSMCB200_PATCH_PROCESSOR || pi.type == DSMCB200_PATCH_PROCESSORCYCLIC) {
            int slot = -1;
            for (size_t k = 0; k < c->nbrProcs.size(); ++k) if (c->nbrProcs[k] == pi.neighbProcNo) slot = int(k);
            if (slot < 0) { slot = int(c->nbrProcs.size()); c->nbrProcs.push_back(pi.neighbProcNo); c->ordinalToPatch.emplace_back(); }
            if (slot >= MAX_NEIGHBOURS) return fail(c, DSMCB200_ERR_CAPACITY, "more than 16 neighbour processors");
            d.nbrSlot = slot;
            d.nbrOrdinal = int(c->ordinalToPatch[slot].size());
            c->ordinalToPatch[slot].push_back(p);
        }
    }
    std::vector<int32_t> measIndex(M.nFaces - M.nInternalFaces, -1);
    c->nMeasFaces = 0;
    c->cllWalls = false;
    for (const dsmcb200_patch_model& pm : c->patchModels) {
        if (pm.patch < 0 || pm.patch >= P.nPatches) return fail(c, DSMCB200_ERR_INVALID, "patch model refers to an unknown patch");
        DevPatch& d = P.patch[pm.patch];
        if (d.type != DSMCB200_PATCH_WALL && d.type != DSMCB200_PATCH_PATCH)
            return fail(c, DSMCB200_ERR_INVALID, "Patch: " + M.patches[pm.patch].name + " must be of type wall or patch to carry a dsmcPatchBoundary model");
        if (pm.model < DSMCB200_BND_DIFFUSE_WALL || pm.model > DSMCB200_BND_CLL_WALL) return fail(c, DSMCB200_ERR_UNSUPPORTED, "unknown dsmcPatchBoundary type");
        d.model = pm.model; d.T = pm.temperature; d.diffuseFraction = pm.diffuseFraction;
        d.linearT = pm.linearTemperature != 0; d.depthAxis = pm.depthAxis; d.Tformation = pm.formationLevelTemperature;
        if (pm.model == DSMCB200_BND_CLL_WALL) {
            const double aN = pm.normalAccommodationCoefficient, aT = pm.tangentialAccommodationCoefficient, aR = pm.rotationalEnergyAccommodationCoefficient;
            if (!(aN >= 0 && aN <= 1) || !(aT >= 0 && aT <= 2) || !(aR >= 0 && aR <= 1))
                return fail(c, DSMCB200_ERR_INVALID, "dsmcCLLWallPatch: accommodation coefficients out of range (normal, rotational in [0, 1], tangential in [0, 2])");
            d.alphaN = aN; d.alphaT = aT * (2.0 - aT); d.alphaR = aR;   // dsmcCLLWallPatch.C:131-133
            d.linearT = 0;
            c->cllWalls = true;
        }
        if (d.linearT) {
            if (pm.depthAxis < 0 || pm.depthAxis > 2) return fail(c, DSMCB200_ERR_INVALID, "dsmcDiffuseWallPatch: depthAxis must be x, y or z");
            d.maxDepth = comp(M.boundsMax, pm.depthAxis);                       // mesh.bounds() (dsmcDiffuseWallPatch.C:181-184)
            d.lengthPatch = d.maxDepth - comp(M.boundsMin, pm.depthAxis);
        }
        d.vel[0] = pm.velocity[0]; d.vel[1] = pm.velocity[1]; d.vel[2] = pm.velocity[2];
        if (pm.model != DSMCB200_BND_DELETION)
            for (int i = 0; i < d.size; ++i) measIndex[d.start - M.nInternalFaces + i] = c->nMeasFaces++;
    }
    // dsmcBoundaries::checkPatchBoundaryModels (dsmcBoundaries.C:455-520): every wall/patch needs a model
    for (int p = 0; p < P.nPatches; ++p) {
        const DevPatch& d = P.patch[p];
        if ((d.type == DSMCB200_PATCH_WALL || d.type == DSMCB200_PATCH_PATCH) && d.size > 0 && d.model == DSMCB200_BND_NONE)
            return fail(c, DSMCB200_ERR_INVALID, "patch '" + M.patches[p].name + "' has no dsmcPatchBoundary model (one model per non-coupled patch is mandatory)");
    }
    for (const dsmcb200_inflow& in : c->inflows) {
        if (in.patch < 0 || in.patch >= P.nPatches) return fail(c, DSMCB200_ERR_INVALID, "inflow refers to an unknown patch");
        if (in.nTypes <= 0 || in.nTypes > MAX_SPECIES) return fail(c, DSMCB200_ERR_INVALID, "Cannot have zero typeIds being inserted.");
        for (int i = 0; i < in.nTypes; ++i)
            if (in.typeIds[i] < 0 || in.typeIds[i] >= P.nSpecies) return fail(c, DSMCB200_ERR_INVALID, "inflow typeId out of range");
    }

    // ---- dsmcReactions: <model>::setProperties checks and the typeId-pair addressing (dsmcReactions.C:137-165)
    P.nReactions = int(c->reactions.size());
    for (int i = 0; i < MAX_SPECIES; ++i) for (int j = 0; j < MAX_SPECIES; ++j) P.pairReaction[i][j] = -1;
    if (P.nReactions > MAX_REACTIONS) return fail(c, DSMCB200_ERR_CAPACITY, "more than 32 reactions");
    for (int k = 0; k < P.nReactions; ++k) {
        const dsmcb200_reaction& in = c->reactions[k];
        DevReaction& R = P.reactions[k];
        const std::string head = "For reaction number " + std::to_string(k) + "\n";
        if (in.model < DSMCB200_REACT_DISSOCIATION_QK || in.model > DSMCB200_REACT_DISSOCIATION_EXCHANGE_QK)
            return fail(c, DSMCB200_ERR_UNSUPPORTED, "dsmcReaction::New(const dictionary&) : \n    unknown dsmc reaction model type " + std::to_string(in.model) +
                        ", constructor not in hash table\n\n    Valid reaction types are :\n3(dissociationQK exchangeQK dissociationExchangeQK)");
        R.model = in.model; R.allowSplitting = in.allowSplitting != 0; R.posMolReactant = -1;
        int rType[2];
        for (int r = 0; r < 2; ++r) {
            if (in.reactants[r] < 0 || in.reactants[r] >= P.nSpecies) return fail(c, DSMCB200_ERR_INVALID, head + "Cannot find type id: " + std::to_string(in.reactants[r]));
            R.reactants[r] = in.reactants[r];
            rType[r] = P.sp[R.reactants[r]].type;
            R.heatDissJ[r] = 0.0;
            R.dissProd[r][0] = in.dissociationProducts[r][0]; R.dissProd[r][1] = in.dissociationProducts[r][1];
        }
        if (R.model != DSMCB200_REACT_EXCHANGE_QK) {   // dissociationQK::setProperties, dissociationQK.C:44-195
            bool moleculeFound = false;
            for (int r = 0; r < 2; ++r) {
                const bool mol = rType[r] == 20 || rType[r] == 30;
                if (mol) { moleculeFound = true; R.heatDissJ[r] = P.kB * P.sp[R.reactants[r]].thetaD; }
            }
            if (!moleculeFound) return fail(c, DSMCB200_ERR_INVALID, head + "None of the reactants is a molecule.");
            for (int r = 0; r < 2; ++r) {
                const bool mol = rType[r] == 20 || rType[r] == 30;
                const bool hasProducts = R.dissProd[r][0] >= 0 || R.dissProd[r][1] >= 0;
                if (!mol && hasProducts) return fail(c, DSMCB200_ERR_INVALID, head + "Reactant " + std::to_string(R.reactants[r]) + " is not a molecule \nand therefore, there should be no dissociation products");
                if (mol && !(R.dissProd[r][0] >= 0 && R.dissProd[r][1] >= 0)) return fail(c, DSMCB200_ERR_INVALID, head + "Reactant " + std::to_string(R.reactants[r]) + " is a molecule \nand therefore, it should have dissociation products");
                if (!mol) continue;
                for (int q = 0; q < 2; ++q) {
                    const int pi = R.dissProd[r][q];
                    if (pi >= P.nSpecies) return fail(c, DSMCB200_ERR_INVALID, head + "Cannot find type id: " + std::to_string(pi));
                    const int pt = P.sp[pi].type;
                    if (rType[r] == 20 && pt != 10) return fail(c, DSMCB200_ERR_INVALID, head + "Dissociation product of a diatomic molecule must be an atom: " + std::to_string(pi));
                    if (rType[r] == 30 && pt != 20 && pt != 30 && pt != 10) return fail(c, DSMCB200_ERR_INVALID, head + "Dissociation product of a polyatomic molecule must be a diatomic/polyatomic molecule instead of " + std::to_string(pi));
                }
            }
        }
        if (R.model != DSMCB200_REACT_DISSOCIATION_QK) {   // exchangeQK::setProperties, exchangeQK.C:44-176
            bool mol = false, atom = false;
            for (int r = 0; r < 2; ++r) {
                if (rType[r] >= 20) { mol = true; R.posMolReactant = r; }
                else if (rType[r] == 10 || rType[r] == 11) atom = true;
                else return fail(c, DSMCB200_ERR_INVALID, head + "Reactant " + std::to_string(R.reactants[r]) + " is neither a molecule nor an atom");
            }
            if (!mol) return fail(c, DSMCB200_ERR_INVALID, head + "None of the reactants is a molecule.");
            if (!atom) return fail(c, DSMCB200_ERR_INVALID, head + "None of the reactants is an atom.");
            R.exchProd[0] = R.exchProd[1] = -1;
            for (int r = 0; r < 2; ++r) {
                const int pi = in.exchangeProducts[r];
                if (pi < 0 || pi >= P.nSpecies) return fail(c, DSMCB200_ERR_INVALID, head + "Cannot find type id: " + std::to_string(pi));
                const int pt = P.sp[pi].type;
                if (pt >= 20) R.exchProd[0] = pi;
                else if (pt == 10 || pt == 11) R.exchProd[1] = pi;
                else return fail(c, DSMCB200_ERR_INVALID, head + "Product " + std::to_string(pi) + " is neither a molecule nor an atom");
            }
            if (R.exchProd[0] < 0) return fail(c, DSMCB200_ERR_INVALID, head + "None of the products is a molecule.");
            if (R.exchProd[1] < 0) return fail(c, DSMCB200_ERR_INVALID, head + "None of the products is an atom.");
            R.heatExchJ = in.heatOfReactionExchange * P.kB;
            R.bCoeff = in.bCoeff;
            const double omegaPQ = 0.5 * (P.sp[R.reactants[0]].omega + P.sp[R.reactants[1]].omega);
            const double chiB = 2.5 - omegaPQ;
            R.aDash = in.aCoeff * (std::pow(chiB, in.bCoeff) * std::exp(std::lgamma(chiB)) / std::exp(std::lgamma(chiB + in.bCoeff)));
        }
    }
    for (int i = 0; i < P.nSpecies; ++i)
        for (int j = i; j < P.nSpecies; ++j) {
            int nModels = 0;
            for (int r = 0; r < P.nReactions; ++r) {
                const DevReaction& R = P.reactions[r];
                const int pi = R.reactants[0] == i ? 0 : (R.reactants[1] == i ? 1 : -1);   // findIndex: the first match
                const int qi = R.reactants[0] == j ? 0 : (R.reactants[1] == j ? 1 : -1);
                bool yes = false;
                if (pi != -1 && qi != -1) {
                    if (R.model == DSMCB200_REACT_EXCHANGE_QK) yes = pi != qi;
                    else yes = (pi == qi && R.reactants[0] == R.reactants[1]) || (pi != qi && R.reactants[0] != R.reactants[1]);
                }
                if (yes) { P.pairReaction[i][j] = int8_t(r); P.pairReaction[j][i] = int8_t(r); ++nModels; }
            }
            if (nModels > 1) return fail(c, DSMCB200_ERR_INVALID, "There is more than one reaction model specified for the typeId pair: " + std::to_string(i) + " and " + std::to_string(j));
        }
    c->reactionTotals.assign(size_t(3) * P.nReactions, 0);

    {
        // 1/Zv of dsmcCloud::postCollisionVibrationalEnergyLevel (dsmcCloud.C:1429-1504) depends on the species, the
        // partner (through omegaPQ) and the integer iMax only: tabulated on the host with the reference's expression
        std::vector<double> tab(size_t(P.nSpecies) * P.nSpecies * MAX_MODES * ZV_TABLE, 1.0);
        for (int s1 = 0; s1 < P.nSpecies; ++s1)
            for (int s2 = 0; s2 < P.nSpecies; ++s2)
                for (int m = 0; m < P.sp[s1].nVib; ++m)
                    for (int iMax = 1; iMax < ZV_TABLE; ++iMax) {
                        const DevSpecies& S = P.sp[s1];
                        const double omega = P.omegaPQ[s1][s2];
                        const double T = iMax * S.thetaV[m] / (3.5 - omega);
                        const double pow1 = std::pow(S.thetaD / T, 1. / 3.) - 1.0;
                        const double pow2 = std::pow(S.thetaD / S.TrefZv[m], 1. / 3.) - 1.0;
                        const double ZvP1 = std::pow(S.thetaD / T, omega);
                        const double ZvP2 = std::pow(S.Zref[m] * std::pow(S.thetaD / S.TrefZv[m], -omega), pow1 / pow2);
                        const double Zv = ZvP1 * ZvP2;
                        tab[((size_t(s1) * P.nSpecies + s2) * MAX_MODES + m) * ZV_TABLE + iMax] = P.invZvFormulation == 2 ? 1.0 / (5.0 * Zv) : 1.0 / Zv;
                    }
        double* dTab = nullptr;
        CK(upload(&dTab, tab));
        c->dZvTab = dTab;
        P.invZvTab = dTab;
    }
    CK(devAlloc(&c->dP, 1));
    CK(cudaMemcpy(c->dP, &P, sizeof(P), cudaMemcpyHostToDevice));

    // ---- bake and upload the tracking tables in chunks
    {
        // window of the move kernel: two blocks per SM share the 227 KB of shared memory
        // ring of MOVE_NBUF windows in the 227 KB of one SM (one persistent block per SM)
        int st = moveMaxStageTets();
        if (const char* e = std::getenv("DSMCB200_STAGE_TETS")) st = std::atoi(e);
        c->stageTets = std::max(0, std::min(st, moveMaxStageTets()));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, c->device));
        c->moveBlocks = prop.multiProcessorCount;
        M.buildStageGroups(std::max(1, c->stageTets));
        c->nGroups = int32_t(M.stageGroupCell.size()) - 1;
        CK(upload(&c->dGroupCell, M.stageGroupCell));
        CK(devAlloc(&c->dPlanSub, size_t(c->nGroups) + 2)); CK(devAlloc(&c->dPlanBase, size_t(c->nGroups) + 4));
    }
    const int64_t nT = M.nTets();
    CK(devAlloc(&c->dTets, size_t(nT)));
    {
        const int64_t chunk = 1 << 20;
        std::vector<TetRec> h(size_t(std::min<int64_t>(chunk, nT)));
        for (int64_t t0 = 0; t0 < nT; t0 += chunk) {
            const int64_t n = std::min<int64_t>(chunk, nT - t0);
            M.bakeTets(t0, n, h.data());
            CK(cudaMemcpy(c->dTets + t0, h.data(), size_t(n) * sizeof(TetRec), cudaMemcpyHostToDevice));
        }
    }
    {
        std::vector<BFaceRec> bf;
        M.bakeBFaces(bf);
        for (size_t i = 0; i < bf.size(); ++i) bf[i].measIndex = measIndex[i];
        CK(upload(&c->dBFaces, bf));
        std::vector<double> ar(bf.size() * 3);
        for (size_t i = 0; i < bf.size(); ++i) {
            const V3& s = M.faceAreas[M.nInternalFaces + i];
            ar[3 * i] = s.x; ar[3 * i + 1] = s.y; ar[3 * i + 2] = s.z;
        }
        CK(upload(&c->dBFaceArea, ar));
    }
    {
        auto flat = [](const std::vector<V3>& v) { std::vector<double> o(v.size() * 3); for (size_t i = 0; i < v.size(); ++i) { o[3 * i] = v[i].x; o[3 * i + 1] = v[i].y; o[3 * i + 2] = v[i].z; } return o; };
        CK(upload(&c->dPoints, flat(M.points)));
        CK(upload(&c->dCellCentres, flat(M.cellCentres)));
        CK(upload(&c->dFaceCentres, flat(M.faceCentres)));
        CK(upload(&c->dFaceAreas, flat(M.faceAreas)));
        CK(upload(&c->dCellVolumes, M.cellVolumes));
        CK(upload(&c->dFaceOffsets, M.faceOffsets));
        CK(upload(&c->dFacePoints, M.facePoints));
        CK(upload(&c->dOwner, M.owner));
        {
            double r2 = 0.0;
            for (int32_t f = 0; f < M.nFaces; ++f)
                for (int32_t k = M.faceOffsets[f]; k < M.faceOffsets[f + 1]; ++k) {
                    const V3 pt = M.points[M.facePoints[k]];
                    const V3 d0 = pt - M.cellCentres[M.owner[f]];
                    r2 = std::max(r2, dot(d0, d0));
                    if (f < M.nInternalFaces) { const V3 d1 = pt - M.cellCentres[M.neighbour[f]]; r2 = std::max(r2, dot(d1, d1)); }
                }
            c->cellRadius2Max = r2 * (1.0 + 1e-9);
        }
        if (!c->newOfOld.empty()) { CK(upload(&c->dNewOfOld, c->newOfOld)); CK(upload(&c->dOldOfNew, c->oldOfNew)); }
        CK(upload(&c->dNeighbour, M.neighbour));
        CK(upload(&c->dTetBasePtIs, M.tetBasePtIs));
        CK(upload(&c->dFaceTet0, M.faceTet0));
        CK(upload(&c->dCellTetStart, M.cellTetStart));
        CK(upload(&c->dCellFaceOffsets, M.cellFaceOffsets));
        CK(upload(&c->dCellFaces, M.cellFaces));
    }
    // ---- per-cell state and accumulators
    const size_t nC = size_t(M.nCells);
    CK(devAlloc(&c->dCellCount, nC + 1)); CK(devAlloc(&c->dCellOffset, nC + 1)); CK(devAlloc(&c->dCursor, nC + 1));
    CK(devAlloc(&c->dScanScratch, size_t(scanScratchInts(int32_t(nC + 1))) + 8));
    CK(devAlloc(&c->dSigma, nC)); CK(devAlloc(&c->dRem, nC)); CK(devAlloc(&c->dNColls, nC)); CK(devAlloc(&c->dCollSep, nC));
    CK(cudaMemset(c->dSigma, 0, nC * 8)); CK(cudaMemset(c->dNColls, 0, nC * 8)); CK(cudaMemset(c->dCollSep, 0, nC * 8));
    CK(cudaMemset(c->dCellOffset, 0, (nC + 1) * 4));
    fillRemainder<<<GRID(nC), 0, c->stream>>>(c->dRem, int32_t(nC), P.seed);
    c->nQ = 5 + (internal ? 2 + nModes : 0) + (P.measureFlux ? 12 : 0) + (P.measureClass ? 3 : 0);
    CK(devAlloc(&c->dAcc, nC * P.nSpecies * c->nQ)); CK(devAlloc(&c->dCollCum, nC * 2));
    CK(cudaMemset(c->dAcc, 0, nC * P.nSpecies * c->nQ * 8)); CK(cudaMemset(c->dCollCum, 0, nC * 16));
    c->nWallQ = WQ_EVIBMOD0 + nModes;
    CK(devAlloc(&c->dWallAcc, size_t(c->nMeasFaces) * P.nSpecies * c->nWallQ));
    CK(cudaMemset(c->dWallAcc, 0, std::max<size_t>(1, size_t(c->nMeasFaces) * P.nSpecies * c->nWallQ) * 8));
    if (c->setIntervals.size() > 1) {
        const size_t nW = size_t(c->nMeasFaces) * P.nSpecies * c->nWallQ;
        c->models.sampleInterval = c->setIntervals[0];
        c->extraSets.assign(c->setIntervals.size() - 1, dsmcb200_ctx::SampleSet{});
        for (size_t k = 0; k < c->extraSets.size(); ++k) {
            auto& e = c->extraSets[k];
            e.interval = std::max(1, int(c->setIntervals[k + 1]));
            CK(devAlloc(&e.dAcc, nC * P.nSpecies * c->nQ)); CK(devAlloc(&e.dCollCum, nC * 2)); CK(devAlloc(&e.dWallAcc, nW));
            CK(cudaMemset(e.dAcc, 0, nC * P.nSpecies * c->nQ * 8)); CK(cudaMemset(e.dCollCum, 0, nC * 16)); CK(cudaMemset(e.dWallAcc, 0, std::max<size_t>(1, nW) * 8));
        }
        CK(devAlloc(&c->dWallStep, nW));
        CK(cudaMemset(c->dWallStep, 0, std::max<size_t>(1, nW) * 8));
    }
    if (c->models.trackFaceFluxes) {
        CK(devAlloc(&c->dFaceFlux, size_t(2) * P.nSpecies * M.nFaces));
        CK(cudaMemset(c->dFaceFlux, 0, size_t(2) * P.nSpecies * M.nFaces * 8));
    }
    CK(devAlloc(&c->dCounters, 1)); CK(cudaMemset(c->dCounters, 0, sizeof(DevCounters)));
    CK(devAlloc(&c->dGiantList, size_t(GIANT_LIST)));
    { int r = uploadCellFields(c); if (r) return r; }
    CK(devAlloc(&c->dBad, 1));
    CK(devAlloc(&c->dInfo, 8)); CK(devAlloc(&c->dInfoScratch, size_t(infoScratchDoubles())));
    // inflow state
    size_t maxInflow = 1;
    for (const dsmcb200_inflow& in : c->inflows) {
        const size_t n = size_t(in.nTypes) * M.patches[in.patch].size;
        double* acc = nullptr; int32_t* cnt = nullptr;
        CK(devAlloc(&acc, n)); CK(cudaMemset(acc, 0, std::max<size_t>(n, 1) * 8));
        CK(devAlloc(&cnt, n + 1));
        c->dInflowAcc.push_back(acc); c->dInflowCounts.push_back(cnt);
        maxInflow = std::max(maxInflow, n + 1);
    }
    CK(devAlloc(&c->dInflowScan, maxInflow + 1));
    // migration buffers
    if (!c->nbrProcs.empty()) {
        std::vector<int32_t> o2p(size_t(MAX_NEIGHBOURS) * MAX_PATCHES, -1);
        for (size_t s = 0; s < c->ordinalToPatch.size(); ++s)
            for (size_t k = 0; k < c->ordinalToPatch[s].size(); ++k) o2p[s * MAX_PATCHES + k] = c->ordinalToPatch[s][k];
        CK(upload(&c->dOrdinalToPatch, o2p));
        CK(devAlloc(&c->dCountsMatrix, size_t(c->nRanks) * c->nRanks + c->nRanks));
    }
    CK(cudaStreamSynchronize(c->stream));
    c->ready = true;
    return 0;
}

int ensureMigBuffers(dsmcb200_ctx* c) {
    if (c->nbrProcs.empty()) return 0;
    const int32_t want = int32_t(std::max<int64_t>(1 << 16, c->capacity / 8));
    if (want <= c->migCapacity) return 0;
    devFree(c->dMigSend); devFree(c->dMigRecv); devFree(c->dMigKey); devFree(c->dMigWork); devFree(c->dMigRwfSend); devFree(c->dMigRwfRecv);
    if (c->dMigTemp) { cudaFree(c->dMigTemp); c->dMigTemp = nullptr; }
    CK(devAlloc(&c->dMigSend, size_t(want) * MAX_NEIGHBOURS));
    CK(devAlloc(&c->dMigRecv, size_t(want) * MAX_NEIGHBOURS));
    CK(devAlloc(&c->dMigKey, size_t(want) * MAX_NEIGHBOURS));
    CK(devAlloc(&c->dMigWork, size_t(want) * 3));
    if (c->useRwf) { CK(devAlloc(&c->dMigRwfSend, size_t(want) * MAX_NEIGHBOURS)); CK(devAlloc(&c->dMigRwfRecv, size_t(want) * MAX_NEIGHBOURS)); }
    c->migTempBytes = orderMigrantsTempBytes(want);
    CK(cudaMalloc(&c->dMigTemp, std::max<size_t>(c->migTempBytes, 16)));
    c->migCapacity = want;
    return 0;
}

// ---- stage 2: dsmcCloud::buildCellOccupancy ----
// parcels this step's dissociations may add to a cloud of n (the collide stage fails with a capacity error beyond it)
int64_t bornHeadroom(int64_t n) { return n / 8 + 4096; }

int ensureBornBuffers(dsmcb200_ctx* c) {
    const int64_t want = bornHeadroom(c->N);
    if (want <= c->bornCap) return 0;
    devFree(c->dBorn); devFree(c->dBornKeys); devFree(c->dBornIdx);
    if (c->dBornTemp) { cudaFree(c->dBornTemp); c->dBornTemp = nullptr; }
    const int32_t cap = int32_t(std::min<int64_t>(want + want / 4, (int64_t(1) << 30)));
    CK(devAlloc(&c->dBorn, size_t(cap)));
    CK(devAlloc(&c->dBornKeys, size_t(cap) * 2));
    CK(devAlloc(&c->dBornIdx, size_t(cap) * 2));
    c->bornTempBytes = orderBornTempBytes(cap);
    CK(cudaMalloc(&c->dBornTemp, std::max<size_t>(c->bornTempBytes, 16)));
    c->bornCap = cap;
    return 0;
}

int stageSort(dsmcb200_ctx* c, bool histogramDone) {
    const int32_t nCells = c->mesh.nCells;
    if (!c->reactions.empty()) {   // room for the products of this step's dissociations: the buffers must not move after the sort
        { int r = ensureCapacity(c, c->N + bornHeadroom(c->N)); if (r) return r; }
        { int r = ensureBornBuffers(c); if (r) return r; }
    }
    const int32_t nIn = int32_t(c->N);
    ParcelArrays& src = c->buf[c->cur].a;
    if (!histogramDone) {
        KT t(c, "histogram");
        CK(cudaMemsetAsync(c->dCellCount, 0, size_t(nCells + 1) * 4, c->stream));
        CK(launchHistogram(src.cell, nIn, c->dCellCount, c->stream));
    }
    { KT t(c, "scan"); CK(launchExclusiveScan(c->dCellCount, c->dCellOffset, c->dCursor, nCells, c->dScanScratch, c->stream)); }
    int32_t nOut = 0;
    CK(cudaMemcpyAsync(&nOut, c->dCellOffset + nCells, 4, cudaMemcpyDeviceToHost, c->stream));
    { KT t(c, "scatterIndex"); CK(launchScatterIndex(src.cell, nIn, c->dCursor, c->dPerm, c->stream)); }
    { KT t(c, "segmentSort"); CK(launchSegmentSort(c->dCellOffset, nCells, c->dPerm, c->dCounters, c->dCursor, c->dGiantList, c->stream)); }
    int32_t nGiant = 0;
    CK(cudaMemcpyAsync(&nGiant, &c->dCounters->giantSortCells, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // nOut, nGiant
    c->nGiant = nGiant;
    if (nGiant > 0) {   // cells of more than 65 536 parcels (a heat bath in one cell)
        const int64_t words = int64_t(nIn) / 32 + 2;
        if (words > c->giantWords) {
            devFree(c->dGiantBitmap);
            c->giantWords = words + words / 4;
            CK(devAlloc(&c->dGiantBitmap, size_t(c->giantWords) * GIANT_SORT_BLOCKS));
        }
        KT t(c, "giantSort");
        CK(launchGiantSort(c->dCellOffset, c->dPerm, c->dGiantList, nGiant, c->dGiantBitmap, c->giantWords, c->stream));
    }
    ParcelArrays& dst = c->buf[1 - c->cur].a;
    { KT t(c, "gather"); CK(launchGather(src, dst, c->dPerm, c->dCellCentres, c->dOctKey, nOut, c->nModes, c->internal, c->stream)); }
    c->cur = 1 - c->cur;
    c->N = nOut;
    c->sortedN = nOut;
    c->occupancyValid = true;
    c->csrValid = true;
    return 0;
}

int stageInflow(dsmcb200_ctx* c, int64_t tailStart) {
    HostMesh& M = c->mesh;
    for (size_t k = 0; k < c->inflows.size(); ++k) {
        const dsmcb200_inflow& in = c->inflows[k];
        const PatchInfo& pi = M.patches[in.patch];
        if (pi.size == 0) continue;
        InflowArgs a{};
        a.nFaces = pi.size; a.patch = in.patch; a.patchStart = pi.start;
        a.faceOffsets = c->dFaceOffsets; a.facePoints = c->dFacePoints; a.owner = c->dOwner; a.tetBasePtIs = c->dTetBasePtIs;
        a.bfaces = c->dBFaces; a.nInternalFaces = M.nInternalFaces; a.points = c->dPoints; a.faceCentres = c->dFaceCentres; a.faceAreas = c->dFaceAreas;
        a.origProc = c->rank; a.cf = cellFields(c);
        a.P = c->dP; a.nTypes = in.nTypes; a.faceFlux = c->dFaceFlux; a.nFacesAll = M.nFaces;
        for (int i = 0; i < in.nTypes; ++i) { a.typeIds[i] = in.typeIds[i]; a.numberDensities[i] = in.numberDensities[i]; }
        for (int d = 0; d < 3; ++d) a.velocity[d] = in.velocity[d];
        a.Ttra = in.translationalTemperature; a.Trot = in.rotationalTemperature; a.Tvib = in.vibrationalTemperature; a.Telec = in.electronicTemperature;
        a.accumulator = c->dInflowAcc[k]; a.counts = c->dInflowCounts[k];
        a.counters = c->dCounters; a.step = c->step;
        const int32_t n = in.nTypes * pi.size;
        { KT t(c, "inflowCount"); CK(launchInflow(a, 0, c->stream)); }
        CK(launchExclusiveScan(a.counts, c->dInflowScan, nullptr, n, c->dScanScratch, c->stream));
        int32_t total = 0;
        CK(cudaMemcpyAsync(&total, c->dInflowScan + n, 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (total <= 0) continue;
        { int r = ensureCapacity(c, c->N + total); if (r) return r; }
        { int r = ensureSfTail(c, c->N + total - tailStart); if (r) return r; }
        a.p = c->buf[c->cur].a;
        a.counts = c->dInflowScan; a.base = int32_t(c->N); a.capacity = int32_t(c->capacity);
        a.sfTail = c->dSfTail; a.tailStart = int32_t(tailStart); a.origIdBase = int32_t(c->nextOrigId & 0x7fffffff);
        { KT t(c, "inflowInsert"); CK(launchInflow(a, 1, c->stream)); }
        c->N += total;
        c->nextOrigId += total;
        c->last.inserted += total;
    }
    return 0;
}

MoveArgs moveArgs(dsmcb200_ctx* c, int32_t tailStart) {
    MoveArgs a{};
    a.p = c->buf[c->cur].a; a.plan = c->dPlan; a.planTotal = c->dPlanBase + c->nGroups + 1; a.gridBlocks = c->moveBlocks; a.stageTets = c->stageTets;
    a.tailStart = tailStart; a.sfTail = c->dSfTail; a.cf = cellFields(c); a.weighted = c->useRwf ? 1 : 0;
    a.tets = c->dTets; a.bfaces = c->dBFaces; a.bfaceArea = c->dBFaceArea; a.P = c->dP; a.wallAcc = c->dWallAcc; a.nWallQ = c->nWallQ;
    // boundaryMeas_ is cleaned every step (dsmcCloud.C:924) but only folded into the fields on sampled steps (dsmcVolFields.C:1081,1292)
    a.wallsDue = c->sampleCounter + 1 >= std::max(1, c->models.sampleInterval);
    if (!c->extraSets.empty()) {   // several sample sets: the step's measurements go to dWallStep and from there to the sets that sample it
        for (auto& e : c->extraSets) a.wallsDue = a.wallsDue || (e.counter + 1 >= e.interval);
        a.wallAcc = c->dWallStep;
    }
    a.faceFlux = c->dFaceFlux; a.faceAreas = c->dFaceAreas; a.nFacesAll = c->mesh.nFaces;
    a.migBuf = c->dMigSend; a.migRwf = c->dMigRwfSend; a.migKey = c->dMigKey; a.migCapacity = c->migCapacity; a.cellCount = c->dCellCount; a.counters = c->dCounters; a.step = c->step;
    a.cllWalls = c->cllWalls ? 1 : 0;
    return a;
}

// Move parcels [first, last).  useCsr: parcels [0, sortedN) are still in the order of the last sort and take the per-step work list
// with shared-memory windows (requires first == 0); everything else is walked in plain pieces.
int launchMoveRange(dsmcb200_ctx* c, int32_t first, int32_t last, bool useCsr, int64_t tailStart, const char* timerName = "move") {
    if (last <= first) return 0;
    const int32_t sorted = (useCsr && first == 0 && c->csrValid && c->stageTets > 0) ? int32_t(std::min<int64_t>(c->sortedN, last)) : 0;
    const int32_t nGroups = sorted > 0 ? c->nGroups : 0;
    const int64_t need = int64_t(nGroups) + sorted / MOVE_PMAX + (last - std::max(first, sorted)) / MOVE_TAIL + 4;
    if (need > c->planCap) {
        devFree(c->dPlan);
        c->planCap = need + need / 4 + 64;
        CK(devAlloc(&c->dPlan, size_t(c->planCap)));
    }
    MovePlanArgs m{};
    m.groupCell = c->dGroupCell; m.nGroups = nGroups; m.cellOffset = c->dCellOffset; m.cellTetStart = c->dCellTetStart;
    m.nSub = c->dPlanSub; m.subBase = c->dPlanBase; m.maxTets = c->stageTets; m.plan = c->dPlan; m.planTotal = c->dPlanBase + c->nGroups + 1;
    m.tailBeg = std::max(first, sorted); m.tailEnd = last;
    { KT t(c, "movePlan"); CK(launchMovePlan(m, c->dScanScratch, c->stream)); }
    KT t(c, timerName);
    CK(launchMove(moveArgs(c, int32_t(tailStart)), c->stream));
    return 0;
}

int ncclFail(dsmcb200_ctx* c, int r, const char* what) {
    c->err = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
    return DSMCB200_ERR_NCCL;
}

// ---- stage 1 incl. processor-patch migration rounds ----
int stageMove(dsmcb200_ctx* c, int64_t tailStart) {
    const int32_t nCells = c->mesh.nCells;
    { int r = ensureMigBuffers(c); if (r) return r; }
    CK(cudaMemsetAsync(c->dCellCount, 0, size_t(nCells + 1) * 4, c->stream));
    CK(cudaMemsetAsync(c->dCounters->nMig, 0, sizeof(int32_t) * MAX_NEIGHBOURS, c->stream));
    if (tailStart >= c->N) { int r = ensureSfTail(c, 1); if (r) return r; }
    { int r = launchMoveRange(c, 0, int32_t(c->N), true, tailStart); if (r) return r; }
    c->csrValid = false;
    if (c->nRanks <= 1 || c->nbrProcs.empty()) return 0;
    if (!c->comm) return fail(c, DSMCB200_ERR_STATE, "mesh has processor patches but dsmcb200_init_comm was not called");

    const int R = c->nRanks;
    std::vector<int32_t> sendTo(R), matrix(size_t(R) * R);
    for (int round = 0; round < 1000; ++round) {
        KT t(c, "migrate");
        int32_t nMig[MAX_NEIGHBOURS];
        CK(cudaMemcpyAsync(nMig, c->dCounters->nMig, sizeof(nMig), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        std::fill(sendTo.begin(), sendTo.end(), 0);
        {
        KT tOrder(c, "migOrder");
        for (size_t s = 0; s < c->nbrProcs.size(); ++s) {
            if (nMig[s] > c->migCapacity) return fail(c, DSMCB200_ERR_CAPACITY, "migration buffer overflow");
            sendTo[c->nbrProcs[s]] = nMig[s];
            c->last.migratedTo[s] += nMig[s];
            // particleTransferLists[neighbour] in cloud-list order (Cloud.C:283-306); the receive slab of the slot is free until the
            // exchange below and serves as scratch
            CK(orderMigrants(c->dMigSend + s * c->migCapacity, c->dMigRecv + s * c->migCapacity, c->dMigKey + s * c->migCapacity, c->dMigWork,
                             c->dMigTemp, c->migTempBytes, nMig[s], c->stream,
                             c->dMigRwfSend ? c->dMigRwfSend + s * c->migCapacity : nullptr, c->dMigRwfRecv ? c->dMigRwfRecv + s * c->migCapacity : nullptr));
        }
        }
        cudaEvent_t evCounts0 = nullptr, evCounts1 = nullptr;
        if (c->timeKernels) { evCounts0 = getEvent(c); evCounts1 = getEvent(c); cudaEventRecord(evCounts0, c->stream); }
        // pBufs.finishedSends(allNTrans) + reduce(transfered, orOp) -> one all-gather of the send-count rows
        int32_t* dRow = c->dCountsMatrix + size_t(R) * R;
        CK(cudaMemcpyAsync(dRow, sendTo.data(), size_t(R) * 4, cudaMemcpyHostToDevice, c->stream));
        int r = g_nccl.AllGather(dRow, c->dCountsMatrix, size_t(R), NCCL_INT32, c->comm, c->stream);
        if (r) return ncclFail(c, r, "ncclAllGather");
        CK(cudaMemcpyAsync(matrix.data(), c->dCountsMatrix, size_t(R) * R * 4, cudaMemcpyDeviceToHost, c->stream));
        if (c->timeKernels) { cudaEventRecord(evCounts1, c->stream); c->pending.push_back({"migCounts", {evCounts0, evCounts1}}); }
        CK(cudaStreamSynchronize(c->stream));
        bool any = false;
        for (int32_t v : matrix) if (v) { any = true; break; }
        if (!any) break;
        c->last.migrationRounds += 1;
        int64_t nRecvTotal = 0;
        std::vector<int32_t> recvFrom(c->nbrProcs.size());
        for (size_t s = 0; s < c->nbrProcs.size(); ++s) {
            recvFrom[s] = matrix[size_t(c->nbrProcs[s]) * R + c->rank];
            c->last.migratedFrom[s] += recvFrom[s];
            if (recvFrom[s] > c->migCapacity) return fail(c, DSMCB200_ERR_CAPACITY, "migration receive buffer overflow");
            nRecvTotal += recvFrom[s];
        }
        { int rr = ensureCapacity(c, c->N + nRecvTotal); if (rr) return rr; }
        { int rr = ensureSfTail(c, c->N + nRecvTotal - tailStart + 1); if (rr) return rr; }
        {
        KT tExchange(c, "migExchange");
        r = g_nccl.GroupStart();
        if (r) return ncclFail(c, r, "ncclGroupStart");
        for (size_t s = 0; s < c->nbrProcs.size(); ++s) {
            if (nMig[s]) { r = g_nccl.Send(c->dMigSend + s * c->migCapacity, size_t(nMig[s]) * sizeof(MigRec), NCCL_CHAR, c->nbrProcs[s], c->comm, c->stream); if (r) return ncclFail(c, r, "ncclSend"); }
            if (recvFrom[s]) { r = g_nccl.Recv(c->dMigRecv + s * c->migCapacity, size_t(recvFrom[s]) * sizeof(MigRec), NCCL_CHAR, c->nbrProcs[s], c->comm, c->stream); if (r) return ncclFail(c, r, "ncclRecv"); }
            if (c->dMigRwfSend) {   // the radial weights travel next to the records
                if (nMig[s]) { r = g_nccl.Send(c->dMigRwfSend + s * c->migCapacity, size_t(nMig[s]) * sizeof(double), NCCL_CHAR, c->nbrProcs[s], c->comm, c->stream); if (r) return ncclFail(c, r, "ncclSend"); }
                if (recvFrom[s]) { r = g_nccl.Recv(c->dMigRwfRecv + s * c->migCapacity, size_t(recvFrom[s]) * sizeof(double), NCCL_CHAR, c->nbrProcs[s], c->comm, c->stream); if (r) return ncclFail(c, r, "ncclRecv"); }
            }
        }
        r = g_nccl.GroupEnd();
        if (r) return ncclFail(c, r, "ncclGroupEnd");
        }
        const int64_t firstNew = c->N;
        {
        KT tUnpack(c, "migUnpack");
        for (size_t s = 0; s < c->nbrProcs.size(); ++s) {
            if (!recvFrom[s]) continue;
            UnpackArgs u{};
            u.p = c->buf[c->cur].a; u.recv = c->dMigRecv + s * c->migCapacity; u.nRecv = recvFrom[s]; u.base = int32_t(c->N);
            u.recvRwf = c->dMigRwfRecv ? c->dMigRwfRecv + s * c->migCapacity : nullptr;
            u.sfTail = c->dSfTail; u.tailStart = int32_t(tailStart); u.ordinalToPatch = c->dOrdinalToPatch + s * MAX_PATCHES;
            u.bfaces = c->dBFaces; u.P = c->dP;
            CK(launchUnpack(u, c->stream));
            c->N += recvFrom[s];
            c->last.migratedIn += recvFrom[s];
        }
        }
        CK(cudaMemsetAsync(c->dCounters->nMig, 0, sizeof(int32_t) * MAX_NEIGHBOURS, c->stream));
        if (nRecvTotal) { int rr = launchMoveRange(c, int32_t(firstNew), int32_t(firstNew + nRecvTotal), false, tailStart, "moveArrivals"); if (rr) return rr; }
    }
    return 0;
}

// coordSystem().evolve() (dsmcCloud.C:884) = dsmcAxisymmetric::axisymmetricWeighting + reBuildCellOccupancy (dsmcAxisymmetric.C:477-487)
int stageWeighting(dsmcb200_ctx* c) {
    if (!c->useRwf || !c->dRWF || c->N == 0) return 0;
    if (!c->occupancyValid) return fail(c, DSMCB200_ERR_STATE, "radial weighting needs the cell occupancy: run the sort stage first");
    const int32_t n = int32_t(c->N);
    if (n + 1 > c->weightCountsCap) {
        devFree(c->dWeightCounts);
        c->weightCountsCap = int64_t(n) + n / 4 + 4096;
        CK(devAlloc(&c->dWeightCounts, size_t(c->weightCountsCap) * 2 + size_t(scanScratchInts(int32_t(c->weightCountsCap))) + 8));
    }
    int32_t* counts = c->dWeightCounts;
    int32_t* offsets = counts + c->weightCountsCap;
    int32_t* scratch = offsets + c->weightCountsCap;
    WeightArgs a{};
    a.p = c->buf[c->cur].a; a.cf = cellFields(c); a.n = n; a.base = n; a.capacity = int32_t(c->capacity);
    a.angularCoordinate = c->hP.coordinateSystem == DSMCB200_COORD_AXISYMMETRIC ? c->hP.angularCoordinate : -1; a.counts = counts; a.origIdBase = int32_t(c->nextOrigId & 0x7fffffff); a.origProc = c->rank;
    a.nModes = c->nModes; a.P = c->dP; a.counters = c->dCounters; a.step = c->step;
    CK(cudaMemsetAsync(&c->dCounters->weightDeleted, 0, sizeof(int32_t), c->stream));
    { KT t(c, "weighting"); CK(launchWeighting(a, 0, c->stream)); }
    CK(launchExclusiveScan(counts, offsets, nullptr, n, scratch, c->stream));
    int32_t total = 0, nDel = 0;
    CK(cudaMemcpyAsync(&total, offsets + n, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&nDel, &c->dCounters->weightDeleted, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (total == 0 && nDel == 0) return 0;
    if (total > 0) {
        if (c->N + total > c->capacity) {
            // the buffers move: the counts stay valid, the sorted order is kept by the copy
            { int r = ensureCapacity(c, c->N + total); if (r) return r; }
            a.p = c->buf[c->cur].a; a.capacity = int32_t(c->capacity);
        }
        a.counts = offsets;
        KT t(c, "weighting");
        CK(launchWeighting(a, 1, c->stream));
        c->N += total;
        c->nextOrigId += total;
    }
    c->cloned += total; c->weightDeleted += nDel;
    c->last.cloned = total; c->weightDeletedStep = nDel;
    return stageSort(c, false);   // cloud_.reBuildCellOccupancy()
}

int stageCollide(dsmcb200_ctx* c) {
    if (!c->occupancyValid) return fail(c, DSMCB200_ERR_STATE, "cell occupancy is stale: run the sort stage first");
    CollideArgs a{};
    a.p = c->buf[c->cur].a; a.cellOffset = c->dCellOffset; a.nCells = c->mesh.nCells; a.cellCentres = c->dCellCentres;
    a.cellVolumes = c->dCellVolumes; a.sigmaTcRMax = c->dSigma; a.remainder = c->dRem; a.nCollsStep = c->dNColls; a.collSepStep = c->dCollSep;
    a.overallT = c->hP.invZvFormulation == 1 ? c->dOverallT : nullptr;
    a.cf = cellFields(c);
    a.nModes = c->internal ? c->nModes : 0;
    a.bigScratch = c->dPerm; a.bigList = c->dCursor; a.octKey = c->dOctKey; a.giantList = c->dGiantList; a.nGiant = c->nGiant; a.P = c->dP; a.counters = c->dCounters; a.step = c->step;
    const bool chem = !c->reactions.empty();
    if (chem) {
        if (c->N != c->sortedN) return fail(c, DSMCB200_ERR_STATE, "collide stage with chemistry: the cloud was modified since the sort stage");
        { int r = ensureBornBuffers(c); if (r) return r; }
        a.born = c->dBorn;
        a.bornCapacity = int32_t(std::min<int64_t>(c->bornCap, c->capacity - c->N));
        CK(cudaMemsetAsync(&c->dCounters->nBorn, 0, sizeof(int32_t), c->stream));
        CK(cudaMemsetAsync(&c->dCounters->nReact[0][0], 0, sizeof(c->dCounters->nReact), c->stream));
    }
    CK(cudaMemsetAsync(&c->dCounters->bigCells, 0, sizeof(int32_t), c->stream));
    {
        KT t(c, "collide");
        CK(launchCollide(a, c->stream));
    }
    if (chem) {
        // dsmcCloud::addNewParcel: the second products join the cloud behind the sorted part, in (cell, candidate) order
        struct { int32_t nBorn; } h{};
        CK(cudaMemcpyAsync(&h.nBorn, &c->dCounters->nBorn, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        std::vector<unsigned long long> nReact(size_t(3) * MAX_REACTIONS);
        CK(cudaMemcpyAsync(nReact.data(), &c->dCounters->nReact[0][0], sizeof(unsigned long long) * 3 * MAX_REACTIONS, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (size_t k = 0; k < c->reactionTotals.size(); ++k) c->reactionTotals[k] += int64_t(nReact[k]);
        if (h.nBorn > a.bornCapacity) return fail(c, DSMCB200_ERR_CAPACITY, "more parcels created by dissociations in one step (" + std::to_string(h.nBorn) + ") than the reserve of " + std::to_string(a.bornCapacity));
        if (h.nBorn > 0) {
            KT t(c, "appendBorn");
            CK(launchAppendBorn(c->buf[c->cur].a, c->dBorn, h.nBorn, int32_t(c->N), int32_t(c->nextOrigId & 0x7fffffff), c->rank, c->nModes, c->dBornKeys,
                                c->dBornIdx, c->dBornTemp, c->bornTempBytes, c->stream));
            c->N += h.nBorn;
            c->nextOrigId += h.nBorn;
            c->last.inserted += 0;
        }
    }
    return 0;
}

int sampleInto(dsmcb200_ctx* c, double* acc, double* collCum, double& nTimeSteps) {
    if (!c->occupancyValid) return fail(c, DSMCB200_ERR_STATE, "cell occupancy is stale: run the sort stage first");
    SampleArgs a{};
    a.p = c->buf[c->cur].a; a.cellOffset = c->dCellOffset; a.nCells = c->mesh.nCells; a.acc = acc; a.nQ = c->nQ; a.nSpecies = int(c->species.size()); a.nParcels = int32_t(c->sortedN); a.nCloud = int32_t(c->N);
    a.collCum = collCum; a.nCollsStep = c->dNColls; a.collSepStep = c->dCollSep; a.P = c->dP;
    KT t(c, "sample");
    CK(launchSample(a, c->stream));
    nTimeSteps += 1.0;
    // cellMeas_.clean() (dsmcCloud.C:925): the per-step arrays are rewritten by the next collide stage
    return 0;
}

// the sample stage on its own (dsmcb200_stage): every set takes the sample
int stageSample(dsmcb200_ctx* c) {
    { int r = sampleInto(c, c->dAcc, c->dCollCum, c->nTimeSteps); if (r) return r; }
    for (auto& e : c->extraSets) { int r = sampleInto(c, e.dAcc, e.dCollCum, e.nTimeSteps); if (r) return r; }
    return 0;
}

// several sample sets: the wall measurements of a step are collected apart ...
int beginWallStep(dsmcb200_ctx* c) {
    if (c->extraSets.empty() || !c->nMeasFaces) return 0;
    CK(cudaMemsetAsync(c->dWallStep, 0, size_t(c->nMeasFaces) * c->hP.nSpecies * c->nWallQ * 8, c->stream));
    return 0;
}
// ... and added to the sets whose sampleInterval makes this step a sampled one (all of them for a move stage run on its own)
int endWallStep(dsmcb200_ctx* c, bool everySet) {
    if (c->extraSets.empty() || !c->nMeasFaces) return 0;
    const int32_t n = int32_t(size_t(c->nMeasFaces) * c->hP.nSpecies * c->nWallQ);
    if (everySet || c->sampleCounter + 1 >= std::max(1, c->models.sampleInterval)) addDoubles<<<GRID(n), 0, c->stream>>>(c->dWallAcc, c->dWallStep, n);
    for (auto& e : c->extraSets)
        if (everySet || e.counter + 1 >= e.interval) addDoubles<<<GRID(n), 0, c->stream>>>(e.dWallAcc, c->dWallStep, n);
    return 0;
}

int fetchCounters(dsmcb200_ctx* c) {
    CK(cudaMemcpyAsync(&c->hCounters, c->dCounters, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int dsmcb200_abi_version(void) { return DSMCB200_ABI_VERSION; }

int dsmcb200_create(dsmcb200_ctx** out, int device, int rank, int nRanks) {
    if (!out) return DSMCB200_ERR_INVALID;
    *out = nullptr;
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) return DSMCB200_ERR_CUDA;  // no CPU fallback
    if (device < 0 || device >= nDev) return DSMCB200_ERR_INVALID;
    dsmcb200_ctx* c = new dsmcb200_ctx();
    c->device = device; c->rank = rank; c->nRanks = nRanks < 1 ? 1 : nRanks;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return DSMCB200_ERR_CUDA;
    }
    *out = c;
    return 0;
}

void dsmcb200_destroy(dsmcb200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (int k = 0; k < 2; ++k) { devFree(c->buf[k].dslab); devFree(c->buf[k].islab); devFree(c->buf[k].bslab); }
    devFree(c->dP); devFree(c->dTets); devFree(c->dBFaces); devFree(c->dBFaceArea); devFree(c->dPoints); devFree(c->dCellCentres);
    devFree(c->dCellVolumes); devFree(c->dFaceCentres); devFree(c->dFaceAreas); devFree(c->dFaceOffsets); devFree(c->dFacePoints);
    devFree(c->dOwner); devFree(c->dNeighbour); devFree(c->dTetBasePtIs); devFree(c->dFaceTet0); devFree(c->dCellTetStart); devFree(c->dGroupCell); devFree(c->dPlanSub); devFree(c->dPlanBase); devFree(c->dPlan); devFree(c->dCellFaceOffsets); devFree(c->dCellFaces);
    devFree(c->dCellCount); devFree(c->dCellOffset); devFree(c->dCursor); devFree(c->dPerm); devFree(c->dOctKey); devFree(c->dScanScratch);
    devFree(c->dSigma); devFree(c->dRem); devFree(c->dNColls); devFree(c->dCollSep); devFree(c->dAcc); devFree(c->dCollCum);
    devFree(c->dOverallT); devFree(c->dFaceFlux); devFree(c->dWallAcc); devFree(c->dSfTail); devFree(c->dInfo); devFree(c->dInfoScratch); devFree(c->dCounters); devFree(c->dBad);
    devFree(c->dMigRwfSend); devFree(c->dMigRwfRecv);
    devFree(c->dZvTab); devFree(c->dMigSend); devFree(c->dMigRecv); devFree(c->dMigKey); devFree(c->dMigWork); devFree(c->dInflowScan);
    if (c->dMigTemp) cudaFree(c->dMigTemp); devFree(c->dOrdinalToPatch); devFree(c->dCountsMatrix);
    devFree(c->dBorn); devFree(c->dBornKeys); devFree(c->dBornIdx); if (c->dBornTemp) cudaFree(c->dBornTemp);
    devFree(c->dNPts); devFree(c->dDt); devFree(c->dRWF); devFree(c->dWeightCounts); devFree(c->dGiantBitmap); devFree(c->dGiantList);
    devFree(c->dNewOfOld); devFree(c->dOldOfNew); devFree(c->dWallStep);
    for (auto& e : c->extraSets) { devFree(e.dAcc); devFree(e.dCollCum); devFree(e.dWallAcc); }
    for (auto& p : c->dInflowAcc) devFree(p);
    for (auto& p : c->dInflowCounts) devFree(p);
    for (cudaEvent_t e : c->evPool) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

const char* dsmcb200_last_error(const dsmcb200_ctx* c) { return c ? c->err.c_str() : "null context"; }

int dsmcb200_nccl_unique_id(void* id128) {
    std::string err;
    if (!id128 || !g_nccl.load(err)) return DSMCB200_ERR_NCCL;
    NcclUniqueId id;
    if (g_nccl.GetUniqueId(&id)) return DSMCB200_ERR_NCCL;
    std::memcpy(id128, &id, sizeof(id));
    return 0;
}

int dsmcb200_init_comm(dsmcb200_ctx* c, const void* id128) {
    if (!c || !id128) return DSMCB200_ERR_INVALID;
    if (!g_nccl.load(c->err)) return DSMCB200_ERR_NCCL;
    cudaSetDevice(c->device);
    NcclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    int r = g_nccl.CommInitRank(&c->comm, c->nRanks, id, c->rank);
    if (r) return ncclFail(c, r, "ncclCommInitRank");
    return 0;
}

int dsmcb200_set_mesh(dsmcb200_ctx* c, const dsmcb200_mesh* m) {
    if (!c || !m) return DSMCB200_ERR_INVALID;
    if (c->ready) return fail(c, DSMCB200_ERR_STATE, "mesh cannot change after the engine has been finalised");
    c->newOfOld.clear(); c->oldOfNew.clear();
    if (c->cellOrderMode == DSMCB200_CELL_ORDER_AS_GIVEN) {
        std::string e = c->mesh.build(*m);
        if (!e.empty()) return fail(c, DSMCB200_ERR_INVALID, e);
        c->haveMesh = true;
        return 0;
    }
    // The engine works on the same polyMesh with its cells relabelled (owner / neighbour entries only: faces, points, patches and the
    // face lists of every cell stay as they are), so that neighbours in space are neighbours in memory; every entry point that takes
    // or returns cell labels or per-cell rows translates (what renumberMesh would do to the case files, without touching them).
    const int32_t nC = m->nCells;
    if (nC < 0 || m->nFaces < 0 || (m->nFaces && (!m->owner || !m->faceOffsets || !m->facePoints || !m->points)))
        return fail(c, DSMCB200_ERR_INVALID, "set_mesh: incomplete mesh");
    std::vector<int32_t> newOfOld(size_t(nC), 0);
    if (c->cellOrderMode == DSMCB200_CELL_ORDER_GIVEN) {
        if (c->userOrder.size() != size_t(nC)) return fail(c, DSMCB200_ERR_INVALID, "set_cell_order: the permutation has not the mesh's number of cells");
        newOfOld = c->userOrder;
    } else {
        // z-order curve through the cell centres (the caller's, or the mean of the face-point means of the cell's faces), 2^10 per direction
        std::vector<double> cc(size_t(nC) * 3, 0.0);
        if (m->cellCentres) cc.assign(m->cellCentres, m->cellCentres + size_t(nC) * 3);
        else {
            std::vector<double> cnt(size_t(nC), 0.0);
            for (int32_t f = 0; f < m->nFaces; ++f) {
                double fc[3] = {0, 0, 0};
                const int32_t b = m->faceOffsets[f], e = m->faceOffsets[f + 1];
                for (int32_t k = b; k < e; ++k) for (int d = 0; d < 3; ++d) fc[d] += m->points[3 * size_t(m->facePoints[k]) + d];
                for (int d = 0; d < 3; ++d) fc[d] /= double(std::max(1, e - b));
                const int32_t o = m->owner[f], nb = f < m->nInternalFaces ? m->neighbour[f] : -1;
                if (o < 0 || o >= nC || nb >= nC) return fail(c, DSMCB200_ERR_INVALID, "set_mesh: owner / neighbour out of range");
                for (int d = 0; d < 3; ++d) cc[3 * size_t(o) + d] += fc[d];
                cnt[o] += 1.0;
                if (nb >= 0) { for (int d = 0; d < 3; ++d) cc[3 * size_t(nb) + d] += fc[d]; cnt[nb] += 1.0; }
            }
            for (int32_t k = 0; k < nC; ++k) for (int d = 0; d < 3; ++d) cc[3 * size_t(k) + d] /= std::max(1.0, cnt[k]);
        }
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int32_t k = 0; k < nC; ++k) for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], cc[3 * size_t(k) + d]); hi[d] = std::max(hi[d], cc[3 * size_t(k) + d]); }
        double span = 0.0;
        for (int d = 0; d < 3; ++d) span = std::max(span, hi[d] - lo[d]);
        const double scale = span > 0 ? 1023.0 / span : 0.0;   // one scale for the three directions: the curve's bricks are cubes
        std::vector<std::pair<uint32_t, int32_t>> key(static_cast<size_t>(nC));
        for (int32_t k = 0; k < nC; ++k) {
            uint32_t q[3], z = 0;
            for (int d = 0; d < 3; ++d) q[d] = uint32_t(std::min(1023.0, std::max(0.0, std::floor((cc[3 * size_t(k) + d] - lo[d]) * scale))));
            for (int b = 0; b < 10; ++b) for (int d = 0; d < 3; ++d) z |= ((q[d] >> b) & 1u) << (3 * b + d);
            key[k] = {z, k};
        }
        std::sort(key.begin(), key.end());   // ties keep the caller's order (the label is the second key)
        for (int32_t r = 0; r < nC; ++r) newOfOld[key[r].second] = r;
    }
    std::vector<int32_t> oldOfNew(size_t(nC), -1);
    for (int32_t k = 0; k < nC; ++k) {
        const int32_t r = newOfOld[k];
        if (r < 0 || r >= nC || oldOfNew[r] >= 0) return fail(c, DSMCB200_ERR_INVALID, "set_cell_order: not a permutation of the cell labels");
        oldOfNew[r] = k;
    }
    dsmcb200_mesh mm = *m;
    std::vector<int32_t> owner(static_cast<size_t>(m->nFaces)), neighbour(static_cast<size_t>(m->nInternalFaces));
    for (int32_t f = 0; f < m->nFaces; ++f) {
        if (m->owner[f] < 0 || m->owner[f] >= nC) return fail(c, DSMCB200_ERR_INVALID, "set_mesh: owner out of range");
        owner[f] = newOfOld[m->owner[f]];
    }
    for (int32_t f = 0; f < m->nInternalFaces; ++f) {
        if (m->neighbour[f] < 0 || m->neighbour[f] >= nC) return fail(c, DSMCB200_ERR_INVALID, "set_mesh: neighbour out of range");
        neighbour[f] = newOfOld[m->neighbour[f]];
    }
    mm.owner = owner.data(); mm.neighbour = neighbour.data();
    std::vector<double> centres, volumes;
    if (m->cellCentres) {
        centres.resize(size_t(nC) * 3);
        for (int32_t k = 0; k < nC; ++k) for (int d = 0; d < 3; ++d) centres[3 * size_t(newOfOld[k]) + d] = m->cellCentres[3 * size_t(k) + d];
        mm.cellCentres = centres.data();
    }
    if (m->cellVolumes) {
        volumes.resize(static_cast<size_t>(nC));
        for (int32_t k = 0; k < nC; ++k) volumes[newOfOld[k]] = m->cellVolumes[k];
        mm.cellVolumes = volumes.data();
    }
    std::string e = c->mesh.build(mm);
    if (!e.empty()) return fail(c, DSMCB200_ERR_INVALID, e);
    c->newOfOld.swap(newOfOld); c->oldOfNew.swap(oldOfNew);
    c->haveMesh = true;
    return 0;
}

int dsmcb200_set_cell_order(dsmcb200_ctx* c, int mode, const int32_t* newOfOld, int32_t nCells) {
    if (!c) return DSMCB200_ERR_INVALID;
    if (c->haveMesh) return fail(c, DSMCB200_ERR_STATE, "set_cell_order: call it before set_mesh");
    if (mode != DSMCB200_CELL_ORDER_AS_GIVEN && mode != DSMCB200_CELL_ORDER_Z_CURVE && mode != DSMCB200_CELL_ORDER_GIVEN)
        return fail(c, DSMCB200_ERR_INVALID, "set_cell_order: unknown mode");
    if (mode == DSMCB200_CELL_ORDER_GIVEN) {
        if (!newOfOld || nCells < 0) return fail(c, DSMCB200_ERR_INVALID, "set_cell_order: a permutation is required");
        c->userOrder.assign(newOfOld, newOfOld + nCells);
    } else c->userOrder.clear();
    c->cellOrderMode = mode;
    return 0;
}

int dsmcb200_download_cell_order(dsmcb200_ctx* c, int32_t* newOfOld) {
    if (!c || !newOfOld) return DSMCB200_ERR_INVALID;
    if (!c->haveMesh) return fail(c, DSMCB200_ERR_STATE, "download_cell_order: call set_mesh first");
    for (int32_t k = 0; k < c->mesh.nCells; ++k) newOfOld[k] = c->newOfOld.empty() ? k : c->newOfOld[k];
    return 0;
}

int dsmcb200_set_species(dsmcb200_ctx* c, int nSpecies, const dsmcb200_species* sp) {
    if (!c || !sp) return DSMCB200_ERR_INVALID;
    if (c->ready) return fail(c, DSMCB200_ERR_STATE, "species cannot change after the engine has been finalised");
    if (nSpecies < 1 || nSpecies > MAX_SPECIES) return fail(c, DSMCB200_ERR_CAPACITY, "between 1 and 8 species are supported");
    for (int s = 0; s < nSpecies; ++s) {
        const dsmcb200_species& h = sp[s];
        if (!(h.mass > 0) || !(h.diameter > 0)) return fail(c, DSMCB200_ERR_INVALID, "species mass and diameter must be positive");
        if (h.nVibrationalModes < 0 || h.nVibrationalModes > MAX_MODES) return fail(c, DSMCB200_ERR_CAPACITY, "at most 3 vibrational modes per species");
        if (h.nElectronicLevels > MAX_ELEC) return fail(c, DSMCB200_ERR_CAPACITY, "at most 16 electronic levels per species");
        if (h.charge < -1 || h.charge > 1) return fail(c, DSMCB200_ERR_INVALID, "Charge value should be 0 for neutrals, 1 for ions, or -1 for electrons");
    }
    c->species.assign(sp, sp + nSpecies);
    c->haveSpecies = true;
    return 0;
}

int dsmcb200_set_models(dsmcb200_ctx* c, const dsmcb200_models* m) {
    if (!c || !m) return DSMCB200_ERR_INVALID;
    if (c->ready) return fail(c, DSMCB200_ERR_STATE, "models cannot change after the engine has been finalised");
    if (m->collisionModel < DSMCB200_COLL_NONE || m->collisionModel > DSMCB200_COLL_LB_VSS)
        return fail(c, DSMCB200_ERR_UNSUPPORTED, "unknown BinaryCollisionModel type; valid types are: NoBinaryCollision VariableHardSphere LarsenBorgnakkeVariableHardSphere");
    if (!(m->nEquivalentParticles > 0) || !(m->deltaT > 0)) return fail(c, DSMCB200_ERR_INVALID, "nEquivalentParticles and deltaT must be positive");
    c->models = *m;
    c->patchModels.assign(m->patchModels, m->patchModels + (m->patchModels ? m->nPatchModels : 0));
    c->inflows.assign(m->inflows, m->inflows + (m->inflows ? m->nInflows : 0));
    c->models.patchModels = nullptr; c->models.inflows = nullptr;
    c->haveModels = true;
    return 0;
}

int dsmcb200_set_cell_fields(dsmcb200_ctx* c, const double* nParticles, const double* deltaT, const double* RWF) {
    if (!c) return DSMCB200_ERR_INVALID;
    if (!c->haveMesh) return fail(c, DSMCB200_ERR_STATE, "set_cell_fields: call set_mesh first");
    const size_t nC = size_t(c->mesh.nCells);
    auto take = [&](std::vector<double>& h, const double* src, const char* what) -> int {
        h.clear();
        if (!src) return 0;
        for (size_t k = 0; k < nC; ++k) if (!(src[k] > 0)) return fail(c, DSMCB200_ERR_INVALID, std::string("set_cell_fields: ") + what + " must be positive in every cell");
        h.assign(src, src + nC);
        return 0;
    };
    { int r = take(c->hNPts, nParticles, "nParticles"); if (r) return r; }
    { int r = take(c->hDt, deltaT, "deltaT"); if (r) return r; }
    { int r = take(c->hRWF, RWF, "RWF"); if (r) return r; }
    if (c->ready) { cudaSetDevice(c->device); CK(cudaStreamSynchronize(c->stream)); return uploadCellFields(c); }
    return 0;
}

int dsmcb200_download_cell_fields(dsmcb200_ctx* c, double* nParticles, double* deltaT, double* RWF) {
    if (!c) return DSMCB200_ERR_INVALID;
    if (!c->haveMesh || !c->haveModels) return fail(c, DSMCB200_ERR_STATE, "download_cell_fields: call set_mesh and set_models first");
    const size_t nC = size_t(c->mesh.nCells);
    for (size_t k = 0; k < nC; ++k) {
        if (nParticles) nParticles[k] = c->hNPts.empty() ? c->models.nEquivalentParticles : c->hNPts[k];
        if (deltaT) deltaT[k] = c->hDt.empty() ? c->models.deltaT : c->hDt[k];
        if (RWF) RWF[k] = c->hRWF.empty() ? 1.0 : c->hRWF[k];
    }
    return 0;
}

int dsmcb200_set_reactions(dsmcb200_ctx* c, int n, const dsmcb200_reaction* reactions) {
    if (!c || n < 0 || (n > 0 && !reactions)) return DSMCB200_ERR_INVALID;
    if (c->ready) return fail(c, DSMCB200_ERR_STATE, "reactions cannot change after the engine has been finalised");
    if (n > MAX_REACTIONS) return fail(c, DSMCB200_ERR_CAPACITY, "more than 32 reactions");
    c->reactions.assign(reactions, reactions + n);   // checked against the species in finalize (dsmcReactions::initialConfiguration)
    return 0;
}

int dsmcb200_reaction_counts(dsmcb200_ctx* c, int n, int64_t* counts3n) {
    if (!c || !counts3n || n < 0) return DSMCB200_ERR_INVALID;
    if (n != int(c->reactions.size())) return fail(c, DSMCB200_ERR_INVALID, "reaction_counts: n differs from the number of reactions set");
    for (size_t k = 0; k < size_t(3) * n; ++k) counts3n[k] = k < c->reactionTotals.size() ? c->reactionTotals[k] : 0;
    return 0;
}

int dsmcb200_reserve(dsmcb200_ctx* c, int64_t maxParcels) {
    if (!c) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    return ensureCapacity(c, maxParcels);
}

int dsmcb200_upload_parcels(dsmcb200_ctx* c, int64_t n, const dsmcb200_parcels_soa* h) {
    if (!c || !h || n < 0) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    if (n && (!h->position || !h->U || !h->cell || !h->typeId)) return fail(c, DSMCB200_ERR_INVALID, "position, U, cell and typeId are required");
    { int r = ensureCapacity(c, std::max<int64_t>(n, 1)); if (r) return r; }
    c->N = 0;
    const int32_t n32 = int32_t(n);
    ParcelArrays& a = c->buf[c->cur].a;
    ParcelBuffer& st = c->buf[1 - c->cur];  // staging: the double slab holds 7*cap doubles, the int slab 6*cap ints
    cudaStream_t s = c->stream;
    c->fillsDone = 0;
    if (n == 0) { c->nextOrigId = 0; c->occupancyValid = false; return stageSort(c, false); }
    const int32_t *tetFace = h->tetFace, *tetPt = h->tetPt;
    const bool locate = !tetFace || !tetPt;   // the caller has no tet indices: located on the device below
    CK(cudaMemcpyAsync(st.dslab, h->position, size_t(n) * 24, cudaMemcpyHostToDevice, s));
    deinterleave3<<<GRID(n), 0, s>>>(st.dslab, a.px, a.py, a.pz, n32);
    CK(cudaMemcpyAsync(st.dslab, h->U, size_t(n) * 24, cudaMemcpyHostToDevice, s));
    deinterleave3<<<GRID(n), 0, s>>>(st.dslab, a.ux, a.uy, a.uz, n32);
    CK(cudaMemcpyAsync(a.cell, h->cell, size_t(n) * 4, cudaMemcpyHostToDevice, s));
    if (c->dNewOfOld) mapLabels<<<GRID(n), 0, s>>>(a.cell, a.cell, c->dNewOfOld, n32, c->mesh.nCells);   // the caller's cell labels -> the engine's
    int32_t* sf = st.islab;                 // staging rows
    int32_t* sp2 = st.islab + c->capacity;
    CK(cudaMemsetAsync(c->dBad, 0, 4, s));
    if (locate) {
        // particle::initCellFacePtOrDeleteLostParticle (BASIC/particle/particleI.H:851-996): first tet of the given cell that
        // contains the position; lost parcels get cell -1 and are dropped by the sort below
        LocateArgs l{};
        l.px = a.px; l.py = a.py; l.pz = a.pz; l.cell = a.cell; l.tet = a.tet; l.n = n32; l.nCells = c->mesh.nCells;
        l.cellFaceOffsets = c->dCellFaceOffsets; l.cellFaces = c->dCellFaces; l.faceOffsets = c->dFaceOffsets; l.facePoints = c->dFacePoints;
        l.owner = c->dOwner; l.neighbour = c->dNeighbour; l.nInternalFaces = c->mesh.nInternalFaces; l.tetBasePtIs = c->dTetBasePtIs; l.cellTetStart = c->dCellTetStart; l.points = c->dPoints;
        l.cellCentres = c->dCellCentres; l.lost = &c->dCounters->deleted;
        l.pending = st.islab; l.nPending = &c->dCounters->nPendingLocate; l.searchRadius2 = c->cellRadius2Max;
        CK(cudaMemsetAsync(&c->dCounters->deleted, 0, sizeof(unsigned long long), s));
        CK(cudaMemsetAsync(&c->dCounters->nPendingLocate, 0, sizeof(int32_t), s));
        CK(launchLocate(l, s));
        int32_t nPending = 0;
        CK(cudaMemcpyAsync(&nPending, &c->dCounters->nPendingLocate, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (nPending > LOCATE_GLOBAL_MAX)
            return fail(c, DSMCB200_ERR_INVALID, "upload_parcels: " + std::to_string(nPending) + " parcels are neither in nor next to the cell their label names "
                        "(the mesh-wide search of polyMesh::findCellFacePt is limited to " + std::to_string(LOCATE_GLOBAL_MAX) + " parcels per upload): wrong mesh or cell labels?");
        CK(launchLocateGlobal(l, nPending, s));
        unsigned long long lost = 0;
        CK(cudaMemcpyAsync(&lost, &c->dCounters->deleted, sizeof(lost), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        c->last.deleted += int64_t(lost);
    } else {
        CK(cudaMemcpyAsync(sf, tetFace, size_t(n) * 4, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(sp2, tetPt, size_t(n) * 4, cudaMemcpyHostToDevice, s));
        toTetId<<<GRID(n), 0, s>>>(a.cell, sf, sp2, c->dFaceTet0, c->dFaceOffsets, c->dOwner, c->dTets, a.tet, n32, c->mesh.nFaces, c->mesh.nCells, c->dBad);
    }
    CK(cudaMemcpyAsync(sf, h->typeId, size_t(n) * 4, cudaMemcpyHostToDevice, s));
    i32ToU8Checked<<<GRID(n), 0, s>>>(sf, a.typeId, n32, c->hP.nSpecies, c->dBad);
    if (h->origId) CK(cudaMemcpyAsync(a.origId, h->origId, size_t(n) * 4, cudaMemcpyHostToDevice, s));
    else iotaKernel<<<GRID(n), 0, s>>>(a.origId, n32, 0);
    if (c->internal) {
        if (h->ERot) CK(cudaMemcpyAsync(a.erot, h->ERot, size_t(n) * 8, cudaMemcpyHostToDevice, s));
        else CK(cudaMemsetAsync(a.erot, 0, size_t(n) * 8, s));
        for (int m = 0; m < c->nModes; ++m) {
            if (h->vibLevel && m < h->maxModes) {
                if (h->maxModes == 1) CK(cudaMemcpyAsync(a.vib[m], h->vibLevel, size_t(n) * 4, cudaMemcpyHostToDevice, s));
                else {
                    // parcel-major [n][maxModes] staged through the double slab of the other buffer
                    int32_t* stv = reinterpret_cast<int32_t*>(st.dslab);
                    CK(cudaMemcpyAsync(stv, h->vibLevel, size_t(n) * 4 * h->maxModes, cudaMemcpyHostToDevice, s));
                    stridedToMode<<<GRID(n), 0, s>>>(stv, a.vib[m], n32, h->maxModes, m);
                }
            } else CK(cudaMemsetAsync(a.vib[m], 0, size_t(n) * 4, s));
        }
        if (h->ELevel) { CK(cudaMemcpyAsync(sf, h->ELevel, size_t(n) * 4, cudaMemcpyHostToDevice, s)); i32ToU8<<<GRID(n), 0, s>>>(sf, a.elevel, n32); }
        else CK(cudaMemsetAsync(a.elevel, 0, size_t(n), s));
    }
    if (a.origProc) {
        if (h->origProc) { CK(cudaMemcpyAsync(sf, h->origProc, size_t(n) * 4, cudaMemcpyHostToDevice, s)); i32ToU8<<<GRID(n), 0, s>>>(sf, a.origProc, n32); }
        else CK(cudaMemsetAsync(a.origProc, c->rank & 0xff, size_t(n), s));
    }
    if (a.cls) {
        if (h->classification) { CK(cudaMemcpyAsync(sf, h->classification, size_t(n) * 4, cudaMemcpyHostToDevice, s)); i32ToU8<<<GRID(n), 0, s>>>(sf, a.cls, n32); }
        else CK(cudaMemsetAsync(a.cls, 0, size_t(n), s));
    }
    if (a.rwf) {   // dsmcParcel::RWF_: the lagrangian field radialWeight, or the weight of the parcel's cell
        if (h->radialWeight) CK(cudaMemcpyAsync(a.rwf, h->radialWeight, size_t(n) * 8, cudaMemcpyHostToDevice, s));
        else rwfOfCell<<<GRID(n), 0, s>>>(a.cell, c->dRWF, a.rwf, n32);
    }
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, c->dBad, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (bad == 2) return fail(c, DSMCB200_ERR_INVALID, "upload_parcels: typeId not defined in typeIdList");
    if (bad) return fail(c, DSMCB200_ERR_INVALID, "upload_parcels: cell / tetFace / tetPt out of range");
    c->N = n;
    int64_t maxId = -1;
    if (h->origId) {   // 2.5e8 entries in the end-to-end path: reduced by the host's cores, not by one of them
        int32_t mx = -1;
        const int32_t* ids = h->origId;
#pragma omp parallel for reduction(max : mx) schedule(static) num_threads(hostThreads(c))
        for (int64_t i = 0; i < n; ++i) mx = std::max(mx, ids[i]);
        maxId = mx;
    } else maxId = n32 - 1;
    c->nextOrigId = maxId + 1;
    c->occupancyValid = false;
    return stageSort(c, false);  // buildCellOccupancyFromScratch (dsmcCloud.C:677)
}

int dsmcb200_download_parcels(dsmcb200_ctx* c, int64_t capacity, int64_t* nOut, dsmcb200_parcels_soa* h) {
    if (!c || !nOut) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    *nOut = c->N;
    if (!h) return 0;
    if (capacity < c->N) return fail(c, DSMCB200_ERR_CAPACITY, "download_parcels: caller buffer too small");
    const int64_t n = c->N;
    if (n == 0) return 0;
    const int32_t n32 = int32_t(n);
    ParcelArrays& a = c->buf[c->cur].a;
    ParcelBuffer& st = c->buf[1 - c->cur];
    cudaStream_t s = c->stream;
    if (h->position) { interleave3<<<GRID(n), 0, s>>>(a.px, a.py, a.pz, st.dslab, n32); CK(cudaMemcpyAsync(h->position, st.dslab, size_t(n) * 24, cudaMemcpyDeviceToHost, s)); }
    if (h->U) { interleave3<<<GRID(n), 0, s>>>(a.ux, a.uy, a.uz, st.dslab + 3 * c->capacity, n32); CK(cudaMemcpyAsync(h->U, st.dslab + 3 * c->capacity, size_t(n) * 24, cudaMemcpyDeviceToHost, s)); }
    int32_t* r0 = st.islab; int32_t* r1 = st.islab + c->capacity; int32_t* r2 = r1 + c->capacity; int32_t* r3 = r2 + c->capacity; int32_t* r4 = r3 + c->capacity;
    if (h->cell) {
        if (c->dOldOfNew) { mapLabels<<<GRID(n), 0, s>>>(a.cell, r0, c->dOldOfNew, n32, c->mesh.nCells); CK(cudaMemcpyAsync(h->cell, r0, size_t(n) * 4, cudaMemcpyDeviceToHost, s)); }
        else CK(cudaMemcpyAsync(h->cell, a.cell, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
    }
    if (h->tetFace || h->tetPt) {
        fromTetId<<<GRID(n), 0, s>>>(a.tet, c->dTets, r0, r1, n32);
        if (h->tetFace) CK(cudaMemcpyAsync(h->tetFace, r0, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
        if (h->tetPt) CK(cudaMemcpyAsync(h->tetPt, r1, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
    }
    if (h->typeId) { u8ToI32<<<GRID(n), 0, s>>>(a.typeId, r2, n32); CK(cudaMemcpyAsync(h->typeId, r2, size_t(n) * 4, cudaMemcpyDeviceToHost, s)); }
    if (h->origId) CK(cudaMemcpyAsync(h->origId, a.origId, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
    if (h->ERot) { if (c->internal) CK(cudaMemcpyAsync(h->ERot, a.erot, size_t(n) * 8, cudaMemcpyDeviceToHost, s)); else std::memset(h->ERot, 0, size_t(n) * 8); }
    if (h->ELevel) { if (c->internal) { u8ToI32<<<GRID(n), 0, s>>>(a.elevel, r3, n32); CK(cudaMemcpyAsync(h->ELevel, r3, size_t(n) * 4, cudaMemcpyDeviceToHost, s)); } else std::memset(h->ELevel, 0, size_t(n) * 4); }
    if (h->classification) { if (a.cls) { u8ToI32<<<GRID(n), 0, s>>>(a.cls, r4, n32); CK(cudaMemcpyAsync(h->classification, r4, size_t(n) * 4, cudaMemcpyDeviceToHost, s)); } else std::memset(h->classification, 0, size_t(n) * 4); }
    if (h->origProc) {
        if (a.origProc) { CK(cudaStreamSynchronize(s)); u8ToI32<<<GRID(n), 0, s>>>(a.origProc, r4, n32); CK(cudaMemcpyAsync(h->origProc, r4, size_t(n) * 4, cudaMemcpyDeviceToHost, s)); }
        else {
            int32_t* op = h->origProc; const int32_t rk = c->rank;
#pragma omp parallel for schedule(static) num_threads(hostThreads(c))
            for (int64_t i = 0; i < n; ++i) op[i] = rk;
        }
    }
    if (h->newParcel) {
        int32_t* np_ = h->newParcel;
#pragma omp parallel for schedule(static) num_threads(hostThreads(c))
        for (int64_t i = 0; i < n; ++i) np_[i] = -1;
    }
    if (h->radialWeight) {
        if (a.rwf) CK(cudaMemcpyAsync(h->radialWeight, a.rwf, size_t(n) * 8, cudaMemcpyDeviceToHost, s));
        else {
            double* rw = h->radialWeight;
#pragma omp parallel for schedule(static) num_threads(hostThreads(c))
            for (int64_t i = 0; i < n; ++i) rw[i] = 1.0;
        }
    }
    if (h->vibLevel && h->maxModes > 0) {
        CK(cudaStreamSynchronize(s));
        int32_t* stv = reinterpret_cast<int32_t*>(st.dslab);  // [n][maxModes]
        CK(cudaMemsetAsync(stv, 0, size_t(n) * 4 * h->maxModes, s));
        for (int m = 0; m < std::min(c->nModes, h->maxModes); ++m) modeToStrided<<<GRID(n), 0, s>>>(a.vib[m], stv, n32, h->maxModes, m);
        CK(cudaMemcpyAsync(h->vibLevel, stv, size_t(n) * 4 * h->maxModes, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return 0;
}

int dsmcb200_upload_cellstate(dsmcb200_ctx* c, const double* sigma, const double* rem) {
    if (!c) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    const size_t nC = size_t(c->mesh.nCells);
    (void)nC;
    if (sigma) CK(cellRowsToDevice(c, c->dSigma, sigma, 8));
    if (rem) CK(cellRowsToDevice(c, c->dRem, rem, 8));
    return 0;
}

int dsmcb200_download_cellstate(dsmcb200_ctx* c, double* sigma, double* rem) {
    if (!c) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    const size_t nC = size_t(c->mesh.nCells);
    CK(cudaStreamSynchronize(c->stream));
    (void)nC;
    if (sigma) CK(cellRowsToHost(c, sigma, c->dSigma, 8));
    if (rem) CK(cellRowsToHost(c, rem, c->dRem, 8));
    return 0;
}

// dsmcMeshFill (zone == nullptr: the cloud is replaced) and dsmcZoneFill (the cells of one cellZone in the zone's order, appended)
static int fillCells(dsmcb200_ctx* c, const int32_t* zone, int64_t nZone, int nTypes, const int32_t* typeIds, const double* numberDensities, double Ttra,
                     double Trot, double Tvib, double Telec, const double velocity[3]) {
    if (!c || !typeIds || !numberDensities || nTypes < 1 || nTypes > MAX_SPECIES) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    for (int i = 0; i < nTypes; ++i)
        if (typeIds[i] < 0 || typeIds[i] >= c->hP.nSpecies) return fail(c, DSMCB200_ERR_INVALID, "mesh_fill: typeId not defined");
    const HostMesh& M = c->mesh;
    int32_t* dZone = nullptr;
    if (zone) {
        if (nZone < 0 || nZone > M.nCells) return fail(c, DSMCB200_ERR_INVALID, "zone_fill: more zone cells than cells");
        std::vector<int32_t> z(zone, zone + nZone);
        for (int32_t& k : z) {
            if (k < 0 || k >= M.nCells) return fail(c, DSMCB200_ERR_INVALID, "zone_fill: cell label outside the mesh");
            if (!c->newOfOld.empty()) k = c->newOfOld[k];   // dsmcb200_set_cell_order: the caller's label -> the engine's
        }
        if (nZone > 0) {
            CK(cudaMalloc(&dZone, size_t(nZone) * 4));
            CK(cudaMemcpyAsync(dZone, z.data(), size_t(nZone) * 4, cudaMemcpyHostToDevice, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
    } else {
        c->N = 0; c->nextOrigId = 0; c->fillsDone = 0;
    }
    const int32_t nFill = zone ? int32_t(nZone) : M.nCells;
    const int64_t base = c->N;
    FillArgs a{};
    a.nCells = M.nCells; a.cellFaceOffsets = c->dCellFaceOffsets; a.cellFaces = c->dCellFaces; a.faceOffsets = c->dFaceOffsets;
    a.facePoints = c->dFacePoints; a.owner = c->dOwner; a.tetBasePtIs = c->dTetBasePtIs; a.cellTetStart = c->dCellTetStart;
    a.points = c->dPoints; a.cellCentres = c->dCellCentres; a.P = c->dP; a.nTypes = nTypes;
    for (int i = 0; i < nTypes; ++i) { a.typeIds[i] = typeIds[i]; a.numberDensities[i] = numberDensities[i]; }
    a.Ttra = Ttra; a.Trot = Trot; a.Tvib = Tvib; a.Telec = Telec;
    for (int d = 0; d < 3; ++d) a.velocity[d] = velocity ? velocity[d] : 0.0;
    a.cellCount = c->dCellCount; a.origIdBase = int32_t(c->nextOrigId & 0x7fffffff); a.origProc = c->rank; a.cf = cellFields(c);
    a.nFill = nFill; a.cellList = dZone; a.slotBase = int32_t(base); a.fillIndex = uint32_t(c->fillsDone++);
    a.p = c->buf[c->cur].a;
    int32_t total = 0;
    if (nFill > 0) {
        CK(launchFill(a, 0, c->stream));
        CK(launchExclusiveScan(c->dCellCount, c->dCellOffset, nullptr, nFill, c->dScanScratch, c->stream));
        CK(cudaMemcpyAsync(&total, c->dCellOffset + nFill, 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    { int r = ensureCapacity(c, std::max<int64_t>(base + total, 1)); if (r) { if (dZone) cudaFree(dZone); return r; } }
    a.p = c->buf[c->cur].a;
    a.cellCount = c->dCellOffset;
    if (nFill > 0) CK(launchFill(a, 1, c->stream));
    c->N = base + total;
    c->nextOrigId += total;
    // sigmaTcRMax = sigmaT(most abundant) * most probable speed (dsmcMeshFill.C:224-235; dsmcZoneFill.C:246-268: the zone's cells only)
    int most = 0;
    for (int i = 1; i < nTypes; ++i) if (numberDensities[i] > numberDensities[most]) most = i;
    // the reference indexes constProps by dictionary position (SURVEY 8a quirk list); the typeId is used here
    const DevSpecies& S = c->hP.sp[typeIds[most]];
    const double sig = PI * S.d * S.d * std::sqrt(2.0 * c->hP.kB * Ttra / S.mass);
    if (!zone) fillDouble<<<GRID(M.nCells), 0, c->stream>>>(c->dSigma, M.nCells, sig);
    else if (nFill > 0) fillDoubleAt<<<GRID(nFill), 0, c->stream>>>(c->dSigma, dZone, nFill, sig);
    if (dZone) { CK(cudaStreamSynchronize(c->stream)); cudaFree(dZone); }
    // the fill writes the cloud in cell order; the sort leaves that order as it is and produces the occupancy arrays and sub-cell keys
    c->occupancyValid = false; c->csrValid = false;
    return stageSort(c, false);
}

int dsmcb200_mesh_fill(dsmcb200_ctx* c, int nTypes, const int32_t* typeIds, const double* numberDensities, double Ttra, double Trot,
                       double Tvib, double Telec, const double velocity[3]) {
    return fillCells(c, nullptr, 0, nTypes, typeIds, numberDensities, Ttra, Trot, Tvib, Telec, velocity);
}

int dsmcb200_zone_fill(dsmcb200_ctx* c, int64_t nZoneCells, const int32_t* zoneCells, int nTypes, const int32_t* typeIds, const double* numberDensities,
                       double Ttra, double Trot, double Tvib, double Telec, const double velocity[3]) {
    if (!zoneCells && nZoneCells > 0) return DSMCB200_ERR_INVALID;
    static const int32_t none = 0;
    return fillCells(c, zoneCells ? zoneCells : &none, nZoneCells, nTypes, typeIds, numberDensities, Ttra, Trot, Tvib, Telec, velocity);
}

int dsmcb200_set_step(dsmcb200_ctx* c, uint32_t step) { if (!c) return DSMCB200_ERR_INVALID; c->step = step; return 0; }

int dsmcb200_stage(dsmcb200_ctx* c, int stage) {
    if (!c) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    int r = 0;
    switch (stage) {
        case DSMCB200_STAGE_INFLOW: r = stageInflow(c, c->N); c->occupancyValid = false; break;  // new parcels are in no cell list; step fractions are only kept until the next stage call
        case DSMCB200_STAGE_MOVE: r = beginWallStep(c); if (!r) r = stageMove(c, c->N); if (!r) r = endWallStep(c, true); c->occupancyValid = false; break;
        case DSMCB200_STAGE_SORT: r = stageSort(c, false); if (!r) r = stageWeighting(c); break;   // the occupancy the collide stage sees
        case DSMCB200_STAGE_COLLIDE: r = stageCollide(c); break;
        case DSMCB200_STAGE_SAMPLE: r = stageSample(c); break;
        default: return fail(c, DSMCB200_ERR_INVALID, "unknown stage");
    }
    if (r) return r;
    CK(cudaStreamSynchronize(c->stream));
    resolveTimers(c);
    return fetchCounters(c);
}

int dsmcb200_evolve(dsmcb200_ctx* c, int nSteps) {
    if (!c || nSteps < 0) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    for (int it = 0; it < nSteps; ++it) {
        cudaEvent_t e0 = getEvent(c), e1 = getEvent(c), e2 = getEvent(c), e3 = getEvent(c), e4 = getEvent(c), e5 = getEvent(c);
        c->last.inserted = 0; c->last.migratedIn = 0; c->last.migrationRounds = 0; c->last.cloned = 0; c->weightDeletedStep = 0;
        c->last.nNeighbours = int32_t(c->nbrProcs.size());
        for (int k = 0; k < MAX_NEIGHBOURS; ++k) { c->last.neighbourProc[k] = k < int(c->nbrProcs.size()) ? c->nbrProcs[k] : -1; c->last.migratedTo[k] = 0; c->last.migratedFrom[k] = 0; }
        CK(cudaMemsetAsync(c->dCounters, 0, sizeof(DevCounters), c->stream));
        const int64_t tailStart = c->N;
        // trackingInfo_.clean() (dsmcCloud.C:923): the face tracker holds one step
        if (c->dFaceFlux) CK(cudaMemsetAsync(c->dFaceFlux, 0, size_t(2) * c->hP.nSpecies * c->mesh.nFaces * 8, c->stream));
        cudaEventRecord(e0, c->stream);
        { int r = stageInflow(c, tailStart); if (r) return r; }    // boundaries_.controlBeforeMove()
        cudaEventRecord(e1, c->stream);
        { int r = beginWallStep(c); if (r) return r; }
        { int r = stageMove(c, tailStart); if (r) return r; }      // Cloud<dsmcParcel>::move
        { int r = endWallStep(c, false); if (r) return r; }
        cudaEventRecord(e2, c->stream);
        { int r = stageSort(c, true); if (r) return r; }           // buildCellOccupancy()
        { int r = stageWeighting(c); if (r) return r; }            // coordSystem().evolve()
        cudaEventRecord(e3, c->stream);
        { int r = stageCollide(c); if (r) return r; }              // collisions()
        cudaEventRecord(e4, c->stream);
        // dsmcVolFields::calculateField samples when sampleInterval_ <= ++sampleCounter_ (dsmcVolFields.C:1073-1081,1362)
        if (++c->sampleCounter >= std::max(1, c->models.sampleInterval)) {
            { int r = sampleInto(c, c->dAcc, c->dCollCum, c->nTimeSteps); if (r) return r; }           // fields_.calculateFields()
            c->sampleCounter = 0;
        }
        for (auto& e : c->extraSets)   // field{} entries with another sampleInterval
            if (++e.counter >= e.interval) {
                { int r = sampleInto(c, e.dAcc, e.dCollCum, e.nTimeSteps); if (r) return r; }
                e.counter = 0;
            }
        cudaEventRecord(e5, c->stream);
        c->step++;
        { int r = fetchCounters(c); if (r) return r; }
        float ms[5] = {0, 0, 0, 0, 0};
        cudaEventElapsedTime(&ms[0], e0, e1); cudaEventElapsedTime(&ms[1], e1, e2); cudaEventElapsedTime(&ms[2], e2, e3);
        cudaEventElapsedTime(&ms[3], e3, e4); cudaEventElapsedTime(&ms[4], e4, e5);
        c->last.stageMs[0] = ms[0]; c->last.stageMs[1] = ms[1]; c->last.stageMs[2] = 0; c->last.stageMs[3] = ms[2];
        c->last.stageMs[4] = ms[3]; c->last.stageMs[5] = ms[4]; c->last.stageMs[6] = 0;
        c->last.stageMs[7] = ms[0] + ms[1] + ms[2] + ms[3] + ms[4];
        resolveTimers(c);
        if (c->hCounters.overflow) return fail(c, DSMCB200_ERR_CAPACITY, "a fixed-capacity device buffer overflowed during the step");
        if (c->hCounters.trackingFailures) return fail(c, DSMCB200_ERR_STATE, "the move stage gave up on " + std::to_string(c->hCounters.trackingFailures) + " parcel(s) after 200000 tet visits: the tet table is inconsistent with the cloud");
    }
    return 0;
}

int dsmcb200_download_occupancy(dsmcb200_ctx* c, int32_t* cellOffsets) {
    if (!c || !cellOffsets) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    if (!c->occupancyValid) return fail(c, DSMCB200_ERR_STATE, "cell occupancy is stale: run the sort stage first");
    CK(cudaStreamSynchronize(c->stream));
    if (!c->newOfOld.empty()) {
        // the cloud is cell-major in the engine's labels; in the caller's labels the offsets are those of the same cloud ordered by
        // the caller's cells (a stable sort of the downloaded cloud by `cell` has this occupancy)
        std::vector<int32_t> off(size_t(c->mesh.nCells) + 1);
        CK(cudaMemcpy(off.data(), c->dCellOffset, off.size() * 4, cudaMemcpyDeviceToHost));
        cellOffsets[0] = 0;
        for (int32_t k = 0; k < c->mesh.nCells; ++k) { const int32_t r = c->newOfOld[k]; cellOffsets[k + 1] = cellOffsets[k] + (off[r + 1] - off[r]); }
        return 0;
    }
    CK(cudaMemcpy(cellOffsets, c->dCellOffset, size_t(c->mesh.nCells + 1) * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int dsmcb200_set_sample_sets(dsmcb200_ctx* c, int nSets, const int32_t* sampleIntervals) {
    if (!c || nSets < 1 || nSets > 8 || !sampleIntervals) return DSMCB200_ERR_INVALID;
    if (c->ready) return fail(c, DSMCB200_ERR_STATE, "set_sample_sets: the engine has been finalised");
    for (int k = 0; k < nSets; ++k) if (sampleIntervals[k] < 1) return fail(c, DSMCB200_ERR_INVALID, "set_sample_sets: sampleInterval must be at least 1");
    c->setIntervals.assign(sampleIntervals, sampleIntervals + nSets);
    return 0;
}

int dsmcb200_select_sample_set(dsmcb200_ctx* c, int set) {
    if (!c) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    if (set < 0 || set > int(c->extraSets.size())) return fail(c, DSMCB200_ERR_INVALID, "select_sample_set: no such set");
    c->selectedSet = set;
    return 0;
}

int dsmcb200_accum_info_get(dsmcb200_ctx* c, dsmcb200_accum_info* o) {
    if (!c || !o) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    o->nCells = c->mesh.nCells; o->nSpecies = c->hP.nSpecies; o->nQuantities = c->nQ; o->nModes = c->internal ? c->nModes : -1;
    o->nTimeSteps = selSteps(c);
    return 0;
}

int dsmcb200_download_accumulators(dsmcb200_ctx* c, double* acc, double* coll) {
    if (!c) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    CK(cudaStreamSynchronize(c->stream));
    const size_t nC = size_t(c->mesh.nCells);
    (void)nC;
    if (acc) CK(cellRowsToHost(c, acc, selAcc(c), size_t(c->hP.nSpecies) * c->nQ * 8));
    if (coll) CK(cellRowsToHost(c, coll, selColl(c), 16));
    return 0;
}

int dsmcb200_upload_accumulators(dsmcb200_ctx* c, const double* acc, const double* coll, double nTimeSteps) {
    if (!c) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    const size_t nC = size_t(c->mesh.nCells);
    (void)nC;
    if (acc) CK(cellRowsToDevice(c, selAcc(c), acc, size_t(c->hP.nSpecies) * c->nQ * 8));
    if (coll) CK(cellRowsToDevice(c, selColl(c), coll, 16));
    selSteps(c) = nTimeSteps;
    return 0;
}

int dsmcb200_upload_wall_accumulators(dsmcb200_ctx* c, const double* wall) {
    if (!c || !wall) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    if (c->nMeasFaces) CK(cudaMemcpy(selWall(c), wall, size_t(c->nMeasFaces) * c->hP.nSpecies * c->nWallQ * 8, cudaMemcpyHostToDevice));
    return 0;
}

int dsmcb200_reset_accumulators(dsmcb200_ctx* c) {
    if (!c) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    const size_t nC = size_t(c->mesh.nCells);
    CK(cudaMemsetAsync(selAcc(c), 0, nC * c->hP.nSpecies * c->nQ * 8, c->stream));
    CK(cudaMemsetAsync(selColl(c), 0, nC * 16, c->stream));
    if (c->nMeasFaces) CK(cudaMemsetAsync(selWall(c), 0, size_t(c->nMeasFaces) * c->hP.nSpecies * c->nWallQ * 8, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    selSteps(c) = 0;
    return 0;
}

int dsmcb200_wall_info(dsmcb200_ctx* c, int32_t* nFaces, int32_t* nWallQ) {
    if (!c) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    if (nFaces) *nFaces = c->nMeasFaces;
    if (nWallQ) *nWallQ = c->nWallQ;
    return 0;
}

int dsmcb200_upload_overall_temperature(dsmcb200_ctx* c, const double* Tov) {
    if (!c || !Tov) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    const size_t nC = size_t(c->mesh.nCells);
    if (!c->dOverallT) CK(devAlloc(&c->dOverallT, nC));
    CK(cudaStreamSynchronize(c->stream));
    CK(cellRowsToDevice(c, c->dOverallT, Tov, 8));
    return 0;
}

int dsmcb200_download_face_fluxes(dsmcb200_ctx* c, double* parcelIdFlux, double* massIdFlux) {
    if (!c || !parcelIdFlux || !massIdFlux) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    if (!c->dFaceFlux) return fail(c, DSMCB200_ERR_STATE, "face fluxes are not tracked: set models.trackFaceFluxes");
    CK(cudaStreamSynchronize(c->stream));
    const size_t n = size_t(c->hP.nSpecies) * c->mesh.nFaces;
    CK(cudaMemcpy(parcelIdFlux, c->dFaceFlux, n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(massIdFlux, c->dFaceFlux + n, n * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int dsmcb200_download_wall_accumulators(dsmcb200_ctx* c, double* wall) {
    if (!c || !wall) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    CK(cudaStreamSynchronize(c->stream));
    if (c->nMeasFaces) CK(cudaMemcpy(wall, selWall(c), size_t(c->nMeasFaces) * c->hP.nSpecies * c->nWallQ * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int dsmcb200_get_counters(dsmcb200_ctx* c, dsmcb200_counters* o) {
    if (!c || !o) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    // dsmcCloud::info(): one pass over the cloud for the energy sums
    double e5[6] = {0, 0, 0, 0, 0, 0};
    if (c->N > 0) {
        CK(launchInfo(c->buf[c->cur].a, cellFields(c), int32_t(c->N), c->dP, c->dInfo, c->dInfoScratch, c->stream));
        CK(cudaMemcpyAsync(e5, c->dInfo, sizeof(e5), cudaMemcpyDeviceToHost, c->stream));
    }
    { int r = fetchCounters(c); if (r) return r; }
    dsmcb200_counters& L = c->last;
    L.nParcels = c->N; L.collisions = int64_t(c->hCounters.collisions); L.collisionCandidates = int64_t(c->hCounters.candidates);
    L.trackingRescues = int64_t(c->hCounters.rescues); L.deleted = int64_t(c->hCounters.deleted) + c->weightDeletedStep;
    L.migratedOut = int64_t(c->hCounters.migratedOut); L.unsortedLargeCells = int64_t(c->hCounters.unsortedLargeCells);
    L.mass = e5[0]; L.linearKineticEnergy = e5[1]; L.rotationalEnergy = e5[2]; L.vibrationalEnergy = e5[3]; L.electronicEnergy = e5[4];
    L.nMolecules = e5[5];
    *o = L;
    return 0;
}

int dsmcb200_kernel_times(dsmcb200_ctx* c, int capacity, int* n, char* names, float* ms, int64_t* launches) {
    if (!c || !n) return DSMCB200_ERR_INVALID;
    int k = 0;
    for (auto& e : c->ktimes) {
        if (k >= capacity) break;
        if (names) { std::memset(names + size_t(k) * DSMCB200_NAME_LEN, 0, DSMCB200_NAME_LEN); std::strncpy(names + size_t(k) * DSMCB200_NAME_LEN, e.first.c_str(), DSMCB200_NAME_LEN - 1); }
        if (ms) ms[k] = float(e.second.first);
        if (launches) launches[k] = e.second.second;
        ++k;
    }
    *n = k;
    if (capacity == 0) c->ktimes.clear();  // capacity 0 resets the table
    return 0;
}

static int allreduceDoubles(dsmcb200_ctx* c, double* vals, int n, int op) {
    if (!c || !vals || n < 0 || n > 8) return DSMCB200_ERR_INVALID;
    if (c->nRanks <= 1 || !c->comm || n == 0) return 0;
    cudaSetDevice(c->device);
    { int r = finalize(c); if (r) return r; }
    CK(cudaMemcpyAsync(c->dInfo, vals, size_t(n) * 8, cudaMemcpyHostToDevice, c->stream));
    const int NCCL_FLOAT64 = 8;
    int r = g_nccl.AllReduce(c->dInfo, c->dInfo, size_t(n), NCCL_FLOAT64, op, c->comm, c->stream);
    if (r) return ncclFail(c, r, "ncclAllReduce");
    CK(cudaMemcpyAsync(vals, c->dInfo, size_t(n) * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}
int dsmcb200_allreduce_sum(dsmcb200_ctx* c, double* vals, int n) { return allreduceDoubles(c, vals, n, /* ncclSum */ 0); }
int dsmcb200_allreduce_min(dsmcb200_ctx* c, double* vals, int n) { return allreduceDoubles(c, vals, n, /* ncclMin */ 3); }

int dsmcb200_timer_start(dsmcb200_ctx* c) {
    if (!c) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    if (!c->tmr0) { CK(cudaEventCreate(&c->tmr0)); CK(cudaEventCreate(&c->tmr1)); }
    CK(cudaEventRecord(c->tmr0, c->stream));
    return 0;
}

int dsmcb200_timer_stop(dsmcb200_ctx* c, float* ms) {
    if (!c || !ms || !c->tmr0) return DSMCB200_ERR_INVALID;
    cudaSetDevice(c->device);
    CK(cudaEventRecord(c->tmr1, c->stream));
    CK(cudaEventSynchronize(c->tmr1));
    CK(cudaEventElapsedTime(ms, c->tmr0, c->tmr1));
    return 0;
}

int dsmcb200_download_geometry(dsmcb200_ctx* c, double* cellCentres, double* cellVolumes, double* faceCentres, double* faceAreas,
                               int32_t* tetBasePtIs) {
    if (!c || !c->haveMesh) return DSMCB200_ERR_STATE;
    const HostMesh& M = c->mesh;
    for (int k = 0; k < M.nCells; ++k) {
        const int i = c->newOfOld.empty() ? k : c->newOfOld[k];   // the caller's cell k
        if (cellCentres) { cellCentres[3 * k] = M.cellCentres[i].x; cellCentres[3 * k + 1] = M.cellCentres[i].y; cellCentres[3 * k + 2] = M.cellCentres[i].z; }
        if (cellVolumes) cellVolumes[k] = M.cellVolumes[i];
    }
    for (int f = 0; f < M.nFaces; ++f) {
        if (faceCentres) { faceCentres[3 * f] = M.faceCentres[f].x; faceCentres[3 * f + 1] = M.faceCentres[f].y; faceCentres[3 * f + 2] = M.faceCentres[f].z; }
        if (faceAreas) { faceAreas[3 * f] = M.faceAreas[f].x; faceAreas[3 * f + 1] = M.faceAreas[f].y; faceAreas[3 * f + 2] = M.faceAreas[f].z; }
        if (tetBasePtIs) tetBasePtIs[f] = M.tetBasePtIs[f];
    }
    return 0;
}

}  // extern "C"
