// kernels_move.cu -- stage 1: free flight with face-crossing tracking on the tet decomposition
// of the polyMesh, boundary interactions, and (fused) the cell histogram of stage 2.
//
// Follows, per parcel, Cloud<T>::move (BASIC/Cloud/Cloud.C:204-312), dsmcParcel::move
// (DSMC/parcels/dsmcParcel.C:38-148) and particle::trackToFace(end, td, DSMC=true)
// (BASIC/particle/particleTemplates.C:727-1241) with findTris/tetLambda
// (BASIC/particle/particleI.H:31-140).  The reference walks mesh topology on every tet hop
// (tetNeighbour / crossEdgeConnectedFace, particleI.H:339-601) and recomputes the four
// normalised face-area vectors; here both are look-ups in a 240-byte TetRec baked by
// host_mesh.cpp with the same arithmetic, so the FP64 comparisons see identical operands.
// Compiled with --fmad=false: x86 gcc -O3 without -march does not contract to FMA either.
//
// Execution shape.  The nesting of the reference (dsmcParcel::move loop around the trackToFace
// do-while) is flattened into one loop whose iteration is "one tetrahedron" (move_core.h).  The cloud
// enters the stage sorted by cell, and tet ids are cell-major, so a run of cells is a contiguous
// piece of both: a block takes one run (<= stageTets records, <= MOVE_PMAX parcels; planned per step
// by planMoveKernel), brings its tet records into shared memory with one bulk copy (cp.async.bulk,
// completion on an mbarrier) and walks its parcels from a block-wide queue -- a lane that finishes a
// parcel takes the next one, so lanes stay busy although parcels need 1-8 visits.  A visit reads its
// record from shared memory when the tet belongs to the run and the copy has landed, else from the
// global table (parcels that left the run, the unsorted tail of inflow / migration arrivals).
#include <cub/device/device_radix_sort.cuh>

#include "device_models.cuh"
#include "engine.h"
#include "move_core.h"

namespace dsmc {

namespace {

#ifndef MOVE_BLOCK_SZ
#define MOVE_BLOCK_SZ 256
#endif
constexpr int MOVE_BLOCK = MOVE_BLOCK_SZ;
#ifndef MOVE_MIN_BLOCKS
#define MOVE_MIN_BLOCKS 2
#endif

__device__ __forceinline__ void ldg2(const double* __restrict__ p, double& a, double& b) {
    asm("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}
__device__ __forceinline__ void lds2(uint32_t addr, double& a, double& b) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ int32_t loInt(double w) { return int32_t(__double_as_longlong(w) & 0xffffffffLL); }
__device__ __forceinline__ int32_t hiInt(double w) { return int32_t(__double_as_longlong(w) >> 32); }

// host_mesh.h TetRec -> registers: fifteen 16-byte loads (the trailing {cell, face, tetPt} words are not needed here)
__device__ __forceinline__ void loadRecGlobal(const TetRec* __restrict__ tets, int32_t tet, TetRegs& r) {
    const double* __restrict__ R = reinterpret_cast<const double*>(tets + tet);
    double w0, w1, w2, dummy;
    ldg2(R + 24, r.Ct.x, r.Ct.y); ldg2(R + 26, r.Ct.z, w1);
    ldg2(R + 0, r.N0.x, r.N0.y); ldg2(R + 2, r.N0.z, r.numC0);
    ldg2(R + 4, r.N1.x, r.N1.y); ldg2(R + 6, r.N1.z, r.numC1);
    ldg2(R + 8, r.N2.x, r.N2.y); ldg2(R + 10, r.N2.z, r.numC2);
    ldg2(R + 12, r.N3.x, r.N3.y); ldg2(R + 14, r.N3.z, r.numC3);
    ldg2(R + 16, r.base.x, r.base.y); ldg2(R + 18, r.base.z, r.tol);
    ldg2(R + 20, r.pA.x, r.pA.y); ldg2(R + 22, r.pA.z, w0);
    ldg2(R + 28, w2, dummy);
    r.across = loInt(w0); r.nbrCell = hiInt(w0); r.nbr1 = loInt(w1); r.nbr2 = hiInt(w1); r.nbr3 = loInt(w2);
}
__device__ __forceinline__ void loadRecShared(uint32_t addr, TetRegs& r) {
    double w0, w1, w2, dummy;
    lds2(addr + 192, r.Ct.x, r.Ct.y); lds2(addr + 208, r.Ct.z, w1);
    lds2(addr + 0, r.N0.x, r.N0.y); lds2(addr + 16, r.N0.z, r.numC0);
    lds2(addr + 32, r.N1.x, r.N1.y); lds2(addr + 48, r.N1.z, r.numC1);
    lds2(addr + 64, r.N2.x, r.N2.y); lds2(addr + 80, r.N2.z, r.numC2);
    lds2(addr + 96, r.N3.x, r.N3.y); lds2(addr + 112, r.N3.z, r.numC3);
    lds2(addr + 128, r.base.x, r.base.y); lds2(addr + 144, r.base.z, r.tol);
    lds2(addr + 160, r.pA.x, r.pA.y); lds2(addr + 176, r.pA.z, w0);
    lds2(addr + 224, w2, dummy);
    r.across = loInt(w0); r.nbrCell = hiInt(w0); r.nbr1 = loInt(w1); r.nbr2 = hiInt(w1); r.nbr3 = loInt(w2);
}

// ---- shared-memory window: mbarrier + bulk copy (cp.async.bulk, sm_90+) ----
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCopyG2S(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool mbarTest(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

struct WallCtx {
    const DevParams* P;
    double* wallAcc;
    int32_t nWallQ;
    const double* bfaceArea;
};

struct Internal {  // internal energy state: read and written in place by the rare consumers (wall models, migration)
    double ERot;
    int32_t vib0, vib1, vib2;
    int elevel;
};

// dsmcPatchBoundary::measurePropertiesBeforeControl / AfterControl accumulation,
// DSMC/boundaries/basic/dsmcPatchBoundary/dsmcPatchBoundary.C:263-356,358-482
__device__ void wallMeasure(const WallCtx& w, int32_t measIndex, int32_t bfi, int sp, const V3& U, const Internal& in, double& IE, V3& IMom) {
    const DevParams& P = *w.P;
    const DevSpecies& S = P.sp[sp];
    V3 Sf = mk(w.bfaceArea[3 * bfi], w.bfaceArea[3 * bfi + 1], w.bfaceArea[3 * bfi + 2]);
    const double fA = mag(Sf);
    V3 nw = Sf;
    nw /= mag(nw);
    const double m = S.mass;
    const double U_dot_nw = dot(U, nw);
    const V3 Ut = U - U_dot_nw * nw;
    const double rwf = 1.0 / fmax(fabs(U_dot_nw) * fA * P.deltaT, SMALL);
    const double ev0 = S.nVib > 0 ? in.vib0 * P.kB * S.thetaV[0] : 0.0;
    const double ev1 = S.nVib > 1 ? in.vib1 * P.kB * S.thetaV[1] : 0.0;
    const double ev2 = S.nVib > 2 ? in.vib2 * P.kB * S.thetaV[2] : 0.0;
    const double EVib = ev0 + ev1 + ev2;
    const double EEle = S.eElec[in.elevel];
    const double UU = dot(U, U);
    if (measIndex >= 0) {
        double* a = w.wallAcc + (size_t(measIndex) * P.nSpecies + sp) * w.nWallQ;
        atomicAdd(a + WQ_RHON, rwf);
        if (S.rotDof > 0) atomicAdd(a + WQ_RHON_INT, rwf);
        if (S.nElec > 1) atomicAdd(a + WQ_RHON_ELEC, rwf);
        atomicAdd(a + WQ_RHOM, m * rwf);
        atomicAdd(a + WQ_LINKE, 0.5 * m * UU * rwf);
        atomicAdd(a + WQ_MCC, m * UU * rwf);
        atomicAdd(a + WQ_MOMX, m * Ut.x * rwf);
        atomicAdd(a + WQ_MOMY, m * Ut.y * rwf);
        atomicAdd(a + WQ_MOMZ, m * Ut.z * rwf);
        atomicAdd(a + WQ_EROT, in.ERot * rwf);
        atomicAdd(a + WQ_ZETAROT, S.rotDof * rwf);
        atomicAdd(a + WQ_EVIB, EVib * rwf);
        if (S.nVib > 0) atomicAdd(a + WQ_EVIBMOD0 + 0, ev0 * rwf);
        if (S.nVib > 1) atomicAdd(a + WQ_EVIBMOD0 + 1, ev1 * rwf);
        if (S.nVib > 2) atomicAdd(a + WQ_EVIBMOD0 + 2, ev2 * rwf);
        atomicAdd(a + WQ_EELEC, EEle * rwf);
    }
    IE = 0.5 * m * UU + in.ERot + EVib + EEle;
    IMom = m * U;
}

__device__ void wallMeasureDelta(const WallCtx& w, int32_t measIndex, int32_t bfi, int sp, double preIE, const V3& preIMom,
                                 double postIE, const V3& postIMom) {
    if (measIndex < 0) return;
    const DevParams& P = *w.P;
    V3 Sf = mk(w.bfaceArea[3 * bfi], w.bfaceArea[3 * bfi + 1], w.bfaceArea[3 * bfi + 2]);
    const double fA = mag(Sf);
    const double nParticle = 1.0 * P.nParticles;  // RWF * nParticles(patch, face)
    const double deltaQ = nParticle * (preIE - postIE + (0.0 * P.kB)) / (P.deltaT * fA);
    const V3 deltaFD = nParticle * (preIMom - postIMom) / (P.deltaT * fA);
    double* a = w.wallAcc + (size_t(measIndex) * P.nSpecies + sp) * w.nWallQ;
    atomicAdd(a + WQ_Q, deltaQ);
    atomicAdd(a + WQ_FDX, deltaFD.x);
    atomicAdd(a + WQ_FDY, deltaFD.y);
    atomicAdd(a + WQ_FDZ, deltaFD.z);
}

// dsmcFaceTracker::trackFaceTransition (DSMC/faceTracker/dsmcFaceTracker.C:124-198), RWF = 1.  The parcel's face() is the boundary
// face bfi when that is >= 0, else the internal tetFace of its tet.
__device__ __noinline__ void trackFaceTransition(const MoveArgs& a, const DevParams& P, int typeId, const V3& U, int32_t tet, int32_t bfi) {
    int32_t face, target;
    double unsignedCredit = 0.0;
    if (bfi >= 0) {
        face = target = P.nInternalFaces + bfi;
        const DevPatch& pt = P.patch[a.bfaces[bfi].patch];
        if (pt.type == DSMCB200_PATCH_CYCLIC) {   // credited to the coupled face without the sign (:160-170)
            target = P.patch[pt.nbrPatch].start + (face - pt.start);
            unsignedCredit = 1.0;
        }
    } else {
        face = target = a.tets[tet].face;   // the crossed face is the tetFace of either side's tet
    }
    const V3 Sf = mk(a.faceAreas[3 * size_t(face)], a.faceAreas[3 * size_t(face) + 1], a.faceAreas[3 * size_t(face) + 2]);
    const double sgn = unsignedCredit != 0.0 ? 1.0 : (dot(U, Sf) >= 0 ? 1.0 : -1.0);
    atomicAdd(a.faceFlux + size_t(typeId) * a.nFacesAll + target, sgn);
    atomicAdd(a.faceFlux + (size_t(P.nSpecies) + typeId) * a.nFacesAll + target, sgn * P.sp[typeId].mass);
}

__device__ __forceinline__ void loadInternal(const MoveArgs& a, const DevParams& P, int32_t i, Internal& in) {
    in.ERot = 0.0; in.vib0 = in.vib1 = in.vib2 = 0; in.elevel = 0;
    if (!P.hasInternalEnergy) return;
    in.ERot = a.p.erot[i];
    if (P.nModes > 0) in.vib0 = a.p.vib[0][i];
    if (P.nModes > 1) in.vib1 = a.p.vib[1][i];
    if (P.nModes > 2) in.vib2 = a.p.vib[2][i];
    in.elevel = a.p.elevel[i];
}

// dsmcParcel::hitWallPatch / hitPatch -> dsmc{Diffuse,Specular}WallPatch::controlParticle
__device__ __noinline__ V3 wallInteraction(const MoveArgs& a, int32_t i, int sp, int patch, int32_t measIndex, int32_t bfi, V3 nw, V3 U,
                                           double depthPosition, int* wallHits) {
    const DevParams& P = *a.P;
    const DevPatch& pt = P.patch[patch];
    WallCtx wctx{a.P, a.wallAcc, a.nWallQ, a.bfaceArea};
    Internal in;
    loadInternal(a, P, i, in);
    double preIE, postIE;
    V3 preIMom, postIMom;
    wallMeasure(wctx, measIndex, bfi, sp, U, in, preIE, preIMom);
    // the k-th hit of a parcel on a wall that draws random numbers within a step owns the Philox stream ((origProc, origId), k, step)
    Rng wallRng;
    bool specular = pt.model == DSMCB200_BND_SPECULAR_WALL;
    if (!specular) {
        wallRng.init(P.seed, uint32_t(a.p.origId[i]), uint32_t(*wallHits) | (uint32_t(a.p.origProc ? a.p.origProc[i] : 0) << 16), a.step, STREAM_WALL);
        *wallHits += 1;
        // dsmcDiffuseSpecularWallPatch::controlParticle (mixed/dsmcDiffuseSpecularWallPatch.C:97-115): Maxwell's model
        if (pt.model == DSMCB200_BND_DIFFUSE_SPECULAR_WALL) specular = !(pt.diffuseFraction > wallRng.sample01());
    }
    if (specular) {
        // dsmcSpecularWallPatch::performSpecularReflection
        const double U_dot_nw = dot(U, nw);
        if (U_dot_nw > 0.0) U -= 2.0 * U_dot_nw * nw;
    } else {
        // dsmcDiffuseWallPatch::performDiffuseReflection
        const DevSpecies& S = P.sp[sp];
        // dsmcPatchBoundary::calculateWallUnitVectors
        double U_dot_nw = dot(U, nw);
        V3 Ut = U - U_dot_nw * nw;
        while (mag(Ut) < SMALL) {
            double r0 = wallRng.sample01(), r1 = wallRng.sample01(), r2 = wallRng.sample01();
            U = mk(U.x * (0.8 + 0.2 * r0), U.y * (0.8 + 0.2 * r1), U.z * (0.8 + 0.2 * r2));
            U_dot_nw = dot(U, nw);
            Ut = U - U_dot_nw * nw;
            if (magSqr(U) == 0.0) { Ut = mk(nw.y, -nw.x, 0.0); if (mag(Ut) < SMALL) Ut = mk(0.0, nw.z, -nw.y); break; }
        }
        const V3 tw1 = Ut / mag(Ut);
        const V3 tw2 = cross(nw, tw1);
        // dsmcDiffuseWallPatch::getLocalTemperature(p.position()[depthAxis_]), dsmcDiffuseWallPatch.C:141-148
        double Tw = pt.T;
        if (pt.linearT) Tw = pt.T + (depthPosition - pt.maxDepth) * (pt.T - pt.Tformation) / pt.lengthPatch;
        const double g1 = wallRng.gaussNormal();
        const double g2 = wallRng.gaussNormal();
        const double r = wallRng.sample01();
        U = sqrt(P.kB * Tw / S.mass) * (g1 * tw1 + g2 * tw2 - sqrt(-2.0 * log(fmax(1 - r, VSMALL))) * nw);
        in.ERot = equipartitionRotationalEnergy(wallRng, P.kB, Tw, S.rotDof);
        if (S.nVib > 0) in.vib0 = equipartitionVibrationalEnergyLevel(wallRng, Tw, S.thetaV[0]);
        if (S.nVib > 1) in.vib1 = equipartitionVibrationalEnergyLevel(wallRng, Tw, S.thetaV[1]);
        if (S.nVib > 2) in.vib2 = equipartitionVibrationalEnergyLevel(wallRng, Tw, S.thetaV[2]);
        in.elevel = equipartitionElectronicLevel(wallRng, P.kB, Tw, S);
        U += mk(pt.vel[0], pt.vel[1], pt.vel[2]);
        if (P.hasInternalEnergy) {
            a.p.erot[i] = in.ERot;
            if (P.nModes > 0) a.p.vib[0][i] = in.vib0;
            if (P.nModes > 1) a.p.vib[1][i] = in.vib1;
            if (P.nModes > 2) a.p.vib[2][i] = in.vib2;
            a.p.elevel[i] = uint8_t(in.elevel);
        }
    }
    wallMeasure(wctx, measIndex, bfi, sp, U, in, postIE, postIMom);
    wallMeasureDelta(wctx, measIndex, bfi, sp, preIE, preIMom, postIE, postIMom);
    return U;
}

}  // namespace

// The visits that visitFast hands back (a denominator inside the tolerance band, the rescue correction): every branch of the
// reference, out of line so that the hot loop does not carry its registers.
struct SlowOut { V3 pos; double trackFraction; int32_t packed; };  // packed = code | (triI + 1) << 4 | needRescue << 8
__device__ __noinline__ SlowOut slowVisit(const TetRec* __restrict__ tets, int32_t tet, V3 pos, V3 end, double trackFraction, bool rescuePending) {
    TetRegs R;
    loadRecGlobal(tets, tet, R);
    const VisitOut v = visitSlow(R, pos, end, trackFraction, rescuePending);
    SlowOut o;
    o.pos = pos; o.trackFraction = trackFraction;
    o.packed = v.code | ((v.triI + 1) << 4) | (v.needRescue ? 256 : 0);
    return o;
}

// Cloud<T>::move transfer list + particle::prepareForParallelTransfer, fused with the packing
__device__ __noinline__ void packMigrant(const MoveArgs& a, const DevParams& P, int32_t i, int32_t faceBfi, int32_t tet, V3 pos, V3 U, double stepFraction) {
    const BFaceRec bf = a.bfaces[faceBfi];
    const DevPatch& pt = P.patch[bf.patch];
    const int slot = pt.nbrSlot;
    const int32_t k = atomicAdd(&a.counters->nMig[slot], 1);
    if (k < a.migCapacity) {
        Internal in;
        loadInternal(a, P, i, in);
        MigRec r;
        r.pos[0] = pos.x; r.pos[1] = pos.y; r.pos[2] = pos.z;
        r.U[0] = U.x; r.U[1] = U.y; r.U[2] = U.z;
        r.erot = in.ERot; r.stepFraction = stepFraction;
        r.patchOrdinal = pt.nbrOrdinal;
        r.patchFace = faceBfi - (pt.start - P.nInternalFaces);
        r.tetLocal = tet - bf.tet0;
        r.origId = a.p.origId[i];
        r.vib[0] = in.vib0; r.vib[1] = in.vib1; r.vib[2] = in.vib2;
        r.typeId = a.p.typeId[i]; r.elevel = uint8_t(in.elevel); r.cls = a.p.cls ? a.p.cls[i] : 0; r.origProc = a.p.origProc ? a.p.origProc[i] : 0;
        a.migBuf[size_t(slot) * a.migCapacity + k] = r;
        a.migKey[size_t(slot) * a.migCapacity + k] = i;
    } else {
        atomicAdd(&a.counters->overflow, 1ULL);
    }
    atomicAdd(&a.counters->migratedOut, 1ULL);
}

// per-step work list of moveKernel: one entry per block {parcelBeg, parcelEnd, tetBeg, nTets}; a run of cells with more than
// MOVE_PMAX parcels is split over several blocks that stage the same records
__global__ void planCountKernel(const int32_t* __restrict__ groupCell, int32_t nGroups, const int32_t* __restrict__ cellOffset, int32_t* nSub) {
    const int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > nGroups) return;
    if (g == nGroups) { nSub[g] = 0; return; }
    const int32_t cnt = cellOffset[groupCell[g + 1]] - cellOffset[groupCell[g]];
    nSub[g] = (cnt + MOVE_PMAX - 1) / MOVE_PMAX;
}
__global__ void planFillKernel(const int32_t* __restrict__ groupCell, int32_t nGroups, const int32_t* __restrict__ cellOffset,
                               const int32_t* __restrict__ cellTetStart, const int32_t* __restrict__ subBase, int32_t maxTets, int4* plan) {
    const int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nGroups) return;
    const int32_t c0 = groupCell[g], c1 = groupCell[g + 1];
    const int32_t p0 = cellOffset[c0], p1 = cellOffset[c1];
    const int32_t t0 = cellTetStart[c0];
    int32_t nT = cellTetStart[c1] - t0;
    if (nT > maxTets) nT = 0;   // a single cell larger than the window: its parcels read the global table
    int32_t k = subBase[g];
    for (int32_t p = p0; p < p1; p += MOVE_PMAX, ++k) plan[k] = make_int4(p, min(p + MOVE_PMAX, p1), t0, nT);
}

cudaError_t launchMovePlan(const MovePlanArgs& m, int32_t* scanScratch, cudaStream_t s) {
    const int n = m.nGroups + 1;
    planCountKernel<<<(n + 255) / 256, 256, 0, s>>>(m.groupCell, m.nGroups, m.cellOffset, m.nSub);
    cudaError_t e = launchExclusiveScan(m.nSub, m.subBase, nullptr, m.nGroups, scanScratch, s);
    if (e != cudaSuccess) return e;
    planFillKernel<<<(m.nGroups + 255) / 256, 256, 0, s>>>(m.groupCell, m.nGroups, m.cellOffset, m.cellTetStart, m.subBase, m.maxTets, m.plan);
    return cudaGetLastError();
}

// TRACK: the dsmcFaceTracker hook compiled in (its cold call costs the hot loop 1.2 % even when it is never taken: measured A/B)
template <bool TRACK>
__global__ void __launch_bounds__(MOVE_BLOCK, MOVE_MIN_BLOCKS) moveKernel(const __grid_constant__ MoveArgs a) {
    extern __shared__ __align__(16) unsigned char smRaw[];   // [0,16): mbarrier + queue head, then the window of tet records
    __shared__ int32_t sQueue;
    const DevParams& P = *a.P;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;

    // ---- this block's entry of the work list ----
    int32_t pBeg, pEnd, tetBeg = 0, nStaged = 0;
    if (int32_t(blockIdx.x) < a.nPlanBlocks) {
        if (int32_t(blockIdx.x) >= a.planTotal[0]) return;
        const int4 e = a.plan[blockIdx.x];
        pBeg = e.x; pEnd = e.y; tetBeg = e.z; nStaged = e.w;
    } else {
        const int32_t k = int32_t(blockIdx.x) - a.nPlanBlocks;
        pBeg = a.tailBeg + k * MOVE_PMAX;
        pEnd = min(pBeg + MOVE_PMAX, a.tailEnd);
    }
    if (pBeg >= pEnd) return;
    const uint32_t bar = smemAddr(smRaw);
    const uint32_t win = bar + 16;
    if (threadIdx.x == 0) {
        sQueue = pBeg;
        if (nStaged > 0) {
            mbarInit(bar, 1);
            mbarExpectTx(bar, uint32_t(nStaged) * uint32_t(sizeof(TetRec)));
            bulkCopyG2S(win, a.tets + tetBeg, uint32_t(nStaged) * uint32_t(sizeof(TetRec)), bar);
        }
    }
    __syncthreads();
    bool windowReady = false;   // the bulk copy has landed (checked once per iteration until it has)
    bool drained = false;       // warp-uniform: the block's queue is empty

    const double deltaT = P.deltaT;
    const bool constrained = P.solutionD[0] == -1 || P.solutionD[1] == -1 || P.solutionD[2] == -1;

    // per-lane parcel state
    bool active = false;
    int32_t i = -1, cell = -1, tet = 0;
    V3 pos = mk(0, 0, 0), U = mk(0, 0, 0), endPosition = mk(0, 0, 0);
    double tEnd = 0.0, trackFraction = 0.0;
    bool inCall = false, rescuePending = false, faceSet = false, Udirty = false;
    bool keepParticle = true, switchProcessor = false;
    int32_t faceBfi = -1;
    int wallHits = 0;
    int guard = 0;

    // Every lane runs through every section of the loop body and the sections are separated by __syncwarp(), so
    // the warp is converged again at each section head whatever happened in the (divergent) section before it.
    while (true) {
        // ---- section 0: refill idle lanes from the block's queue ----
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle) {
            if (drained) {
                if (idle == FULL) break;
            } else {
                int32_t base = 0;
                if (lane == 0) base = atomicAdd(&sQueue, __popc(idle));
                base = __shfl_sync(FULL, base, 0);
                if (base + __popc(idle) >= pEnd) drained = true;
                const int32_t mine = base + __popc(idle & ((1u << lane) - 1u));
                if (!active && mine < pEnd) {
                    i = mine;
                    cell = a.p.cell[i];
                    if (cell >= 0) {
                        tet = a.p.tet[i];
                        pos = mk(a.p.px[i], a.p.py[i], a.p.pz[i]);
                        U = mk(a.p.ux[i], a.p.uy[i], a.p.uz[i]);
                        const double stepFraction = (a.sfTail != nullptr && i >= a.tailStart) ? a.sfTail[i - a.tailStart] : 0.0;
                        tEnd = (1.0 - stepFraction) * deltaT;
                        inCall = false; rescuePending = false; faceSet = false; Udirty = false;
                        keepParticle = true; switchProcessor = false; faceBfi = -1;
                        wallHits = 0;
                        guard = 0;
                        if (tEnd > ROOTVSMALL) active = true;
                        else if (a.cellCount) atomicAdd(&a.cellCount[cell], 1);  // nothing left to move (stepFraction == 1)
                    }
                }
            }
        }
        if (nStaged > 0 && !windowReady) windowReady = mbarTest(bar, 0);
        __syncwarp();

        // ---- section 1: one tetrahedron ----
        bool finished = false;
        double retVal = 1.0;
        if (active) {
            if (++guard > 200000) { keepParticle = false; atomicAdd(&a.counters->trackingFailures, 1ULL); }  // corrupt tet table: reported as an error by the host
            if (!inCall) {
                // dsmcParcel::move loop body up to the trackToFace call (DSMC/parcels/dsmcParcel.C:74-92)
                V3 Utracking = U;
                if (constrained) {
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                        if (P.solutionD[d] == -1) { setComp(pos, d, P.centre[d]); setComp(Utracking, d, 0.0); }
                }
                endPosition = pos + tEnd * Utracking;  // dt = tEnd
                trackFraction = 0.0;
                inCall = true; rescuePending = false; faceSet = false; faceBfi = -1;
            }
            TetRegs R;
            const uint32_t rel = uint32_t(tet - tetBeg);
            if (windowReady && rel < uint32_t(nStaged)) loadRecShared(win + rel * uint32_t(sizeof(TetRec)), R);
            else loadRecGlobal(a.tets, tet, R);
            VisitOut v;
            v.code = VISIT_SLOW; v.triI = -1; v.needRescue = false;
            if (keepParticle) {
                if (!rescuePending) v = visitFast(R, pos, endPosition, trackFraction);
                if (v.code == VISIT_SLOW) {
                    const SlowOut so = slowVisit(a.tets, tet, pos, endPosition, trackFraction, rescuePending);
                    pos = so.pos; trackFraction = so.trackFraction;
                    v.code = so.packed & 15; v.triI = ((so.packed >> 4) & 15) - 1; v.needRescue = (so.packed & 256) != 0;
                    if (v.code == VISIT_RESCUED) atomicAdd(&a.counters->rescues, 1ULL);   // rare: counted where it happens
                }
                if (v.code != VISIT_RESCUED) {
                    const bool onFace = v.triI == 0;
                    faceSet = onFace;
                    faceBfi = (onFace && R.across < 0) ? (-1 - R.across) : -1;
                }
            }
            finished = !keepParticle || v.code == VISIT_RESCUED || v.code == VISIT_END;
            retVal = v.code == VISIT_RESCUED ? trackFraction : 1.0;
            if (keepParticle && v.code == VISIT_MOVE) {
                if (v.triI > 0) {
                    // particle::tetNeighbour: enter the adjacent tet of the same cell
                    tet = v.triI == 1 ? R.nbr1 : (v.triI == 2 ? R.nbr2 : R.nbr3);
                    rescuePending = v.needRescue;
                } else {
                    if (R.across >= 0) {
                        cell = R.nbrCell;  // internal face: the same face triangle seen from the other cell
                        tet = R.across;
                    } else {
                        const int32_t bfi = -1 - R.across;
                        const BFaceRec bf = a.bfaces[bfi];
                        const DevPatch& pt = P.patch[bf.patch];
                        switch (pt.type) {
                            case DSMCB200_PATCH_PROCESSOR:
                            case DSMCB200_PATCH_PROCESSORCYCLIC:
                                switchProcessor = true;  // dsmcParcel::hitProcessorPatch
                                break;
                            case DSMCB200_PATCH_SYMMETRYPLANE:
                            case DSMCB200_PATCH_SYMMETRY:
                            case DSMCB200_PATCH_WEDGE: {
                                // transformProperties(I - 2.0*nf*nf), particleTemplates.C:1474-1522
                                const V3 nf = R.N0;
                                const V3 t2 = 2.0 * nf;
                                const double xx = 1.0 - t2.x * nf.x, xy = 0.0 - t2.x * nf.y, xz = 0.0 - t2.x * nf.z;
                                const double yx = 0.0 - t2.y * nf.x, yy = 1.0 - t2.y * nf.y, yz = 0.0 - t2.y * nf.z;
                                const double zx = 0.0 - t2.z * nf.x, zy = 0.0 - t2.z * nf.y, zz = 1.0 - t2.z * nf.z;
                                U = mk(xx * U.x + xy * U.y + xz * U.z, yx * U.x + yy * U.y + yz * U.z, zx * U.x + zy * U.y + zz * U.z);
                                Udirty = true;
                                break;
                            }
                            case DSMCB200_PATCH_CYCLIC: {
                                // particle::hitCyclicPatch, particleTemplates.C:1525-1570
                                const int32_t k = tet - bf.tet0;
                                tet = bf.coupledTet0 + (bf.nPts - 3) - k;
                                cell = bf.coupledCell;
                                const DevPatch& rp = P.patch[pt.nbrPatch];
                                pos -= mk(rp.sep[0], rp.sep[1], rp.sep[2]);
                                faceBfi = bfi - (pt.start - P.nInternalFaces) + (rp.start - P.nInternalFaces);
                                break;
                            }
                            case DSMCB200_PATCH_WALL:
                            case DSMCB200_PATCH_PATCH:
                                if (pt.model == DSMCB200_BND_DELETION) {
                                    keepParticle = false;  // dsmcDeletionPatch::controlParticle
                                } else if (pt.model != DSMCB200_BND_NONE) {
                                    U = wallInteraction(a, i, a.p.typeId[i], bf.patch, a.wallsDue ? bf.measIndex : -1, bfi, R.N0, U,
                                                            pt.linearT ? comp(pos, pt.depthAxis) : 0.0, &wallHits);
                                    Udirty = true;
                                }
                                break;
                            default:  // empty patches cannot be hit by constrained tracks
                                break;
                        }
                    }
                    if (v.needRescue) {
                        rescuePending = true;  // correction towards the new tet's centre, then return trackFraction
                    } else {
                        retVal = trackFraction;
                        finished = true;
                    }
                }
            }
        }
        __syncwarp();

        // ---- section 2: trackToFace returned -- back in dsmcParcel::move (DSMC/parcels/dsmcParcel.C:92-118) ----
        if (finished) {
            if constexpr (TRACK) {
                if (faceSet) trackFaceTransition(a, P, a.p.typeId[i], U, tet, faceBfi);  // dsmcParcel.C:106-111
            }
            if (keepParticle) {
                const double dt = tEnd * retVal;
                tEnd -= dt;  // stepFraction = 1 - tEnd/deltaT is only consumed by a processor transfer: evaluated there
                if (faceSet && faceBfi >= 0) {
                    const int ptype = P.patch[a.bfaces[faceBfi].patch].type;
                    if (ptype == DSMCB200_PATCH_PROCESSOR || ptype == DSMCB200_PATCH_PROCESSORCYCLIC) {
                        switchProcessor = true;  // the patch face of the transfer is faceBfi
                    }
                }
            }
            inCall = false;
            if (!(keepParticle && !switchProcessor && tEnd > ROOTVSMALL)) {
                // ---- this parcel is done: write it back ----
                active = false;
                if (!keepParticle) {
                    a.p.cell[i] = -1;
                    atomicAdd(&a.counters->deleted, 1ULL);
                } else if (switchProcessor) {
                    packMigrant(a, P, i, faceBfi, tet, pos, U, 1.0 - tEnd / deltaT);
                    a.p.cell[i] = -1;
                } else {
                    a.p.px[i] = pos.x; a.p.py[i] = pos.y; a.p.pz[i] = pos.z;
                    a.p.cell[i] = cell;
                    a.p.tet[i] = tet;
                    if (Udirty) { a.p.ux[i] = U.x; a.p.uy[i] = U.y; a.p.uz[i] = U.z; }
                    if (a.cellCount) atomicAdd(&a.cellCount[cell], 1);
                }
            }
        }
    }
    // the window must have landed before the block's shared memory is released
    if (threadIdx.x == 0 && nStaged > 0) while (!windowReady) windowReady = mbarTest(bar, 0);
}

cudaError_t launchMove(const MoveArgs& a, cudaStream_t s) {
    const int32_t tailBlocks = a.tailEnd > a.tailBeg ? (a.tailEnd - a.tailBeg + MOVE_PMAX - 1) / MOVE_PMAX : 0;
    const int32_t grid = a.nPlanBlocks + tailBlocks;
    if (grid <= 0) return cudaSuccess;
    const size_t smem = a.nPlanBlocks > 0 ? 16 + size_t(a.stageTets) * sizeof(TetRec) : 16;
    static bool attrSet = false;
    if (!attrSet) {
        cudaFuncSetAttribute(moveKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64);
        cudaFuncSetAttribute(moveKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64);
        attrSet = true;
    }
    if (a.faceFlux) moveKernel<true><<<grid, MOVE_BLOCK, smem, s>>>(a);
    else moveKernel<false><<<grid, MOVE_BLOCK, smem, s>>>(a);
    return cudaGetLastError();
}

// ---- leavers back into cloud-list order (see engine.h) ----
namespace {
__global__ void iotaI32(int32_t* v, int32_t n) {
    const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) v[k] = k;
}
__global__ void permuteMigRecs(const MigRec* __restrict__ in, const int32_t* __restrict__ perm, MigRec* __restrict__ out, int32_t n) {
    // one 16-byte piece of a 96-byte record per thread: coalesced stores, gathers of whole records
    const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int32_t k = int32_t(t / 6), piece = int32_t(t % 6);
    if (k >= n) return;
    static_assert(sizeof(MigRec) == 96, "MigRec is moved in six 16-byte pieces");
    reinterpret_cast<int4*>(out + k)[piece] = reinterpret_cast<const int4*>(in + perm[k])[piece];
}
}  // namespace

size_t orderMigrantsTempBytes(int32_t capacity) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    capacity);
    return bytes;
}

cudaError_t orderMigrants(MigRec* records, MigRec* scratch, const int32_t* keys, int32_t* work, void* temp, size_t tempBytes, int32_t n,
                          cudaStream_t s) {
    if (n <= 1) return cudaSuccess;
    int32_t *keysOut = work, *idxIn = work + n, *idxOut = work + 2 * size_t(n);
    iotaI32<<<(n + 255) / 256, 256, 0, s>>>(idxIn, n);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, tempBytes, keys, keysOut, idxIn, idxOut, n, 0, 32, s);
    if (e != cudaSuccess) return e;
    const int64_t threads = int64_t(n) * 6;
    permuteMigRecs<<<unsigned((threads + 255) / 256), 256, 0, s>>>(records, idxOut, scratch, n);
    e = cudaMemcpyAsync(records, scratch, size_t(n) * sizeof(MigRec), cudaMemcpyDeviceToDevice, s);
    return e != cudaSuccess ? e : cudaGetLastError();
}

// ---- arrivals over a processor patch: particle::correctAfterParallelTransfer,
// BASIC/particle/particleTemplates.C:52-123
__global__ void unpackKernel(const __grid_constant__ UnpackArgs a) {
    const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.nRecv) return;
    const MigRec r = a.recv[k];
    const DevParams& P = *a.P;
    const int32_t patch = a.ordinalToPatch[r.patchOrdinal];
    const DevPatch& pt = P.patch[patch];
    const int32_t bfi = pt.start - P.nInternalFaces + r.patchFace;
    const BFaceRec bf = a.bfaces[bfi];
    const int32_t i = a.base + k;
    V3 pos = mk(r.pos[0], r.pos[1], r.pos[2]);
    pos -= mk(pt.sep[0], pt.sep[1], pt.sep[2]);  // ppp.transformPosition (processorCyclic); zero for processor
    a.p.px[i] = pos.x; a.p.py[i] = pos.y; a.p.pz[i] = pos.z;
    a.p.ux[i] = r.U[0]; a.p.uy[i] = r.U[1]; a.p.uz[i] = r.U[2];
    a.p.cell[i] = bf.owner;
    // tetPtI_ = f.size() - 1 - tetPtI_  <=>  local tet index k -> (nPts-3) - k
    a.p.tet[i] = bf.tet0 + (bf.nPts - 3) - r.tetLocal;
    a.p.origId[i] = r.origId;
    if (a.p.origProc) a.p.origProc[i] = r.origProc;
    a.p.typeId[i] = r.typeId;
    if (P.hasInternalEnergy) {
        a.p.erot[i] = r.erot;
        if (P.nModes > 0) a.p.vib[0][i] = r.vib[0];
        if (P.nModes > 1) a.p.vib[1][i] = r.vib[1];
        if (P.nModes > 2) a.p.vib[2][i] = r.vib[2];
        a.p.elevel[i] = r.elevel;
    }
    if (a.p.cls) a.p.cls[i] = r.cls;
    double sf = r.stepFraction;
    if (sf > (1.0 - SMALL)) sf = 1.0;
    a.sfTail[i - a.tailStart] = sf;
}

cudaError_t launchUnpack(const UnpackArgs& a, cudaStream_t s) {
    if (a.nRecv <= 0) return cudaSuccess;
    unpackKernel<<<(a.nRecv + 255) / 256, 256, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace dsmc
