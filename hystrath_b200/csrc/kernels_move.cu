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
#define DSMC_PHILOX_LOCAL_STATE   // see philox.h: the wall path's random state stays out of the register file
#include <cub/device/device_radix_sort.cuh>

#include "device_models.cuh"
#include "engine.h"
#include "move_core.h"

namespace dsmc {

namespace {

#ifndef MOVE_BLOCK_SZ
#define MOVE_BLOCK_SZ 512
#endif
constexpr int MOVE_BLOCK = MOVE_BLOCK_SZ;   // one persistent block per SM
#ifndef MOVE_BATCH_SZ
#define MOVE_BATCH_SZ 64
#endif
constexpr int MOVE_BATCH = MOVE_BATCH_SZ;   // parcels a warp takes from the block's queue at a time
constexpr int MOVE_WARPSTATE = 16 * (MOVE_BLOCK / 32);   // bytes: per-warp queue piece
constexpr int MOVE_SCRATCH = 7;             // doubles of per-thread shared scratch
#ifndef MOVE_SMEM_KB
#define MOVE_SMEM_KB 227
#endif
constexpr size_t MOVE_SMEM_BUDGET = size_t(MOVE_SMEM_KB) * 1024;

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
// the range [p + first, p + last) of an array towards L2 (16-byte granules): one instruction per array instead of a load per line
template <class T>
__device__ __forceinline__ void prefetchRangeL2(const T* p, int32_t first, int32_t last) {
    const uintptr_t b = reinterpret_cast<uintptr_t>(p + first) & ~uintptr_t(15);
    const uintptr_t e = (reinterpret_cast<uintptr_t>(p + last) + 15) & ~uintptr_t(15);
    if (e > b) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b), "r"(uint32_t(e - b)) : "memory");
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
    double parcelRWF;   // p.RWF(): the weight the parcel carries through the whole move step (dsmcCloud.C:851-862)
    double deltaT;      // cloud_.deltaTValue(p.cell())
    double nPts;        // coordSystem().dtModel().nParticles(patch, face) = the face cell's
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
    const double rwf = w.parcelRWF / fmax(fabs(U_dot_nw) * fA * w.deltaT, SMALL);
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
    const double nParticle = w.parcelRWF * w.nPts;  // p.RWF() * dtModel().nParticles(patch, face), dsmcPatchBoundary.C:456-468
    const double deltaQ = nParticle * (preIE - postIE + (0.0 * P.kB)) / (w.deltaT * fA);
    const V3 deltaFD = nParticle * (preIMom - postIMom) / (w.deltaT * fA);
    double* a = w.wallAcc + (size_t(measIndex) * P.nSpecies + sp) * w.nWallQ;
    atomicAdd(a + WQ_Q, deltaQ);
    atomicAdd(a + WQ_FDX, deltaFD.x);
    atomicAdd(a + WQ_FDY, deltaFD.y);
    atomicAdd(a + WQ_FDZ, deltaFD.z);
}

// dsmcFaceTracker::trackFaceTransition (DSMC/faceTracker/dsmcFaceTracker.C:124-198), RWF = 1.  The parcel's face() is the boundary
// face bfi when that is >= 0, else the internal tetFace of its tet.
__device__ __noinline__ void trackFaceTransition(const MoveArgs& a, const DevParams& P, int typeId, const V3& U, int32_t tet, int32_t bfi, double RWF) {
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
    atomicAdd(a.faceFlux + size_t(typeId) * a.nFacesAll + target, sgn * RWF);
    atomicAdd(a.faceFlux + (size_t(P.nSpecies) + typeId) * a.nFacesAll + target, sgn * RWF * P.sp[typeId].mass);
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

// dsmcCLLWallPatch::controlParticle (patchBoundaries/dsmcCLLWallPatch/dsmcCLLWallPatch.C:100-300): the Cercignani-Lampis-Lord scattering
// kernel with Lord's extension to the rotational energy; vibrational and electronic states pass unchanged (commented out in the reference)
__device__ __noinline__ V3 cllReflection(const DevParams& P, const DevPatch& pt, Rng& rng, int sp, V3 nw, V3 U, Internal& in) {
    const DevSpecies& S = P.sp[sp];
    double U_dot_nw = dot(U, nw);
    V3 Ut = U - U_dot_nw * nw;
    while (mag(Ut) < SMALL) {
        double r0 = rng.sample01(), r1 = rng.sample01(), r2 = rng.sample01();
        U = mk(U.x * (0.8 + 0.2 * r0), U.y * (0.8 + 0.2 * r1), U.z * (0.8 + 0.2 * r2));
        U_dot_nw = dot(U, nw);
        Ut = U - U_dot_nw * nw;
        if (magSqr(U) == 0.0) { Ut = mk(nw.y, -nw.x, 0.0); if (mag(Ut) < SMALL) Ut = mk(0.0, nw.z, -nw.y); break; }
    }
    const V3 tw1 = Ut / mag(Ut);
    const V3 tw2 = cross(nw, tw1);
    const double T = pt.T;
    const double alphaT = pt.alphaT, alphaN = pt.alphaN, alphaR = pt.alphaR;
    const double mostProbableVelocity = sqrt(2.0 * P.kB * T / S.mass);
    const V3 normalisedTangentialVelocity = Ut / mostProbableVelocity;
    const double normalisedNormalVelocity = U_dot_nw / mostProbableVelocity;
    const double twoPi = 2.0 * PI;
    const double thetaNormal = twoPi * rng.sample01();
    const double rNormal = sqrt(-alphaN * log(rng.sample01()));
    const double thetaTangential1 = twoPi * rng.sample01();
    const double rTangential1 = sqrt(-alphaT * log(rng.sample01()));
    const double normalisedIncidentTangentialVelocity1 = mag(normalisedTangentialVelocity);
    const double um = sqrt(1.0 - alphaN) * normalisedNormalVelocity;
    const double normalVelocity = sqrt((rNormal * rNormal) + (um * um) + 2.0 * rNormal * um * cos(thetaNormal));
    const double tangentialVelocity1 = (sqrt(1.0 - alphaT) * fabs(normalisedIncidentTangentialVelocity1) + rTangential1 * cos(thetaTangential1));
    const double tangentialVelocity2 = rTangential1 * sin(thetaTangential1);
    U = mostProbableVelocity * (tangentialVelocity1 * tw1 + tangentialVelocity2 * tw2 - normalVelocity * nw);
    const V3 velocity = mk(pt.vel[0], pt.vel[1], pt.vel[2]);
    const V3 uWallNormal = dot(velocity, nw) * nw;
    const V3 uWallTangential1 = dot(velocity, tw1) * tw1;
    const V3 uWallTangential2 = dot(velocity, tw2) * tw2;
    const V3 UNormal = (dot(U, nw) * nw) + uWallNormal * alphaN;
    const V3 UTangential1 = dot(U, tw1) * tw1 + uWallTangential1 * alphaT;
    const V3 UTangential2 = dot(U, tw2) * tw2 + uWallTangential2 * alphaT;
    U = UNormal + UTangential1 + UTangential2;
    if (S.rotDof == 2.0) {
        const double om = sqrt((in.ERot * (1.0 - alphaR)) / (P.kB * T));
        const double rRot = sqrt(-alphaR * (log(fmax(1.0 - rng.sample01(), VSMALL))));
        const double thetaRot = twoPi * rng.sample01();
        in.ERot = P.kB * T * ((rRot * rRot) + (om * om) + (2.0 * rRot * om * cos(thetaRot)));
    }
    if (S.rotDof == 3.0) {   // "polyatomic case, see Bird's DSMC2.FOR code"
        double X = 0.0, A = 0.0;
        do {
            X = 4.0 * rng.sample01();
            A = 2.7182818 * X * X * exp(-(X * X));
        } while (A < rng.sample01());
        const double om = sqrt((in.ERot * (1.0 - alphaR)) / (P.kB * T));
        const double rRot = sqrt(-alphaR) * X;   // as in the reference: not a number for alphaR > 0
        const double thetaRot = 2.0 * rng.sample01() - 1.0;
        in.ERot = P.kB * T * ((rRot * rRot) + (om * om) + (2.0 * rRot * om * cos(thetaRot)));
    }
    return U;
}

// dsmcParcel::hitWallPatch / hitPatch -> dsmc{Diffuse,Specular}WallPatch::controlParticle
// CLL: the instance for cases with a dsmcCLLWallPatch.  The register need of this function decides what the move kernel keeps live across
// the call (ptxas allocates the call tree as a whole): with the CLL kernel in it the kernel's main loop spills a predicate and the stage is
// 11 % slower (40.0 -> 44.6 ms at 248 M parcels), so cases without such a patch run the instance that does not contain it
template <bool CLL>
__device__ __noinline__ V3 wallInteraction(const MoveArgs& a, int32_t i, int32_t cell, int sp, int patch, int32_t measIndex, int32_t bfi, V3 nw, V3 U,
                                           double depthPosition, int* wallHits) {
    const DevParams& P = *a.P;
    const DevPatch& pt = P.patch[patch];
    WallCtx wctx{a.P, a.wallAcc, a.nWallQ, a.bfaceArea, a.p.rwf ? a.p.rwf[i] : 1.0, a.cf.deltaT(P.deltaT, cell), a.cf.nParticlesTs(P.nParticles, cell)};
    Internal in;
    loadInternal(a, P, i, in);
    double preIE, postIE;
    V3 preIMom, postIMom;
    // dsmcCLLWallPatch::initialConfiguration (dsmcCLLWallPatch.C:82-89): with both coefficients zero the wall is specular and measures nothing
    if constexpr (CLL) { if (pt.model == DSMCB200_BND_CLL_WALL && pt.alphaN < VSMALL && pt.alphaT < VSMALL) measIndex = -1; }
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
    } else if (CLL && pt.model == DSMCB200_BND_CLL_WALL) {
        if constexpr (CLL) {
            U = cllReflection(P, pt, wallRng, sp, nw, U, in);
            if (P.hasInternalEnergy) a.p.erot[i] = in.ERot;
        }
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
        if (a.migRwf) a.migRwf[size_t(slot) * a.migCapacity + k] = a.p.rwf ? a.p.rwf[i] : 1.0;
        a.migKey[size_t(slot) * a.migCapacity + k] = i;
    } else {
        atomicAdd(&a.counters->overflow, 1ULL);
    }
    atomicAdd(&a.counters->migratedOut, 1ULL);
}

// per-step work list of moveKernel: one entry per run of cells {parcelBeg, parcelEnd, tetBeg, nTets}; a run with more than MOVE_PMAX
// parcels is split into several entries that stage the same records; the unsorted tail [tailBeg, tailEnd) (inflow, migration
// arrivals) follows in pieces of MOVE_TAIL parcels without a window
__global__ void planCountKernel(const int32_t* __restrict__ groupCell, int32_t nGroups, const int32_t* __restrict__ cellOffset, int32_t* nSub) {
    const int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > nGroups) return;
    if (g == nGroups) { nSub[g] = 0; return; }
    const int32_t cnt = cellOffset[groupCell[g + 1]] - cellOffset[groupCell[g]];
    nSub[g] = (cnt + MOVE_PMAX - 1) / MOVE_PMAX;
}
__global__ void planFillKernel(const int32_t* __restrict__ groupCell, int32_t nGroups, const int32_t* __restrict__ cellOffset,
                               const int32_t* __restrict__ cellTetStart, const int32_t* __restrict__ subBase, int32_t maxTets, int32_t tailBeg,
                               int32_t tailEnd, int4* plan, int32_t* planTotal) {
    const int32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t nTail = tailEnd > tailBeg ? (tailEnd - tailBeg + MOVE_TAIL - 1) / MOVE_TAIL : 0;
    const int32_t nSorted = nGroups > 0 ? subBase[nGroups] : 0;
    if (g == 0) *planTotal = nSorted + nTail;
    if (g < nGroups) {
        const int32_t c0 = groupCell[g], c1 = groupCell[g + 1];
        const int32_t p0 = cellOffset[c0], p1 = cellOffset[c1];
        const int32_t t0 = cellTetStart[c0];
        int32_t nT = cellTetStart[c1] - t0;
        if (nT > maxTets) nT = 0;   // a single cell larger than the window: its parcels read the global table
        int32_t k = subBase[g];
        for (int32_t p = p0; p < p1; p += MOVE_PMAX, ++k) plan[k] = make_int4(p, min(p + MOVE_PMAX, p1), t0, nT);
    }
    // the tail entries are written by the threads after the groups
    const int32_t t = g - nGroups;
    if (t >= 0 && t < nTail) plan[nSorted + t] = make_int4(tailBeg + t * MOVE_TAIL, min(tailBeg + (t + 1) * MOVE_TAIL, tailEnd), 0, 0);
}

cudaError_t launchMovePlan(const MovePlanArgs& m, int32_t* scanScratch, cudaStream_t s) {
    if (m.nGroups > 0) {
        const int n = m.nGroups + 1;
        planCountKernel<<<(n + 255) / 256, 256, 0, s>>>(m.groupCell, m.nGroups, m.cellOffset, m.nSub);
        cudaError_t e = launchExclusiveScan(m.nSub, m.subBase, nullptr, m.nGroups, scanScratch, s);
        if (e != cudaSuccess) return e;
    }
    const int32_t nTail = m.tailEnd > m.tailBeg ? (m.tailEnd - m.tailBeg + MOVE_TAIL - 1) / MOVE_TAIL : 0;
    const int n = m.nGroups + nTail + 1;
    planFillKernel<<<(n + 255) / 256, 256, 0, s>>>(m.groupCell, m.nGroups, m.cellOffset, m.cellTetStart, m.subBase, m.maxTets, m.tailBeg, m.tailEnd,
                                                   m.plan, m.planTotal);
    return cudaGetLastError();
}

// ---- the persistent move kernel ----
// One block per SM walks the entries b, b + gridDim, b + 2 gridDim, ... of the work list.  A ring of MOVE_NBUF slots holds the entries in
// flight: arming a slot publishes the entry's parcel range as a queue and starts the bulk copy of its tet records into the slot's
// window, so the copy of entry q + MOVE_NBUF - 1 runs while the warps drain entry q.  Lanes take parcels from the slot their warp is
// drawing from and keep the slot of their parcel; an entry is complete when every parcel taken from it has been written back and
// every warp has moved on, and whoever completes it arms the slot with the next entry.  Nobody waits at entry boundaries.
struct MoveSlot {
    int32_t pBeg, pEnd, tetBeg, nStaged;   // nStaged < 0: end of the work list
    int32_t head;      // queue head (atomically advanced)
    int32_t pending;   // parcels not yet written back + warps that have not moved on
    int32_t seq;       // position of the entry in this block's sequence; published last
    int32_t copies;    // bulk copies issued on this slot so far (phase parity of its mbarrier)
};

// arm slot S with the entry of sequence number q of this block (one thread; the previous copy into the window has landed)
__device__ __noinline__ void armSlot(const MoveArgs& a, MoveSlot* S, uint32_t bar, uint32_t win, int32_t q) {
    const int64_t e = int64_t(blockIdx.x) + int64_t(q) * gridDim.x;
    if (e < a.planTotal[0]) {
        const int4 ent = a.plan[e];
        S->pBeg = ent.x; S->pEnd = ent.y; S->tetBeg = ent.z; S->nStaged = ent.w; S->head = ent.x;
        S->pending = (ent.y - ent.x) + MOVE_BLOCK / 32;
        if (ent.w > 0) {
            const uint32_t bytes = uint32_t(ent.w) * uint32_t(sizeof(TetRec));
            mbarExpectTx(bar, bytes);
            bulkCopyG2S(win, a.tets + ent.z, bytes, bar);
            S->copies += 1;
        }
#ifndef MOVE_NO_PREFETCH
        // the entry's rows of the cloud towards L2 while earlier entries are being walked
        prefetchRangeL2(a.p.cell, ent.x, ent.y); prefetchRangeL2(a.p.tet, ent.x, ent.y);
        prefetchRangeL2(a.p.px, ent.x, ent.y); prefetchRangeL2(a.p.py, ent.x, ent.y); prefetchRangeL2(a.p.pz, ent.x, ent.y);
        prefetchRangeL2(a.p.ux, ent.x, ent.y); prefetchRangeL2(a.p.uy, ent.x, ent.y); prefetchRangeL2(a.p.uz, ent.x, ent.y);
#endif
    } else {
        S->pBeg = 0; S->pEnd = 0; S->tetBeg = 0; S->nStaged = -1; S->head = 0; S->pending = 1 << 30;
    }
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smemAddr(&S->seq)), "r"(q) : "memory");
}

// `n` units of slot S are done (parcels written back, or one warp moving on); whoever completes the entry arms the slot with the next one
__device__ __forceinline__ void releaseSlot(const MoveArgs& a, MoveSlot* S, uint32_t bar, uint32_t win, int32_t n) {
    __threadfence_block();   // this thread's reads of the slot and of its window are done before the count that lets the slot be re-armed
    if (atomicSub(&S->pending, n) == n) {
        __threadfence_block();
        const int32_t c = S->copies;
        if (c > 0) {
            int32_t polls = 0;
            while (!mbarTest(bar, uint32_t(c - 1) & 1u)) {
                if (++polls > (1 << 26)) { atomicAdd(&a.counters->trackingFailures, 1ULL << 32); break; }   // watchdog, see the refill section
            }
        }
        armSlot(a, S, bar, win, S->seq + MOVE_NBUF);
    }
}

// TRACK: dsmcFaceTracker counters; CF: per-cell time steps / parcel weights (dsmcb200_set_cell_fields, dsmcAxisymmetric) -- the uniform
// Cartesian instance keeps deltaT in a register and never looks at the cell fields
template <bool TRACK, bool CF, bool CLL>
__global__ void __launch_bounds__(MOVE_BLOCK, 1) moveKernel(const __grid_constant__ MoveArgs a) {
    extern __shared__ __align__(16) unsigned char smRaw[];
    // layout: [0, 8 NBUF) mbarriers | slots | per-thread scratch U.xyz, tEnd | windows
    MoveSlot* sSlot = reinterpret_cast<MoveSlot*>(smRaw + 64);
    double* sScratch = reinterpret_cast<double*>(smRaw + 64 + MOVE_NBUF * sizeof(MoveSlot) + MOVE_WARPSTATE);
    const uint32_t barBase = smemAddr(smRaw);
    const uint32_t winBase = barBase + 64 + MOVE_NBUF * uint32_t(sizeof(MoveSlot)) + MOVE_WARPSTATE + MOVE_SCRATCH * MOVE_BLOCK * 8;
    const uint32_t winBytes = uint32_t(a.stageTets) * uint32_t(sizeof(TetRec));
    const DevParams& P = *a.P;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    double* const myU = sScratch + threadIdx.x;   // [k * MOVE_BLOCK]: U.x, U.y, U.z, tEnd, endPosition.xyz of this lane's parcel (touched once per visit or less)

    if (threadIdx.x == 0) {
        for (int s = 0; s < MOVE_NBUF; ++s) {
            mbarInit(barBase + 8 * s, 1);
            sSlot[s].copies = 0; sSlot[s].seq = -1;
        }
        for (int s = 0; s < MOVE_NBUF; ++s) armSlot(a, &sSlot[s], barBase + 8 * s, winBase + uint32_t(s) * winBytes, s);
    }
    __syncthreads();

    const double deltaT = P.deltaT;
    // dsmcParcel.C:76: the reduced-D corrections of position and tracking velocity apply "but not for axisymmetric cases"
    const bool constrained = (!CF || P.coordinateSystem == DSMCB200_COORD_CARTESIAN) && (P.solutionD[0] == -1 || P.solutionD[1] == -1 || P.solutionD[2] == -1);

    // warp-uniform: the entry this warp draws parcels from
    int32_t wq = 0;             // its sequence number
    // the warp's private piece of a queue {next, end, slot} lives in shared memory (registers are for the parcels)
    volatile int32_t* const wPiece = reinterpret_cast<volatile int32_t*>(smRaw + 64 + MOVE_NBUF * sizeof(MoveSlot)) + 4 * (threadIdx.x >> 5);
    if (lane == 0) { wPiece[0] = 0; wPiece[1] = 0; wPiece[2] = 0; wPiece[3] = 0; }
    __syncwarp();
    bool listDone = false;      // the end of the work list has been reached

    // per-lane parcel state (U and tEnd live in shared memory: they are touched once per trackToFace call)
    // flags of the lane's parcel in one register (eight bools cost the kernel eight registers and their spills)
    constexpr uint32_t F_ACTIVE = 1, F_INCALL = 2, F_RESCUE = 4, F_FACESET = 8, F_UDIRTY = 16, F_KEEP = 32, F_SWITCH = 64, F_WINREADY = 128,
                       F_SLOT_SHIFT = 8;   // bits 8..11: ring slot of the parcel
    uint32_t st = 0;
    int32_t i = -1, cell = -1, tet = 0;
    V3 pos = mk(0, 0, 0);
    double trackFraction = 0.0;
    int32_t faceBfi = -1;
    int32_t hitsAndGuard = 0;   // wall hits that drew random numbers (bits 24..31) | tet visits of this parcel (bits 0..23)

    // Every lane runs through every section of the loop body and the sections are separated by __syncwarp(), so
    // the warp is converged again at each section head whatever happened in the (divergent) section before it.
    while (true) {
        // ---- section 0: refill idle lanes.  The warp holds a private piece [wNext, wEnd) of the queue of slot wSlot and takes the next
        // piece (one shared atomic per MOVE_BATCH parcels) when that runs out ----
        bool written = false;   // this lane's parcel left the kernel in this iteration (counted at the end of the body)
        const unsigned idle = __ballot_sync(FULL, !(st & F_ACTIVE));
        if (idle) {
            int32_t wNext = wPiece[0], wEnd = wPiece[1];
            __syncwarp();
            if (wNext >= wEnd && !listDone) {
                const int s = wq % MOVE_NBUF;
                MoveSlot* const S = &sSlot[s];
                int32_t sq;
                asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(sq) : "r"(smemAddr(&S->seq)) : "memory");
                if (sq == wq) {   // sq < wq: not armed yet (an entry is re-armed only after every warp has moved on, so never sq > wq)
                    const int32_t end = S->pEnd;
                    if (S->nStaged < 0) {
                        listDone = true;
                    } else {
                        int32_t base = 0;
                        if (lane == 0) base = atomicAdd(&S->head, MOVE_BATCH);
                        base = __shfl_sync(FULL, base, 0);
                        wNext = base; wEnd = min(base + MOVE_BATCH, end);
                        if (lane == 0) { wPiece[0] = wNext; wPiece[1] = wEnd; wPiece[2] = s; wPiece[3] = 0; }
                        __syncwarp();
                        if (base + MOVE_BATCH >= end) {   // the queue is empty after this piece: this warp moves on (its piece stays counted
                            if (lane == 0) releaseSlot(a, S, barBase + 8 * s, winBase + uint32_t(s) * winBytes, 1);   // in `pending` through its parcels)
                            wq += 1;
                        }
                    }
                } else {
                    // not armed yet.  A watchdog bounds the wait: a protocol error must end as an error code, not as a hung device
                    const int32_t c = wPiece[3] + 1;
                    __syncwarp();
                    if (lane == 0) wPiece[3] = c;
                    if (c > (1 << 22)) {
                        listDone = true;
                        if (lane == 0) atomicAdd(&a.counters->trackingFailures, 1ULL << 32);
                    }
                }
            }
            if (wNext < wEnd) {
                const int32_t mine = wNext + __popc(idle & ((1u << lane) - 1u));
                const int wSlot = wPiece[2];
                __syncwarp();
                if (lane == 0) wPiece[0] = wNext + __popc(idle);
                if (!(st & F_ACTIVE) && mine < wEnd) {
                    i = mine;
                    st = F_KEEP | (uint32_t(wSlot) << F_SLOT_SHIFT);
                    // the whole row at once: one memory latency (a deleted parcel, cell < 0, only occurs in the unsorted tail)
                    cell = a.p.cell[i];
                    tet = a.p.tet[i];
                    pos = mk(a.p.px[i], a.p.py[i], a.p.pz[i]);
                    const V3 U = mk(a.p.ux[i], a.p.uy[i], a.p.uz[i]);
                    const double stepFraction = (a.sfTail != nullptr && i >= a.tailStart) ? a.sfTail[i - a.tailStart] : 0.0;
                    double dtCell = deltaT;
                    if constexpr (CF) dtCell = a.cf.deltaT(deltaT, cell < 0 ? 0 : cell);   // deltaTValue(orgCell), dsmcParcel.C:62-63
                    const double tEnd = (1.0 - stepFraction) * dtCell;
                    myU[0] = U.x; myU[MOVE_BLOCK] = U.y; myU[2 * MOVE_BLOCK] = U.z; myU[3 * MOVE_BLOCK] = tEnd;
                    faceBfi = -1;
                    hitsAndGuard = 0;
                    written = true;
                    if (cell >= 0) {
                        if (tEnd > ROOTVSMALL) { st |= F_ACTIVE; written = false; }
                        else if (a.cellCount) atomicAdd(&a.cellCount[cell], 1);  // nothing left to move (stepFraction == 1)
                    }
                }
            } else if (listDone && idle == FULL) {
                break;
            }
        }
        __syncwarp();

        // ---- section 1: one tetrahedron ----
        bool finished = false;
        double retVal = 1.0;
        if (st & F_ACTIVE) {
            const int mySlot = int((st >> F_SLOT_SHIFT) & 15u);
            hitsAndGuard += 1;
            if ((hitsAndGuard & 0xffffff) > 200000) { st &= ~F_KEEP; atomicAdd(&a.counters->trackingFailures, 1ULL); }  // corrupt tet table: reported as an error by the host
            if (!(st & F_INCALL)) {
                // dsmcParcel::move loop body up to the trackToFace call (DSMC/parcels/dsmcParcel.C:74-92)
                V3 Utracking = mk(myU[0], myU[MOVE_BLOCK], myU[2 * MOVE_BLOCK]);
                if (constrained) {
#pragma unroll
                    for (int d = 0; d < 3; ++d)
                        if (P.solutionD[d] == -1) { setComp(pos, d, P.centre[d]); setComp(Utracking, d, 0.0); }
                }
                const V3 endNew = pos + myU[3 * MOVE_BLOCK] * Utracking;  // dt = tEnd
                myU[4 * MOVE_BLOCK] = endNew.x; myU[5 * MOVE_BLOCK] = endNew.y; myU[6 * MOVE_BLOCK] = endNew.z;
                trackFraction = 0.0;
                st = (st | F_INCALL) & ~(F_RESCUE | F_FACESET); faceBfi = -1;
            }
            const V3 endPosition = mk(myU[4 * MOVE_BLOCK], myU[5 * MOVE_BLOCK], myU[6 * MOVE_BLOCK]);
            TetRegs R;
            {
                const MoveSlot& S = sSlot[mySlot];
                const uint32_t rel = uint32_t(tet - S.tetBeg);
                const bool inWin = rel < uint32_t(S.nStaged);   // nStaged = 0: no window
                if (inWin && !(st & F_WINREADY) && mbarTest(barBase + 8 * mySlot, uint32_t(S.copies - 1) & 1u)) st |= F_WINREADY;
                if (inWin && (st & F_WINREADY)) loadRecShared(winBase + uint32_t(mySlot) * winBytes + rel * uint32_t(sizeof(TetRec)), R);
                else loadRecGlobal(a.tets, tet, R);
            }
            VisitOut v;
            v.code = VISIT_SLOW; v.triI = -1; v.needRescue = false;
            if (st & F_KEEP) {
                if (!(st & F_RESCUE)) v = visitFast(R, pos, endPosition, trackFraction);
                if (v.code == VISIT_SLOW) {
                    const SlowOut so = slowVisit(a.tets, tet, pos, endPosition, trackFraction, (st & F_RESCUE) != 0);
                    pos = so.pos; trackFraction = so.trackFraction;
                    v.code = so.packed & 15; v.triI = ((so.packed >> 4) & 15) - 1; v.needRescue = (so.packed & 256) != 0;
                    if (v.code == VISIT_RESCUED) atomicAdd(&a.counters->rescues, 1ULL);   // rare: counted where it happens
                }
                if (v.code != VISIT_RESCUED) {
                    const bool onFace = v.triI == 0;
                    st = onFace ? (st | F_FACESET) : (st & ~F_FACESET);
                    faceBfi = (onFace && R.across < 0) ? (-1 - R.across) : -1;
                }
            }
            finished = !(st & F_KEEP) || v.code == VISIT_RESCUED || v.code == VISIT_END;
            retVal = v.code == VISIT_RESCUED ? trackFraction : 1.0;
            if ((st & F_KEEP) && v.code == VISIT_MOVE) {
                if (v.triI > 0) {
                    // particle::tetNeighbour: enter the adjacent tet of the same cell
                    tet = v.triI == 1 ? R.nbr1 : (v.triI == 2 ? R.nbr2 : R.nbr3);
                    st = v.needRescue ? (st | F_RESCUE) : (st & ~F_RESCUE);
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
                                st |= F_SWITCH;  // dsmcParcel::hitProcessorPatch
                                break;
                            case DSMCB200_PATCH_SYMMETRYPLANE:
                            case DSMCB200_PATCH_SYMMETRY:
                            case DSMCB200_PATCH_WEDGE: {
                                // transformProperties(I - 2.0*nf*nf), particleTemplates.C:1474-1522
                                const V3 U = mk(myU[0], myU[MOVE_BLOCK], myU[2 * MOVE_BLOCK]);
                                const V3 nf = R.N0;
                                const V3 t2 = 2.0 * nf;
                                const double xx = 1.0 - t2.x * nf.x, xy = 0.0 - t2.x * nf.y, xz = 0.0 - t2.x * nf.z;
                                const double yx = 0.0 - t2.y * nf.x, yy = 1.0 - t2.y * nf.y, yz = 0.0 - t2.y * nf.z;
                                const double zx = 0.0 - t2.z * nf.x, zy = 0.0 - t2.z * nf.y, zz = 1.0 - t2.z * nf.z;
                                myU[0] = xx * U.x + xy * U.y + xz * U.z;
                                myU[MOVE_BLOCK] = yx * U.x + yy * U.y + yz * U.z;
                                myU[2 * MOVE_BLOCK] = zx * U.x + zy * U.y + zz * U.z;
                                st |= F_UDIRTY;
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
                                    st &= ~F_KEEP;  // dsmcDeletionPatch::controlParticle
                                } else if (pt.model != DSMCB200_BND_NONE) {
                                    int wallHits = int(uint32_t(hitsAndGuard) >> 24);
                                    const V3 U = wallInteraction<CLL>(a, i, cell, a.p.typeId[i], bf.patch, a.wallsDue ? bf.measIndex : -1, bfi, R.N0,
                                                                 mk(myU[0], myU[MOVE_BLOCK], myU[2 * MOVE_BLOCK]),
                                                                 pt.linearT ? comp(pos, pt.depthAxis) : 0.0, &wallHits);
                                    hitsAndGuard = (hitsAndGuard & 0xffffff) | (wallHits << 24);
                                    myU[0] = U.x; myU[MOVE_BLOCK] = U.y; myU[2 * MOVE_BLOCK] = U.z;
                                    st |= F_UDIRTY;
                                }
                                break;
                            default:  // empty patches cannot be hit by constrained tracks
                                break;
                        }
                    }
                    if (v.needRescue) {
                        st |= F_RESCUE;  // correction towards the new tet's centre, then return trackFraction
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
                if (st & F_FACESET) trackFaceTransition(a, P, a.p.typeId[i], mk(myU[0], myU[MOVE_BLOCK], myU[2 * MOVE_BLOCK]), tet, faceBfi, a.p.rwf ? a.p.rwf[i] : 1.0);  // dsmcParcel.C:106-111
            }
            double tEnd = myU[3 * MOVE_BLOCK];
            if (st & F_KEEP) {
                const double dt = tEnd * retVal;
                tEnd -= dt;  // stepFraction = 1 - tEnd/deltaT is only consumed by a processor transfer: evaluated there
                myU[3 * MOVE_BLOCK] = tEnd;
                if ((st & F_FACESET) && faceBfi >= 0) {
                    const int ptype = P.patch[a.bfaces[faceBfi].patch].type;
                    if (ptype == DSMCB200_PATCH_PROCESSOR || ptype == DSMCB200_PATCH_PROCESSORCYCLIC) {
                        st |= F_SWITCH;  // the patch face of the transfer is faceBfi
                    }
                }
            }
            st &= ~F_INCALL;
            if (!((st & F_KEEP) && !(st & F_SWITCH) && tEnd > ROOTVSMALL)) {
                // ---- this parcel is done: write it back ----
                st &= ~F_ACTIVE;
                written = true;
                if (!(st & F_KEEP)) {
                    a.p.cell[i] = -1;
                    atomicAdd(&a.counters->deleted, 1ULL);
                } else if (st & F_SWITCH) {
                    packMigrant(a, P, i, faceBfi, tet, pos, mk(myU[0], myU[MOVE_BLOCK], myU[2 * MOVE_BLOCK]), 1.0 - tEnd / (CF ? a.cf.deltaT(deltaT, cell) : deltaT));   // stepFraction = 1 - tEnd / deltaTValue(orgCell): the face cell (dsmcParcel.C:97)
                    a.p.cell[i] = -1;
                } else {
                    a.p.px[i] = pos.x; a.p.py[i] = pos.y; a.p.pz[i] = pos.z;
                    a.p.cell[i] = cell;
                    a.p.tet[i] = tet;
                    if (st & F_UDIRTY) { a.p.ux[i] = myU[0]; a.p.uy[i] = myU[MOVE_BLOCK]; a.p.uz[i] = myU[2 * MOVE_BLOCK]; }
                    if (a.cellCount) atomicAdd(&a.cellCount[cell], 1);
                }
            }
        }
        // parcels that left in this iteration are taken off their entries: one atomic per slot
        if (written) {
            const int mySlot = int((st >> F_SLOT_SHIFT) & 15u);
            const unsigned grp = __match_any_sync(__activemask(), mySlot);
            if (lane == __ffs(grp) - 1) releaseSlot(a, &sSlot[mySlot], barBase + 8 * mySlot, winBase + uint32_t(mySlot) * winBytes, __popc(grp));
        }
    }
}

size_t moveSharedBytes(int32_t stageTets) { return 64 + MOVE_NBUF * sizeof(MoveSlot) + MOVE_WARPSTATE + size_t(MOVE_SCRATCH) * MOVE_BLOCK * 8 + size_t(MOVE_NBUF) * stageTets * sizeof(TetRec); }
int32_t moveMaxStageTets() { return int32_t((MOVE_SMEM_BUDGET - 64 - MOVE_NBUF * sizeof(MoveSlot) - MOVE_WARPSTATE - size_t(MOVE_SCRATCH) * MOVE_BLOCK * 8) / (MOVE_NBUF * sizeof(TetRec))); }

cudaError_t launchMove(const MoveArgs& a, cudaStream_t s) {
    if (a.gridBlocks <= 0) return cudaSuccess;
    const size_t smem = moveSharedBytes(a.stageTets);
    static bool attrSet = false;
    if (!attrSet) {
        cudaFuncSetAttribute(moveKernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(moveKernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(moveKernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(moveKernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(moveKernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(moveKernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(moveKernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(moveKernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attrSet = true;
    }
    const bool cf = a.cf.nPts || a.cf.dt || a.cf.rwf || a.p.rwf || a.weighted;
    void (*k)(MoveArgs) = nullptr;
    if (a.cllWalls) {   // some patch carries a dsmcCLLWallPatch (engine.cu finalize)
        if (a.faceFlux) k = cf ? moveKernel<true, true, true> : moveKernel<true, false, true>;
        else k = cf ? moveKernel<false, true, true> : moveKernel<false, false, true>;
    } else {
        if (a.faceFlux) k = cf ? moveKernel<true, true, false> : moveKernel<true, false, false>;
        else k = cf ? moveKernel<false, true, false> : moveKernel<false, false, false>;
    }
    k<<<a.gridBlocks, MOVE_BLOCK, smem, s>>>(a);
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

__global__ void permuteDoubles(const double* __restrict__ in, const int32_t* __restrict__ perm, double* __restrict__ out, int32_t n) {
    const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = in[perm[k]];
}

cudaError_t orderMigrants(MigRec* records, MigRec* scratch, const int32_t* keys, int32_t* work, void* temp, size_t tempBytes, int32_t n,
                          cudaStream_t s, double* rwf, double* rwfScratch) {
    if (n <= 1) return cudaSuccess;
    int32_t *keysOut = work, *idxIn = work + n, *idxOut = work + 2 * size_t(n);
    iotaI32<<<(n + 255) / 256, 256, 0, s>>>(idxIn, n);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, tempBytes, keys, keysOut, idxIn, idxOut, n, 0, 32, s);
    if (e != cudaSuccess) return e;
    const int64_t threads = int64_t(n) * 6;
    permuteMigRecs<<<unsigned((threads + 255) / 256), 256, 0, s>>>(records, idxOut, scratch, n);
    if (rwf) {
        permuteDoubles<<<unsigned((n + 255) / 256), 256, 0, s>>>(rwf, idxOut, rwfScratch, n);
        e = cudaMemcpyAsync(rwf, rwfScratch, size_t(n) * sizeof(double), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return e;
    }
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
    if (a.p.rwf) a.p.rwf[i] = a.recvRwf ? a.recvRwf[k] : 1.0;
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
