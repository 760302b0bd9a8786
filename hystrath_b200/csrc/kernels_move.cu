// kernels_move.cu -- stage 1: free flight with face-crossing tracking on the tet decomposition
// of the polyMesh, boundary interactions, and (fused) the cell histogram of stage 2.
//
// Follows, per parcel, Cloud<T>::move (BASIC/Cloud/Cloud.C:204-312), dsmcParcel::move
// (DSMC/parcels/dsmcParcel.C:38-148) and particle::trackToFace(end, td, DSMC=true)
// (BASIC/particle/particleTemplates.C:727-1241) with findTris/tetLambda
// (BASIC/particle/particleI.H:31-140).  The reference walks mesh topology on every tet hop
// (tetNeighbour / crossEdgeConnectedFace, particleI.H:339-601) and recomputes the four
// normalised face-area vectors; here both are look-ups in a 224-byte TetRec baked by
// host_mesh.cpp with the same arithmetic, so the FP64 comparisons see identical operands.
// Compiled with --fmad=false: x86 gcc -O3 without -march does not contract to FMA either.
//
// Execution shape.  The nesting of the reference (dsmcParcel::move loop around the trackToFace
// do-while) is flattened into one loop whose iteration is "one tetrahedron": load its record, decide
// which plane (if any) the remaining track crosses, advance.  Parcels need different numbers of
// iterations, so each warp owns a chunk of MOVE_CHUNK*32 consecutive parcels and a lane that finishes
// its parcel immediately takes the next unprocessed one of the chunk (warp-private work queue); the
// lanes of a warp therefore stay busy and memory accesses stay inside the chunk's cache lines.
#include <cub/device/device_radix_sort.cuh>

#include "device_models.cuh"
#include "engine.h"

namespace dsmc {

namespace {

constexpr double kTrackingCorrectionTol = 1.0e-5;  // BASIC/particle/particle.C:33
#ifndef MOVE_CHUNK_SZ
#define MOVE_CHUNK_SZ 16
#endif
constexpr int MOVE_CHUNK = MOVE_CHUNK_SZ;  // parcels per lane in a warp work queue
#ifndef MOVE_BLOCK_SZ
#define MOVE_BLOCK_SZ 64
#endif
constexpr int MOVE_BLOCK = MOVE_BLOCK_SZ;
#ifndef MOVE_MIN_BLOCKS
#define MOVE_MIN_BLOCKS 8
#endif

// one 32-byte sector per instruction (sm_100 LDG.256): halves the L1 tag look-ups of the scattered record reads
__device__ __forceinline__ void ld4(const double* __restrict__ p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ int32_t loInt(double w) { return int32_t(__double_as_longlong(w) & 0xffffffffLL); }
__device__ __forceinline__ int32_t hiInt(double w) { return int32_t(__double_as_longlong(w) >> 32); }

// findTris without the division: (lambda > 0 && lambda < 1) for lambda = num/den is decided from the signs and
// magnitudes of num and den.  For IEEE doubles RN(num/den) < 1 <=> |num| < |den| and RN(num/den) > 0 <=> same
// sign and num != 0 (quotients of the magnitudes met here cannot underflow), so the result is identical to
// particle::findTris + tetLambda (BASIC/particle/particleI.H:31-140) while the FP64 divider is left to the
// one or two lambdas that are actually needed.  num = (planeBase - tetCentre) & n is baked into the record.
__device__ __forceinline__ bool planeCrossed(double num, const V3& toMinusCt, const V3& n, double tol) {
    double den = dot(toMinusCt, n);
    if (fabs(den) < tol) {
        if (fabs(num) < tol) return false;                   // lambda = 0
        if (mag(toMinusCt) < tol / mag(n)) return false;     // lambda = GREAT
        den = (den >= 0 ? 1.0 : -1.0) * SMALL;
    }
    return den > 0 ? (num > 0 && num < den) : (num < 0 && num > den);
}

// particle::tetLambda, BASIC/particle/particleI.H:68-140 (static mesh branch)
__device__ __forceinline__ double tetLambda(const V3& from, const V3& toMinusFrom, const V3& n, const V3& base, double tol, bool crossed) {
    double lambdaNumerator = dot(base - from, n);
    double lambdaDenominator = dot(toMinusFrom, n);
    // The quotient of a plane that is not crossed is discarded.  A parcel that has just entered through a face sits on
    // that plane (numerator == 0 after cancellation) and 0/x sends the FP64 divide to its slow path: feed it 1/1 instead.
    if (!crossed) { lambdaNumerator = 1.0; lambdaDenominator = 1.0; }
    if (fabs(lambdaDenominator) < tol) {
        if (fabs(lambdaNumerator) < tol) return 0.0;
        if (mag(toMinusFrom) < tol / mag(n)) return GREAT;
        lambdaDenominator = (lambdaDenominator >= 0 ? 1.0 : -1.0) * SMALL;
    }
    return lambdaNumerator / lambdaDenominator;
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
// face bfi when that is >= 0, else the internal face that owns the face-triangle pair of its tet.
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
        const int32_t pair = tet >> 1;
        int32_t lo = 0, hi = a.nFacesAll;   // last f with faceTetPair0[f] <= pair
        while (hi - lo > 1) {
            const int32_t mid = (lo + hi) >> 1;
            if (a.faceTetPair0[mid] <= pair) lo = mid; else hi = mid;
        }
        face = target = lo;
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
    // the k-th hit of a parcel on a wall that draws random numbers within a step owns the Philox stream (origId, k, step)
    Rng wallRng;
    bool specular = pt.model == DSMCB200_BND_SPECULAR_WALL;
    if (!specular) {
        wallRng.init(P.seed, uint32_t(a.p.origId[i]), uint32_t(*wallHits), a.step, STREAM_WALL);
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

// TRACK: the dsmcFaceTracker hook compiled in (its cold call costs the hot loop 1.2 % even when it is never taken: measured A/B)
template <bool TRACK>
__global__ void __launch_bounds__(MOVE_BLOCK, MOVE_MIN_BLOCKS) moveKernel(const __grid_constant__ MoveArgs a) {
    const DevParams& P = *a.P;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int32_t warpGlobal = (blockIdx.x * MOVE_BLOCK + threadIdx.x) >> 5;
    const int32_t chunkBeg = a.first + warpGlobal * (32 * MOVE_CHUNK);
    int32_t chunkEnd = chunkBeg + 32 * MOVE_CHUNK;
    if (chunkEnd > a.first + a.count) chunkEnd = a.first + a.count;
    if (chunkBeg >= chunkEnd) return;
    int32_t warpNext = chunkBeg;  // warp-uniform: next unassigned parcel of the chunk

    const double deltaT = P.deltaT;
    const bool constrained = P.solutionD[0] == -1 || P.solutionD[1] == -1 || P.solutionD[2] == -1;
    const double* __restrict__ tetBase = reinterpret_cast<const double*>(a.tets);

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
        // ---- section 0: refill idle lanes from the warp's queue ----
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle) {
            const int32_t avail = chunkEnd - warpNext;
            if (avail <= 0) {
                if (idle == FULL) break;
            } else {
                const int r = __popc(idle & ((1u << lane) - 1u));
                if (!active && r < avail) {
                    i = warpNext + r;
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
                const int nIdle = __popc(idle);
                warpNext += nIdle < avail ? nIdle : avail;
            }
        }
        __syncwarp();

        // ---- section 1: one tetrahedron -- its record and the planes crossed by (tet centre -> end position) ----
        bool finished = false;
        double retVal = 1.0;
        if (active) {
            if (++guard > 200000) keepParticle = false;  // corrupt tet table: drop the parcel rather than hang
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
            // the whole record up front: seven independent 32-byte sectors in flight
            const double* __restrict__ R = tetBase + size_t(tet) * 28;
            double n0x, n0y, n0z, numC0, n1x, n1y, n1z, numC1, n2x, n2y, n2z, numC2, n3x, n3y, n3z, numC3;
            double bx, by, bz, ax, ay, az, ctx, cty, ctz, tol, nb01, nb23;
            ld4(R + 24, ctx, cty, ctz, nb23);
            ld4(R + 16, bx, by, bz, tol);
            ld4(R + 0, n0x, n0y, n0z, numC0);
            ld4(R + 4, n1x, n1y, n1z, numC1);
            ld4(R + 8, n2x, n2y, n2z, numC2);
            ld4(R + 12, n3x, n3y, n3z, numC3);
            ld4(R + 20, ax, ay, az, nb01);
            const V3 Ct = mk(ctx, cty, ctz);
            const V3 N0 = mk(n0x, n0y, n0z), N1 = mk(n1x, n1y, n1z), N2 = mk(n2x, n2y, n2z), N3 = mk(n3x, n3y, n3z);
            const V3 base = mk(bx, by, bz), pA = mk(ax, ay, az);
            const V3 toMinusCt = endPosition - Ct;
            const bool c0 = planeCrossed(numC0, toMinusCt, N0, tol);
            const bool c1 = planeCrossed(numC1, toMinusCt, N1, tol);
            const bool c2 = planeCrossed(numC2, toMinusCt, N2, tol);
            const bool c3 = planeCrossed(numC3, toMinusCt, N3, tol);
            // ---- outcome of the visit, decided with predicates and selects; only a boundary face is a real branch.
            // (The nested if/else form of the reference merges position, fraction and flags at every join and the compiler
            // pays for each join with register copies.)
            const V3 toMinusFrom = endPosition - pos;
            const double l0 = tetLambda(pos, toMinusFrom, N0, base, tol, c0);
            const double l1 = tetLambda(pos, toMinusFrom, N1, pA, tol, c1);
            const double l2 = tetLambda(pos, toMinusFrom, N2, base, tol, c2);
            const double l3 = tetLambda(pos, toMinusFrom, N3, base, tol, c3);
            int triI = -1;
            double lambdaMin = VGREAT;
            if (c0 && l0 < lambdaMin) { lambdaMin = l0; triI = 0; }
            if (c1 && l1 < lambdaMin) { lambdaMin = l1; triI = 1; }
            if (c2 && l2 < lambdaMin) { lambdaMin = l2; triI = 2; }
            if (c3 && l3 < lambdaMin) { lambdaMin = l3; triI = 3; }
            const bool none = !(c0 | c1 | c2 | c3);
            const bool live = keepParticle && !rescuePending;
            const bool kResc = keepParticle && rescuePending;         // lambdaMin < SMALL last time: correction towards this tet's centre
            const bool gtS = lambdaMin > SMALL, le1 = lambdaMin <= 1.0;
            const bool kEnd = live && (none || (gtS && !le1));         // the end position lies in this tet
            const bool kAdv = live && !none && gtS && le1;             // advance to the nearest crossed plane
            const bool needRescue = live && !none && !gtS;             // lambdaMin = 0.0
            const bool moving = kAdv || needRescue;
            if (kResc) atomicAdd(&a.counters->rescues, 1ULL);   // rare: counted where it happens, no per-lane counter register
            {
                // rescue and advance have the same form: pos += f * (X - pos)
                const V3 X = kResc ? Ct : endPosition;
                const double f = kResc ? kTrackingCorrectionTol : lambdaMin;
                const V3 stepped = pos + f * (X - pos);
                if (kResc || kAdv) pos = stepped;
                if (kEnd) pos = endPosition;
                if (kAdv) trackFraction += lambdaMin * (1 - trackFraction);
            }
            const int32_t nb0 = loInt(nb01);
            if (live) {
                const bool onFace = !none && triI == 0;
                faceSet = onFace;
                faceBfi = (onFace && nb0 < 0) ? (-1 - nb0) : -1;
            }
            finished = !keepParticle || kResc || kEnd;
            retVal = kResc ? trackFraction : 1.0;
            const bool hop = moving && triI > 0, face = moving && triI == 0;
            if (hop) {
                // particle::tetNeighbour: enter the adjacent tet of the same cell
                tet = triI == 1 ? hiInt(nb01) : (triI == 2 ? loInt(nb23) : hiInt(nb23));
                rescuePending = needRescue;
            }
            if (face) {
                if (nb0 >= 0) {
                    cell = nb0;  // internal face: the same face triangle seen from the other cell
                    tet ^= 1;
                } else {
                        const int32_t bfi = -1 - nb0;
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
                                const V3 nf = N0;
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
                                const int32_t k = (tet >> 1) - bf.tetPair0;
                                tet = 2 * (bf.coupledTetPair0 + (bf.nPts - 3) - k);
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
                                    U = wallInteraction(a, i, a.p.typeId[i], bf.patch, a.wallsDue ? bf.measIndex : -1, bfi, N0, U,
                                                            pt.linearT ? comp(pos, pt.depthAxis) : 0.0, &wallHits);
                                    Udirty = true;
                                }
                                break;
                            default:  // empty patches cannot be hit by constrained tracks
                                break;
                        }
                }
                if (needRescue) {
                    rescuePending = true;  // correction towards the new tet's centre, then return trackFraction
                } else {
                    retVal = trackFraction;
                    finished = true;
                }
            }
        }
        __syncwarp();

        // ---- section 3: trackToFace returned -- back in dsmcParcel::move (DSMC/parcels/dsmcParcel.C:92-118) ----
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
                    // Cloud<T>::move transfer list + particle::prepareForParallelTransfer, fused with the packing
                    const BFaceRec bf = a.bfaces[faceBfi];
                    const DevPatch& pt = P.patch[bf.patch];
                    const int slot = pt.nbrSlot;
                    const double stepFraction = 1.0 - tEnd / deltaT;
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
                        r.tetLocal = (tet >> 1) - bf.tetPair0;
                        r.origId = a.p.origId[i];
                        r.vib[0] = in.vib0; r.vib[1] = in.vib1; r.vib[2] = in.vib2;
                        r.typeId = a.p.typeId[i]; r.elevel = uint8_t(in.elevel); r.cls = a.p.cls ? a.p.cls[i] : 0; r.pad_ = 0;
                        a.migBuf[size_t(slot) * a.migCapacity + k] = r;
                        a.migKey[size_t(slot) * a.migCapacity + k] = i;
                    } else {
                        atomicAdd(&a.counters->overflow, 1ULL);
                    }
                    a.p.cell[i] = -1;
                    atomicAdd(&a.counters->migratedOut, 1ULL);
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
}

cudaError_t launchMove(const MoveArgs& a, cudaStream_t s) {
    if (a.count <= 0) return cudaSuccess;
    const int perBlock = MOVE_BLOCK * MOVE_CHUNK;
    const int grid = (a.count + perBlock - 1) / perBlock;
    if (a.faceFlux) moveKernel<true><<<grid, MOVE_BLOCK, 0, s>>>(a);
    else moveKernel<false><<<grid, MOVE_BLOCK, 0, s>>>(a);
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
    a.p.tet[i] = 2 * (bf.tetPair0 + (bf.nPts - 3) - r.tetLocal);
    a.p.origId[i] = r.origId;
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
