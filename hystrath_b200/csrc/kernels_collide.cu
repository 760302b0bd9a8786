// kernels_collide.cu -- stages 3+4 (noTimeCounter candidate selection, VHS / Larsen-Borgnakke
// binary collisions) and stage 5 (per-cell sampling).
//
// Stage 3 follows noTimeCounter::collide (DSMC/collisionPartnerSelection/derived/noTimeCounter/
// noTimeCounter.C:81-339); stage 4 VariableHardSphere (DSMC/collisions/derived/VariableHardSphere/
// VariableHardSphere.C:80-228) and LarsenBorgnakkeVariableHardSphere (.../LarsenBorgnakkeVariableHardSphere.C:
// 110-279) with the cloud helpers postCollision{Rotational,Vibrational,Electronic}* (DSMC/clouds/
// dsmcCloud.C:1327-1656).  The reference processes the candidates of a cell serially; a later candidate
// sees the velocities written by an earlier accepted one.  Every candidate owns a Philox stream keyed by
// (cell, candidate index, step), so its partners (P,Q) are known independently of the others.
//   collideLaneKernel      one LANE per cell: the lane walks its cell's candidates in order (cells <= 255 parcels)
//   collideBigCellsKernel  one WARP per cell: batches of 32 candidates resolved in dependency order (a candidate
//                          waits for every earlier candidate that shares a parcel with it)
// Both yield exactly the serial result.
//
// Stage 5 follows the per-parcel accumulation of dsmcVolFields::calculateField (DSMC/macroscopicProperties/
// derived/combined/dsmcVolFields/dsmcVolFields.C:1115-1237): parcels are cell-sorted, so each warp reduces
// the parcels of its cell and adds one row of per-species moment sums (single writer per accumulator element).
#include <cub/device/device_radix_sort.cuh>

#include "device_models.cuh"
#include "engine.h"

namespace dsmc {

namespace {

constexpr int COL_WARPS = 4;
constexpr int BIG_CELL_THRESHOLD = 255;  // = LANE_CELL_MAX: larger cells are processed by collideBigCellsKernel

struct CellView {  // the parcels of one cell, where they lie in the sorted cloud
    double *ux, *uy, *uz, *erot;
    int32_t* vib[MAX_MODES];
    uint8_t *typ, *elev;
    const double* tMacro;  // &overallT[cell] or null
};

struct WarpSmem {   // collideBigCellsKernel works on the parcels where they lie; only the sub-cell offsets are staged
    int32_t subStart[9];
};

__device__ __noinline__ double powNI(double a, double b) { return pow(a, b); }

// VariableHardSphere::sigmaTcR
__device__ __forceinline__ double sigmaTcR(const DevParams& P, int tP, int tQ, double cR) {
    if (cR < VSMALL) return 0.0;
    const double sigmaTPQ = P.vhsA[tP][tQ] * powNI(2.0 * P.kB * P.Tref / (P.mR[tP][tQ] * (cR * cR)), P.omegaPQ[tP][tQ] - 0.5) / P.vhsG[tP][tQ];
    return sigmaTPQ * cR;
}

// VariableHardSphere::postCollisionVelocities
__device__ __forceinline__ void postCollisionVelocities(const DevParams& P, Rng& rng, int tP, int tQ, V3& UP, V3& UQ, double cR) {
    if (cR == -1) cR = mag(UP - UQ);
    const double mP = P.sp[tP].mass, mQ = P.sp[tQ].mass;
    const V3 Ucm = (mP * UP + mQ * UQ) / (mP + mQ);
    if (P.collisionModel == DSMCB200_COLL_VSS || P.collisionModel == DSMCB200_COLL_LB_VSS) {
        // VariableSoftSphere::postCollisionVelocities, VariableSoftSphere.C:195-262 (Bird, equation 2.22)
        const double alphaPQ = 0.5 * (P.sp[tP].alpha + P.sp[tQ].alpha);
        const V3 cRComponents = UP - UQ;
        const double cosTheta = 2.0 * (pow(rng.sample01(), 1.0 / alphaPQ)) - 1.0;
        const double sinTheta = sqrt(1.0 - cosTheta * cosTheta);
        const double phi = TWO_PI * rng.sample01();
        const double D = sqrt(cRComponents.y * cRComponents.y + cRComponents.z * cRComponents.z);
        const V3 postCollisionRelU =
            mk(cosTheta * cRComponents.x + sinTheta * sin(phi) * D,
               cosTheta * cRComponents.y + sinTheta * (cR * cRComponents.z * cos(phi) - cRComponents.x * cRComponents.y * sin(phi)) / D,
               cosTheta * cRComponents.z - sinTheta * (cR * cRComponents.y * cos(phi) + cRComponents.x * cRComponents.z * sin(phi)) / D);
        UP = Ucm + postCollisionRelU * mQ / (mP + mQ);
        UQ = Ucm - postCollisionRelU * mP / (mP + mQ);
        return;
    }
    const double cosTheta = 2.0 * rng.sample01() - 1.0;
    const double sinTheta = sqrt(1.0 - cosTheta * cosTheta);
    const double phi = TWO_PI * rng.sample01();
    const V3 postCollisionRelativeU = cR * mk(cosTheta, sinTheta * cos(phi), sinTheta * sin(phi));
    UP = Ucm + postCollisionRelativeU * mQ / (mP + mQ);
    UQ = Ucm - postCollisionRelativeU * mP / (mP + mQ);
}

// dsmcCloud::postCollisionRotationalEnergy
__device__ double postCollisionRotationalEnergy(Rng& rng, double rotationalDof, double ChiB) {
    double energyRatio = 0.0;
    if (rotationalDof == 2.0) {
        energyRatio = 1.0 - powNI(rng.sample01(), 1.0 / ChiB);
    } else {
        const double ChiA = 0.5 * rotationalDof;
        const double ChiAMinusOne = ChiA - 1., ChiBMinusOne = ChiB - 1.;
        if (ChiAMinusOne < SMALL && ChiBMinusOne < SMALL) return rng.sample01();
        double Pp = 0.0;
        do {
            energyRatio = rng.sample01();
            if (ChiAMinusOne < SMALL) Pp = powNI(1.0 - energyRatio, ChiBMinusOne);
            else if (ChiBMinusOne < SMALL) Pp = powNI(1.0 - energyRatio, ChiAMinusOne);
            else
                Pp = powNI((ChiAMinusOne + ChiBMinusOne) * energyRatio / ChiAMinusOne, ChiAMinusOne) *
                     powNI((ChiAMinusOne + ChiBMinusOne) * (1 - energyRatio) / ChiBMinusOne, ChiBMinusOne);
        } while (Pp < rng.sample01());
    }
    return energyRatio;
}

// dsmcCloud::postCollisionVibrationalEnergyLevel; postReaction: no relaxation-number test (dsmcCloud.C:1407-1424)
// ZV2008: inverseZvFormulation "2008" compiled in (the common formulations keep the kernel free of the temperature pointer)
template <bool ZV2008>
__device__ int32_t postCollisionVibrationalEnergyLevel(const DevParams& P, Rng& rng, int32_t vibLevel, int32_t iMax, double thetaV,
                                                       double thetaD, double refTempZv, double omega, double Zref, double Ec,
                                                       const double* zvRow, const double* tMacro, bool postReaction = false) {
    int32_t iDash = vibLevel;
    if (postReaction) {
        double func, EVib;
        do {
            iDash = rng.randomLabel(0, iMax);
            EVib = iDash * P.kB * thetaV;
            func = powNI(1.0 - EVib / Ec, 1.5 - omega);
        } while (func < rng.sample01());
        return iDash;
    }
    double inverseVibrationalCollisionNumber = 1.0;
    const double fixedZv = P.Zvib;
    // invZvFormulation 0 and 2 use the quantised collision temperature; formulation 1 ("2008") the macroscopic overall
    // temperature of the cell, fields().overallT(cellI), and falls back to the former while that is not available
    // (TMacro <= SMALL, dsmcCloud.C:1441-1456)
    double TMacro = 0.0;
    if constexpr (ZV2008) {
        if (fixedZv == 0 && P.invZvFormulation == 1 && tMacro != nullptr) TMacro = *tMacro;
    }
    if (fixedZv == 0 && iMax < ZV_TABLE && !(TMacro > SMALL)) {
        inverseVibrationalCollisionNumber = zvRow[iMax];  // host-tabulated value of the expression below
    } else if (fixedZv == 0) {
        const double T = TMacro > SMALL ? TMacro : iMax * thetaV / (3.5 - omega);
        const double pow1 = powNI(thetaD / T, 1. / 3.) - 1.0;
        const double pow2 = powNI(thetaD / refTempZv, 1. / 3.) - 1.0;
        const double ZvP1 = powNI(thetaD / T, omega);
        const double ZvP2 = powNI(Zref * powNI(thetaD / refTempZv, -omega), pow1 / pow2);
        const double Zv = ZvP1 * ZvP2;
        if (P.invZvFormulation == 2) inverseVibrationalCollisionNumber = 1.0 / (5.0 * Zv);
        else inverseVibrationalCollisionNumber = 1.0 / Zv;
    } else {
        inverseVibrationalCollisionNumber = 1.0 / fixedZv;
    }
    if (inverseVibrationalCollisionNumber > rng.sample01()) {
        double func, EVib;
        do {
            iDash = rng.randomLabel(0, iMax);
            EVib = iDash * P.kB * thetaV;
            func = powNI(1.0 - EVib / Ec, 1.5 - omega);
        } while (func < rng.sample01());
    }
    return iDash;
}

// dsmcCloud::postCollisionElectronicEnergyLevel
__device__ int32_t postCollisionElectronicEnergyLevel(Rng& rng, double Ec, double omega, const DevSpecies& S) {
    int jSelectA = 0, jSelectB = 0;
    double gMax = 0.0;
    for (int i = 0; i < S.nElec; ++i) {
        if (S.eElec[i] > Ec) break;
        jSelectA = i;
        const double g = S.gElec[i] * powNI(Ec - S.eElec[i], 1.5 - omega);
        if (gMax < g) { gMax = g; jSelectB = i; }
    }
    const int jSelect = jSelectA < jSelectB ? jSelectA : jSelectB;
    const double denomMax = S.gElec[jSelect] * powNI(Ec - S.eElec[jSelect], 1.5 - omega);
    int jDash = 0;
    double prob;
    do {
        jDash = rng.randomLabel(0, jSelectA);
        prob = S.gElec[jDash] * powNI(Ec - S.eElec[jDash], 1.5 - omega) / denomMax;
    } while (prob < rng.sample01());
    return jDash;
}

// LarsenBorgnakkeVariableHardSphere::redistribute (postReaction = false)
__device__ __noinline__ void redistribute(const DevParams& P, Rng& rng, const CellView& v, int j, int tOther, double& translationalEnergy, double omegaPQ) {
    const DevSpecies& S = P.sp[v.typ[j]];
    if (S.type == 0) return;  // electron
    // electronic mode
    if (P.invZelec > rng.sample01()) {
        const double EcP = translationalEnergy + S.eElec[v.elev[j]];
        const int lvl = postCollisionElectronicEnergyLevel(rng, EcP, omegaPQ, S);
        v.elev[j] = uint8_t(lvl);
        translationalEnergy = EcP - S.eElec[lvl];
    }
    // vibrational modes
    if (S.nVib > 0) {
        double preEVib[MAX_MODES];
#pragma unroll
        for (int m = 0; m < MAX_MODES; ++m) preEVib[m] = m < S.nVib ? v.vib[m][j] * P.kB * S.thetaV[m] : 0.0;
#pragma unroll
        for (int m = 0; m < MAX_MODES; ++m) {
            if (m >= S.nVib) break;
            const double EcP = translationalEnergy + preEVib[m];
            const int32_t iMaxP = int32_t(EcP / (P.kB * S.thetaV[m]));
            if (iMaxP > 0) {
                const int32_t lvl = postCollisionVibrationalEnergyLevel<true>(P, rng, v.vib[m][j], iMaxP, S.thetaV[m], S.thetaD, S.TrefZv[m],
                                                                        omegaPQ, S.Zref[m], EcP,
                                                                        P.invZvTab + ((size_t(v.typ[j]) * P.nSpecies + tOther) * MAX_MODES + m) * ZV_TABLE, v.tMacro);
                v.vib[m][j] = lvl;
                translationalEnergy = EcP - lvl * P.kB * S.thetaV[m];
            }
        }
    }
    // rotational mode
    if (S.rotDof > 0) {
        const double preCollisionERotP = v.erot[j];
        if (P.invZrot > rng.sample01()) {
            const double EcP = translationalEnergy + preCollisionERotP;
            const double ChiB = 2.5 - omegaPQ;
            const double energyRatio = postCollisionRotationalEnergy(rng, S.rotDof, ChiB);
            v.erot[j] = energyRatio * EcP;
            translationalEnergy = EcP - v.erot[j];
        }
    }
}

__device__ __forceinline__ double warpSumOrdered(double v) {
    // fixed-shape tree: the same order on every run
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return __shfl_sync(0xffffffffu, v, 0);
}

__device__ bool reactPair(const CollideArgs& a, const DevParams& P, Rng& rng, int ri, int32_t gp0, int32_t gq0, int32_t cell, int32_t cand);

struct CellTally {   // what the candidates of one thread add to their cell
    double newMax, nColl, sepSum, nReacted;   // nReacted: accepted pairs a reaction took (counted as collisions, not measured)
};

// candidate `cand` of cell c picks its pair (noTimeCounter.C:160-230): P anywhere in the cell, Q from P's sub-cell when that holds another parcel
__device__ __forceinline__ void pickPair(const CollideArgs& a, const DevParams& P, Rng& rng, const int32_t* subStart, int32_t b, int32_t nC, int32_t c,
                                         int32_t cand, int32_t& cp, int32_t& cq) {
    rng.init(P.seed, uint32_t(c), uint32_t(cand), a.step, STREAM_COLLIDE);
    cp = rng.randomLabel(0, nC - 1);
    const int sub = a.octKey[b + cp];
    const int32_t s0 = subStart[sub];
    const int32_t nSC = subStart[sub + 1] - s0;
    if (nSC > 1) {
        do {
            const int32_t k = s0 + rng.randomLabel(0, nSC - 1);
            cq = a.bigScratch[b + k];
        } while (cp == cq);
    } else {
        do { cq = rng.randomLabel(0, nC - 1); } while (cp == cq);
    }
}

// acceptance, reaction and collision of one candidate pair (noTimeCounter.C:232-320); the caller orders candidates that share a parcel
__device__ __forceinline__ void collideCandidate(const CollideArgs& a, const DevParams& P, const CellView& v, bool LB, Rng& rng, int32_t b, int32_t c,
                                                 int32_t cand, int32_t cp, int32_t cq, double sigmaTcRMaxLatched, CellTally& t) {
    const int tP = v.typ[cp], tQ = v.typ[cq];
    if (!(P.sp[tP].charge == -1 && P.sp[tQ].charge == -1)) {
        V3 UP = mk(v.ux[cp], v.uy[cp], v.uz[cp]);
        V3 UQ = mk(v.ux[cq], v.uy[cq], v.uz[cq]);
        const double cR0 = mag(UP - UQ);
        const double sTcR = sigmaTcR(P, tP, tQ, cR0);
        if (sTcR > t.newMax) t.newMax = sTcR;
        bool relax = (sTcR / sigmaTcRMaxLatched) > rng.sample01();
        if (relax && P.nReactions > 0) {
            // chemical reactions (noTimeCounter.C:250-303)
            const int rMId = P.pairReaction[tP][tQ];
            if (rMId >= 0) {
                relax = reactPair(a, P, rng, rMId, b + cp, b + cq, c, cand);
                if (!relax) t.nReacted += 1.0;
            }
        }
        if (relax) {
            double cR = -1;
            if (LB) {
                // LarsenBorgnakkeVariableHardSphere::collide
                const double mR = P.mR[tP][tQ];
                const double cRsqr = magSqr(UP - UQ);
                double translationalEnergy = 0.5 * mR * cRsqr;
                const double omegaPQ = P.omegaPQ[tP][tQ];
                redistribute(P, rng, v, cp, tQ, translationalEnergy, omegaPQ);
                redistribute(P, rng, v, cq, tP, translationalEnergy, omegaPQ);
                cR = sqrt(2.0 * translationalEnergy / mR);
            }
            postCollisionVelocities(P, rng, tP, tQ, UP, UQ, cR);
            v.ux[cp] = UP.x; v.uy[cp] = UP.y; v.uz[cp] = UP.z;
            v.ux[cq] = UQ.x; v.uy[cq] = UQ.y; v.uz[cq] = UQ.z;
            // cellMeasurements (VariableHardSphere.C:154-162)
            const int32_t gp = b + cp, gq = b + cq;
            const double dx = a.p.px[gp] - a.p.px[gq], dy = a.p.py[gp] - a.p.py[gq], dz = a.p.pz[gp] - a.p.pz[gq];
            t.sepSum += sqrt(dx * dx + dy * dy + dz * dz);
            t.nColl += 1.0;
            // classification promotion (VariableHardSphere.C:164-187)
            if (a.p.cls) {
                const int clP = a.p.cls[gp], clQ = a.p.cls[gq];
                if (clP == 0 && (clQ == 1 || clQ == 2)) a.p.cls[gp] = 2;
                if (clQ == 0 && (clP == 1 || clP == 2)) a.p.cls[gq] = 2;
            }
        }
    }
}

}  // namespace

__global__ void __launch_bounds__(COL_WARPS * 32) collideBigCellsKernel(const __grid_constant__ CollideArgs a) {
    int32_t* const bigScratch = a.bigScratch;
    __shared__ WarpSmem smAll[COL_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    WarpSmem& sm = smAll[w];
    const DevParams& P = *a.P;
    const bool LB = P.collisionModel == DSMCB200_COLL_LB_VHS || P.collisionModel == DSMCB200_COLL_LB_VSS;
    const int32_t nWarps = gridDim.x * COL_WARPS;
    unsigned long long totColl = 0, totCand = 0;

    const int32_t nBig = a.counters->bigCells;  // the cells collideLaneKernel left to this kernel
    for (int32_t ib = blockIdx.x * COL_WARPS + w; ib < nBig; ib += nWarps) {
        const int32_t c = a.bigList[ib];
        const int32_t b = a.cellOffset[c];
        const int32_t nC = a.cellOffset[c + 1] - b;
        if (nC <= BIG_CELL_THRESHOLD || nC > GIANT_SORT || P.collisionModel == DSMCB200_COLL_NONE) continue;  // collideLaneKernel / collideGiantCellsKernel
        CellView v;
        v.tMacro = a.overallT ? a.overallT + c : nullptr;
        v.ux = a.p.ux + b; v.uy = a.p.uy + b; v.uz = a.p.uz + b; v.erot = a.p.erot ? a.p.erot + b : nullptr;
        for (int m = 0; m < MAX_MODES; ++m) v.vib[m] = a.p.vib[m] ? a.p.vib[m] + b : nullptr;
        v.typ = a.p.typeId + b; v.elev = a.p.elevel ? a.p.elevel + b : nullptr;
        __syncwarp();

        // ---- the 8 Cartesian sub-cells (noTimeCounter.C:112-138): stable counting sort of parcel indices by octant
        int32_t cnt[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) cnt[s] = 0;
        // the sub-cell key of every sorted parcel was written by the sort's gather (the lane kernel reads the same bytes); four
        // rows of 32 keys are in flight per iteration: a cell of 1e5 parcels is 6000 dependent round trips otherwise
        for (int j0 = 0; j0 < nC; j0 += 128) {
            int o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int j = j0 + 32 * u + lane; o[u] = j < nC ? int(a.octKey[b + j]) : -1; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int s = 0; s < 8; ++s) cnt[s] += __popc(__ballot_sync(0xffffffffu, o[u] == s));
        }
        int32_t start[9];
        start[0] = 0;
#pragma unroll
        for (int s = 0; s < 8; ++s) start[s + 1] = start[s] + cnt[s];
        if (lane < 9) {
            int32_t val = 0;
#pragma unroll
            for (int s = 0; s < 9; ++s) if (lane == s) val = start[s];
            sm.subStart[lane] = val;
        }
        {
            int32_t run[8];
#pragma unroll
            for (int s = 0; s < 8; ++s) run[s] = start[s];
            for (int j0 = 0; j0 < nC; j0 += 128) {
                int o[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int j = j0 + 32 * u + lane; o[u] = j < nC ? int(a.octKey[b + j]) : -1; }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u + lane;
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const unsigned m = __ballot_sync(0xffffffffu, o[u] == s);
                        if (o[u] == s) {
                            const int32_t posn = run[s] + __popc(m & ((1u << lane) - 1u));
                            bigScratch[b + posn] = j;
                        }
                        run[s] += __popc(m);
                    }
                }
            }
        }
        __syncwarp();

        // ---- number of candidate pairs (noTimeCounter.C:142-155)
        const double sigmaTcRMaxLatched = a.sigmaTcRMax[c];
        const double selectedPairs =
            a.remainder[c] + 0.5 * nC * (nC - 1) * a.cf.nParticles(P.nParticles, c) * sigmaTcRMaxLatched * a.cf.deltaT(P.deltaT, c) / a.cellVolumes[c];
        const int32_t nCandidates = int32_t(selectedPairs);
        if (lane == 0) a.remainder[c] = selectedPairs - nCandidates;
        totCand += (lane == 0) ? (unsigned long long)(nCandidates > 0 ? nCandidates : 0) : 0ULL;

        CellTally t{sigmaTcRMaxLatched, 0.0, 0.0, 0.0};

        for (int32_t c0 = 0; c0 < nCandidates; c0 += 32) {
            const int32_t cand = c0 + lane;
            const bool active = cand < nCandidates;
            int32_t cp = -1, cq = -2;
            Rng rng;
            if (active) {
                pickPair(a, P, rng, sm.subStart, b, nC, c, cand, cp, cq);
            }
            const unsigned activeMask = __ballot_sync(0xffffffffu, active);
            // earlier candidates of the batch that touch one of my parcels
            unsigned confl = 0;
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
                const int32_t pj = __shfl_sync(0xffffffffu, cp, j);
                const int32_t qj = __shfl_sync(0xffffffffu, cq, j);
                if (j < lane && (pj == cp || pj == cq || qj == cp || qj == cq)) confl |= 1u << j;
            }
            confl &= activeMask;
            unsigned done = ~activeMask;
            while (done != 0xffffffffu) {
                const bool ready = active && !((done >> lane) & 1u) && ((confl & ~done) == 0u);
                if (ready) {
                    collideCandidate(a, P, v, LB, rng, b, c, cand, cp, cq, sigmaTcRMaxLatched, t);
                }
                __syncwarp();
                done |= __ballot_sync(0xffffffffu, ready);
            }
        }

        // ---- per-cell results
        double mx = t.newMax;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const double nCollTot = warpSumOrdered(t.nColl);
        const double sepTot = warpSumOrdered(t.sepSum);
        const double nReactedTot = warpSumOrdered(t.nReacted);
        if (lane == 0) {
            a.sigmaTcRMax[c] = mx;
            a.nCollsStep[c] = nCollTot;
            a.collSepStep[c] = sepTot;
            totColl += (unsigned long long)nCollTot + (unsigned long long)nReactedTot;
        }
        __syncwarp();
    }
    if (lane == 0 && (totColl | totCand)) {
        atomicAdd(&a.counters->collisions, totColl);
        atomicAdd(&a.counters->candidates, totCand);
    }
}

// ------------------------------------------------------------------------------------------------
// Cells of more than GIANT_SORT parcels (a heat bath in one cell: 2e5 parcels, 2e4 candidates a step): one block per cell.
// The same algorithm as collideBigCellsKernel with batches of GIANT_THREADS candidates: a candidate runs once every EARLIER
// candidate of its batch that shares a parcel with it has run, so the outcome is that of the reference's serial loop.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int GIANT_THREADS = 512, GIANT_WARPS = GIANT_THREADS / 32;
}

__global__ void __launch_bounds__(GIANT_THREADS) collideGiantCellsKernel(const __grid_constant__ CollideArgs a) {
    __shared__ int32_t subStart[9];
    __shared__ int32_t warpCnt[GIANT_WARPS][8];
    __shared__ int32_t cpS[GIANT_THREADS], cqS[GIANT_THREADS];
    __shared__ __align__(16) unsigned doneS[GIANT_WARPS];   // own 16-byte granules: the words are read with vector loads
    __shared__ double red[GIANT_WARPS][4];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const DevParams& P = *a.P;
    if (P.collisionModel == DSMCB200_COLL_NONE) return;
    const bool LB = P.collisionModel == DSMCB200_COLL_LB_VHS || P.collisionModel == DSMCB200_COLL_LB_VSS;
    const int32_t c = a.giantList[blockIdx.x];
    const int32_t b = a.cellOffset[c];
    const int32_t nC = a.cellOffset[c + 1] - b;
    CellView v;
    v.tMacro = a.overallT ? a.overallT + c : nullptr;
    v.ux = a.p.ux + b; v.uy = a.p.uy + b; v.uz = a.p.uz + b; v.erot = a.p.erot ? a.p.erot + b : nullptr;
    for (int m = 0; m < MAX_MODES; ++m) v.vib[m] = a.p.vib[m] ? a.p.vib[m] + b : nullptr;
    v.typ = a.p.typeId + b; v.elev = a.p.elevel ? a.p.elevel + b : nullptr;

    // ---- the 8 sub-cells: stable counting sort of the parcel indices by octant; warp w owns the w-th contiguous piece of the cell
    const int32_t piece = ((nC + GIANT_WARPS * 128 - 1) / (GIANT_WARPS * 128)) * 128;
    const int32_t j0w = w * piece, j1w = min(nC, j0w + piece);
    {
        int32_t cnt[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) cnt[s] = 0;
        for (int j0 = j0w; j0 < j1w; j0 += 128) {
            int o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int j = j0 + 32 * u + lane; o[u] = j < j1w ? int(a.octKey[b + j]) : -1; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int s = 0; s < 8; ++s) cnt[s] += __popc(__ballot_sync(0xffffffffu, o[u] == s));
        }
        if (lane < 8) {
            int32_t val = 0;
#pragma unroll
            for (int s = 0; s < 8; ++s) if (lane == s) val = cnt[s];
            warpCnt[w][lane] = val;
        }
    }
    __syncthreads();
    {
        int32_t run[8];
        int32_t acc = 0;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            int32_t before = 0, total = 0;
#pragma unroll
            for (int ww = 0; ww < GIANT_WARPS; ++ww) { const int32_t k = warpCnt[ww][s]; total += k; if (ww < w) before += k; }
            run[s] = acc + before;
            if (tid == 0) subStart[s] = acc;
            acc += total;
        }
        if (tid == 0) subStart[8] = acc;
        for (int j0 = j0w; j0 < j1w; j0 += 128) {
            int o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int j = j0 + 32 * u + lane; o[u] = j < j1w ? int(a.octKey[b + j]) : -1; }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 32 * u + lane;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const unsigned m = __ballot_sync(0xffffffffu, o[u] == s);
                    if (o[u] == s) a.bigScratch[b + run[s] + __popc(m & ((1u << lane) - 1u))] = j;
                    run[s] += __popc(m);
                }
            }
        }
    }
    __syncthreads();

    // ---- number of candidate pairs (noTimeCounter.C:142-155)
    const double sigmaTcRMaxLatched = a.sigmaTcRMax[c];
    const double selectedPairs =
        a.remainder[c] + 0.5 * nC * (nC - 1) * a.cf.nParticles(P.nParticles, c) * sigmaTcRMaxLatched * a.cf.deltaT(P.deltaT, c) / a.cellVolumes[c];
    const int32_t nCandidates = int32_t(selectedPairs);
    __syncthreads();   // every thread has read remainder[c]
    if (tid == 0) a.remainder[c] = selectedPairs - nCandidates;

    CellTally t{sigmaTcRMaxLatched, 0.0, 0.0, 0.0};
    for (int32_t c0 = 0; c0 < nCandidates; c0 += GIANT_THREADS) {
        const int32_t cand = c0 + tid;
        const bool active = cand < nCandidates;
        int32_t cp = -1, cq = -2;
        Rng rng;
        if (active) pickPair(a, P, rng, subStart, b, nC, c, cand, cp, cq);
        cpS[tid] = cp; cqS[tid] = cq;
        __syncthreads();
        // earlier candidates of the batch that touch one of my parcels, one word per warp of the batch
        unsigned confl[GIANT_WARPS];
#pragma unroll
        for (int k = 0; k < GIANT_WARPS; ++k) {
            confl[k] = 0;
            if (k <= w && active) {
                for (int jj = 0; jj < 32; ++jj) {
                    const int j = 32 * k + jj;
                    const int32_t pj = cpS[j], qj = cqS[j];
                    if (j < tid && (pj == cp || pj == cq || qj == cp || qj == cq)) confl[k] |= 1u << jj;
                }
            }
        }
        const unsigned activeMask = __ballot_sync(0xffffffffu, active);
        if (lane == 0) doneS[w] = ~activeMask;
        __syncthreads();
        bool mine = !active;   // this thread's candidate has run
        for (;;) {
            unsigned all = 0xffffffffu, pending = 0;
#pragma unroll
            for (int k = 0; k < GIANT_WARPS; ++k) { const unsigned d = doneS[k]; all &= d; pending |= confl[k] & ~d; }
            if (all == 0xffffffffu) break;   // the same words for every thread: the block leaves together
            const bool ready = !mine && pending == 0u;
            if (ready) { collideCandidate(a, P, v, LB, rng, b, c, cand, cp, cq, sigmaTcRMaxLatched, t); mine = true; }
            __syncthreads();   // doneS was read by everyone, the parcels written by this round are visible to the block
            const unsigned r = __ballot_sync(0xffffffffu, ready);
            if (lane == 0 && r) doneS[w] |= r;
            __syncthreads();
        }
    }

    // ---- per-cell results: fixed-shape trees, the same order on every run
    double mx = t.newMax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const double nCollW = warpSumOrdered(t.nColl), sepW = warpSumOrdered(t.sepSum), nReactedW = warpSumOrdered(t.nReacted);
    if (lane == 0) { red[w][0] = mx; red[w][1] = nCollW; red[w][2] = sepW; red[w][3] = nReactedW; }
    __syncthreads();
    if (tid == 0) {
        double m = red[0][0], nCollTot = red[0][1], sepTot = red[0][2], nReactedTot = red[0][3];
        for (int k = 1; k < GIANT_WARPS; ++k) { m = fmax(m, red[k][0]); nCollTot += red[k][1]; sepTot += red[k][2]; nReactedTot += red[k][3]; }
        a.sigmaTcRMax[c] = m;
        a.nCollsStep[c] = nCollTot;
        a.collSepStep[c] = sepTot;
        atomicAdd(&a.counters->collisions, (unsigned long long)nCollTot + (unsigned long long)nReactedTot);
        atomicAdd(&a.counters->candidates, (unsigned long long)(nCandidates > 0 ? nCandidates : 0));
    }
}

// ------------------------------------------------------------------------------------------------
// Lane-per-cell kernel: a warp takes 32 consecutive cells and every lane walks the candidates of ITS cell
// in the reference's serial order.  Candidates of one cell are then ordered by construction (no owner table,
// no atomics) and every lane of the warp executes the expensive parts -- Philox, pow() of sigmaTcR, the
// Larsen-Borgnakke exchange -- at the same time.  Only the per-parcel octant keys, species and the sub-cell
// index lists live in shared memory (3 bytes per parcel); velocities and internal energies of the ~2 parcels a
// candidate touches are read and written in place.  Cells with more than LANE_CELL_MAX parcels are left to
// collideBigCellsKernel.
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int LANE_WARPS = 4;
constexpr int LANE_CAP = 2048;       // parcels of one pass (<= 32 cells) per warp
constexpr int LANE_CELL_MAX = 255;   // indices within a cell are stored in a byte

struct LaneSmem {
    uint8_t key[LANE_CAP], typ[LANE_CAP], sub[LANE_CAP];
    uint8_t cnt[8][32];    // [octant][lane]: running fill position of the lane's cell
    uint8_t start[9][32];  // [octant][lane]: first slot of the octant in the cell's sub-list
};

struct InPlace {  // accessor of the parcels of one cell where they lie in the sorted cloud
    const ParcelArrays& p;
    int32_t b;
    __device__ __forceinline__ int elev(int j) const { return p.elevel[b + j]; }
    __device__ __forceinline__ void setElev(int j, int v) const { p.elevel[b + j] = uint8_t(v); }
    __device__ __forceinline__ int32_t vib(int m, int j) const { return p.vib[m][b + j]; }
    __device__ __forceinline__ void setVib(int m, int j, int32_t v) const { p.vib[m][b + j] = v; }
    __device__ __forceinline__ double erot(int j) const { return p.erot[b + j]; }
    __device__ __forceinline__ void setErot(int j, double v) const { p.erot[b + j] = v; }
};

// LarsenBorgnakkeVariableHardSphere::redistribute on parcel j of the view
// NM: vibrational modes stored per parcel (P.nModes), a compile-time bound of the mode loops
template <bool ZV2008, int NM>
__device__ __forceinline__ void redistributeInPlace(const DevParams& P, Rng& rng, const InPlace v, int j, int tSelf, int tOther,
                                                 double& translationalEnergy, double omegaPQ, const double* tMacro, bool postReaction = false) {
    const DevSpecies& S = P.sp[tSelf];
    if (S.type == 0) return;  // electron
    if (P.invZelec > rng.sample01()) {
        const double EcP = translationalEnergy + S.eElec[v.elev(j)];
        const int lvl = postCollisionElectronicEnergyLevel(rng, EcP, omegaPQ, S);
        v.setElev(j, lvl);
        translationalEnergy = EcP - S.eElec[lvl];
    }
    if (NM > 0 && S.nVib > 0) {
        double preEVib[NM > 0 ? NM : 1];
        int32_t lvl0[NM > 0 ? NM : 1];
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            lvl0[m] = m < S.nVib ? v.vib(m, j) : 0;
            preEVib[m] = m < S.nVib ? lvl0[m] * P.kB * S.thetaV[m] : 0.0;
        }
#pragma unroll
        for (int m = 0; m < NM; ++m) {
            if (m >= S.nVib) break;
            const double EcP = translationalEnergy + preEVib[m];
            const int32_t iMaxP = int32_t(EcP / (P.kB * S.thetaV[m]));
            if (iMaxP > 0) {
                const int32_t lvl = postCollisionVibrationalEnergyLevel<ZV2008>(P, rng, lvl0[m], iMaxP, S.thetaV[m], S.thetaD, S.TrefZv[m], omegaPQ,
                                                                        S.Zref[m], EcP,
                                                                        P.invZvTab + ((size_t(tSelf) * P.nSpecies + tOther) * MAX_MODES + m) * ZV_TABLE, tMacro,
                                                                        postReaction);
                if (lvl != lvl0[m]) v.setVib(m, j, lvl);
                translationalEnergy = EcP - lvl * P.kB * S.thetaV[m];
            }
        }
    }
    if (S.rotDof > 0) {
        const double preCollisionERotP = v.erot(j);
        if (P.invZrot > rng.sample01()) {
            const double EcP = translationalEnergy + preCollisionERotP;
            const double ChiB = 2.5 - omegaPQ;
            const double energyRatio = postCollisionRotationalEnergy(rng, S.rotDof, ChiB);
            const double e = energyRatio * EcP;
            v.setErot(j, e);
            translationalEnergy = EcP - e;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Quantum-kinetic chemistry inside the candidate loop (noTimeCounter.C:250-303): the reaction model of the typeId pair runs before the
// conventional collision and says whether that still takes place (relax()).  Parcels are read and written where they lie in the cloud.
//   dissociationQK          DSMC/reactions/derived/dissociationQK/dissociationQK.C:197-383, 481-583
//   exchangeQK              DSMC/reactions/derived/exchangeQK/exchangeQK.C:178-362, 457-511
//   dissociationExchangeQK  DSMC/reactions/derived/mixed/dissociationExchangeQK/dissociationExchangeQK.C:107-264
// ------------------------------------------------------------------------------------------------
struct PairRef {   // the two parcels of the candidate, by cloud index
    int32_t g[2];
};

__device__ __forceinline__ double eVibTot(const DevParams& P, const ParcelArrays& p, int32_t g, const DevSpecies& S) {
    double e = 0.0;
#pragma unroll
    for (int m = 0; m < MAX_MODES; ++m) if (m < S.nVib) e += p.vib[m][g] * P.kB * S.thetaV[m];
    return e;
}
__device__ __forceinline__ void clearInternal(const DevParams& P, const ParcelArrays& p, int32_t g) {
    if (p.erot) p.erot[g] = 0.0;
#pragma unroll
    for (int m = 0; m < MAX_MODES; ++m) if (m < P.nModes) p.vib[m][g] = 0;
    if (p.elevel) p.elevel[g] = 0;
}

// VariableHardSphere::postReactionVelocities, VariableHardSphere.C:231-262 (UP enters as the centre-of-mass velocity)
__device__ void postReactionVelocities(const DevParams& P, Rng& rng, int tP, int tQ, V3& UP, V3& UQ, double cR) {
    const double mP = P.sp[tP].mass, mQ = P.sp[tQ].mass;
    const double cosTheta = 2.0 * rng.sample01() - 1.0;
    const double sinTheta = sqrt(1.0 - cosTheta * cosTheta);
    const double phi = TWO_PI * rng.sample01();
    const V3 rel = cR * mk(cosTheta, sinTheta * cos(phi), sinTheta * sin(phi));
    UQ = UP - rel * mP / (mP + mQ);
    UP = UP + rel * mQ / (mP + mQ);
}

// dissociationQK::testDissociation
__device__ void testDissociation(const DevParams& P, const ParcelArrays& p, int32_t g, double translationalEnergy, int& vibModeDisso,
                                 double& collisionEnergy, double& total, double& prob) {
    const DevSpecies& S = P.sp[p.typeId[g]];
    if (S.type == 20 || S.type == 30) {
#pragma unroll
        for (int m = 0; m < MAX_MODES; ++m) {
            if (m >= S.nVib) break;
            const double EVibP_m = p.vib[m][g] * P.kB * S.thetaV[m];
            const int32_t idP = int32_t(S.thetaD / S.thetaV[m]);   // charDissQuantumLevel_m (dsmcParcelI.H:152-156)
            collisionEnergy = translationalEnergy + EVibP_m;
            const int32_t imaxP = int32_t(collisionEnergy / (P.kB * S.thetaV[m]));
            if (imaxP > idP) { prob = 1.0; total += prob; vibModeDisso = m; break; }
        }
    }
}

// exchangeQK::testExchange
__device__ void testExchange(const DevParams& P, const DevReaction& R, const ParcelArrays& p, int32_t g, double translationalEnergy, double omegaPQ,
                             double& collisionEnergy, double& total, double& prob) {
    const DevSpecies& S = P.sp[p.typeId[g]];
    const double chiB = 2.5 - omegaPQ;
    const double TColl = translationalEnergy / (P.kB * chiB);
    double activationEnergy = R.aDash * powNI(TColl / 273.0, R.bCoeff) * fabs(R.heatExchJ);
    double summation = 1.0;
    if (R.heatExchJ < 0.0) activationEnergy -= R.heatExchJ;
    const int nV = S.nVib;
    if (nV == 0) { total += prob; return; }
    int m = 0;
    do {
        const double kBByThetaVP = P.kB * S.thetaV[m];
        const double EVibP_m = p.vib[m][g] * P.kB * S.thetaV[m];
        collisionEnergy = translationalEnergy + EVibP_m;
        if (collisionEnergy > activationEnergy) {
            if (activationEnergy > kBByThetaVP) {
                summation = 0.0;
                const int32_t iaP = int32_t(collisionEnergy / kBByThetaVP);
                for (int32_t i = 0; i <= iaP; ++i) summation += powNI(1.0 - (i * P.kB * S.thetaV[m]) / collisionEnergy, 1.5 - omegaPQ);
            }
            prob = powNI(1.0 - activationEnergy / collisionEnergy, 1.5 - omegaPQ) / summation;
            m = nV;
        }
        m += 1;
    } while (m < nV);
    total += prob;
}

// dissociationQK::dissociateParticleByPartner: parcel gP splits into products[nR], gQ is the partner
__device__ void dissociateParticleByPartner(const CollideArgs& a, const DevParams& P, Rng& rng, const DevReaction& R, int ri, int32_t gP, int32_t gQ,
                                            int nR, int vibModeDisso, double collisionEnergy, int32_t cell, int32_t cand, bool& relax) {
    const ParcelArrays& p = a.p;
    const int tP = p.typeId[gP], tQ = p.typeId[gQ];
    const int nReac = tP == tQ ? 0 : nR;
    atomicAdd(&a.counters->nReact[ri][nReac], 1ULL);
    if (!R.allowSplitting) return;
    relax = false;
    collisionEnergy -= R.heatDissJ[nR];
    const double omegaPQ = 0.5 * (P.sp[tP].omega + P.sp[tQ].omega);
    const double* tMacro = a.overallT ? a.overallT + cell : nullptr;
    redistributeInPlace<true, MAX_MODES>(P, rng, InPlace{p, 0}, gQ, tQ, tP, collisionEnergy, omegaPQ, tMacro, true);
    const double mP = P.sp[tP].mass, mQ = P.sp[tQ].mass, mR = mP * mQ / (mP + mQ);
    const double relVelNonDissoParticle = sqrt(2.0 * collisionEnergy / mR);
    V3 UP = mk(p.ux[gP], p.uy[gP], p.uz[gP]), UQ = mk(p.ux[gQ], p.uy[gQ], p.uz[gQ]);
    postCollisionVelocities(P, rng, tP, tQ, UP, UQ, relVelNonDissoParticle);
    const int typeId1 = R.dissProd[nR][0], typeId2 = R.dissProd[nR][1];
    const double mP1 = P.sp[typeId1].mass, mP2 = P.sp[typeId2].mass, mRproducts = mP1 * mP2 / (mP1 + mP2);
    const DevSpecies& SP = P.sp[tP];
    const double ERotP = p.erot ? p.erot[gP] : 0.0;
    const double EVibP_tot = eVibTot(P, p, gP, SP);
    const double EVibP_mdisso = p.vib[vibModeDisso][gP] * P.kB * SP.thetaV[vibModeDisso];
    const double EVibP_nondisso = EVibP_tot - EVibP_mdisso;
    const double EEleP = SP.eElec[p.elevel ? p.elevel[gP] : 0];
    const double translationalEnergyLeft = ERotP + EVibP_nondisso + EEleP;
    const double cRproducts = sqrt(2.0 * translationalEnergyLeft / mRproducts);
    V3 UP2 = mk(0.0, 0.0, 0.0);
    postReactionVelocities(P, rng, typeId1, typeId2, UP, UP2, cRproducts);
    p.typeId[gP] = uint8_t(typeId1);
    clearInternal(P, p, gP);
    p.ux[gP] = UP.x; p.uy[gP] = UP.y; p.uz[gP] = UP.z;
    p.ux[gQ] = UQ.x; p.uy[gQ] = UQ.y; p.uz[gQ] = UQ.z;
    // cloud_.addNewParcel(position, UP2, ..., cell, tetFace, tetPt, typeId2, -1, classification, vibLevel 0)
    const int32_t k = atomicAdd(&a.counters->nBorn, 1);
    if (k < a.bornCapacity) {
        BornRec b;
        b.key = (static_cast<unsigned long long>(uint32_t(cell)) << 32) | uint32_t(cand);
        b.pos[0] = p.px[gP]; b.pos[1] = p.py[gP]; b.pos[2] = p.pz[gP];
        b.U[0] = UP2.x; b.U[1] = UP2.y; b.U[2] = UP2.z;
        b.rwf = p.rwf ? p.rwf[gP] : 1.0;
        b.cell = cell; b.tet = p.tet[gP];
        b.typeId = uint8_t(typeId2); b.cls = p.cls ? p.cls[gP] : 0;
        for (int i = 0; i < 6; ++i) b.pad_[i] = 0;
        a.born[k] = b;
    } else {
        atomicAdd(&a.counters->overflow, 1ULL);
    }
}

// exchangeQK::exchange (gP: the molecule, becomes the atom; gQ: the atom, becomes the molecule)
__device__ void exchangeParcels(const CollideArgs& a, const DevParams& P, Rng& rng, const DevReaction& R, int ri, int32_t gP, int32_t gQ, int32_t cell,
                                bool& relax) {
    const ParcelArrays& p = a.p;
    atomicAdd(&a.counters->nReact[ri][2], 1ULL);
    if (!R.allowSplitting) return;
    relax = false;
    const int tP = p.typeId[gP], tQ = p.typeId[gQ];
    const V3 UP0 = mk(p.ux[gP], p.uy[gP], p.uz[gP]), UQ0 = mk(p.ux[gQ], p.uy[gQ], p.uz[gQ]);
    const double mP = P.sp[tP].mass, mQ = P.sp[tQ].mass, mR = mP * mQ / (mP + mQ);
    const double cRsqr = magSqr(UP0 - UQ0);
    double translationalEnergy = 0.5 * mR * cRsqr;
    const V3 Ucm = (mP * UP0 + mQ * UQ0) / (mP + mQ);
    const int typeIdMol = R.exchProd[0], typeIdAtom = R.exchProd[1];
    const double mPExch = P.sp[typeIdAtom].mass, mQExch = P.sp[typeIdMol].mass, mRExch = mPExch * mQExch / (mPExch + mQExch);
    const double omegaExch = 0.5 * (P.sp[typeIdAtom].omega + P.sp[typeIdAtom].omega);   // the atom twice, as in the reference (:296-301)
    const double EVibP = eVibTot(P, p, gP, P.sp[tP]);
    const double EEleP = P.sp[tP].eElec[p.elevel ? p.elevel[gP] : 0], EEleQ = P.sp[tQ].eElec[p.elevel ? p.elevel[gQ] : 0];
    translationalEnergy += (p.erot ? p.erot[gP] : 0.0) + EVibP + EEleP + EEleQ + R.heatExchJ;
    p.typeId[gP] = uint8_t(typeIdAtom); clearInternal(P, p, gP);
    p.typeId[gQ] = uint8_t(typeIdMol); clearInternal(P, p, gQ);
    const double* tMacro = a.overallT ? a.overallT + cell : nullptr;
    redistributeInPlace<true, MAX_MODES>(P, rng, InPlace{p, 0}, gQ, typeIdMol, typeIdAtom, translationalEnergy, omegaExch, tMacro, true);
    const double relVelExchMol = sqrt(2.0 * translationalEnergy / mRExch);
    V3 UP = Ucm, UQ = UQ0;
    postReactionVelocities(P, rng, typeIdAtom, typeIdMol, UP, UQ, relVelExchMol);
    p.ux[gP] = UP.x; p.uy[gP] = UP.y; p.uz[gP] = UP.z;
    p.ux[gQ] = UQ.x; p.uy[gQ] = UQ.y; p.uz[gQ] = UQ.z;
}

// <model>::reaction(p, q); returns relax().  (gp0, gq0): the candidate pair in the order noTimeCounter passes it
__device__ __noinline__ bool reactPair(const CollideArgs& a, const DevParams& P, Rng& rng, int ri, int32_t gp0, int32_t gq0, int32_t cell, int32_t cand) {
    const ParcelArrays& p = a.p;
    const DevReaction& R = P.reactions[ri];
    bool relax = true;
    if (R.model == DSMCB200_REACT_EXCHANGE_QK) {
        int32_t gP = gp0, gQ = gq0;   // P must be the molecule
        { const int ty = P.sp[p.typeId[gp0]].type; if (ty == 10 || ty == 11) { gP = gq0; gQ = gp0; } }
        const int tP = p.typeId[gP], tQ = p.typeId[gQ];
        const double mP = P.sp[tP].mass, mQ = P.sp[tQ].mass, mR = mP * mQ / (mP + mQ);
        const double omegaPQ = 0.5 * (P.sp[tP].omega + P.sp[tQ].omega);
        const double translationalEnergy = 0.5 * mR * magSqr(mk(p.ux[gP], p.uy[gP], p.uz[gP]) - mk(p.ux[gQ], p.uy[gQ], p.uz[gQ]));
        double total = 0.0, prob = 0.0, Ecoll = 0.0;
        testExchange(P, R, p, gP, translationalEnergy, omegaPQ, Ecoll, total, prob);
        if (total > rng.sample01()) exchangeParcels(a, P, rng, R, ri, gP, gQ, cell, relax);
        return relax;
    }
    int32_t gP = gp0, gQ = gq0;
    if (p.typeId[gp0] != R.reactants[0]) { gP = gq0; gQ = gp0; }
    const int tP = p.typeId[gP], tQ = p.typeId[gQ];
    const double mP = P.sp[tP].mass, mQ = P.sp[tQ].mass, mR = mP * mQ / (mP + mQ);
    const double omegaPQ = 0.5 * (P.sp[tP].omega + P.sp[tQ].omega);
    const double translationalEnergy = 0.5 * mR * magSqr(mk(p.ux[gP], p.uy[gP], p.uz[gP]) - mk(p.ux[gQ], p.uy[gQ], p.uz[gQ]));
    const bool mixed = R.model == DSMCB200_REACT_DISSOCIATION_EXCHANGE_QK;
    const int nPoss = mixed ? 3 : 2;
    double total = 0.0, pr0 = 0.0, pr1 = 0.0, pr2 = 0.0, Ec0 = 0.0, Ec1 = 0.0, Ec2 = 0.0;
    int vibModeDissoP = -1, vibModeDissoQ = -1;
    testDissociation(P, p, gP, translationalEnergy, vibModeDissoP, Ec0, total, pr0);
    if (mixed || tP != tQ) testDissociation(P, p, gQ, translationalEnergy, vibModeDissoQ, Ec1, total, pr1);
    if (mixed) testExchange(P, R, p, R.posMolReactant == 0 ? gP : gQ, translationalEnergy, omegaPQ, Ec2, total, pr2);
    if (total > rng.sample01()) {
        const double n0 = pr0 / total, n1 = pr1 / total, n2 = pr2 / total;
        // dsmcReaction::decreasing_sort_indices (dsmcReaction.C:176-205): one random number per entry breaks ties
        const double r0 = rng.sample01(), r1 = rng.sample01(), r2 = mixed ? rng.sample01() : 0.0;
        auto val = [&](int i) { return i == 0 ? n0 : (i == 1 ? n1 : n2); };
        auto rnd = [&](int i) { return i == 0 ? r0 : (i == 1 ? r1 : r2); };
        int idx0 = 0, idx1 = 1, idx2 = 2;
        auto before = [&](int b, int a2) { return (val(b) == val(a2)) ? (rnd(b) > rnd(a2)) : (val(b) > val(a2)); };   // b sorts in front of a2
        if (before(idx1, idx0)) { const int t = idx0; idx0 = idx1; idx1 = t; }
        if (nPoss == 3) {
            if (before(idx2, idx1)) { const int t = idx1; idx1 = idx2; idx2 = t; if (before(idx1, idx0)) { const int t2 = idx0; idx0 = idx1; idx1 = t2; } }
        }
        double cumulative = 0.0;
        for (int k = 0; k < nPoss; ++k) {
            const int i = k == 0 ? idx0 : (k == 1 ? idx1 : idx2);
            const double ni = val(i);
            if (!(ni > SMALL)) break;
            cumulative += ni;
            if (cumulative > rng.sample01()) {
                if (i == 0) dissociateParticleByPartner(a, P, rng, R, ri, gP, gQ, 0, vibModeDissoP, Ec0, cell, cand, relax);
                else if (i == 1) dissociateParticleByPartner(a, P, rng, R, ri, gQ, gP, 1, vibModeDissoQ, Ec1, cell, cand, relax);
                else if (R.posMolReactant == 0) exchangeParcels(a, P, rng, R, ri, gP, gQ, cell, relax);
                else exchangeParcels(a, P, rng, R, ri, gQ, gP, cell, relax);
                break;
            }
        }
    }
    return relax;
}

__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
}  // namespace

template <bool ZV2008, int NM, bool CHEM>
__global__ void __launch_bounds__(LANE_WARPS * 32, 4) collideLaneKernel(const __grid_constant__ CollideArgs a) {
    __shared__ LaneSmem smAll[LANE_WARPS];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    LaneSmem& sm = smAll[w];
    const DevParams& P = *a.P;
    const bool LB = P.collisionModel == DSMCB200_COLL_LB_VHS || P.collisionModel == DSMCB200_COLL_LB_VSS;
    const bool internal = P.hasInternalEnergy != 0;
    const bool none = P.collisionModel == DSMCB200_COLL_NONE;
    const int32_t nWarps = gridDim.x * LANE_WARPS;
    const int32_t nGroups = (a.nCells + 31) / 32;
    unsigned long long totColl = 0, totCand = 0;

    for (int32_t grp = blockIdx.x * LANE_WARPS + w; grp < nGroups; grp += nWarps) {
        const int32_t c = grp * 32 + lane;  // this lane's cell
        const bool haveCell = c < a.nCells;
        const int32_t off = haveCell ? a.cellOffset[c] : a.cellOffset[a.nCells];
        const int32_t n = haveCell ? a.cellOffset[c + 1] - off : 0;
        const int nG = a.nCells - grp * 32 < 32 ? a.nCells - grp * 32 : 32;
        if (none) {
            if (haveCell) { a.nCollsStep[c] = 0.0; a.collSepStep[c] = 0.0; }
            continue;
        }
        if (n > LANE_CELL_MAX) a.bigList[atomicAdd(&a.counters->bigCells, 1)] = c;  // one warp per such cell, afterwards

        for (int g0 = 0; g0 < nG;) {
            // ---- the pass: the longest run of cells [g0, g1) whose parcels fit the shared-memory lists ----
            const int32_t passBeg = __shfl_sync(FULL, off, g0);
            const unsigned fitMask = __ballot_sync(FULL, lane >= g0 && lane < nG && off + n - passBeg <= LANE_CAP);
            const unsigned notFit = ~(fitMask >> g0);  // bit r: cell g0 + r does not fit any more
            const int g1 = g0 + (notFit ? __ffs(notFit) - 1 : 32);
            if (g1 == g0) {  // a single cell larger than the lists: collideBigCellsKernel
                ++g0;
                continue;
            }
            const bool inPass = lane >= g0 && lane < g1;
            const bool mine = inPass && n <= LANE_CELL_MAX;  // larger cells: collideBigCellsKernel
            const int32_t rel = off - passBeg;
            __syncwarp();
            // ---- octant key (from the sort) and species of every parcel of the pass: one coalesced sweep ----
            {
                const int32_t passEnd = __shfl_sync(FULL, off + n, g1 - 1);
                const int32_t nPass = passEnd - passBeg;
                for (int j = lane; j < nPass; j += 32) {
                    sm.key[j] = a.octKey[passBeg + j];
                    sm.typ[j] = a.p.typeId[passBeg + j];
                    if ((j & 3) == 0) {  // one 32-byte sector of each velocity component towards L2 ahead of the candidate loop
                        const int32_t g = passBeg + j;
                        prefetchL2(a.p.ux + g); prefetchL2(a.p.uy + g); prefetchL2(a.p.uz + g);
                        if (internal) prefetchL2(a.p.erot + g);
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < 8; ++s) sm.cnt[s][lane] = 0;
            __syncwarp();
            // ---- sub-cell lists: every lane does the stable counting sort of its own cell ----
            const int nMine = (mine && n > 1) ? n : 0;
            int nMax = nMine;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nMax = max(nMax, __shfl_xor_sync(FULL, nMax, o));
            for (int j = 0; j < nMax; ++j)
                if (j < nMine) sm.cnt[sm.key[rel + j]][lane] += 1;
            {
                int acc = 0;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int cs = sm.cnt[s][lane];
                    sm.start[s][lane] = uint8_t(acc);
                    sm.cnt[s][lane] = uint8_t(acc);
                    acc += cs;
                }
                sm.start[8][lane] = uint8_t(acc);
            }
            for (int j = 0; j < nMax; ++j) {
                if (j < nMine) {
                    const int k = sm.key[rel + j];
                    const int posn = sm.cnt[k][lane];
                    sm.cnt[k][lane] = uint8_t(posn + 1);
                    sm.sub[rel + posn] = uint8_t(j);
                }
            }
            // ---- candidate count (noTimeCounter.C:142-155) ----
            double sigmaL = 0.0;
            int32_t nCand = 0;
            if (nMine) {
                sigmaL = a.sigmaTcRMax[c];
                const double selectedPairs = a.remainder[c] + 0.5 * n * (n - 1) * a.cf.nParticles(P.nParticles, c) * sigmaL * a.cf.deltaT(P.deltaT, c) / a.cellVolumes[c];
                nCand = int32_t(selectedPairs);
                a.remainder[c] = selectedPairs - nCand;
                if (nCand < 0) nCand = 0;
            }
            double cellMax = sigmaL, cellColl = 0.0, cellSep = 0.0;
            const InPlace v{a.p, off};
            // ---- the candidates of the lane's cell, in order ----
            for (int32_t k = 0; __any_sync(FULL, k < nCand); ++k) {
                if (k < nCand) {
                    Rng rng;
                    rng.init(P.seed, uint32_t(c), uint32_t(k), a.step, STREAM_COLLIDE);
                    const int32_t cp = rng.randomLabel(0, n - 1);
                    int32_t cq;
                    const int sub = sm.key[rel + cp];
                    const int32_t s0 = sm.start[sub][lane];
                    const int32_t nSC = int32_t(sm.start[sub + 1][lane]) - s0;
                    if (nSC > 1) {
                        do { cq = sm.sub[rel + s0 + rng.randomLabel(0, nSC - 1)]; } while (cp == cq);
                    } else {
                        do { cq = rng.randomLabel(0, n - 1); } while (cp == cq);
                    }
                    const int tP = sm.typ[rel + cp], tQ = sm.typ[rel + cq];
                    if (!(P.sp[tP].charge == -1 && P.sp[tQ].charge == -1)) {
                        const int32_t gp = off + cp, gq = off + cq;
                        V3 UP = mk(a.p.ux[gp], a.p.uy[gp], a.p.uz[gp]);
                        V3 UQ = mk(a.p.ux[gq], a.p.uy[gq], a.p.uz[gq]);
                        const double sTcR = sigmaTcR(P, tP, tQ, mag(UP - UQ));
                        if (sTcR > cellMax) cellMax = sTcR;
                        bool relax = (sTcR / sigmaL) > rng.sample01();
                        if constexpr (CHEM) {
                            if (relax) {   // chemical reactions (noTimeCounter.C:250-303)
                                const int rMId = P.pairReaction[tP][tQ];
                                if (rMId >= 0) {
                                    relax = reactPair(a, P, rng, rMId, gp, gq, c, k);
                                    if (!relax) {   // later candidates of the cell see the new species
                                        sm.typ[rel + cp] = a.p.typeId[gp]; sm.typ[rel + cq] = a.p.typeId[gq];
                                        totColl += 1;
                                    }
                                }
                            }
                        }
                        if (relax) {
                            double cR = -1;
                            if (LB) {
                                const double mR = P.mR[tP][tQ];
                                const double cRsqr = magSqr(UP - UQ);
                                double translationalEnergy = 0.5 * mR * cRsqr;
                                const double omegaPQ = P.omegaPQ[tP][tQ];
                                const double* tMacro = nullptr;
                                if constexpr (ZV2008) tMacro = a.overallT ? a.overallT + c : nullptr;
                                redistributeInPlace<ZV2008, NM>(P, rng, v, cp, tP, tQ, translationalEnergy, omegaPQ, tMacro);
                                redistributeInPlace<ZV2008, NM>(P, rng, v, cq, tQ, tP, translationalEnergy, omegaPQ, tMacro);
                                cR = sqrt(2.0 * translationalEnergy / mR);
                            }
                            postCollisionVelocities(P, rng, tP, tQ, UP, UQ, cR);
                            a.p.ux[gp] = UP.x; a.p.uy[gp] = UP.y; a.p.uz[gp] = UP.z;
                            a.p.ux[gq] = UQ.x; a.p.uy[gq] = UQ.y; a.p.uz[gq] = UQ.z;
                            // cellMeasurements (VariableHardSphere.C:154-162)
                            const double dx = a.p.px[gp] - a.p.px[gq], dy = a.p.py[gp] - a.p.py[gq], dz = a.p.pz[gp] - a.p.pz[gq];
                            cellSep += sqrt(dx * dx + dy * dy + dz * dz);
                            cellColl += 1.0;
                            // classification promotion (VariableHardSphere.C:164-187)
                            if (a.p.cls) {
                                const int clP = a.p.cls[gp], clQ = a.p.cls[gq];
                                if (clP == 0 && (clQ == 1 || clQ == 2)) a.p.cls[gp] = 2;
                                if (clQ == 0 && (clP == 1 || clP == 2)) a.p.cls[gq] = 2;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            if (inPass && n <= LANE_CELL_MAX) {
                if (n > 1) a.sigmaTcRMax[c] = cellMax;
                a.nCollsStep[c] = cellColl;
                a.collSepStep[c] = cellSep;
                totColl += (unsigned long long)cellColl;
                totCand += (unsigned long long)nCand;
            }
            __syncwarp();
            g0 = g1;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        totColl += __shfl_xor_sync(FULL, totColl, o);
        totCand += __shfl_xor_sync(FULL, totCand, o);
    }
    if (lane == 0 && (totColl | totCand)) {
        atomicAdd(&a.counters->collisions, totColl);
        atomicAdd(&a.counters->candidates, totCand);
    }
}

cudaError_t launchCollide(const CollideArgs& a, cudaStream_t s) {
    {
        const int nGroups = (a.nCells + 31) / 32;
        int grid = (nGroups + LANE_WARPS - 1) / LANE_WARPS;
        if (grid > 148 * 4) grid = 148 * 4;  // persistent: 4 resident blocks per SM, grid-stride over the cell groups
        if (grid < 1) grid = 1;
        // engine.cu passes overallT only for inverseZvFormulation "2008"
        if (a.born) collideLaneKernel<true, MAX_MODES, true><<<grid, LANE_WARPS * 32, 0, s>>>(a);   // with chemistry: the general instance
        else if (a.overallT) collideLaneKernel<true, MAX_MODES, false><<<grid, LANE_WARPS * 32, 0, s>>>(a);
        else if (a.nModes <= 0) collideLaneKernel<false, 0, false><<<grid, LANE_WARPS * 32, 0, s>>>(a);
        else if (a.nModes == 1) collideLaneKernel<false, 1, false><<<grid, LANE_WARPS * 32, 0, s>>>(a);
        else if (a.nModes == 2) collideLaneKernel<false, 2, false><<<grid, LANE_WARPS * 32, 0, s>>>(a);
        else collideLaneKernel<false, MAX_MODES, false><<<grid, LANE_WARPS * 32, 0, s>>>(a);
    }
    int gridBig = (a.nCells + COL_WARPS - 1) / COL_WARPS;
    if (gridBig > 148 * 4) gridBig = 148 * 4;
    if (gridBig < 1) gridBig = 1;
    collideBigCellsKernel<<<gridBig, COL_WARPS * 32, 0, s>>>(a);
    if (a.nGiant > 0) collideGiantCellsKernel<<<a.nGiant, GIANT_THREADS, 0, s>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Stage 5
// ------------------------------------------------------------------------------------------------
namespace {
constexpr int SMP_THREADS = 256;
#ifndef SMP_TILE_MAX
#define SMP_TILE_MAX 512       // parcels of one chunk
#endif
#ifndef SMP_TILE_BYTES
#define SMP_TILE_BYTES (40 * 1024)
#endif
#ifndef SMP_GROUP_PARCELS
#define SMP_GROUP_PARCELS 1024 // parcels of the consecutive cells one block takes at a time
#endif
#ifndef SMP_BLOCKS_PER_SM
#define SMP_BLOCKS_PER_SM 4    // persistent grid: blocks per SM
#endif

// sum[sp] += x with sp a run-time index and sum[] in registers: one predicated add per species
template <int I, int S>
__device__ __forceinline__ void addToSpecies(double (&sum)[S], int sp, double x) {
    if constexpr (I < S) {
        asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\t@p add.rn.f64 %0, %0, %3;\n\t}" : "+d"(sum[I]) : "r"(sp), "n"(I), "d"(x));
        addToSpecies<I + 1, S>(sum, sp, x);
    }
}
}

// A block takes a group of consecutive cells and walks their (contiguous, sorted) parcels in chunks.
// Phase 1, thread = parcel: coalesced loads, the parcel's row of moment contributions goes to shared memory
// (row stride nQ|1 doubles: conflict free both ways).  Phase 2, thread = (sub-range, cell, quantity): adds the rows of
// its cell that lie in the chunk, in parcel order, into its own accumulator selected by the row's species
// (accS[species][thread], conflict free).  After the last chunk the sub-ranges are folded in a fixed order and thread
// (cell, quantity) adds sum[s] to the cell's accumulator row with one RED per non-zero element: a single writer per
// element per step, so the sums are run-to-run deterministic.
template <int S>
__global__ void __launch_bounds__(SMP_THREADS) sampleKernel(const __grid_constant__ SampleArgs a, const int cellsPerGroup, const int nSub,
                                                            const int tile) {
    extern __shared__ double smemD[];
    const DevParams& P = *a.P;
    const int nQ = a.nQ;
    const int stride = nQ | 1;
    double* __restrict__ val = smemD;                                                 // [tile][stride]
    double* __restrict__ accS = val + size_t(tile) * stride;                          // [S][SMP_THREADS]
    int32_t* __restrict__ offS = reinterpret_cast<int32_t*>(accS + S * SMP_THREADS);  // [cellsPerGroup + 1]
    uint8_t* __restrict__ spS = reinterpret_cast<uint8_t*>(offS + cellsPerGroup + 1); // [tile]
    const bool internal = P.hasInternalEnergy != 0;
    const int qFlux = 5 + (internal ? 2 + P.nModes : 0);
    const int qClass = qFlux + (P.measureFlux ? 12 : 0);
    const int tid = threadIdx.x;
    // phase-2 role
    const int perSub = cellsPerGroup * nQ;
    const int sub = tid / perSub;
    const int rem = tid - sub * perSub;
    const int cl = rem / nQ, q = rem - cl * nQ;
    const int32_t nGroups = (a.nCells + cellsPerGroup - 1) / cellsPerGroup;

    for (int32_t grp = blockIdx.x; grp < nGroups; grp += gridDim.x) {
        const int32_t c0 = grp * cellsPerGroup;
        const int nc = a.nCells - c0 < cellsPerGroup ? a.nCells - c0 : cellsPerGroup;
        __syncthreads();  // the previous group is done with the shared arrays
        for (int k = tid; k <= nc; k += SMP_THREADS) offS[k] = a.cellOffset[c0 + k];
        for (int k = tid; k < 2 * nc; k += SMP_THREADS) {
            // fold this step's cellMeasurements into the cumulative pair (dsmcVolFields.C:1244-1248)
            const int32_t c = c0 + (k >> 1);
            const double add = (k & 1) ? a.collSepStep[c] : a.nCollsStep[c];
            if (add != 0.0) a.collCum[2 * size_t(c0) + k] += add;
        }
        __syncthreads();
        double sum[S];  // my accumulator per species (registers: the species of a row selects by predicate, not by address)
#pragma unroll
        for (int s = 0; s < S; ++s) sum[s] = 0.0;
        const int32_t pBeg = offS[0], pEnd = offS[nc];
        const bool role = sub < nSub && cl < nc;
        const int32_t myLo = role ? offS[cl] : 0, myHi = role ? offS[cl + 1] : 0;

        for (int32_t pos = pBeg; pos < pEnd; pos += tile) {
            const int n = pEnd - pos < tile ? pEnd - pos : tile;
            // ---- phase 1: one parcel per thread
            for (int j = tid; j < n; j += SMP_THREADS) {
                const int32_t g = pos + j;
                const int mySp = a.p.typeId[g];
                const double ux = a.p.ux[g], uy = a.p.uy[g], uz = a.p.uz[g];
                double* row = val + j * stride;
                const double cc = ux * ux + uy * uy + uz * uz;
                row[0] = 1.0; row[1] = ux; row[2] = uy; row[3] = uz; row[4] = cc;
                spS[j] = uint8_t(mySp);
                double Eint = 0.0;
                if (internal) {
                    const DevSpecies& Sp = P.sp[mySp];
                    const double er = a.p.erot[g];
                    row[5] = er;
                    row[6] = Sp.eElec[a.p.elevel[g]];
                    Eint = er;
#pragma unroll
                    for (int m = 0; m < MAX_MODES; ++m) {
                        if (m < P.nModes) {
                            const double ev = (m < Sp.nVib) ? a.p.vib[m][g] * P.kB * Sp.thetaV[m] : 0.0;
                            row[7 + m] = ev;
                            Eint += ev;
                        }
                    }
                }
                if (P.measureFlux) {
                    double* f = row + qFlux;
                    f[0] = ux * ux; f[1] = ux * uy; f[2] = ux * uz; f[3] = uy * uy; f[4] = uy * uz; f[5] = uz * uz;
                    f[6] = cc * ux; f[7] = cc * uy; f[8] = cc * uz;
                    f[9] = Eint * ux; f[10] = Eint * uy; f[11] = Eint * uz;
                }
                if (P.measureClass) {
                    const int cls = a.p.cls ? a.p.cls[g] : 0;
                    row[qClass] = cls == 0 ? 1.0 : 0.0; row[qClass + 1] = cls == 1 ? 1.0 : 0.0; row[qClass + 2] = cls == 2 ? 1.0 : 0.0;
                }
            }
            __syncthreads();
            // ---- phase 2: the rows of my cell inside this chunk, every nSub-th one starting at sub
            if (role) {
                int lo = (myLo > pos ? myLo : pos), hi = (myHi < pos + n ? myHi : pos + n);
                if (nSub > 1) lo += ((sub - (lo - myLo)) % nSub + nSub) % nSub;  // first index >= lo with (index - myLo) = sub (mod nSub)
                const double* v = val + q + (lo - pos) * stride;
                const uint8_t* ps = spS + (lo - pos);
                const int vStep = nSub * stride;
                for (int32_t g = lo; g < hi; g += nSub, v += vStep, ps += nSub) {
                    const int sp = *ps;
                    const double x = *v;
                    // one predicated add per species (a select chain would cost three instructions per species)
                    addToSpecies<0, S>(sum, sp, x);
                }
            }
            __syncthreads();
        }
        // ---- fold the sub-ranges in a fixed order; one writer per accumulator element
        if (nSub > 1) {
#pragma unroll
            for (int s = 0; s < S; ++s) accS[s * SMP_THREADS + tid] = sum[s];
            __syncthreads();
            if (role && sub == 0) {
#pragma unroll
                for (int s = 0; s < S; ++s)
                    for (int k = 1; k < nSub; ++k) sum[s] += accS[s * SMP_THREADS + k * perSub + rem];
            }
        }
        if (role && sub == 0) {
            double* row = a.acc + (size_t(c0 + cl) * S) * nQ + q;
#pragma unroll
            for (int s = 0; s < S; ++s)
                if (sum[s] != 0.0) atomicAdd(&row[size_t(s) * nQ], sum[s]);  // fire-and-forget RED
        }
    }
}

// The parcels this step's dissociations created are part of the cloud (dsmcVolFields loops over the cloud, dsmcVolFields.C:1115) but in
// no cell list until the next buildCellOccupancy: one thread per such parcel adds its row.
__global__ void sampleTailKernel(const __grid_constant__ SampleArgs a) {
    const int32_t g = a.nParcels + blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.nCloud) return;
    const DevParams& P = *a.P;
    const int32_t cell = a.p.cell[g];
    if (cell < 0) return;
    const bool internal = P.hasInternalEnergy != 0;
    const int qFlux = 5 + (internal ? 2 + P.nModes : 0);
    const int qClass = qFlux + (P.measureFlux ? 12 : 0);
    const int mySp = a.p.typeId[g];
    double* row = a.acc + (size_t(cell) * a.nSpecies + mySp) * a.nQ;
    const double ux = a.p.ux[g], uy = a.p.uy[g], uz = a.p.uz[g];
    const double cc = ux * ux + uy * uy + uz * uz;
    atomicAdd(row + 0, 1.0); atomicAdd(row + 1, ux); atomicAdd(row + 2, uy); atomicAdd(row + 3, uz); atomicAdd(row + 4, cc);
    double Eint = 0.0;
    if (internal) {
        const DevSpecies& Sp = P.sp[mySp];
        const double er = a.p.erot[g];
        atomicAdd(row + 5, er);
        atomicAdd(row + 6, Sp.eElec[a.p.elevel[g]]);
        Eint = er;
#pragma unroll
        for (int m = 0; m < MAX_MODES; ++m) {
            if (m < P.nModes) {
                const double ev = (m < Sp.nVib) ? a.p.vib[m][g] * P.kB * Sp.thetaV[m] : 0.0;
                atomicAdd(row + 7 + m, ev);
                Eint += ev;
            }
        }
    }
    if (P.measureFlux) {
        double* f = row + qFlux;
        atomicAdd(f + 0, ux * ux); atomicAdd(f + 1, ux * uy); atomicAdd(f + 2, ux * uz); atomicAdd(f + 3, uy * uy); atomicAdd(f + 4, uy * uz);
        atomicAdd(f + 5, uz * uz); atomicAdd(f + 6, cc * ux); atomicAdd(f + 7, cc * uy); atomicAdd(f + 8, cc * uz);
        atomicAdd(f + 9, Eint * ux); atomicAdd(f + 10, Eint * uy); atomicAdd(f + 11, Eint * uz);
    }
    if (P.measureClass) atomicAdd(row + qClass + (a.p.cls ? a.p.cls[g] : 0), 1.0);
}

cudaError_t launchSample(const SampleArgs& a, cudaStream_t s) {
    // group size: about 1000 parcels of consecutive cells, at most SMP_THREADS / nQ cells (one phase-2 thread per (cell, quantity));
    // spare threads split each cell's parcels into nSub interleaved sub-ranges (few, crowded cells)
    const int maxCells = SMP_THREADS / a.nQ > 0 ? SMP_THREADS / a.nQ : 1;
    const double ppc = a.nCells > 0 ? double(a.nParcels) / a.nCells : 1.0;
    int cpg = int(double(SMP_GROUP_PARCELS) / (ppc > 1.0 ? ppc : 1.0));
    if (cpg > maxCells) cpg = maxCells;
    if (cpg < 1) cpg = 1;
    int nSub = SMP_THREADS / (cpg * a.nQ);
    if (nSub < 1) nSub = 1;
    const int stride = a.nQ | 1;
    int tile = (SMP_TILE_BYTES / (stride * 8)) & ~31;
    if (tile > SMP_TILE_MAX) tile = SMP_TILE_MAX;
    if (tile < 32) tile = 32;
    const size_t smem = (size_t(tile) * stride + size_t(a.nSpecies) * SMP_THREADS) * sizeof(double) + size_t(cpg + 1) * 4 + tile;
    const int nGroups = (a.nCells + cpg - 1) / cpg;
    int grid = nGroups < 148 * SMP_BLOCKS_PER_SM ? nGroups : 148 * SMP_BLOCKS_PER_SM;
    if (grid < 1) grid = 1;
    auto go = [&](auto kernel) -> cudaError_t {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        kernel<<<grid, SMP_THREADS, smem, s>>>(a, cpg, nSub, tile);
        if (a.nCloud > a.nParcels) sampleTailKernel<<<(a.nCloud - a.nParcels + 127) / 128, 128, 0, s>>>(a);
        return cudaGetLastError();
    };
    switch (a.nSpecies) {
        case 1: return go(sampleKernel<1>);
        case 2: return go(sampleKernel<2>);
        case 3: return go(sampleKernel<3>);
        case 4: return go(sampleKernel<4>);
        case 5: return go(sampleKernel<5>);
        case 6: return go(sampleKernel<6>);
        case 7: return go(sampleKernel<7>);
        default: return go(sampleKernel<8>);
    }
}

// ------------------------------------------------------------------------------------------------
// dsmcCloud::addNewParcel for the second products of this step's dissociations: the kernels stage them in atomic order; they join the
// cloud in the order the reference's serial candidate loop creates them (cell, then candidate), so a run is reproducible.
// ------------------------------------------------------------------------------------------------
__global__ void bornKeysKernel(const BornRec* born, int32_t n, unsigned long long* keys, int32_t* idx) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = born[i].key; idx[i] = i; }
}
__global__ void appendBornKernel(const __grid_constant__ ParcelArrays p, const BornRec* born, const int32_t* order, int32_t n, int32_t base,
                                 int32_t origIdBase, int32_t origProc, int32_t nModes) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const BornRec b = born[order[r]];
    const int32_t g = base + r;
    p.px[g] = b.pos[0]; p.py[g] = b.pos[1]; p.pz[g] = b.pos[2];
    p.ux[g] = b.U[0]; p.uy[g] = b.U[1]; p.uz[g] = b.U[2];
    if (p.erot) p.erot[g] = 0.0;
    p.cell[g] = b.cell; p.tet[g] = b.tet;
    p.origId[g] = int32_t((uint32_t(origIdBase) + uint32_t(r)) & 0x7fffffffu);
    for (int m = 0; m < MAX_MODES; ++m) if (m < nModes && p.vib[m]) p.vib[m][g] = 0;
    p.typeId[g] = b.typeId;
    if (p.elevel) p.elevel[g] = 0;
    if (p.cls) p.cls[g] = b.cls;
    if (p.origProc) p.origProc[g] = uint8_t(origProc);
    if (p.rwf) p.rwf[g] = b.rwf;
}
size_t orderBornTempBytes(int32_t capacity) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const int32_t*)nullptr,
                                    (int32_t*)nullptr, capacity);
    return bytes;
}
cudaError_t launchAppendBorn(const ParcelArrays& p, const BornRec* born, int32_t n, int32_t base, int32_t origIdBase, int32_t origProc, int32_t nModes,
                             unsigned long long* keyWork, int32_t* idxWork, void* temp, size_t tempBytes, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    bornKeysKernel<<<(n + 255) / 256, 256, 0, s>>>(born, n, keyWork, idxWork);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, tempBytes, keyWork, keyWork + n, idxWork, idxWork + n, n, 0, 64, s);
    if (e != cudaSuccess) return e;
    appendBornKernel<<<(n + 255) / 256, 256, 0, s>>>(p, born, idxWork + n, n, base, origIdBase, origProc, nModes);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// dsmcCloud::info / infoMeasurements (DSMC/clouds/dsmcCloud.C:935-985, dsmcCloudI.H:268-297):
// mass, linear KE, rotational, vibrational, electronic energy of the whole cloud (real molecules).
// ------------------------------------------------------------------------------------------------
namespace { constexpr int INFO_BLOCKS = 296; constexpr int INFO_THREADS = 256; }

__global__ void __launch_bounds__(INFO_THREADS) infoKernel(const __grid_constant__ ParcelArrays p, const CellFields cf, int32_t n, const DevParams* Pp, double* scratch) {
    __shared__ double red[6][INFO_THREADS / 32];
    const DevParams& P = *Pp;
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int32_t cell = p.cell[i];
        if (cell < 0) continue;
        const DevSpecies& S = P.sp[p.typeId[i]];
        const double ux = p.ux[i], uy = p.uy[i], uz = p.uz[i];
        const double w = cf.nParticles(P.nParticles, cell);   // this->nParticles(p.cell()), dsmcCloudI.H:278
        v[5] += w;   // infoMeasurements[6]: free molecules
        v[0] += S.mass * w;
        v[1] += 0.5 * S.mass * (ux * ux + uy * uy + uz * uz) * w;
        if (P.hasInternalEnergy) {
            v[2] += p.erot[i] * w;
            double ev = 0.0;
#pragma unroll
            for (int m = 0; m < MAX_MODES; ++m) if (m < S.nVib) ev += p.vib[m][i] * P.kB * S.thetaV[m];
            v[3] += ev * w;
            v[4] += S.eElec[p.elevel[i]] * w;
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) red[k][w] = x;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double x = 0;
        for (int k = 0; k < INFO_THREADS / 32; ++k) x += red[threadIdx.x][k];
        scratch[blockIdx.x * 6 + threadIdx.x] = x;
    }
}

__global__ void infoFinalKernel(const double* scratch, const DevParams* Pp, double* out5) {
    if (threadIdx.x < 6) {
        double x = 0;
        for (int b = 0; b < INFO_BLOCKS; ++b) x += scratch[b * 6 + threadIdx.x];
        out5[threadIdx.x] = x;
    }
}

int32_t infoScratchDoubles() { return INFO_BLOCKS * 6; }

cudaError_t launchInfo(const ParcelArrays& p, const CellFields& cf, int32_t n, const DevParams* P, double* out5, double* scratch, cudaStream_t s) {
    infoKernel<<<INFO_BLOCKS, INFO_THREADS, 0, s>>>(p, cf, n, P, scratch);
    infoFinalKernel<<<1, 32, 0, s>>>(scratch, P, out5);
    return cudaGetLastError();
}

}  // namespace dsmc
