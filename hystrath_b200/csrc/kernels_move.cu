// kernels_move.cu -- stage 1: free flight with face-crossing tracking on the tet decomposition
// of the polyMesh, boundary interactions, and (fused) the cell histogram of stage 2.
//
// Follows, per parcel, Cloud<T>::move (BASIC/Cloud/Cloud.C:204-312), dsmcParcel::move
// (DSMC/parcels/dsmcParcel.C:38-148) and particle::trackToFace(end, td, DSMC=true)
// (BASIC/particle/particleTemplates.C:727-1241) with findTris/tetLambda
// (BASIC/particle/particleI.H:31-140).  The reference walks mesh topology on every tet hop
// (tetNeighbour / crossEdgeConnectedFace, particleI.H:339-601) and recomputes the four
// normalised face-area vectors; here both are table look-ups in a 192-byte TetRec baked by
// host_mesh.cpp with the same arithmetic, so the FP64 comparisons see identical operands.
// Compiled with --fmad=false: x86 gcc -O3 without -march does not contract to FMA either.
#include "device_models.cuh"
#include "engine.h"

namespace dsmc {

namespace {

constexpr double kTrackingCorrectionTol = 1.0e-5;  // BASIC/particle/particle.C:33

struct Tet {
    double d[24];
    __device__ __forceinline__ V3 n(int i) const { return mk(d[3 * i], d[3 * i + 1], d[3 * i + 2]); }
    __device__ __forceinline__ V3 base() const { return mk(d[12], d[13], d[14]); }
    __device__ __forceinline__ V3 pA() const { return mk(d[15], d[16], d[17]); }
    __device__ __forceinline__ V3 ct() const { return mk(d[18], d[19], d[20]); }
    __device__ __forceinline__ double tol() const { return d[21]; }
    __device__ __forceinline__ int32_t nbr(int i) const {
        long long w = __double_as_longlong(d[22 + (i >> 1)]);
        return (i & 1) ? int32_t(w >> 32) : int32_t(w & 0xffffffffLL);
    }
};

__device__ __forceinline__ void loadTet(const TetRec* __restrict__ tets, int32_t id, Tet& t) {
    const double2* s = reinterpret_cast<const double2*>(tets + id);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        double2 v = __ldg(s + i);
        t.d[2 * i] = v.x;
        t.d[2 * i + 1] = v.y;
    }
}

// particle::tetLambda, BASIC/particle/particleI.H:68-140 (static mesh branch)
__device__ __forceinline__ double tetLambda(const V3& from, const V3& to, const V3& n, const V3& base, double tol) {
    double lambdaNumerator = dot(base - from, n);
    double lambdaDenominator = dot(to - from, n);
    if (fabs(lambdaDenominator) < tol) {
        if (fabs(lambdaNumerator) < tol) return 0.0;
        if (mag(to - from) < tol / mag(n)) return GREAT;
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

// dsmcPatchBoundary::measurePropertiesBeforeControl / AfterControl accumulation,
// DSMC/boundaries/basic/dsmcPatchBoundary/dsmcPatchBoundary.C:263-356,358-482
__device__ void wallMeasure(const WallCtx& w, int32_t measIndex, int32_t bfi, int sp, const V3& U, double ERot,
                            const int32_t* vib, int elevel, double& IE, V3& IMom) {
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
    double EVib = 0.0;
    for (int mo = 0; mo < S.nVib; ++mo) EVib += vib[mo] * P.kB * S.thetaV[mo];
    const double EEle = S.eElec[elevel];
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
        atomicAdd(a + WQ_EROT, ERot * rwf);
        atomicAdd(a + WQ_ZETAROT, S.rotDof * rwf);
        atomicAdd(a + WQ_EVIB, EVib * rwf);
        for (int mo = 0; mo < S.nVib; ++mo) atomicAdd(a + WQ_EVIBMOD0 + mo, vib[mo] * P.kB * S.thetaV[mo] * rwf);
        atomicAdd(a + WQ_EELEC, EEle * rwf);
    }
    IE = 0.5 * m * UU + ERot + EVib + EEle;
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

}  // namespace

__global__ void __launch_bounds__(256) moveKernel(MoveArgs a) {
    const int32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= a.count) return;
    const int32_t i = a.first + li;
    const DevParams& P = *a.P;

    int32_t cell = a.p.cell[i];
    if (cell < 0) return;
    int32_t tet = a.p.tet[i];
    V3 pos = mk(a.p.px[i], a.p.py[i], a.p.pz[i]);
    V3 U = mk(a.p.ux[i], a.p.uy[i], a.p.uz[i]);
    double stepFraction = (a.sfTail != nullptr && i >= a.tailStart) ? a.sfTail[i - a.tailStart] : 0.0;
    const double deltaT = P.deltaT;

    // internal state is only touched by wall models
    const int sp = a.p.typeId[i];
    bool internalDirty = false, Udirty = false;
    double ERot = 0.0;
    int32_t vib[MAX_MODES] = {0, 0, 0};
    int elevel = 0;
    bool internalLoaded = false;
    Rng wallRng;
    bool wallRngInit = false;
    WallCtx wctx{a.P, a.wallAcc, a.nWallQ, a.bfaceArea};

    bool keepParticle = true, switchProcessor = false;
    int32_t procBfi = -1;
    unsigned rescues = 0;

    double tEnd = (1.0 - stepFraction) * deltaT;
    Tet T;
    int guard = 0;

    while (keepParticle && !switchProcessor && tEnd > ROOTVSMALL) {
        if (++guard > 100000) break;  // cannot happen on a valid mesh; keeps a corrupt one from hanging the GPU
        V3 Utracking = U;
        // meshTools::constrainToMeshCentre / constrainDirection (DSMC/parcels/dsmcParcel.C:76-85)
#pragma unroll
        for (int d = 0; d < 3; ++d)
            if (P.solutionD[d] == -1) { setComp(pos, d, P.centre[d]); setComp(Utracking, d, 0.0); }

        double dt = tEnd;
        const V3 endPosition = pos + dt * Utracking;

        // ---------------- particle::trackToFace ----------------
        double trackFraction = 0.0;
        int triI = -1;
        double lambdaMin = VGREAT;
        bool faceSet = false;       // faceI_ >= 0
        int32_t faceBfi = -1;       // boundary-face index of faceI_ when it is a boundary face
        bool returned = false;
        double retVal = 0.0;
        Tet cur;                    // tet on which the face was hit
        do {
            if (++guard > 100000) {  // lost in a corrupt tet table: drop the parcel rather than hang
                keepParticle = false; returned = true; retVal = 1.0;
                break;
            }
            if (triI != -1) tet = T.nbr(triI);  // particle::tetNeighbour (triI in 1..3 here)
            loadTet(a.tets, tet, T);
            if (lambdaMin < SMALL) {
                // tracking correction towards the tet centre
                pos += kTrackingCorrectionTol * (T.ct() - pos);
                ++rescues;
                returned = true; retVal = trackFraction;
                break;
            }
            const double tol = T.tol();
            // findTris: which planes does the ray tetCentre -> end cross
            const V3 Ct = T.ct();
            unsigned tris = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const V3 base = (k == 1) ? T.pA() : T.base();
                const double lambda = tetLambda(Ct, endPosition, T.n(k), base, tol);
                if (lambda > 0.0 && lambda < 1.0) tris |= 1u << k;
            }
            triI = -1;
            lambdaMin = VGREAT;
            if (tris == 0) {  // (faceI_ < 0 always holds here: hitWallFaces is inactive for DSMC)
                pos = endPosition;
                faceSet = false; faceBfi = -1;
                returned = true; retVal = 1.0;
                break;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (tris & (1u << k)) {
                    const V3 base = (k == 1) ? T.pA() : T.base();
                    const double lam = tetLambda(pos, endPosition, T.n(k), base, tol);
                    if (lam < lambdaMin) { lambdaMin = lam; triI = k; }
                }
            }
            if (triI == 0) {
                faceSet = true;
                faceBfi = T.nbr(0) < 0 ? (-1 - T.nbr(0)) : -1;
            } else if (triI > 0) {
                faceSet = false; faceBfi = -1;
            }
            if (lambdaMin > SMALL) {
                if (lambdaMin <= 1.0) {
                    trackFraction += lambdaMin * (1 - trackFraction);
                    pos += lambdaMin * (endPosition - pos);
                } else {
                    pos = endPosition;
                    returned = true; retVal = 1.0;
                    break;
                }
            } else {
                lambdaMin = 0.0;
            }
        } while (!faceSet);

        if (!returned) {
            // a cell face has been hit on tri 0 of tet T
            const int32_t nb0 = T.nbr(0);
            if (nb0 >= 0) {
                cell = nb0;   // internal face: the same face-triangle seen from the other cell
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
                        const V3 nf = T.n(0);
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
                    case DSMCB200_PATCH_PATCH: {
                        // dsmcParcel::hitWallPatch / hitPatch -> patch model controlParticle
                        if (pt.model == DSMCB200_BND_DELETION) {
                            keepParticle = false;  // dsmcDeletionPatch::controlParticle
                        } else if (pt.model == DSMCB200_BND_SPECULAR_WALL || pt.model == DSMCB200_BND_DIFFUSE_WALL) {
                            if (!internalLoaded) {
                                if (P.hasInternalEnergy) {
                                    ERot = a.p.erot[i];
                                    for (int mo = 0; mo < P.nModes; ++mo) vib[mo] = a.p.vib[mo][i];
                                    elevel = a.p.elevel[i];
                                }
                                internalLoaded = true;
                            }
                            double preIE, postIE;
                            V3 preIMom, postIMom;
                            wallMeasure(wctx, bf.measIndex, bfi, sp, U, ERot, vib, elevel, preIE, preIMom);
                            const V3 nw = T.n(0);
                            if (pt.model == DSMCB200_BND_SPECULAR_WALL) {
                                // dsmcSpecularWallPatch::performSpecularReflection
                                const double U_dot_nw = dot(U, nw);
                                if (U_dot_nw > 0.0) U -= 2.0 * U_dot_nw * nw;
                            } else {
                                // dsmcDiffuseWallPatch::performDiffuseReflection
                                if (!wallRngInit) {
                                    wallRng.init(P.seed, uint32_t(a.p.origId[i]), 0u, a.step, STREAM_WALL);
                                    wallRngInit = true;
                                }
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
                                const double Tw = pt.T;
                                const double g1 = wallRng.gaussNormal();
                                const double g2 = wallRng.gaussNormal();
                                const double r = wallRng.sample01();
                                U = sqrt(P.kB * Tw / S.mass) * (g1 * tw1 + g2 * tw2 - sqrt(-2.0 * log(fmax(1 - r, VSMALL))) * nw);
                                ERot = equipartitionRotationalEnergy(wallRng, P.kB, Tw, S.rotDof);
                                for (int mo = 0; mo < S.nVib; ++mo) vib[mo] = equipartitionVibrationalEnergyLevel(wallRng, Tw, S.thetaV[mo]);
                                elevel = equipartitionElectronicLevel(wallRng, P.kB, Tw, S);
                                U += mk(pt.vel[0], pt.vel[1], pt.vel[2]);
                                internalDirty = true;
                            }
                            Udirty = true;
                            wallMeasure(wctx, bf.measIndex, bfi, sp, U, ERot, vib, elevel, postIE, postIMom);
                            wallMeasureDelta(wctx, bf.measIndex, bfi, sp, preIE, preIMom, postIE, postIMom);
                        }
                        break;
                    }
                    default:  // empty patches cannot be hit by constrained tracks
                        break;
                }
            }
            if (lambdaMin < SMALL) {
                // tracking correction towards the centre of the tet now occupied
                Tet C;
                loadTet(a.tets, tet, C);
                pos += kTrackingCorrectionTol * (C.ct() - pos);
                ++rescues;
            }
            retVal = trackFraction;
        }
        // ---------------- back in dsmcParcel::move ----------------
        dt *= retVal;
        tEnd -= dt;
        stepFraction = 1.0 - tEnd / deltaT;
        if (faceSet && faceBfi >= 0 && keepParticle) {
            const int ptype = P.patch[a.bfaces[faceBfi].patch].type;
            if (ptype == DSMCB200_PATCH_PROCESSOR || ptype == DSMCB200_PATCH_PROCESSORCYCLIC) {
                switchProcessor = true;
                procBfi = faceBfi;
            }
        }
    }

    if (rescues) atomicAdd(&a.counters->rescues, (unsigned long long)rescues);

    if (!keepParticle) {
        a.p.cell[i] = -1;
        atomicAdd(&a.counters->deleted, 1ULL);
        return;
    }
    if (switchProcessor) {
        // Cloud<T>::move transfer list + particle::prepareForParallelTransfer, fused with the packing
        const BFaceRec bf = a.bfaces[procBfi];
        const DevPatch& pt = P.patch[bf.patch];
        const int slot = pt.nbrSlot;
        const int32_t k = atomicAdd(&a.counters->nMig[slot], 1);
        if (k < a.migCapacity) {
            MigRec r;
            r.pos[0] = pos.x; r.pos[1] = pos.y; r.pos[2] = pos.z;
            r.U[0] = U.x; r.U[1] = U.y; r.U[2] = U.z;
            if (!internalLoaded && P.hasInternalEnergy) {
                ERot = a.p.erot[i];
                for (int mo = 0; mo < P.nModes; ++mo) vib[mo] = a.p.vib[mo][i];
                elevel = a.p.elevel[i];
            }
            r.erot = ERot; r.stepFraction = stepFraction;
            r.patchOrdinal = pt.nbrOrdinal;
            r.patchFace = procBfi - (pt.start - P.nInternalFaces);
            r.tetLocal = (tet >> 1) - bf.tetPair0;
            r.origId = a.p.origId[i];
            for (int mo = 0; mo < MAX_MODES; ++mo) r.vib[mo] = vib[mo];
            r.typeId = uint8_t(sp); r.elevel = uint8_t(elevel); r.cls = a.p.cls ? a.p.cls[i] : 0; r.pad_ = 0;
            a.migBuf[size_t(slot) * a.migCapacity + k] = r;
        } else {
            atomicAdd(&a.counters->overflow, 1ULL);
        }
        a.p.cell[i] = -1;
        atomicAdd(&a.counters->migratedOut, 1ULL);
        return;
    }

    a.p.px[i] = pos.x; a.p.py[i] = pos.y; a.p.pz[i] = pos.z;
    a.p.cell[i] = cell;
    a.p.tet[i] = tet;
    if (Udirty) { a.p.ux[i] = U.x; a.p.uy[i] = U.y; a.p.uz[i] = U.z; }
    if (internalDirty && P.hasInternalEnergy) {
        a.p.erot[i] = ERot;
        for (int mo = 0; mo < P.nModes; ++mo) a.p.vib[mo][i] = vib[mo];
        a.p.elevel[i] = uint8_t(elevel);
    }
    if (a.cellCount) atomicAdd(&a.cellCount[cell], 1);
}

cudaError_t launchMove(const MoveArgs& a, cudaStream_t s) {
    if (a.count <= 0) return cudaSuccess;
    const int block = 256;
    const int grid = (a.count + block - 1) / block;
    moveKernel<<<grid, block, 0, s>>>(a);
    return cudaGetLastError();
}

// ---- arrivals over a processor patch: particle::correctAfterParallelTransfer,
// BASIC/particle/particleTemplates.C:52-123
__global__ void unpackKernel(UnpackArgs a) {
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
        for (int mo = 0; mo < P.nModes; ++mo) a.p.vib[mo][i] = r.vib[mo];
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
