// kernels_init.cu -- parcel generation on the device: the dsmcMeshFill initialiser
// (DSMC/initialiseDsmcParcels/derived/dsmcMeshFill/dsmcMeshFill.C:70-240) used for synthetic loads,
// and dsmcFreeStreamInflowPatch::controlParcelsBeforeMove
// (DSMC/boundaries/derived/generalBoundaries/dsmcFreeStreamInflowPatch/dsmcFreeStreamInflowPatch.C:86-375).
// Both run in two passes around an exclusive scan: pass 0 decides how many parcels each
// (cell | face, species) inserts, pass 1 generates them into the slots the scan assigned.
#include "device_models.cuh"
#include "engine.h"

namespace dsmc {

namespace {

struct FaceView {
    const int32_t* pts;
    int n, base;
};

// tetIndices::tet for (cell, face, tetPt): Cc, basePt, pA, pB
__device__ __forceinline__ void tetPointsDev(const double* points, const FaceView& f, bool own, int tetPt, V3& b, V3& c, V3& d) {
    const int facePtI = (tetPt + f.base) % f.n;
    const int otherFacePtI = (facePtI + 1) % f.n;
    const int fPtAI = own ? facePtI : otherFacePtI;
    const int fPtBI = own ? otherFacePtI : facePtI;
    const int32_t lb = f.pts[f.base], la = f.pts[fPtAI], lbb = f.pts[fPtBI];
    b = mk(points[3 * lb], points[3 * lb + 1], points[3 * lb + 2]);
    c = mk(points[3 * la], points[3 * la + 1], points[3 * la + 2]);
    d = mk(points[3 * lbb], points[3 * lbb + 1], points[3 * lbb + 2]);
}

__device__ __forceinline__ void storeParcel(const ParcelArrays& p, const DevParams& P, int32_t slot, const V3& pos, const V3& U, double ERot,
                                            const int32_t* vib, int elevel, int32_t cell, int32_t tet, int typeId, int32_t origId, int origProc, double RWF) {
    p.px[slot] = pos.x; p.py[slot] = pos.y; p.pz[slot] = pos.z;
    p.ux[slot] = U.x; p.uy[slot] = U.y; p.uz[slot] = U.z;
    p.cell[slot] = cell; p.tet[slot] = tet; p.origId[slot] = origId;
    if (p.origProc) p.origProc[slot] = uint8_t(origProc);
    p.typeId[slot] = uint8_t(typeId);
    if (P.hasInternalEnergy) {
        p.erot[slot] = ERot;
        for (int m = 0; m < P.nModes; ++m) p.vib[m][slot] = vib[m];
        p.elevel[slot] = uint8_t(elevel);
    }
    if (p.cls) p.cls[slot] = 0;
    if (p.rwf) p.rwf[slot] = RWF;
}

}  // namespace

// pass 0: cellCount[cell] = parcels to insert; pass 1: cellCount holds exclusive offsets
__global__ void __launch_bounds__(128) fillKernel(const __grid_constant__ FillArgs a, int pass) {
    // entry t of the fill: cell t of the mesh (dsmcMeshFill) or the t-th cell of the zone (dsmcZoneFill, cellList)
    const int32_t entry = blockIdx.x * blockDim.x + threadIdx.x;
    if (entry >= a.nFill) return;
    const int32_t cell = a.cellList ? a.cellList[entry] : entry;
    const DevParams& P = *a.P;
    const V3 Cc = mk(a.cellCentres[3 * cell], a.cellCentres[3 * cell + 1], a.cellCentres[3 * cell + 2]);
    int32_t count = 0;
    int32_t slot = pass == 1 ? a.slotBase + a.cellCount[entry] : 0;
    int tetLocal = 0;
    for (int k = a.cellFaceOffsets[cell]; k < a.cellFaceOffsets[cell + 1]; ++k) {
        const int32_t face = a.cellFaces[k];
        FaceView f{a.facePoints + a.faceOffsets[face], a.faceOffsets[face + 1] - a.faceOffsets[face], a.tetBasePtIs[face]};
        const bool own = a.owner[face] == cell;
        for (int tetPt = 1; tetPt < f.n - 1; ++tetPt, ++tetLocal) {
            V3 b, c, d;
            tetPointsDev(a.points, f, own, tetPt, b, c, d);
            const double tetVolume = (1.0 / 6.0) * dot(cross(b - Cc, c - Cc), d - Cc);  // tetrahedron::mag
            for (int i = 0; i < a.nTypes; ++i) {
                const int typeId = a.typeIds[i];
                const DevSpecies& S = P.sp[typeId];
                Rng rng;
                rng.init(P.seed, uint32_t(cell), uint32_t(tetLocal * MAX_SPECIES + i), a.fillIndex, STREAM_FILL);
                const double particlesRequired = a.numberDensities[i] * tetVolume / a.cf.nParticles(P.nParticles, cell);   // cloud_.nParticles(cellI), dsmcMeshFill.C:146
                int32_t nParticlesToInsert = int32_t(particlesRequired);
                if ((particlesRequired - nParticlesToInsert) > rng.sample01()) nParticlesToInsert++;
                if (pass == 0) { count += nParticlesToInsert; continue; }
                const int32_t tet = a.cellTetStart[cell] + tetLocal;
                for (int32_t pI = 0; pI < nParticlesToInsert; ++pI) {
                    const V3 pos = tetRandomPoint(rng, Cc, b, c, d);
                    V3 U = equipartitionLinearVelocity(rng, P.kB, a.Ttra, S.mass);
                    const double ERot = equipartitionRotationalEnergy(rng, P.kB, a.Trot, S.rotDof);
                    int32_t vib[MAX_MODES] = {0, 0, 0};
                    for (int m = 0; m < S.nVib; ++m) vib[m] = equipartitionVibrationalEnergyLevel(rng, a.Tvib, S.thetaV[m]);
                    const int elevel = equipartitionElectronicLevel(rng, P.kB, a.Telec, S);
                    U += mk(a.velocity[0], a.velocity[1], a.velocity[2]);
                    storeParcel(a.p, P, slot, pos, U, ERot, vib, elevel, cell, tet, typeId, a.origIdBase + (slot - a.slotBase), a.origProc, a.cf.RWF(cell));
                    ++slot;
                }
            }
        }
    }
    if (pass == 0) a.cellCount[entry] = count;
}

cudaError_t launchFill(const FillArgs& a, int pass, cudaStream_t s) {
    if (a.nFill <= 0) return cudaSuccess;
    fillKernel<<<(a.nFill + 127) / 128, 128, 0, s>>>(a, pass);
    return cudaGetLastError();
}

// ---- cloud load without tet indices: particle::initCellFacePtOrDeleteLostParticle (BASIC/particle/particleI.H:851-996) ----
namespace {

// polyMesh::findTetFacePt: the first tet of the cell (cell faces in cells[] order, tetPt ascending) whose tetrahedron::inside(p)
// holds -- "inside unless definitively shown otherwise": ((p - pt) & n) > SMALL with n = S/(mag(S) + VSMALL) for the four faces
__device__ bool findTetFacePtDev(const LocateArgs& a, int32_t cell, const V3& p, int32_t& tet) {
    const V3 A = mk(a.cellCentres[3 * cell], a.cellCentres[3 * cell + 1], a.cellCentres[3 * cell + 2]);
    int32_t tetLocal = 0;
    for (int k = a.cellFaceOffsets[cell]; k < a.cellFaceOffsets[cell + 1]; ++k) {
        const int32_t face = a.cellFaces[k];
        FaceView f{a.facePoints + a.faceOffsets[face], a.faceOffsets[face + 1] - a.faceOffsets[face], a.tetBasePtIs[face]};
        const bool own = a.owner[face] == cell;
        for (int tetPt = 1; tetPt < f.n - 1; ++tetPt, ++tetLocal) {
            V3 b, c, d;
            tetPointsDev(a.points, f, own, tetPt, b, c, d);
            V3 nn = 0.5 * cross(c - b, d - b); nn /= (mag(nn) + VSMALL);          // Sa = triNormal(b, c, d)
            if (dot(p - b, nn) > SMALL) continue;
            nn = 0.5 * cross(d - A, c - A); nn /= (mag(nn) + VSMALL);              // Sb = triNormal(a, d, c)
            if (dot(p - c, nn) > SMALL) continue;
            nn = 0.5 * cross(b - A, d - A); nn /= (mag(nn) + VSMALL);              // Sc = triNormal(a, b, d)
            if (dot(p - b, nn) > SMALL) continue;
            nn = 0.5 * cross(c - A, b - A); nn /= (mag(nn) + VSMALL);              // Sd = triNormal(a, c, b)
            if (dot(p - b, nn) > SMALL) continue;
            tet = a.cellTetStart[cell] + tetLocal;
            return true;
        }
    }
    return false;
}

// polyMesh::pointInCellBB(p, cell, 0.1): bounding box of the cell's points inflated by 10 % of its extent (particleI.H:892-902)
__device__ bool pointInCellBBDev(const LocateArgs& a, int32_t cell, const V3& p, double inflationFraction) {
    V3 mn = mk(VGREAT, VGREAT, VGREAT), mx = mk(-VGREAT, -VGREAT, -VGREAT);
    for (int k = a.cellFaceOffsets[cell]; k < a.cellFaceOffsets[cell + 1]; ++k) {
        const int32_t face = a.cellFaces[k];
        for (int j = a.faceOffsets[face]; j < a.faceOffsets[face + 1]; ++j) {
            const int32_t l = a.facePoints[j];
            const V3 q = mk(a.points[3 * l], a.points[3 * l + 1], a.points[3 * l + 2]);
            mn = mk(fmin(mn.x, q.x), fmin(mn.y, q.y), fmin(mn.z, q.z));
            mx = mk(fmax(mx.x, q.x), fmax(mx.y, q.y), fmax(mx.z, q.z));
        }
    }
    if (inflationFraction > SMALL) {
        const V3 inflationVec = (mx - mn) * inflationFraction;
        mn = mn - inflationVec;
        mx = mx + inflationVec;
    }
    return p.x >= mn.x && p.x <= mx.x && p.y >= mn.y && p.y <= mx.y && p.z >= mn.z && p.z <= mx.z;
}

// The position is in the (slightly extended) bound-box of the cell but in none of its tets (written with too few digits next to a
// boundary face, another base-point decision): particleI.H:927-976 moves it towards the cell centre in steps of
// trackingCorrectionTol*(cC - position), at most 1/tol + 1 of them, until a tet of the cell claims it, and keeps the new position.  The
// first 256 steps are taken one by one as the reference takes them; the tets of a cell all have the cell centre as a vertex, so once the
// segment is inside it stays inside, and the first claiming step beyond 256 is found by bisection.
__device__ bool walkToCellCentre(const LocateArgs& a, int32_t i, int32_t cell, const V3& p, int32_t& tet) {
    if (!pointInCellBBDev(a, cell, p, 0.1)) return false;
    const V3 cc = mk(a.cellCentres[3 * cell], a.cellCentres[3 * cell + 1], a.cellCentres[3 * cell + 2]);
    const V3 step = 1.0e-5 * (cc - p);
    const int trap = 100001;
    V3 q = p;
    int it = 0;
    bool ok = false;
    while (!ok && it < 256) { q += step; ++it; ok = findTetFacePtDev(a, cell, q, tet); }
    if (!ok) {
        int32_t tHi = 0;
        if (findTetFacePtDev(a, cell, p + double(trap) * step, tHi)) {
            int lo = it, hi = trap;   // not inside at lo, inside at hi
            while (hi - lo > 1) {
                const int mid = lo + (hi - lo) / 2;
                int32_t tm = 0;
                if (findTetFacePtDev(a, cell, p + double(mid) * step, tm)) { hi = mid; tHi = tm; } else lo = mid;
            }
            q = p + double(hi) * step; tet = tHi; ok = true;
        }
    }
    if (ok) { a.px[i] = q.x; a.py[i] = q.y; a.pz[i] = q.z; }   // position_ = newPosition
    return ok;
}

}  // namespace

__global__ void __launch_bounds__(128) locateKernel(const __grid_constant__ LocateArgs a) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int32_t cell = a.cell[i];
    const V3 p = mk(a.px[i], a.py[i], a.pz[i]);
    int32_t tet = 0;
    bool ok = false;
    if (cell >= 0 && cell < a.nCells) {
        ok = findTetFacePtDev(a, cell, p, tet);
        if (!ok) {
            // mesh_.findCellFacePt(position, cell, tetFace, tetPt) (particleI.H:884-890): another cell may hold the point (a position written
            // with 10 digits next to a face, a label from a slightly different mesh).  The reference searches the whole mesh through its
            // octree; here the cells around the given one are tried, the face neighbours and then theirs, in cells[] face order
            int32_t found = -1;
            for (int k = a.cellFaceOffsets[cell]; k < a.cellFaceOffsets[cell + 1] && found < 0; ++k) {
                const int32_t f = a.cellFaces[k];
                if (f >= a.nInternalFaces) continue;
                const int32_t c1 = a.owner[f] == cell ? a.neighbour[f] : a.owner[f];
                if (findTetFacePtDev(a, c1, p, tet)) found = c1;
            }
            for (int k = a.cellFaceOffsets[cell]; k < a.cellFaceOffsets[cell + 1] && found < 0; ++k) {
                const int32_t f = a.cellFaces[k];
                if (f >= a.nInternalFaces) continue;
                const int32_t c1 = a.owner[f] == cell ? a.neighbour[f] : a.owner[f];
                for (int k2 = a.cellFaceOffsets[c1]; k2 < a.cellFaceOffsets[c1 + 1] && found < 0; ++k2) {
                    const int32_t f2 = a.cellFaces[k2];
                    if (f2 >= a.nInternalFaces) continue;
                    const int32_t c2 = a.owner[f2] == c1 ? a.neighbour[f2] : a.owner[f2];
                    if (c2 != cell && findTetFacePtDev(a, c2, p, tet)) found = c2;
                }
            }
            if (found >= 0) { ok = true; a.cell[i] = found; }
        }
        if (!ok && a.pending) {
            // mesh-wide search (locateGlobalKernel), then the walk: in the order of particleI.H:884-976
            a.pending[atomicAdd(a.nPending, 1)] = i;
            a.tet[i] = 0;
            return;
        }
        if (!ok) ok = walkToCellCentre(a, i, cell, p, tet);
    }
    if (!ok) {   // lost: deleted by the sort (cell -1), hyStrath's change at particleI.H:892-902
        a.cell[i] = -1;
        tet = 0;
        atomicAdd(a.lost, 1ULL);
    }
    a.tet[i] = tet;
}

// polyMesh::findCellFacePt for the parcels the cells around their label did not hold: the reference asks its cell octree for a cell that
// contains the point; here a block looks through all cells whose centre is within the mesh's largest centre-to-vertex distance and takes
// the lowest label whose tets claim the point.  No cell: back to the cell of the label for the bound-box test and the walk, else lost.
__global__ void __launch_bounds__(256) locateGlobalKernel(const __grid_constant__ LocateArgs a, int32_t nPending) {
    __shared__ int32_t sBest;
    const int32_t i = a.pending[blockIdx.x];
    const int32_t cell0 = a.cell[i];
    const V3 p = mk(a.px[i], a.py[i], a.pz[i]);
    if (threadIdx.x == 0) sBest = 0x7fffffff;
    __syncthreads();
    for (int32_t c = threadIdx.x; c < a.nCells; c += blockDim.x) {
        if (c == cell0) continue;
        const V3 d = p - mk(a.cellCentres[3 * c], a.cellCentres[3 * c + 1], a.cellCentres[3 * c + 2]);
        if (dot(d, d) > a.searchRadius2) continue;
        int32_t t;
        if (findTetFacePtDev(a, c, p, t)) atomicMin(&sBest, c);
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int32_t tet = 0;
    bool ok = false;
    if (sBest != 0x7fffffff) { ok = findTetFacePtDev(a, sBest, p, tet); if (ok) a.cell[i] = sBest; }
    if (!ok) ok = walkToCellCentre(a, i, cell0, p, tet);
    if (!ok) { a.cell[i] = -1; tet = 0; atomicAdd(a.lost, 1ULL); }
    a.tet[i] = tet;
}

cudaError_t launchLocateGlobal(const LocateArgs& a, int32_t nPending, cudaStream_t s) {
    if (nPending <= 0) return cudaSuccess;
    locateGlobalKernel<<<nPending, 256, 0, s>>>(a, nPending);
    return cudaGetLastError();
}

cudaError_t launchLocate(const LocateArgs& a, cudaStream_t s) {
    if (a.n <= 0) return cudaSuccess;
    locateKernel<<<(a.n + 127) / 128, 128, 0, s>>>(a);
    return cudaGetLastError();
}

// pass 0: accumulator update (Bird eq. 4.22) and the integer number to insert per (species, face)
// pass 1: counts holds exclusive offsets; generate
__global__ void __launch_bounds__(128) inflowKernel(const __grid_constant__ InflowArgs a, int pass) {
    const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nFaces * a.nTypes) return;
    const int32_t m = t / a.nFaces, fl = t % a.nFaces;  // counts layout [species][face]
    const DevParams& P = *a.P;
    const int typeId = a.typeIds[m];
    const DevSpecies& S = P.sp[typeId];
    const int32_t faceI = a.patchStart + fl;
    const V3 sF = mk(a.faceAreas[3 * faceI], a.faceAreas[3 * faceI + 1], a.faceAreas[3 * faceI + 2]);
    const double fA = mag(sF);
    const double mass = S.mass;
    const V3 velocity = mk(a.velocity[0], a.velocity[1], a.velocity[2]);
    const double mostProbableSpeed = sqrt(2.0 * P.kB * a.Ttra / mass);  // maxwellianMostProbableSpeed
    const double sqrtPi = sqrt(PI);
    Rng rng;
    rng.init(P.seed, uint32_t(faceI), uint32_t(a.patch * MAX_SPECIES + m), a.step, STREAM_INFLOW);

    if (pass == 0) {
        const double sCosTheta = dot(velocity, -sF / fA) / mostProbableSpeed;
        double acc = a.accumulator[t];
        const int32_t faceCell = a.owner[faceI];   // deltaTValue(faceCells()[f]), nParticles(patch, f): the cell's (dsmcFreeStreamInflowPatch.C:109-140)
        acc += (fA * a.numberDensities[m] * a.cf.deltaT(P.deltaT, faceCell) * mostProbableSpeed *
                (exp(-(sCosTheta * sCosTheta)) + sqrtPi * sCosTheta * (1 + erf(sCosTheta)))) /
               (2.0 * sqrtPi * a.cf.nParticles(P.nParticles, faceCell));
        int32_t nI = int32_t(acc) > 0 ? int32_t(acc) : 0;
        if ((acc - nI) > rng.sample01()) nI++;
        acc -= nI;
        a.accumulator[t] = acc;
        a.counts[t] = nI;
        return;
    }

    const int32_t nI = a.counts[t + 1] - a.counts[t];
    if (nI <= 0) return;
    rng.idx = 1;  // draw 0 was the insertion decision of pass 0
    int32_t slot = a.base + a.counts[t];
    if (slot + nI > a.capacity) { atomicAdd(&a.counters->overflow, 1ULL); return; }

    const int32_t cellI = a.owner[faceI];
    FaceView f{a.facePoints + a.faceOffsets[faceI], a.faceOffsets[faceI + 1] - a.faceOffsets[faceI], a.tetBasePtIs[faceI]};
    const V3 fC = mk(a.faceCentres[3 * faceI], a.faceCentres[3 * faceI + 1], a.faceCentres[3 * faceI + 2]);
    V3 n = sF;
    n /= -mag(n);
    const int32_t l0 = f.pts[0];
    V3 t1 = fC - mk(a.points[3 * l0], a.points[3 * l0 + 1], a.points[3 * l0 + 2]);
    t1 /= mag(t1);
    V3 t2 = cross(n, t1);
    t2 /= mag(t2);

    for (int32_t i = 0; i < nI; ++i, ++slot) {
        // triangle chosen by cumulative area fraction
        const double triSelection = rng.sample01();
        int selectedTriI = -1;
        double cum = 0.0;
        V3 b, c, d;
        for (int tetPt = 1; tetPt < f.n - 1; ++tetPt) {
            selectedTriI = tetPt;
            tetPointsDev(a.points, f, true, tetPt, b, c, d);
            cum = mag(0.5 * cross(c - b, d - b)) / fA + cum;
            const double frac = (tetPt == f.n - 2) ? 1.0 : cum;
            if (frac >= triSelection) break;
        }
        const V3 pos = triRandomPoint(rng, b, c, d);
        const double sCosTheta = dot(velocity, n) / mostProbableSpeed;
        const double uNormProbCoeffA = sCosTheta + sqrt(sCosTheta * sCosTheta + 2.0);
        const double uNormProbCoeffB = 0.5 * (1.0 + sCosTheta * (sCosTheta - sqrt(sCosTheta * sCosTheta + 2.0)));
        double randomScaling = 3.0;
        if (sCosTheta < -3) randomScaling = fabs(sCosTheta) + 1;
        double Pp = -1, uNormal, uNormalThermal;
        if (fabs(dot(velocity, n)) > VSMALL) {
            do {
                uNormalThermal = randomScaling * (2.0 * rng.sample01() - 1);
                uNormal = uNormalThermal + sCosTheta;
                if (uNormal < 0.0) Pp = -1;
                else Pp = 2.0 * uNormal / uNormProbCoeffA * exp(uNormProbCoeffB - uNormalThermal * uNormalThermal);
            } while (Pp < rng.sample01());
        } else {
            uNormal = sqrt(-log(rng.sample01()));
        }
        const double g1 = rng.gaussNormal(), g2 = rng.gaussNormal();
        const V3 U = sqrt(P.kB * a.Ttra / mass) * (g1 * t1 + g2 * t2) + dot(t1, velocity) * t1 + dot(t2, velocity) * t2 +
                     mostProbableSpeed * uNormal * n;
        const double ERot = equipartitionRotationalEnergy(rng, P.kB, a.Trot, S.rotDof);
        int32_t vib[MAX_MODES] = {0, 0, 0};
        for (int mo = 0; mo < S.nVib; ++mo) vib[mo] = equipartitionVibrationalEnergyLevel(rng, a.Tvib, S.thetaV[mo]);
        const int elevel = equipartitionElectronicLevel(rng, P.kB, a.Telec, S);
        const int32_t tet = a.bfaces[faceI - a.nInternalFaces].tet0 + selectedTriI - 1;
        storeParcel(a.p, P, slot, pos, U, ERot, vib, elevel, cellI, tet, typeId, int32_t((int64_t(a.origIdBase) + (slot - a.base)) & 0x7fffffff), a.origProc, a.cf.RWF(cellI));
        // dsmcParcel::move: a freshly inserted parcel moves a random fraction of the step (dsmcParcel.C:52-59)
        a.sfTail[slot - a.tailStart] = rng.sample01();
        if (a.faceFlux) {
            // dsmcCloud::addNewParcel with newParcel != -1 (dsmcCloud.C:429-437) -> dsmcFaceTracker::trackFaceTransition
            const double sgn = dot(U, sF) >= 0 ? 1.0 : -1.0;
            atomicAdd(a.faceFlux + size_t(typeId) * a.nFacesAll + faceI, sgn * a.cf.RWF(cellI));
            atomicAdd(a.faceFlux + (size_t(P.nSpecies) + typeId) * a.nFacesAll + faceI, sgn * a.cf.RWF(cellI) * mass);
        }
    }
}

// ---- dsmcAxisymmetric::axisymmetricWeighting (DSMC/coordinateSystem/derived/axisymmetric/dsmcAxisymmetric.C:50-209) and
// dsmcSpherical::sphericalWeighting (.../spherical/dsmcSpherical.C:50-216: the same loop, the clone keeps the parent's velocity) ----
// One thread per parcel of the sorted cloud (= the occupancy loop of the reference: cell by cell, list order).  The parcel's draw comes
// from the Philox stream (index in that order, step).  pass 0: the parcel takes its cell's RWF; with a smaller RWF than before it is
// cloned floor(old/new - 1) times plus once more with the remaining probability, with a larger one it is deleted with probability
// 1 - old/new.  pass 1: the clones (same state, the angular velocity component mirrored) are written behind the cloud, parent by parent.
__global__ void __launch_bounds__(256) weightKernel(const __grid_constant__ WeightArgs a, int pass) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const DevParams& P = *a.P;
    const ParcelArrays& p = a.p;
    if (pass == 0) {
        const int32_t cell = p.cell[i];
        int32_t nClones = 0;
        if (cell >= 0) {
            Rng rng;
            rng.init(P.seed, uint32_t(i), 0u, a.step, STREAM_WEIGHT);
            const double oldRadialWeight = p.rwf[i], newRadialWeight = a.cf.RWF(cell);
            p.rwf[i] = newRadialWeight;
            if (oldRadialWeight > newRadialWeight) {
                double prob = (oldRadialWeight / newRadialWeight) - 1.0;
                while (prob > 1.0) { nClones += 1; prob -= 1.0; }
                if (prob > rng.sample01()) nClones += 1;
            } else if (newRadialWeight > oldRadialWeight) {
                if ((oldRadialWeight / newRadialWeight) < rng.sample01()) {
                    p.cell[i] = -1;   // cloud_.deleteParticle(p): dropped by the rebuild of the occupancy
                    atomicAdd(&a.counters->weightDeleted, 1);
                }
            }
        }
        a.counts[i] = nClones;
        return;
    }
    const int32_t first = a.counts[i], nClones = a.counts[i + 1] - first;
    for (int32_t k = 0; k < nClones; ++k) {
        const int32_t g = a.base + first + k;
        if (g >= a.capacity) { atomicAdd(&a.counters->overflow, 1ULL); return; }
        double U[3] = {p.ux[i], p.uy[i], p.uz[i]};
        if (a.angularCoordinate >= 0) U[a.angularCoordinate] *= -1.0;   // dsmcAxisymmetric mirrors the angular component, dsmcSpherical clones as is
        p.px[g] = p.px[i]; p.py[g] = p.py[i]; p.pz[g] = p.pz[i];
        p.ux[g] = U[0]; p.uy[g] = U[1]; p.uz[g] = U[2];
        p.cell[g] = p.cell[i]; p.tet[g] = p.tet[i];
        p.origId[g] = int32_t((uint32_t(a.origIdBase) + uint32_t(first + k)) & 0x7fffffffu);
        p.typeId[g] = p.typeId[i];
        if (p.erot) p.erot[g] = p.erot[i];
        for (int m = 0; m < MAX_MODES; ++m) if (m < a.nModes && p.vib[m]) p.vib[m][g] = p.vib[m][i];
        if (p.elevel) p.elevel[g] = p.elevel[i];
        if (p.cls) p.cls[g] = p.cls[i];
        if (p.origProc) p.origProc[g] = uint8_t(a.origProc);
        p.rwf[g] = p.rwf[i];
    }
}

cudaError_t launchWeighting(const WeightArgs& a, int pass, cudaStream_t s) {
    if (a.n <= 0) return cudaSuccess;
    weightKernel<<<(a.n + 255) / 256, 256, 0, s>>>(a, pass);
    return cudaGetLastError();
}

cudaError_t launchInflow(const InflowArgs& a, int pass, cudaStream_t s) {
    const int n = a.nFaces * a.nTypes;
    if (n <= 0) return cudaSuccess;
    inflowKernel<<<(n + 127) / 128, 128, 0, s>>>(a, pass);
    return cudaGetLastError();
}

}  // namespace dsmc
