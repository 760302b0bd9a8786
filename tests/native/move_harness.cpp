// move_harness.cpp -- CPU harness of the move kernel's per-visit code (TEST INFRASTRUCTURE, never loaded by the product).
//
// Runs hystrath_b200/csrc/move_core.h (visitFast / visitSlow) on the tet table baked by hystrath_b200/csrc/host_mesh.cpp, parcel
// by parcel, with the sequencing of moveKernel's sections (kernels_move.cu): what the GPU executes per lane, minus the launch
// machinery.  tests/test_move_core.py compares it bit for bit with the oracle's restatement of particle::trackToFace
// (oracle/oracle.cpp), so the tables, the cell-major tet numbering and the division-free plane tests are pinned in the CPU suite.
// Boundary handling is limited to what needs no species data: internal faces, cyclic, symmetry(Plane)/wedge, specular walls
// (patch type wall) and deletion (patch type patch).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../hystrath_b200/csrc/host_mesh.h"
#include "../../hystrath_b200/csrc/move_core.h"

using namespace dsmc;

namespace {
TetRegs regs(const TetRec& t) {
    TetRegs r;
    r.N0 = mk(t.plane[0][0], t.plane[0][1], t.plane[0][2]); r.numC0 = t.plane[0][3];
    r.N1 = mk(t.plane[1][0], t.plane[1][1], t.plane[1][2]); r.numC1 = t.plane[1][3];
    r.N2 = mk(t.plane[2][0], t.plane[2][1], t.plane[2][2]); r.numC2 = t.plane[2][3];
    r.N3 = mk(t.plane[3][0], t.plane[3][1], t.plane[3][2]); r.numC3 = t.plane[3][3];
    r.base = mk(t.base[0], t.base[1], t.base[2]); r.pA = mk(t.pA[0], t.pA[1], t.pA[2]); r.Ct = mk(t.ct[0], t.ct[1], t.ct[2]);
    r.tol = t.tol; r.across = t.across; r.nbrCell = t.nbrCell; r.nbr1 = t.nbr1; r.nbr2 = t.nbr2; r.nbr3 = t.nbr3;
    return r;
}
}  // namespace

extern "C" {

// in/out: pos[3n], U[3n], cell[n], tetFace[n], tetPt[n]; cell = -1 on return for deleted parcels.
// stats[0] = rescues, stats[1] = visits, stats[2] = visits that took the slow path, stats[3] = tets in the table
int movecheck_run(const dsmcb200_mesh* m, double deltaT, int64_t n, double* pos3, double* U3, int32_t* cellIO, int32_t* tetFaceIO, int32_t* tetPtIO,
                  int forceSlow, int64_t* stats) {
    HostMesh M;
    if (!M.build(*m).empty()) return 1;
    std::vector<TetRec> tets(size_t(M.nTets()));
    M.bakeTets(0, M.nTets(), tets.data());
    std::vector<BFaceRec> bfaces;
    M.bakeBFaces(bfaces);
    const bool constrained = M.solutionD[0] == -1 || M.solutionD[1] == -1 || M.solutionD[2] == -1;
    V3 centre = 0.5 * (M.boundsMin + M.boundsMax);
    int64_t rescues = 0, visits = 0, slow = 0;
    for (int64_t i = 0; i < n; ++i) {
        int32_t cell = cellIO[i];
        if (cell < 0) continue;
        int32_t tet = M.tetId(cell, tetFaceIO[i], tetPtIO[i]);
        if (tets[tet].cell != cell || tets[tet].face != tetFaceIO[i] || tets[tet].tetPt != tetPtIO[i]) return 2;
        V3 pos = mk(pos3[3 * i], pos3[3 * i + 1], pos3[3 * i + 2]), U = mk(U3[3 * i], U3[3 * i + 1], U3[3 * i + 2]);
        double tEnd = deltaT;
        bool keepParticle = true, switchProcessor = false;
        int guard = 0;
        while (keepParticle && !switchProcessor && tEnd > ROOTVSMALL) {
            V3 Utracking = U;
            if (constrained)
                for (int d = 0; d < 3; ++d)
                    if (M.solutionD[d] == -1) { setComp(pos, d, comp(centre, d)); setComp(Utracking, d, 0.0); }
            const V3 endPosition = pos + tEnd * Utracking;
            double trackFraction = 0.0, retVal = 1.0;
            bool rescuePending = false, faceSet = false, finished = false;
            int32_t faceBfi = -1;
            while (!finished) {
                if (++guard > 200000) return 3;
                const TetRegs R = regs(tets[tet]);
                VisitOut v;
                v.code = VISIT_SLOW; v.triI = -1; v.needRescue = false;
                ++visits;
                if (!rescuePending && !forceSlow) v = visitFast(R, pos, endPosition, trackFraction);
                if (v.code == VISIT_SLOW) {
                    ++slow;
                    v = visitSlow(R, pos, endPosition, trackFraction, rescuePending);
                    if (v.code == VISIT_RESCUED) ++rescues;
                }
                if (v.code != VISIT_RESCUED) {
                    const bool onFace = v.triI == 0;
                    faceSet = onFace;
                    faceBfi = (onFace && R.across < 0) ? (-1 - R.across) : -1;
                }
                finished = v.code == VISIT_RESCUED || v.code == VISIT_END;
                retVal = v.code == VISIT_RESCUED ? trackFraction : 1.0;
                if (v.code == VISIT_MOVE) {
                    if (v.triI > 0) {
                        tet = v.triI == 1 ? R.nbr1 : (v.triI == 2 ? R.nbr2 : R.nbr3);
                        rescuePending = v.needRescue;
                    } else {
                        if (R.across >= 0) {
                            cell = R.nbrCell;
                            tet = R.across;
                        } else {
                            const int32_t bfi = -1 - R.across;
                            const BFaceRec& bf = bfaces[bfi];
                            const PatchInfo& pt = M.patches[bf.patch];
                            switch (pt.type) {
                                case DSMCB200_PATCH_PROCESSOR:
                                case DSMCB200_PATCH_PROCESSORCYCLIC: switchProcessor = true; break;
                                case DSMCB200_PATCH_SYMMETRYPLANE:
                                case DSMCB200_PATCH_SYMMETRY:
                                case DSMCB200_PATCH_WEDGE: {
                                    const V3 nf = R.N0;
                                    const V3 t2 = 2.0 * nf;
                                    const double xx = 1.0 - t2.x * nf.x, xy = 0.0 - t2.x * nf.y, xz = 0.0 - t2.x * nf.z;
                                    const double yx = 0.0 - t2.y * nf.x, yy = 1.0 - t2.y * nf.y, yz = 0.0 - t2.y * nf.z;
                                    const double zx = 0.0 - t2.z * nf.x, zy = 0.0 - t2.z * nf.y, zz = 1.0 - t2.z * nf.z;
                                    U = mk(xx * U.x + xy * U.y + xz * U.z, yx * U.x + yy * U.y + yz * U.z, zx * U.x + zy * U.y + zz * U.z);
                                    break;
                                }
                                case DSMCB200_PATCH_CYCLIC: {
                                    const int32_t k = tet - bf.tet0;
                                    tet = bf.coupledTet0 + (bf.nPts - 3) - k;
                                    cell = bf.coupledCell;
                                    const PatchInfo& rp = M.patches[pt.neighbPatch];
                                    pos -= rp.separation;
                                    faceBfi = bfi - (pt.start - M.nInternalFaces) + (rp.start - M.nInternalFaces);
                                    break;
                                }
                                case DSMCB200_PATCH_WALL: {   // dsmcSpecularWallPatch
                                    const double U_dot_nw = dot(U, R.N0);
                                    if (U_dot_nw > 0.0) U -= 2.0 * U_dot_nw * R.N0;
                                    break;
                                }
                                case DSMCB200_PATCH_PATCH: keepParticle = false; break;   // dsmcDeletionPatch
                                default: break;
                            }
                        }
                        if (v.needRescue) rescuePending = true;
                        else { retVal = trackFraction; finished = true; }
                    }
                }
                if (!keepParticle) finished = true;
            }
            if (keepParticle) {
                const double dt = tEnd * retVal;
                tEnd -= dt;
                if (faceSet && faceBfi >= 0) {
                    const int ptype = M.patches[bfaces[faceBfi].patch].type;
                    if (ptype == DSMCB200_PATCH_PROCESSOR || ptype == DSMCB200_PATCH_PROCESSORCYCLIC) switchProcessor = true;
                }
            }
        }
        pos3[3 * i] = pos.x; pos3[3 * i + 1] = pos.y; pos3[3 * i + 2] = pos.z;
        U3[3 * i] = U.x; U3[3 * i + 1] = U.y; U3[3 * i + 2] = U.z;
        cellIO[i] = keepParticle ? cell : -1;
        tetFaceIO[i] = tets[tet].face;
        tetPtIO[i] = tets[tet].tetPt;
    }
    if (stats) { stats[0] = rescues; stats[1] = visits; stats[2] = slow; stats[3] = M.nTets(); }
    return 0;
}

// the baked table itself, for invariants (every in-cell link stays in the cell, across links are mutual, ...)
int movecheck_tets(const dsmcb200_mesh* m, int64_t capacity, int32_t* cell, int32_t* face, int32_t* tetPt, int32_t* across, int32_t* nbr123, int32_t* cellTetStart) {
    HostMesh M;
    if (!M.build(*m).empty()) return -1;
    if (M.nTets() > capacity) return int(M.nTets());
    std::vector<TetRec> tets(size_t(M.nTets()));
    M.bakeTets(0, M.nTets(), tets.data());
    for (int64_t t = 0; t < M.nTets(); ++t) {
        cell[t] = tets[t].cell; face[t] = tets[t].face; tetPt[t] = tets[t].tetPt; across[t] = tets[t].across;
        nbr123[3 * t] = tets[t].nbr1; nbr123[3 * t + 1] = tets[t].nbr2; nbr123[3 * t + 2] = tets[t].nbr3;
    }
    for (int c = 0; c <= M.nCells; ++c) cellTetStart[c] = M.cellTetStart[c];
    return int(M.nTets());
}

}  // extern "C"
