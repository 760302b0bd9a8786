// host_mesh.cpp -- see host_mesh.h.  Compiled with -ffp-contract=off.
#include "host_mesh.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace dsmc {

namespace {

constexpr double kLambdaDistanceToleranceCoeff = 1.0e3 * SMALL;  // BASIC/particle/particle.C:35
constexpr double kMinTetQuality = 1.0e-9;                         // polyMeshTetDecomposition::minTetQuality

inline V3 triNormal(const V3& a, const V3& b, const V3& c) { return 0.5 * cross(b - a, c - a); }

double tetMag(const V3& a, const V3& b, const V3& c, const V3& d) {
    return (1.0 / 6.0) * dot(cross(b - a, c - a), d - a);
}

double tetCircumRadius(const V3& a_, const V3& b_, const V3& c_, const V3& d_) {
    V3 a = b_ - a_, b = c_ - a_, c = d_ - a_;
    double lambda = magSqr(c) - dot(a, c);
    double mu = magSqr(b) - dot(a, b);
    V3 ba = cross(b, a), ca = cross(c, a);
    V3 num = lambda * ba - mu * ca;
    double denom = dot(c, ba);
    if (std::fabs(denom) < ROOTVSMALL) return GREAT;
    return mag(0.5 * (a + num / denom));
}

double tetQuality(const V3& a, const V3& b, const V3& c, const V3& d) {
    double r = std::min(tetCircumRadius(a, b, c, d), GREAT);
    return tetMag(a, b, c, d) / (8.0 / (9.0 * std::sqrt(3.0)) * (r * r * r) + ROOTVSMALL);
}

// face::edgeDirection
int edgeDirection(const int32_t* f, int n, int32_t e0, int32_t e1) {
    for (int i = 0; i < n; ++i) {
        int rc = (i + n - 1) % n, fc = (i + 1) % n;
        if (f[i] == e0) {
            if (f[rc] == e1) return -1;
            if (f[fc] == e1) return 1;
            return 0;
        } else if (f[i] == e1) {
            if (f[rc] == e0) return 1;
            if (f[fc] == e0) return -1;
            return 0;
        }
    }
    return 0;
}

bool sameFace(const int32_t* a, int na, const int32_t* b, int nb) {
    if (na != nb) return false;
    int s = -1;
    for (int i = 0; i < nb; ++i)
        if (b[i] == a[0]) { s = i; break; }
    if (s < 0) return false;
    bool fwd = true, rev = true;
    for (int i = 0; i < na; ++i) {
        if (a[i] != b[(s + i) % nb]) fwd = false;
        if (a[i] != b[(s - i + nb) % nb]) rev = false;
    }
    return fwd || rev;
}

}  // namespace

std::string HostMesh::build(const dsmcb200_mesh& m) {
    if (m.nPoints <= 0 || m.nFaces <= 0 || m.nCells <= 0 || m.nInternalFaces < 0 || m.nInternalFaces > m.nFaces)
        return "mesh: bad sizes";
    if (!m.points || !m.faceOffsets || !m.facePoints || !m.owner || (m.nInternalFaces && !m.neighbour))
        return "mesh: null array";
    nPoints = m.nPoints; nFaces = m.nFaces; nInternalFaces = m.nInternalFaces; nCells = m.nCells;
    points.resize(nPoints);
    for (int i = 0; i < nPoints; ++i) points[i] = mk(m.points[3 * i], m.points[3 * i + 1], m.points[3 * i + 2]);
    faceOffsets.assign(m.faceOffsets, m.faceOffsets + nFaces + 1);
    facePoints.assign(m.facePoints, m.facePoints + faceOffsets[nFaces]);
    owner.assign(m.owner, m.owner + nFaces);
    neighbour.assign(m.neighbour, m.neighbour + nInternalFaces);
    for (int f = 0; f < nFaces; ++f) {
        if (nFacePts(f) < 3) return "mesh: face with fewer than 3 points";
        if (owner[f] < 0 || owner[f] >= nCells) return "mesh: owner out of range";
        if (f < nInternalFaces && (neighbour[f] < 0 || neighbour[f] >= nCells)) return "mesh: neighbour out of range";
    }
    for (int32_t p : facePoints)
        if (p < 0 || p >= nPoints) return "mesh: face point label out of range";

    patches.clear();
    facePatch.assign(nFaces - nInternalFaces, -1);
    for (int p = 0; p < m.nPatches; ++p) {
        const dsmcb200_patch& s = m.patches[p];
        PatchInfo pi;
        pi.name = std::string(s.name, strnlen(s.name, DSMCB200_NAME_LEN));
        pi.type = s.type; pi.start = s.start; pi.size = s.size; pi.neighbPatch = s.neighbPatch;
        pi.myProcNo = s.myProcNo; pi.neighbProcNo = s.neighbProcNo; pi.referPatch = s.referPatch;
        pi.separation = mk(s.separation[0], s.separation[1], s.separation[2]);
        pi.separated = s.hasSeparation != 0;
        pi.userSeparation = s.hasSeparation != 0;
        if (s.size < 0 || s.start < nInternalFaces || s.start + s.size > nFaces) return "mesh: patch '" + pi.name + "' range";
        for (int i = 0; i < s.size; ++i) facePatch[s.start + i - nInternalFaces] = p;
        patches.push_back(pi);
    }
    for (size_t i = 0; i < facePatch.size(); ++i)
        if (facePatch[i] < 0) return "mesh: boundary face not covered by any patch";

    // bounds (boundBox of points)
    boundsMin = boundsMax = points[0];
    for (const V3& p : points) {
        boundsMin = mk(std::min(boundsMin.x, p.x), std::min(boundsMin.y, p.y), std::min(boundsMin.z, p.z));
        boundsMax = mk(std::max(boundsMax.x, p.x), std::max(boundsMax.y, p.y), std::max(boundsMax.z, p.z));
    }

    // ---- face centres and areas (primitiveMesh::makeFaceCentresAndAreas)
    faceCentres.resize(nFaces); faceAreas.resize(nFaces);
    if (m.faceCentres && m.faceAreas) {
        for (int f = 0; f < nFaces; ++f) {
            faceCentres[f] = mk(m.faceCentres[3 * f], m.faceCentres[3 * f + 1], m.faceCentres[3 * f + 2]);
            faceAreas[f] = mk(m.faceAreas[3 * f], m.faceAreas[3 * f + 1], m.faceAreas[3 * f + 2]);
        }
    } else {
#pragma omp parallel for schedule(static)
        for (int f = 0; f < nFaces; ++f) {
            const int32_t* fp = facePts(f);
            int n = nFacePts(f);
            if (n == 3) {
                faceCentres[f] = (1.0 / 3.0) * (points[fp[0]] + points[fp[1]] + points[fp[2]]);
                faceAreas[f] = 0.5 * cross(points[fp[1]] - points[fp[0]], points[fp[2]] - points[fp[0]]);
            } else {
                V3 sumN = mk(0, 0, 0), sumAc = mk(0, 0, 0);
                double sumA = 0.0;
                V3 fCentre = points[fp[0]];
                for (int pi = 1; pi < n; ++pi) fCentre += points[fp[pi]];
                fCentre /= double(n);
                for (int pi = 0; pi < n; ++pi) {
                    const V3& nextPoint = points[fp[(pi + 1) % n]];
                    V3 c = points[fp[pi]] + nextPoint + fCentre;
                    V3 nn = cross(nextPoint - points[fp[pi]], fCentre - points[fp[pi]]);
                    double a = mag(nn);
                    sumN += nn; sumA += a; sumAc += a * c;
                }
                if (sumA < ROOTVSMALL) { faceCentres[f] = fCentre; faceAreas[f] = mk(0, 0, 0); }
                else { faceCentres[f] = (1.0 / 3.0) * sumAc / sumA; faceAreas[f] = 0.5 * sumN; }
            }
        }
    }

    // ---- cells (primitiveMesh::calcCells: owned faces first, then neighbour faces)
    cellFaceOffsets.assign(nCells + 1, 0);
    for (int f = 0; f < nFaces; ++f) cellFaceOffsets[owner[f] + 1]++;
    for (int f = 0; f < nInternalFaces; ++f) cellFaceOffsets[neighbour[f] + 1]++;
    for (int c = 0; c < nCells; ++c) cellFaceOffsets[c + 1] += cellFaceOffsets[c];
    cellFaces.assign(cellFaceOffsets[nCells], -1);
    {
        std::vector<int32_t> cur(cellFaceOffsets.begin(), cellFaceOffsets.end() - 1);
        for (int f = 0; f < nFaces; ++f) cellFaces[cur[owner[f]]++] = f;
        for (int f = 0; f < nInternalFaces; ++f) cellFaces[cur[neighbour[f]]++] = f;
    }

    // ---- cell centres and volumes (primitiveMesh::makeCellCentresAndVols)
    cellCentres.assign(nCells, mk(0, 0, 0)); cellVolumes.assign(nCells, 0.0);
    if (m.cellCentres && m.cellVolumes) {
        for (int c = 0; c < nCells; ++c) {
            cellCentres[c] = mk(m.cellCentres[3 * c], m.cellCentres[3 * c + 1], m.cellCentres[3 * c + 2]);
            cellVolumes[c] = m.cellVolumes[c];
        }
    } else {
        std::vector<V3> cEst(nCells, mk(0, 0, 0));
        std::vector<int32_t> nCellFaces(nCells, 0);
        for (int f = 0; f < nFaces; ++f) { cEst[owner[f]] += faceCentres[f]; nCellFaces[owner[f]] += 1; }
        for (int f = 0; f < nInternalFaces; ++f) { cEst[neighbour[f]] += faceCentres[f]; nCellFaces[neighbour[f]] += 1; }
        for (int c = 0; c < nCells; ++c) cEst[c] /= double(nCellFaces[c]);
        for (int f = 0; f < nFaces; ++f) {
            int c = owner[f];
            double pyr3Vol = dot(faceAreas[f], faceCentres[f] - cEst[c]);
            V3 pc = (3.0 / 4.0) * faceCentres[f] + (1.0 / 4.0) * cEst[c];
            cellCentres[c] += pyr3Vol * pc;
            cellVolumes[c] += pyr3Vol;
        }
        for (int f = 0; f < nInternalFaces; ++f) {
            int c = neighbour[f];
            double pyr3Vol = dot(faceAreas[f], cEst[c] - faceCentres[f]);
            V3 pc = (3.0 / 4.0) * faceCentres[f] + (1.0 / 4.0) * cEst[c];
            cellCentres[c] += pyr3Vol * pc;
            cellVolumes[c] += pyr3Vol;
        }
        for (int c = 0; c < nCells; ++c) {
            if (std::fabs(cellVolumes[c]) > VSMALL) cellCentres[c] /= cellVolumes[c];
            else cellCentres[c] = cEst[c];
            cellVolumes[c] *= (1.0 / 3.0);
        }
    }

    // ---- tet base points (polyMeshTetDecomposition::findSharedBasePoint / findBasePoint)
    tetBasePtIs.assign(nFaces, 0);
    if (m.tetBasePtIs) {
        tetBasePtIs.assign(m.tetBasePtIs, m.tetBasePtIs + nFaces);
    } else {
        bool bad = false;
#pragma omp parallel for schedule(static)
        for (int f = 0; f < nFaces; ++f) {
            const int32_t* fp = facePts(f);
            int n = nFacePts(f);
            const V3& oCc = cellCentres[owner[f]];
            int found = -1;
            for (int faceBasePtI = 0; faceBasePtI < n && found < 0; ++faceBasePtI) {
                double minQ = VGREAT;
                const V3& tetBasePt = points[fp[faceBasePtI]];
                for (int tetPtI = 1; tetPtI < n - 1; ++tetPtI) {
                    int facePtI = (tetPtI + faceBasePtI) % n;
                    int otherFacePtI = (facePtI + 1) % n;
                    double q = tetQuality(oCc, tetBasePt, points[fp[facePtI]], points[fp[otherFacePtI]]);
                    if (q < minQ) minQ = q;
                    if (f < nInternalFaces) {
                        const V3& nCc = cellCentres[neighbour[f]];
                        q = tetQuality(nCc, tetBasePt, points[fp[otherFacePtI]], points[fp[facePtI]]);
                        if (q < minQ) minQ = q;
                    }
                }
                if (minQ > kMinTetQuality) found = faceBasePtI;
            }
            if (found < 0) { found = 0; bad = true; }
            tetBasePtIs[f] = found;
        }
        (void)bad;  // degenerate faces fall back to base point 0 (the reference would abort in tetNeighbour)
    }

    // ---- tet numbering: cell-major, faces in cells() order, tetPt ascending
    cellTetStart.assign(nCells + 1, 0);
    faceTet0.assign(size_t(2) * nFaces, -1);
    {
        int64_t t = 0;
        for (int c = 0; c < nCells; ++c) {
            cellTetStart[c] = int32_t(t);
            for (int k = cellFaceOffsets[c]; k < cellFaceOffsets[c + 1]; ++k) {
                const int f = cellFaces[k];
                faceTet0[2 * size_t(f) + (owner[f] == c ? 0 : 1)] = int32_t(t);
                t += nFacePts(f) - 2;
                if (t >= (int64_t(1) << 31) - 2) return "mesh: too many tets for int32 tet ids";
            }
        }
        cellTetStart[nCells] = int32_t(t);
        nTetsTotal = t;
    }

    // ---- solution directions (polyMesh::calcDirections)
    {
        V3 emptyDirVec = mk(0, 0, 0);
        int nEmpty = 0;
        for (const PatchInfo& p : patches)
            if (p.type == DSMCB200_PATCH_EMPTY && p.size > 0) {
                nEmpty++;
                for (int i = 0; i < p.size; ++i) {
                    const V3& s = faceAreas[p.start + i];
                    emptyDirVec += mk(std::fabs(s.x), std::fabs(s.y), std::fabs(s.z));
                }
            }
        solutionD[0] = solutionD[1] = solutionD[2] = 1;
        if (nEmpty) {
            emptyDirVec /= mag(emptyDirVec);
            for (int d = 0; d < 3; ++d) solutionD[d] = comp(emptyDirVec, d) > 1e-6 ? -1 : 1;
        }
    }

    // ---- cyclic separation (coupledPolyPatch::calcTransformTensors, translational case):
    // separation = (nf & (Cr - Cf)) * nf on each patch, collapsed to one vector when uniform.
    for (size_t p = 0; p < patches.size(); ++p) {
        PatchInfo& pi = patches[p];
        if (pi.type != DSMCB200_PATCH_CYCLIC) continue;
        if (pi.userSeparation) continue;
        if (pi.neighbPatch < 0 || pi.neighbPatch >= (int)patches.size()) return "mesh: cyclic patch '" + pi.name + "' has no neighbourPatch";
        const PatchInfo& nb = patches[pi.neighbPatch];
        if (nb.size != pi.size) return "mesh: cyclic patch sizes differ for '" + pi.name + "'";
        if (pi.size == 0) continue;
        V3 s0 = mk(0, 0, 0);
        for (int i = 0; i < pi.size; ++i) {
            V3 nf = faceAreas[pi.start + i];
            nf /= mag(nf);
            V3 s = dot(nf, faceCentres[nb.start + i] - faceCentres[pi.start + i]) * nf;
            if (i == 0) s0 = s;
            else if (mag(s - s0) > 1e-6 * std::max(mag(s0), 1e-30)) return "mesh: non-uniform cyclic separation on '" + pi.name + "' is not supported";
            // rotational cyclics (non-parallel normals) are outside the scoped path
            V3 nn = faceAreas[nb.start + i];
            nn /= mag(nn);
            if (std::fabs(dot(nf, nn) + 1.0) > 1e-6) return "mesh: rotational cyclic '" + pi.name + "' is not supported";
        }
        pi.separation = s0;
        pi.separated = mag(s0) > 0;
    }
    return "";
}

void HostMesh::tetPoints(int32_t cell, int32_t face, int32_t tetPt, V3& a, V3& b, V3& c, V3& d,
                         int32_t* basePtLabel, int32_t* pALabel) const {
    const int32_t* f = facePts(face);
    int n = nFacePts(face);
    bool own = owner[face] == cell;
    int tetBasePtI = tetBasePtIs[face];
    int facePtI = (tetPt + tetBasePtI) % n;
    int otherFacePtI = (facePtI + 1) % n;
    int fPtAI = own ? facePtI : otherFacePtI;
    int fPtBI = own ? otherFacePtI : facePtI;
    a = cellCentres[cell];
    b = points[f[tetBasePtI]];
    c = points[f[fPtAI]];
    d = points[f[fPtBI]];
    if (basePtLabel) *basePtLabel = f[tetBasePtI];
    if (pALabel) *pALabel = f[fPtAI];
}

void HostMesh::tetNeighbour(int32_t cell, int32_t face, int32_t tetPt, int tri, int32_t& nFace, int32_t& nTetPt) const {
    const int32_t* f = facePts(face);
    int n = nFacePts(face);
    bool own = owner[face] == cell;
    int tetBasePtI = tetBasePtIs[face];
    int facePtI = (tetPt + tetBasePtI) % n;
    int otherFacePtI = (facePtI + 1) % n;
    nFace = face; nTetPt = tetPt;
    int32_t e0 = -1, e1 = -1;
    bool crossEdge = false;
    switch (tri) {
        case 1: e0 = f[facePtI]; e1 = f[otherFacePtI]; crossEdge = true; break;
        case 2:
            if (own) {
                if (tetPt < n - 2) nTetPt = (tetPt + 1) % n;
                else { e0 = f[tetBasePtI]; e1 = f[otherFacePtI]; crossEdge = true; }
            } else {
                if (tetPt > 1) nTetPt = (tetPt + n - 1) % n;
                else { e0 = f[tetBasePtI]; e1 = f[facePtI]; crossEdge = true; }
            }
            break;
        case 3:
            if (own) {
                if (tetPt > 1) nTetPt = (tetPt + n - 1) % n;
                else { e0 = f[tetBasePtI]; e1 = f[facePtI]; crossEdge = true; }
            } else {
                if (tetPt < n - 2) nTetPt = (tetPt + 1) % n;
                else { e0 = f[tetBasePtI]; e1 = f[otherFacePtI]; crossEdge = true; }
            }
            break;
        default: break;
    }
    if (!crossEdge) return;
    // particle::crossEdgeConnectedFace
    for (int k = cellFaceOffsets[cell]; k < cellFaceOffsets[cell + 1]; ++k) {
        int fI = cellFaces[k];
        if (fI == face) continue;
        const int32_t* of = facePts(fI);
        int on = nFacePts(fI);
        int edDir = edgeDirection(of, on, e0, e1);
        if (edDir == 0) continue;
        if (sameFace(f, n, of, on)) continue;
        nFace = fI;
        int32_t target = edDir == 1 ? e0 : e1;
        int eIndex = -1;
        for (int i = 0; i < on; ++i)
            if (of[i] == target) { eIndex = i; break; }
        eIndex -= tetBasePtIs[fI];
        if (eIndex < 0) eIndex = (eIndex + on) % on;
        if (eIndex == 0) nTetPt = 1;
        else if (eIndex == on - 1) nTetPt = on - 2;
        else nTetPt = eIndex;
        break;
    }
}

void HostMesh::buildStageGroups(int32_t maxTets) {
    stageGroupCell.clear();
    stageGroupCell.push_back(0);
    int32_t c = 0;
    while (c < nCells) {
        const int32_t t0 = cellTetStart[c];
        int32_t e = c + 1;
        while (e < nCells && cellTetStart[e + 1] - t0 <= maxTets) ++e;
        stageGroupCell.push_back(e);
        c = e;
    }
}

void HostMesh::bakeTets(int64_t first, int64_t count, TetRec* out) const {
    // the cell of the first tet, then a running cursor per thread chunk
#pragma omp parallel
    {
        int32_t cell = -1;
#pragma omp for schedule(static)
        for (int64_t t = first; t < first + count; ++t) {
            if (cell < 0 || t < cellTetStart[cell] || t >= cellTetStart[cell + 1])
                cell = int32_t(std::upper_bound(cellTetStart.begin(), cellTetStart.end(), int32_t(t)) - cellTetStart.begin()) - 1;
            // (face, tetPt) of the local index
            int32_t face = -1, tetPt = -1;
            {
                int32_t rel = int32_t(t) - cellTetStart[cell];
                for (int k = cellFaceOffsets[cell]; k < cellFaceOffsets[cell + 1]; ++k) {
                    const int32_t f = cellFaces[k];
                    const int32_t nT = nFacePts(f) - 2;
                    if (rel < nT) { face = f; tetPt = rel + 1; break; }
                    rel -= nT;
                }
            }
            TetRec& r = out[t - first];
            std::memset(&r, 0, sizeof(TetRec));
            V3 a, b, c, d;
            tetPoints(cell, face, tetPt, a, b, c, d);
            V3 S[4];
            S[0] = triNormal(b, c, d);  // Sa
            S[1] = triNormal(a, d, c);  // Sb
            S[2] = triNormal(a, b, d);  // Sc
            S[3] = triNormal(a, c, b);  // Sd
            const V3 ct = 0.25 * (a + b + c + d);
            for (int i = 0; i < 4; ++i) {
                S[i] /= (mag(S[i]) + VSMALL);  // BASIC/particle/particleTemplates.C:901-904
                r.plane[i][0] = S[i].x; r.plane[i][1] = S[i].y; r.plane[i][2] = S[i].z;
                const V3& planeBase = (i == 1) ? c : b;  // tetPlaneBasePtIs, particleTemplates.C:907-912
                r.plane[i][3] = dot(planeBase - ct, S[i]);  // lambdaNumerator of findTris (from = tet centre)
            }
            r.base[0] = b.x; r.base[1] = b.y; r.base[2] = b.z;
            r.pA[0] = c.x; r.pA[1] = c.y; r.pA[2] = c.z;
            r.ct[0] = ct.x; r.ct[1] = ct.y; r.ct[2] = ct.z;
            r.tol = kLambdaDistanceToleranceCoeff * cellVolumes[cell];
            r.cell = cell; r.face = face; r.tetPt = tetPt;
            if (face < nInternalFaces) {
                const bool own = owner[face] == cell;
                r.nbrCell = own ? neighbour[face] : owner[face];
                r.across = faceTet0[2 * size_t(face) + (own ? 1 : 0)] + tetPt - 1;  // the same face triangle seen from the other cell
            } else {
                r.nbrCell = -1;
                r.across = -1 - (face - nInternalFaces);
            }
            int32_t nb[4] = {0, 0, 0, 0};
            for (int tri = 1; tri <= 3; ++tri) {
                int32_t nf, np;
                tetNeighbour(cell, face, tetPt, tri, nf, np);
                nb[tri] = tetId(cell, nf, np);
            }
            r.nbr1 = nb[1]; r.nbr2 = nb[2]; r.nbr3 = nb[3];
        }
    }
}

void HostMesh::bakeBFaces(std::vector<BFaceRec>& out) const {
    int nB = nFaces - nInternalFaces;
    out.resize(nB);
    for (int i = 0; i < nB; ++i) {
        int f = nInternalFaces + i;
        BFaceRec& r = out[i];
        r.patch = facePatch[i];
        r.owner = owner[f];
        r.tet0 = faceTet0[2 * size_t(f)];
        r.nPts = nFacePts(f);
        r.coupledTet0 = -1; r.coupledCell = -1; r.measIndex = -1; r.pad_ = 0;
        const PatchInfo& p = patches[r.patch];
        if (p.type == DSMCB200_PATCH_CYCLIC) {
            int cf = f - p.start + patches[p.neighbPatch].start;  // cyclicPolyPatch::transformGlobalFace
            r.coupledTet0 = faceTet0[2 * size_t(cf)];
            r.coupledCell = owner[cf];
        }
    }
}

}  // namespace dsmc
