// host_mesh.h -- host-side polyMesh: primitiveMesh geometry, tet decomposition and
// the "baked" tracking tables consumed by the move kernel.
//
// Upstream algorithms restated here live in OpenFOAM v1706 (not vendored in the
// reference tree): primitiveMeshFaceCentresAndAreas.C,
// primitiveMeshCellCentresAndVols.C, polyMeshTetDecomposition, tetrahedronI.H.
// Reference call sites: BASIC/particle/particleTemplates.C:741-743,830-861,
// BASIC/particle/particleI.H:339-601.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/dsmcb200.h"
#include "vec3.h"

namespace dsmc {

// One tetrahedron (Cc, basePt, pA, pB) of the cell decomposition, with everything the tracker needs to cross it.
// 240 bytes = fifteen 16-byte units: an odd unit stride, so the lanes of a warp that read the same field of different records
// from the shared-memory copy of a cell range hit distinct bank groups (LDS.128 conflict free).
// Tet ids are CELL-MAJOR: the tets of cell c are cellTetStart[c] .. cellTetStart[c+1]-1 in primitiveMesh::cells() face order, tetPt
// ascending -- the records of a run of cells are one contiguous piece of the table (one bulk copy into shared memory).
// Each plane k = 0..3 is {unit normal, (base_k - Ct) . n_k}: the numerator of findTris' lambda from the tet centre is a constant of
// the tet and is baked with it.
struct alignas(16) TetRec {
    double plane[4][4];  // [k] = {Sk/(|Sk| + VSMALL) xyz, (planeBase_k - ct) & n_k}, S = Sa,Sb,Sc,Sd
    double base[3];      // basePt  (plane base point of tris 0,2,3)
    double tol;          // lambdaDistanceToleranceCoeff * cellVolume
    double pA[3];        // pA      (plane base point of tri 1)
    int32_t across;      // tet on the other side of the cell face (tri 0), or -1-boundaryFace
    int32_t nbrCell;     // its cell (internal faces), else -1
    double ct[3];        // tet centre
    int32_t nbr1, nbr2;  // tets entered through tris 1 and 2 (same cell)
    int32_t nbr3;        // tet entered through tri 3 (same cell)
    int32_t cell;        // the cell this tet belongs to
    int32_t face;        // tetFace
    int32_t tetPt;       // tetPt
};
static_assert(sizeof(TetRec) == 240, "TetRec must be 240 bytes");

// Per boundary face (index = face - nInternalFaces).
struct alignas(16) BFaceRec {
    int32_t patch;
    int32_t owner;           // faceCells
    int32_t tet0;            // tet id of (owner, this face, tetPt = 1); tetPt p is tet0 + p - 1
    int32_t nPts;
    int32_t coupledTet0;     // cyclic: tet0 of the coupled face, else -1
    int32_t coupledCell;     // cyclic: owner of the coupled face
    int32_t measIndex;       // row in the wall accumulators or -1
    int32_t pad_;
};
static_assert(sizeof(BFaceRec) == 32, "BFaceRec must be 32 bytes");

struct PatchInfo {
    std::string name;
    int32_t type, start, size, neighbPatch, myProcNo, neighbProcNo, referPatch;
    V3 separation;        // cyclic (translational): position -= separation on the receiving side
    bool separated = false;
    bool userSeparation = false;
};

struct HostMesh {
    int32_t nPoints = 0, nFaces = 0, nInternalFaces = 0, nCells = 0;
    std::vector<V3> points;
    std::vector<int32_t> faceOffsets, facePoints, owner, neighbour;
    std::vector<PatchInfo> patches;
    // derived
    std::vector<V3> faceCentres, faceAreas, cellCentres;
    std::vector<double> cellVolumes;
    std::vector<int32_t> tetBasePtIs;
    std::vector<int32_t> cellFaceOffsets, cellFaces;  // primitiveMesh::cells()
    std::vector<int32_t> cellTetStart;                // [nCells+1] first tet id of each cell
    std::vector<int32_t> faceTet0;                    // [2*nFaces] tet id of (face, tetPt = 1) seen from the owner [2f] / neighbour [2f+1] (-1 on boundary faces)
    std::vector<int32_t> facePatch;                   // per boundary face
    std::vector<int32_t> stageGroupCell;              // [nGroups+1] runs of cells whose tet records fit the move kernel's shared-memory window
    int64_t nTetsTotal = 0;
    V3 boundsMin, boundsMax;
    int32_t solutionD[3] = {1, 1, 1};

    int nFacePts(int f) const { return faceOffsets[f + 1] - faceOffsets[f]; }
    const int32_t* facePts(int f) const { return &facePoints[faceOffsets[f]]; }
    int64_t nTets() const { return nTetsTotal; }
    int32_t tetId(int32_t cell, int32_t tetFace, int32_t tetPt) const {
        return faceTet0[2 * tetFace + (owner[tetFace] != cell ? 1 : 0)] + tetPt - 1;
    }
    // runs of consecutive cells with at most maxTets tets each (a cell with more gets a run of its own)
    void buildStageGroups(int32_t maxTets);

    // Build from the ABI struct; computes whatever geometry the caller did not supply.
    std::string build(const dsmcb200_mesh& m);
    // The four points (Cc, basePt, pA, pB) of tet (cell, face, tetPt): tetIndices::tet().
    void tetPoints(int32_t cell, int32_t face, int32_t tetPt, V3& a, V3& b, V3& c, V3& d,
                   int32_t* basePtLabel = nullptr, int32_t* pALabel = nullptr) const;
    // particle::tetNeighbour for tri 1..3: (face, tetPt) of the tet entered.
    void tetNeighbour(int32_t cell, int32_t face, int32_t tetPt, int tri, int32_t& nFace, int32_t& nTetPt) const;
    // Bake records [first, first+count) of the tet table (tet ids) into out.
    void bakeTets(int64_t first, int64_t count, TetRec* out) const;
    void bakeBFaces(std::vector<BFaceRec>& out) const;
};

}  // namespace dsmc
