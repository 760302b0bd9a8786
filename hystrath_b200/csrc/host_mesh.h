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

// One tetrahedron (Cc, basePt, pA, pB) of the cell decomposition, with everything the
// tracker needs to cross it: 224 bytes (seven 32-byte sectors), one record per (face-tri, side).
// Each plane k = 0..3 is one aligned sector {unit normal, (base_k - Ct) . n_k}: the numerator of
// findTris' lambda from the tet centre is a constant of the tet and is baked with it.
struct alignas(32) TetRec {
    double plane[4][4];  // [k] = {Sk/(|Sk| + VSMALL) xyz, (planeBase_k - ct) & n_k}, S = Sa,Sb,Sc,Sd
    double base[3];      // basePt  (plane base point of tris 0,2,3)
    double tol;          // lambdaDistanceToleranceCoeff * cellVolume
    double pA[3];        // pA      (plane base point of tri 1)
    int32_t nbr01[2];    // [0]: >=0 neighbour cell over an internal face, <0: -1-boundaryFace; [1]: tet entered through tri 1
    double ct[3];        // tet centre
    int32_t nbr23[2];    // tets entered through tris 2 and 3 (same cell)
};
static_assert(sizeof(TetRec) == 224, "TetRec must be 224 bytes");

// Per boundary face (index = face - nInternalFaces).
struct alignas(16) BFaceRec {
    int32_t patch;
    int32_t owner;           // faceCells
    int32_t tetPair0;        // first tet-pair index of this face
    int32_t nPts;
    int32_t coupledTetPair0; // cyclic: first tet-pair index of the coupled face, else -1
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
    std::vector<int32_t> faceTetPair0;                // prefix sum of (nPts-2) per face
    std::vector<int32_t> tetPairFace;                 // inverse of faceTetPair0
    std::vector<int32_t> facePatch;                   // per boundary face
    int64_t nTetPairs = 0;
    V3 boundsMin, boundsMax;
    int32_t solutionD[3] = {1, 1, 1};

    int nFacePts(int f) const { return faceOffsets[f + 1] - faceOffsets[f]; }
    const int32_t* facePts(int f) const { return &facePoints[faceOffsets[f]]; }
    int64_t nTets() const { return 2 * nTetPairs; }
    int32_t tetId(int32_t cell, int32_t tetFace, int32_t tetPt) const {
        return 2 * (faceTetPair0[tetFace] + tetPt - 1) + (owner[tetFace] != cell ? 1 : 0);
    }

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
