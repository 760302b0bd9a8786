// dsmc_cloud.cpp -- see dsmc_cloud.h
#include "dsmc_cloud.h"

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <ctime>
#include <sstream>

namespace dsmcb200 {

namespace {
int gCellOrder = DSMCB200_CELL_ORDER_AS_GIVEN;
}
void dsmcCloud::cellOrder(int mode) { gCellOrder = mode; }

using foam::Dict;
using foam::FoamError;

namespace {
const double SMALL = 1e-15, VSMALL = 1e-300, GREAT = 1e15, VGREAT = 1e300;

std::string joinNames(const std::vector<std::string>& v) {
    std::ostringstream s;
    s << v.size() << "(";
    for (size_t i = 0; i < v.size(); ++i) s << (i ? " " : "") << v[i];
    s << ")";
    return s.str();
}

// the message shape of OpenFOAM's run-time selection failure (BinaryCollisionModel.C:70-85)
[[noreturn]] void unknownType(const std::string& where, const std::string& family, const std::string& name, const std::vector<std::string>& valid) {
    throw FoamError(where + " : \n    unknown " + family + " type " + name + ", constructor not in hash table\n\n    Valid " + family +
                    " types are :\n" + joinNames(valid));
}

void copyName(char* dst, const std::string& s) {
    std::memset(dst, 0, DSMCB200_NAME_LEN);
    std::strncpy(dst, s.c_str(), DSMCB200_NAME_LEN - 1);
}
}  // namespace

int selectBinaryCollisionModel(const std::string& name) {
    if (name == "NoBinaryCollision") return DSMCB200_COLL_NONE;
    if (name == "VariableHardSphere") return DSMCB200_COLL_VHS;
    if (name == "LarsenBorgnakkeVariableHardSphere") return DSMCB200_COLL_LB_VHS;
    if (name == "VariableSoftSphere") return DSMCB200_COLL_VSS;
    if (name == "LarsenBorgnakkeVariableSoftSphere") return DSMCB200_COLL_LB_VSS;
    unknownType("BinaryCollisionModel::New(const dictionary&, CloudType&)", "BinaryCollisionModel", name,
                {"LarsenBorgnakkeVariableHardSphere", "LarsenBorgnakkeVariableSoftSphere", "NoBinaryCollision", "VariableHardSphere", "VariableSoftSphere"});
}
int selectCollisionPartnerSelection(const std::string& name) {
    if (name == "noTimeCounter") return 1;
    unknownType("collisionPartnerSelection::New(const dictionary&)", "collisionPartnerSelection", name, {"noTimeCounter"});
}
int selectPatchBoundaryModel(const std::string& name) {
    if (name == "dsmcDiffuseWallPatch") return DSMCB200_BND_DIFFUSE_WALL;
    if (name == "dsmcSpecularWallPatch") return DSMCB200_BND_SPECULAR_WALL;
    if (name == "dsmcDeletionPatch") return DSMCB200_BND_DELETION;
    if (name == "dsmcCLLWallPatch") return DSMCB200_BND_CLL_WALL;
    if (name == "dsmcDiffuseSpecularWallPatch") return DSMCB200_BND_DIFFUSE_SPECULAR_WALL;
    unknownType("dsmcPatchBoundary::New(const dictionary&)", "patch boundary", name,
                {"dsmcCLLWallPatch", "dsmcDeletionPatch", "dsmcDiffuseSpecularWallPatch", "dsmcDiffuseWallPatch", "dsmcSpecularWallPatch"});
}
void selectGeneralBoundaryModel(const std::string& name) {
    if (name != "dsmcFreeStreamInflowPatch")
        unknownType("dsmcGeneralBoundary::New(const dictionary&)", "general boundary", name, {"dsmcFreeStreamInflowPatch"});
}
void selectFieldModel(const std::string& name) {
    if (name != "dsmcVolFields") unknownType("dsmcField::New(const dictionary&)", "dsmcField", name, {"dsmcVolFields"});
}
int selectCoordinateSystem(const std::string& name) {
    if (name == "dsmcCartesian") return DSMCB200_COORD_CARTESIAN;
    if (name == "dsmcAxisymmetric") return DSMCB200_COORD_AXISYMMETRIC;
    if (name == "dsmcSpherical") return DSMCB200_COORD_SPHERICAL;
    unknownType("dsmcCoordinateSystem::New", "dsmcCoordinateSystem", name, {"dsmcAxisymmetric", "dsmcCartesian", "dsmcSpherical"});
    return 0;
}
// dsmcTimeStepModel::New builds the type name "dsmc" + Keyword + "TimeStepModel" (dsmcTimeStepModel.C:99-107)
bool selectTimeStepModel(const std::string& name) {
    if (name == "constant") return false;
    if (name == "variable") return true;
    std::string type = name.empty() ? name : "dsmc" + std::string(1, char(std::toupper(name[0]))) + name.substr(1) + "TimeStepModel";
    unknownType("dsmcTimeStepModel::New", "dsmcTimeStepModel", type, {"dsmcConstantTimeStepModel", "dsmcVariableTimeStepModel"});
    return false;
}
int patchTypeFromWord(const std::string& t) {
    if (t == "wall") return DSMCB200_PATCH_WALL;
    if (t == "patch") return DSMCB200_PATCH_PATCH;
    if (t == "cyclic") return DSMCB200_PATCH_CYCLIC;
    if (t == "processor") return DSMCB200_PATCH_PROCESSOR;
    if (t == "empty") return DSMCB200_PATCH_EMPTY;
    if (t == "symmetryPlane") return DSMCB200_PATCH_SYMMETRYPLANE;
    if (t == "symmetry") return DSMCB200_PATCH_SYMMETRY;
    if (t == "wedge") return DSMCB200_PATCH_WEDGE;
    if (t == "processorCyclic") return DSMCB200_PATCH_PROCESSORCYCLIC;
    throw FoamError("polyPatch type " + t + " is not supported by the dsmcb200 tracker");
}

void dsmcCloud::check(int rc, const char* what) {
    if (rc != 0) throw FoamError(std::string(what) + ": " + (ctx_ ? dsmcb200_last_error(ctx_) : "no context") + " (" + std::to_string(rc) + ")");
}

dsmcCloud::dsmcCloud(const std::string& caseDir, const std::string& cloudName, int rank, int nRanks, int device, const void* ncclId128, bool dryRun,
                     bool initialise)
    : caseDir_(caseDir), cloudName_(cloudName), rank_(rank), nRanks_(nRanks), dryRun_(dryRun), initialise_(initialise) {
    root_ = nRanks > 1 ? caseDir + "/processor" + std::to_string(rank) : caseDir;
    readControl();
    readMesh();
    readProperties();
    readBoundaries();
    readFieldProperties();
    if (dryRun_) { if (initialise_) initialiseFromDict(); else readCloud(); return; }
    int rc = dsmcb200_create(&ctx_, device, rank, nRanks);
    if (rc != 0) throw FoamError("dsmcb200_create failed: a CUDA device is required, there is no CPU fallback");
    if (nRanks > 1) {
        if (!ncclId128) throw FoamError("parallel run needs an ncclUniqueId");
        check(dsmcb200_init_comm(ctx_, ncclId128), "dsmcb200_init_comm");
    }
    dsmcb200_mesh m{};
    m.nPoints = int32_t(points_.size() / 3); m.nFaces = nFaces_; m.nInternalFaces = nInternal_; m.nCells = nCells_;
    m.nPatches = int32_t(patches_.size());
    m.points = points_.data(); m.faceOffsets = faceOffsets_.data(); m.facePoints = facePoints_.data();
    m.owner = owner_.data(); m.neighbour = neighbour_.data(); m.patches = patches_.data();
    if (gCellOrder != DSMCB200_CELL_ORDER_AS_GIVEN) check(dsmcb200_set_cell_order(ctx_, gCellOrder, nullptr, 0), "dsmcb200_set_cell_order");
    check(dsmcb200_set_mesh(ctx_, &m), "dsmcb200_set_mesh");
    check(dsmcb200_set_species(ctx_, int(species_.size()), species_.data()), "dsmcb200_set_species");
    if (sampleSets_.size() > 1) check(dsmcb200_set_sample_sets(ctx_, int(sampleSets_.size()), sampleSets_.data()), "dsmcb200_set_sample_sets");
    if (!reactions_.empty()) check(dsmcb200_set_reactions(ctx_, int(reactions_.size()), reactions_.data()), "dsmcb200_set_reactions");
    models_.nPatchModels = int32_t(patchModels_.size()); models_.patchModels = patchModels_.data();
    models_.nInflows = int32_t(inflows_.size()); models_.inflows = inflows_.data();
    check(dsmcb200_set_models(ctx_, &models_), "dsmcb200_set_models");
    cellVolumes_.resize(nCells_); cellCentres_.resize(size_t(nCells_) * 3); faceAreas_.resize(size_t(nFaces_) * 3); faceCentres_.resize(size_t(nFaces_) * 3);
    check(dsmcb200_download_geometry(ctx_, cellCentres_.data(), cellVolumes_.data(), faceCentres_.data(), faceAreas_.data(), nullptr), "dsmcb200_download_geometry");
    setCellFields();
    if (initialise_) { initialiseFromDict(); return; }
    readCloud();
    // counter-based RNG streams are keyed by the global time index, so a restarted run does not replay the streams of the first one
    startIndex_ = deltaT_ > 0 ? int64_t(std::llround(startTime_ / deltaT_)) : 0;
    check(dsmcb200_set_step(ctx_, uint32_t(startIndex_)), "dsmcb200_set_step");
    readResumeSampling();
}

// dsmcCloud::nParticles(cell) and deltaTValue(cell): the time-step model's nParticles_ / deltaT_ (dsmcVariableTimeStepModel.C:48-100) and the
// radial weighting factors of dsmcAxisymmetric (radial weighting method "cell", dsmcAxisymmetric.C:236-275, 447-456)
void dsmcCloud::setCellFields() {
    nPtsCell_.assign(size_t(nCells_), models_.nEquivalentParticles);
    dtCell_.assign(size_t(nCells_), deltaT_);
    rwfCell_.assign(size_t(nCells_), 1.0);
    if (variableTimeStep_ && nCells_ > 0) {
        double minVolume = cellVolumes_[0];
        for (double v : cellVolumes_) minVolume = std::min(minVolume, v);
        // findRefCell (dsmcVariableTimeStepModel.C:50-72): the reference cell is the smallest of the whole mesh.  On the ranks that do not
        // own it the reference indexes cell -1; what it means there is what the owning rank computes -- nParticles and deltaT start uniform,
        // so nParticleRef and the nParticle / time-step ratio are the same numbers on every rank and only the volume has to travel
        double vRef = minVolume;
        if (nRanks_ > 1) { check(dsmcb200_allreduce_min(ctx_, &vRef, 1), "dsmcb200_allreduce_min"); }
        else {   // the first cell within SMALL of the minimum, and ITS volume (updatenParticles reads volumeCells[refCell_])
            for (int c = 0; c < nCells_; ++c) if (std::fabs(cellVolumes_[c] - minVolume) < SMALL) { vRef = cellVolumes_[c]; break; }
        }
        const double nParticleRef = models_.nEquivalentParticles;
        for (int c = 0; c < nCells_; ++c) nPtsCell_[c] = nParticleRef * cellVolumes_[c] / vRef;
        const double nParticleTimeStepRatio = nParticleRef / deltaT_;
        for (int c = 0; c < nCells_; ++c) dtCell_[c] = nPtsCell_[c] / nParticleTimeStepRatio;
        if (rank_ == 0) std::printf("Variable time-step model:\n- Initial time-step [sec]\t%g\n\n", dtCell_[0]);
    }
    if (models_.coordinateSystem == DSMCB200_COORD_AXISYMMETRIC) {
        double radialExtent = -VGREAT, lowest = VGREAT;
        for (int f = 0; f < nFaces_; ++f) { radialExtent = std::max(radialExtent, faceCentres_[3 * size_t(f) + polarAxis_]); lowest = std::min(lowest, faceCentres_[3 * size_t(f) + polarAxis_]); }
        if (nRanks_ > 1) {   // gMax / gMin over the decomposed mesh (dsmcAxisymmetric.C:447-456)
            double v[2] = {-radialExtent, lowest};
            check(dsmcb200_allreduce_min(ctx_, v, 2), "dsmcb200_allreduce_min");
            radialExtent = -v[0]; lowest = v[1];
        }
        if (!(radialExtent > 0)) radialExtent = -lowest;
        radialExtent_ = radialExtent;
        for (int c = 0; c < nCells_; ++c) {
            const double radius = std::fabs(cellCentres_[3 * size_t(c) + polarAxis_]);
            double RWF = 1.0;
            RWF += (maxRWF_ - 1.0) * radius / radialExtent;
            rwfCell_[c] = RWF;
        }
        if (rank_ == 0)
            std::printf("\nAxisymmetric simulation:\n- revolution axis label\t%d\n- polar axis label\t%d\n- angular coordinate label\t%d\n"
                        "- radial weighting method\tcell-based\n- radial extent\t%g\n- maximum radial weighting factor\t%g\n\n",
                        3 - polarAxis_ - models_.angularCoordinate, polarAxis_, models_.angularCoordinate, radialExtent, maxRWF_);
    }
    if (models_.coordinateSystem == DSMCB200_COORD_SPHERICAL) {
        // radialExtent = gMax |face centre - origin| (dsmcSpherical.C:355-368), RWF = 1 + (maxRWF - 1) (r / radialExtent)^2 (:262-274)
        auto dist = [&](const double* x) {
            return std::sqrt((x[0] - origin_[0]) * (x[0] - origin_[0]) + (x[1] - origin_[1]) * (x[1] - origin_[1]) + (x[2] - origin_[2]) * (x[2] - origin_[2]));
        };
        double radialExtent = 0.0;
        for (int f = 0; f < nFaces_; ++f) radialExtent = std::max(radialExtent, dist(&faceCentres_[3 * size_t(f)]));
        if (nRanks_ > 1) { double neg = -radialExtent; check(dsmcb200_allreduce_min(ctx_, &neg, 1), "dsmcb200_allreduce_min"); radialExtent = -neg; }   // gMax
        radialExtent_ = radialExtent;
        for (int c = 0; c < nCells_; ++c) {
            const double radius = dist(&cellCentres_[3 * size_t(c)]);
            double RWF = 1.0;
            RWF += (maxRWF_ - 1.0) * (radius / radialExtent) * (radius / radialExtent);
            rwfCell_[c] = RWF;
        }
        if (rank_ == 0)
            std::printf("\nSpherical simulation:\n- coordinate system origin\t(%g %g %g)\n- radial weighting method\tcell-based\n- radial extent\t%g\n"
                        "- maximum radial weighting factor\t%g\n\n", origin_[0], origin_[1], origin_[2], radialExtent, maxRWF_);
    }
    const bool axi = models_.coordinateSystem == DSMCB200_COORD_AXISYMMETRIC || models_.coordinateSystem == DSMCB200_COORD_SPHERICAL;
    if (variableTimeStep_ || axi)
        check(dsmcb200_set_cell_fields(ctx_, variableTimeStep_ ? nPtsCell_.data() : nullptr, variableTimeStep_ ? dtCell_.data() : nullptr,
                                       axi ? rwfCell_.data() : nullptr), "dsmcb200_set_cell_fields");
}

dsmcCloud::~dsmcCloud() {
    if (ctx_) dsmcb200_destroy(ctx_);
}

void dsmcCloud::readControl() {
    Dict c = foam::readDict(caseDir_ + "/system/controlDict");
    deltaT_ = c.scalar("deltaT");
    endTime_ = c.scalar("endTime");
    writeControl_ = c.wordOr("writeControl", "timeStep");
    writeInterval_ = c.scalarOr("writeInterval", 1.0);
    timePrecision_ = int(c.labelOr("timePrecision", 6));
    purgeWrite_ = int(c.labelOr("purgeWrite", 0));
    if (purgeWrite_ < 0) throw FoamError("invalid value for purgeWrite " + std::to_string(purgeWrite_) + ", should be >= 0, setting to 0\nin: " + caseDir_ + "/system/controlDict");
    nTerminalOutputs_ = int(c.labelOr("nTerminalOutputs", 1));  // dsmcCloud.C:612-615
    std::string startFrom = c.wordOr("startFrom", "latestTime");
    std::vector<std::pair<double, std::string>> times;
    for (auto& n : foam::listDir(root_)) {
        char* e = nullptr;
        double v = std::strtod(n.c_str(), &e);
        if (e && *e == '\0' && !n.empty() && foam::exists(root_ + "/" + n + "/lagrangian")) times.push_back({v, n});
    }
    // writePrecision (default 6 in OpenFOAM); dsmcInitialise+ forces 15 (dsmcInitialise+.C:73)
    foam::setWritePrecision(initialise_ ? 15 : int(c.labelOr("writePrecision", 6)));
    // writeFormat ascii | binary (Time::writeFormat_): clouds and fields are written in it; what is read says its format in its header
    {
        const std::string wf = c.wordOr("writeFormat", "ascii");
        if (wf != "ascii" && wf != "binary") throw FoamError("controlDict: writeFormat " + wf + " is not in enumeration: 2(ascii binary)");
        foam::setWriteBinary(wf == "binary");
    }
    if (initialise_) {
        // createTime.H: the start time of controlDict; a fresh case has no time directory with a cloud yet
        startTime_ = startFrom == "startTime" ? c.scalarOr("startTime", 0.0) : 0.0;
        if (startFrom != "startTime") {
            for (auto& n : foam::listDir(root_)) {
                char* e = nullptr;
                const double v = std::strtod(n.c_str(), &e);
                if (e && *e == '\0' && !n.empty() && std::isdigit(static_cast<unsigned char>(n[0])) && (startFrom == "latestTime" ? v > startTime_ : false)) startTime_ = v;
            }
        }
        timeName_ = foam::timeName(startTime_, timePrecision_);
        time_ = startTime_;
        return;
    }
    if (times.empty()) throw FoamError("no time directory with a lagrangian cloud under " + root_);
    std::sort(times.begin(), times.end());
    if (startFrom == "latestTime") { startTime_ = times.back().first; timeName_ = times.back().second; }
    else if (startFrom == "firstTime") { startTime_ = times.front().first; timeName_ = times.front().second; }
    else {
        startTime_ = c.scalarOr("startTime", 0.0);
        timeName_.clear();
        for (auto& t : times) if (std::fabs(t.first - startTime_) <= 1e-12 * std::max(1.0, std::fabs(startTime_))) timeName_ = t.second;
        if (timeName_.empty()) throw FoamError("start time directory not found under " + root_);
    }
    time_ = startTime_;
}

namespace {
uint64_t fnv1a(uint64_t h, const void* data, size_t bytes) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    if (h == 0) h = 1469598103934665603ULL;
    for (size_t i = 0; i < bytes; ++i) { h ^= p[i]; h *= 1099511628211ULL; }
    return h;
}
template <class T>
uint64_t fnv1a(uint64_t h, const std::vector<T>& v) { return fnv1a(h, v.data(), v.size() * sizeof(T)); }
}  // namespace

void dsmcCloud::readMesh() {
    const std::string pm = root_ + "/constant/polyMesh/";
    points_ = foam::readVectorField(pm + "points");
    foam::readFaces(pm + "faces", faceOffsets_, facePoints_);
    owner_ = foam::readLabelField(pm + "owner");
    neighbour_ = foam::readLabelField(pm + "neighbour");
    boundary_ = foam::readBoundary(pm + "boundary");
    meshHash_ = fnv1a(fnv1a(fnv1a(fnv1a(fnv1a(0, points_), faceOffsets_), facePoints_), owner_), neighbour_);
    nFaces_ = int(owner_.size());
    nInternal_ = int(neighbour_.size());
    nCells_ = 0;
    for (int32_t c : owner_) nCells_ = std::max(nCells_, c + 1);
    for (int32_t c : neighbour_) nCells_ = std::max(nCells_, c + 1);
    patches_.clear();
    for (auto& b : boundary_) {
        dsmcb200_patch p{};
        copyName(p.name, b.name);
        p.type = patchTypeFromWord(b.type);
        p.start = b.startFace; p.size = b.nFaces; p.neighbPatch = -1; p.referPatch = -1;
        p.myProcNo = b.myProcNo; p.neighbProcNo = b.neighbProcNo;
        p.hasSeparation = b.hasSeparation ? 1 : 0;
        for (int k = 0; k < 3; ++k) p.separation[k] = b.separation[k];
        patches_.push_back(p);
    }
    for (size_t i = 0; i < boundary_.size(); ++i) {
        auto find = [&](const std::string& n) { for (size_t k = 0; k < boundary_.size(); ++k) if (boundary_[k].name == n) return int(k); return -1; };
        if (!boundary_[i].neighbourPatch.empty()) patches_[i].neighbPatch = find(boundary_[i].neighbourPatch);
        if (!boundary_[i].referPatch.empty()) patches_[i].referPatch = find(boundary_[i].referPatch);
        if (patches_[i].type == DSMCB200_PATCH_CYCLIC && patches_[i].neighbPatch < 0)
            throw FoamError("cyclic patch " + boundary_[i].name + " has no neighbourPatch entry");
    }
}

void dsmcCloud::readProperties() {
    Dict d = foam::readDict(caseDir_ + "/constant/dsmcProperties");
    std::memset(&models_, 0, sizeof(models_));
    models_.nEquivalentParticles = d.scalar("nEquivalentParticles");
    models_.deltaT = deltaT_;
    const std::string bcm = d.word("BinaryCollisionModel");
    models_.collisionModel = selectBinaryCollisionModel(bcm);
    selectCollisionPartnerSelection(d.word("collisionPartnerSelectionModel"));
    models_.coordinateSystem = selectCoordinateSystem(d.wordOr("coordinateSystem", "dsmcCartesian"));
    variableTimeStep_ = selectTimeStepModel(d.wordOr("timeStepModel", "constant"));
    if (d.boolOr("nEquivalentParticlesFromFile", false))
        throw FoamError("nEquivalentParticlesFromFile: reading the nParticles field of a previous run is not supported; the variable time-step model sets it from the cell volumes");
    polarAxis_ = 1; models_.angularCoordinate = 2; maxRWF_ = 1.0;
    if (models_.coordinateSystem == DSMCB200_COORD_AXISYMMETRIC) {
        // dsmcAxisymmetric::checkCoordinateSystemInputs (dsmcAxisymmetric.C:337-470)
        const Dict& ax = d.subDict("axisymmetricProperties");
        const std::string method = ax.wordOr("radialWeightingMethod", "cell");
        if (method != "cell" && method != "particleAverage")
            throw FoamError("The radial weighting method is badly defined. Choices in constant/dsmcProperties are \"cell\" or \"particleAverage\". Please edit the entry: radialWeightingMethod.");
        if (method == "particleAverage")
            throw FoamError("radialWeightingMethod particleAverage is not supported (the sampled sums are weighted per cell after the run); use \"cell\"");
        const std::string rev = ax.wordOr("revolutionAxis", ""), pol = ax.wordOr("polarAxis", "");
        const char* bad = "Revolution and polar axes are badly defined in constant/dsmcProperties axisymmetricProperties{}";
        int polar = 1, ang = 2;
        if (rev == "z") {
            if (pol.empty()) { polar = 0; ang = 1; } else if (pol == "y") { polar = 1; ang = 0; } else if (pol == "x") { polar = 0; ang = 1; } else throw FoamError(bad);
        } else if (rev == "y") {
            if (pol.empty()) { polar = 2; ang = 0; } else if (pol == "x") { polar = 0; ang = 2; } else if (pol == "z") { polar = 2; ang = 0; } else throw FoamError(bad);
        } else if (rev == "x") {
            if (pol == "z") { polar = 2; ang = 1; } else if (pol != "y") throw FoamError(bad);
        }
        polarAxis_ = polar; models_.angularCoordinate = ang;
        maxRWF_ = ax.scalar("maxRadialWeightingFactor");
    }
    if (models_.coordinateSystem == DSMCB200_COORD_SPHERICAL) {
        // dsmcSpherical::checkCoordinateSystemInputs (dsmcSpherical.C:325-386)
        const Dict& sph = d.subDict("sphericalProperties");
        const std::string method = sph.wordOr("radialWeightingMethod", "cell");
        if (method != "cell" && method != "particleAverage")
            throw FoamError("The radial weighting method is badly defined. Choices in constant/dsmcProperties are \"cell\" or \"particleAverage\". Please edit the entry: radialWeightingMethod.");
        if (method == "particleAverage")
            throw FoamError("radialWeightingMethod particleAverage is not supported (the sampled sums are weighted per cell after the run); use \"cell\"");
        maxRWF_ = sph.scalar("maxRadialWeightingFactor");
        const std::vector<double> o = sph.found("origin") ? sph.vector3("origin") : std::vector<double>{0.0, 0.0, 0.0};
        for (int k = 0; k < 3; ++k) origin_[k] = o[k];
    }
    // VariableHardSphere reads Tref from VariableHardSphereCoeffs even under the LB model (VariableHardSphere.C:56-62)
    // (VariableSoftSphere.C:54-60 does the same with VariableSoftSphereCoeffs)
    const bool soft = models_.collisionModel == DSMCB200_COLL_VSS || models_.collisionModel == DSMCB200_COLL_LB_VSS;
    const std::string baseCoeffs = soft ? "VariableSoftSphereCoeffs" : "VariableHardSphereCoeffs";
    const std::string lbCoeffs = soft ? "LarsenBorgnakkeVariableSoftSphereCoeffs" : "LarsenBorgnakkeVariableHardSphereCoeffs";
    models_.Tref = d.isDict(baseCoeffs) ? d.subDict(baseCoeffs).scalarOr("Tref", 273.0) : 273.0;
    models_.rotationalRelaxationCollisionNumber = 5.0;
    models_.vibrationalRelaxationCollisionNumber = 0.0;
    models_.electronicRelaxationCollisionNumber = 500.0;
    models_.invZvFormulation = 2;
    if (d.isDict(lbCoeffs)) {
        const Dict& lb = d.subDict(lbCoeffs);
        models_.rotationalRelaxationCollisionNumber = lb.scalarOr("rotationalRelaxationCollisionNumber", 5.0);
        models_.vibrationalRelaxationCollisionNumber = lb.scalarOr("vibrationalRelaxationCollisionNumber", 0.0);
        models_.electronicRelaxationCollisionNumber = lb.scalarOr("electronicRelaxationCollisionNumber", 500.0);
        const std::string v = lb.wordOr("inverseZvFormulation", "");
        models_.invZvFormulation = v == "pre-2008" ? 0 : (v == "2008" ? 1 : 2);
    }
    // dsmcCloud.C:639-646: seedNumber or clock + 7183*rank
    models_.seed = d.found("seedNumber") ? uint64_t(d.label("seedNumber")) + 7183ull * uint64_t(rank_)
                                         : uint64_t(std::time(nullptr)) + 7183ull * uint64_t(rank_);
    typeIdList_ = d.wordList("typeIdList");
    if (typeIdList_.empty()) throw FoamError("typeIdList is empty in " + d.name);
    const Dict& mp = d.subDict("moleculeProperties");
    species_.clear();
    maxModes_ = 1;
    for (auto& id : typeIdList_) {
        const Dict& s = mp.subDict(id);
        dsmcb200_species sp{};
        copyName(sp.name, id);
        sp.mass = s.scalar("mass"); sp.diameter = s.scalar("diameter"); sp.omega = s.scalar("omega"); sp.alpha = s.scalarOr("alpha", 1.0);
        sp.rotationalDegreesOfFreedom = s.scalarOr("rotationalDegreesOfFreedom", 0.0);
        sp.nVibrationalModes = int32_t(s.labelOr("nVibrationalModes", 0));
        auto thetaV = s.scalarListOr("characteristicVibrationalTemperature", {});
        auto Zref = s.scalarListOr("Zref", {});
        auto TrefZv = s.scalarListOr("referenceTempForZref", {});
        if (int(thetaV.size()) != sp.nVibrationalModes)
            throw FoamError("Number of characteristic vibrational temperatures is " + std::to_string(thetaV.size()) + ", instead of " + std::to_string(sp.nVibrationalModes));
        if (int(Zref.size()) != sp.nVibrationalModes)
            throw FoamError("Number of reference vibrational relaxation numbers is" + std::to_string(Zref.size()) + ", instead of " + std::to_string(sp.nVibrationalModes));
        if (int(TrefZv.size()) != sp.nVibrationalModes)
            throw FoamError("Number of reference temperature for vibrational relaxation is" + std::to_string(TrefZv.size()) + ", instead of " + std::to_string(sp.nVibrationalModes));
        if (sp.nVibrationalModes > DSMCB200_MAX_VIB_MODES) throw FoamError("species " + id + ": more than 3 vibrational modes are not supported");
        for (int m = 0; m < sp.nVibrationalModes; ++m) { sp.thetaV[m] = thetaV[m]; sp.Zref[m] = Zref[m]; sp.TrefZv[m] = TrefZv[m]; }
        sp.thetaD = s.scalarOr("dissociationTemperature", 0.0);
        sp.charge = int32_t(s.labelOr("charge", 0));
        if (sp.charge < -1 || sp.charge > 1) throw FoamError("Charge value should be 0 for neutrals, 1 for ions, or -1 for electrons, instead of " + std::to_string(sp.charge));
        sp.nElectronicLevels = int32_t(s.labelOr("nElectronicLevels", 1));
        auto ee = s.scalarListOr("electronicEnergyList", {0.0});
        auto eg = s.labelListOr("electronicDegeneracyList", {1});
        if (int(eg.size()) != sp.nElectronicLevels) throw FoamError("Number of degeneracy levels should be " + std::to_string(sp.nElectronicLevels) + ", instead of " + std::to_string(eg.size()));
        if (int(ee.size()) != sp.nElectronicLevels) throw FoamError("Number of electronic energy levels should be " + std::to_string(sp.nElectronicLevels) + ", instead of " + std::to_string(ee.size()));
        if (sp.nElectronicLevels > DSMCB200_MAX_ELEC_LEVELS) throw FoamError("species " + id + ": more than 16 electronic levels are not supported");
        for (int l = 0; l < sp.nElectronicLevels; ++l) { sp.electronicEnergyList[l] = ee[l]; sp.electronicDegeneracyList[l] = int32_t(eg[l]); }
        maxModes_ = std::max(maxModes_, int(sp.nVibrationalModes));
        species_.push_back(sp);
    }
    readReactions();
}

// system/chemReactDict -> dsmcb200_reaction list (dsmcReactions ctor, dsmcReactions.C:69-118; dsmcReaction::setProperties, dsmcReaction.C:79-120;
// the dissociationQKProperties / exchangeQKProperties sub-dictionaries of the three quantum-kinetic models)
void dsmcCloud::readReactions() {
    reactions_.clear(); reactionNames_.clear();
    if (!foam::exists(caseDir_ + "/system/chemReactDict")) return;
    Dict cr = foam::readDict(caseDir_ + "/system/chemReactDict");
    const bool master = rank_ == 0;
    if (master) std::printf("\nCreating dsmcReactions\n\n");
    auto typeId = [&](const std::string& reaction, const std::string& n) {
        for (size_t k = 0; k < typeIdList_.size(); ++k) if (typeIdList_[k] == n) return int32_t(k);
        throw FoamError("For reaction named " + reaction + "\nCannot find type id: " + n);
    };
    for (auto& e : cr.dictList("reactions")) {
        const Dict& r = *e.second;
        const std::string name = e.first, model = r.word("reactionModel");
        if (master) std::printf("Selecting the reaction model %s\n", model.c_str());
        dsmcb200_reaction R{};
        if (model == "dissociationQK") R.model = DSMCB200_REACT_DISSOCIATION_QK;
        else if (model == "exchangeQK") R.model = DSMCB200_REACT_EXCHANGE_QK;
        else if (model == "dissociationExchangeQK") R.model = DSMCB200_REACT_DISSOCIATION_EXCHANGE_QK;
        else throw FoamError("dsmcReaction::New(const dictionary&) : \n    unknown dsmc reaction model type " + model +
                             ", constructor not in hash table\n\n    Valid reaction types are :\n3(dissociationQK exchangeQK dissociationExchangeQK)");
        const auto reactants = r.wordList("reactants");
        if (reactants.size() != 2) throw FoamError("For reaction named " + name + "\nThere should be two reactants, instead of " + std::to_string(reactants.size()));
        for (int k = 0; k < 2; ++k) R.reactants[k] = typeId(name, reactants[k]);
        R.allowSplitting = r.boolOr("allowSplitting", true) ? 1 : 0;
        for (int k = 0; k < 2; ++k) R.dissociationProducts[k][0] = R.dissociationProducts[k][1] = -1;
        R.exchangeProducts[0] = R.exchangeProducts[1] = -1;
        if (R.model != DSMCB200_REACT_EXCHANGE_QK) {
            const auto& st = r.subDict("dissociationQKProperties").stream("dissociationProducts");
            if (st.empty() || st[0].kind != foam::Node::LIST) throw FoamError("For reaction named " + name + "\ndissociationProducts must be a list of two word lists");
            const auto& lists = st[0].list;
            if (lists.size() != 2) throw FoamError("For reaction named " + name + "\nThere should be two lists of products, instead of " + std::to_string(lists.size()) + "\nNB: a list can be left empty");
            for (int k = 0; k < 2; ++k) {
                if (lists[k].kind != foam::Node::LIST) throw FoamError("For reaction named " + name + "\ndissociationProducts must be a list of two word lists");
                const auto& prods = lists[k].list;
                if (!prods.empty() && prods.size() != 2)
                    throw FoamError("For reaction named " + name + "\nThere should be 2 dissociation products for molecule " + reactants[k] + " instead of " + std::to_string(prods.size()));
                for (size_t q = 0; q < prods.size(); ++q) R.dissociationProducts[k][q] = typeId(name, prods[q].word);
            }
        }
        if (R.model != DSMCB200_REACT_DISSOCIATION_QK) {
            const Dict& x = r.subDict("exchangeQKProperties");
            const auto prods = x.wordList("exchangeProducts");
            if (prods.size() != 2) throw FoamError("For reaction named " + name + "\nThere should be two products, instead of " + std::to_string(prods.size()));
            for (int k = 0; k < 2; ++k) R.exchangeProducts[k] = typeId(name, prods[k]);
            R.heatOfReactionExchange = x.scalar("heatOfReactionExchange");
            R.aCoeff = x.scalar("aCoeff"); R.bCoeff = x.scalar("bCoeff");
        }
        reactions_.push_back(R);
        reactionNames_.push_back(name);
    }
    if (master) {
        if (!reactions_.empty()) std::printf("Number of reactions created: %d\n", int(reactions_.size()));
        else std::printf("There are no chemical reactions defined.\n");
    }
}

void dsmcCloud::readBoundaries() {
    patchModels_.clear(); inflows_.clear();
    const std::string path = caseDir_ + "/system/boundariesDict";
    if (!foam::exists(path)) return;
    Dict d = foam::readDict(path);
    auto patchId = [&](const std::string& n) {
        for (size_t k = 0; k < boundary_.size(); ++k) if (boundary_[k].name == n) return int(k);
        return -1;
    };
    auto typeId = [&](const std::string& n) {
        for (size_t k = 0; k < typeIdList_.size(); ++k) if (typeIdList_[k] == n) return int(k);
        return -1;
    };
    for (auto& e : d.dictList("dsmcPatchBoundaries")) {
        const Dict& b = *e.second;
        const std::string patchName = b.subDict("patchBoundaryProperties").word("patchName");
        const std::string model = b.word("boundaryModel");
        const int kind = selectPatchBoundaryModel(model);
        const int pid = patchId(patchName);
        if (pid < 0) {
            if (nRanks_ > 1) continue;  // the patch has no faces on this processor
            throw FoamError("Cannot find patch: " + patchName + "\nin: " + path);
        }
        dsmcb200_patch_model pm{};
        pm.patch = pid; pm.model = kind;
        if (kind == DSMCB200_BND_DIFFUSE_WALL || kind == DSMCB200_BND_DIFFUSE_SPECULAR_WALL) {
            const Dict& pr = b.subDict(model + "Properties");
            if (kind == DSMCB200_BND_DIFFUSE_SPECULAR_WALL) pm.diffuseFraction = pr.scalar("diffuseFraction");
            pm.temperature = pr.found("groundLevelTemperature") ? pr.scalar("groundLevelTemperature") : pr.scalar("temperature");
            if (pr.found("formationLevelTemperature")) {   // linear T(depth), dsmcDiffuseWallPatch.C:62-63,141-148,169-179
                pm.linearTemperature = 1;
                pm.formationLevelTemperature = pr.scalar("formationLevelTemperature");
                const std::string ax = pr.wordOr("depthAxis", "y");
                pm.depthAxis = ax == "x" ? 0 : (ax == "z" ? 2 : 1);
            }
            auto v = pr.vector3("velocity");
            for (int k = 0; k < 3; ++k) pm.velocity[k] = v[k];
        }
        if (kind == DSMCB200_BND_CLL_WALL) {   // dsmcCLLWallPatch.C:45-75,330-334: every keyword is mandatory
            const Dict& pr = b.subDict(model + "Properties");
            pm.normalAccommodationCoefficient = pr.scalar("normalAccommodationCoefficient");
            pm.tangentialAccommodationCoefficient = pr.scalar("tangentialAccommodationCoefficient");
            pm.rotationalEnergyAccommodationCoefficient = pr.scalar("rotationalEnergyAccommodationCoefficient");
            (void)pr.scalar("vibrationalEnergyAccommodationCoefficient");   // read by the constructor, used nowhere (the vibrational part is commented out)
            pm.temperature = pr.scalar("temperature");
            auto v = pr.vector3("velocity");
            for (int k = 0; k < 3; ++k) pm.velocity[k] = v[k];
        }
        patchModels_.push_back(pm);
    }
    if (!d.dictList("dsmcCyclicBoundaries").empty())
        throw FoamError("dsmcCyclicBoundaries models are outside the scoped path (only an empty list is supported)");
    for (auto& e : d.dictList("dsmcGeneralBoundaries")) {
        const Dict& b = *e.second;
        const std::string patchName = b.subDict("generalBoundaryProperties").word("patchName");
        const std::string model = b.word("boundaryModel");
        selectGeneralBoundaryModel(model);
        const int pid = patchId(patchName);
        if (pid < 0) {
            if (nRanks_ > 1) continue;
            throw FoamError("Cannot find patch: " + patchName + "\nin: " + path);
        }
        const Dict& pr = b.subDict(model + "Properties");
        dsmcb200_inflow in{};
        in.patch = pid;
        std::vector<std::string> mols;
        for (auto& w : pr.wordList("typeIds")) if (std::find(mols.begin(), mols.end(), w) == mols.end()) mols.push_back(w);
        if (mols.empty()) throw FoamError("Cannot have zero typeIds being inserted.\nin: " + path);
        const Dict& nd = pr.subDict("numberDensities");
        in.nTypes = int32_t(mols.size());
        for (size_t k = 0; k < mols.size(); ++k) {
            const int t = typeId(mols[k]);
            if (t < 0) throw FoamError("Cannot find typeId: " + mols[k] + "\nin: " + path);
            in.typeIds[k] = t;
            in.numberDensities[k] = nd.scalar(mols[k]);
        }
        auto v = pr.vector3("velocity");
        for (int k = 0; k < 3; ++k) in.velocity[k] = v[k];
        in.translationalTemperature = pr.scalar("translationalTemperature");
        in.rotationalTemperature = pr.scalarOr("rotationalTemperature", 0.0);
        in.vibrationalTemperature = pr.scalarOr("vibrationalTemperature", 0.0);
        in.electronicTemperature = pr.scalarOr("electronicTemperature", 0.0);
        inflows_.push_back(in);
    }
}

void dsmcCloud::readFieldProperties() {
    fields_.clear();
    const std::string path = caseDir_ + "/system/fieldPropertiesDict";
    if (!foam::exists(path)) return;
    Dict d = foam::readDict(path);
    for (auto& e : d.dictList("dsmcFields")) {
        const Dict& f = *e.second;
        const std::string model = f.word("fieldModel");
        selectFieldModel(model);
        const Dict& pr = f.subDict(model + "Properties");
        FieldSpec s;
        s.fieldName = pr.word("fieldName");
        for (auto& w : pr.wordList("typeIds")) {
            int t = -1;
            for (size_t k = 0; k < typeIdList_.size(); ++k) if (typeIdList_[k] == w) t = int(k);
            if (t < 0) throw FoamError("Cannot find typeId: " + w + "\nin: " + path);
            if (std::find(s.typeIds.begin(), s.typeIds.end(), t) == s.typeIds.end()) s.typeIds.push_back(t);
        }
        s.measureMeanFreePath = pr.boolOr("measureMeanFreePath", false);
        s.densityOnly = pr.boolOr("densityOnly", false);
        s.measureHeatFluxShearStress = pr.boolOr("measureHeatFluxShearStress", false);
        s.measureClassifications = pr.boolOr("measureClassifications", false);
        s.measureErrors = pr.boolOr("measureErrors", false);
        s.mfpReferenceTemperature = pr.scalarOr("mfpReferenceTemperature", 273.0);
        s.sampleInterval = int(pr.labelOr("sampleInterval", 1));
        s.averagingAcrossManyRuns = pr.boolOr("averagingAcrossManyRuns", false);
        if (f.isDict("timeProperties")) {
            const Dict& tp = f.subDict("timeProperties");
            s.resetAtOutput = tp.boolOr("resetAtOutput", true);
            s.resetAtOutputUntilTime = tp.scalarOr("resetAtOutputUntilTime", 1e300);
        }
        if (s.measureHeatFluxShearStress) models_.measureHeatFluxShearStress = 1;
        if (s.measureClassifications) models_.measureClassifications = 1;
        // every field{} samples on its own cadence (dsmcVolFields.C:1073-1081,1362: calculateField samples when sampleInterval_ <=
        // ++sampleCounter_): fields with the same sampleInterval share one set of the engine's sums (dsmcb200_set_sample_sets)
        {
            const int32_t iv = std::max(1, s.sampleInterval);
            size_t k = 0;
            while (k < sampleSets_.size() && sampleSets_[k] != iv) ++k;
            if (k == sampleSets_.size()) {
                if (sampleSets_.size() == 8) throw FoamError("dsmcVolFields " + s.fieldName + ": more than 8 different sampleIntervals\nin: " + path);
                sampleSets_.push_back(iv);
            }
            s.set = int(k);
        }
        // the reset policy of timeProperties (dsmcField.C:113-152) is per field: a field that stopped resetting keeps averaging from its
        // own baseline of the shared accumulators (write())
        models_.sampleInterval = sampleSets_[0];
        fields_.push_back(s);
    }
}

// dsmcInitialise+ (applications/utilities/preProcessing/dsmc/dsmcInitialise+/dsmcInitialise+.C:37-90): the cloud constructor that
// clears the field of parcels and runs every `configuration` of system/dsmcInitialiseDict (dsmcCloud.C:697-809,
// dsmcAllConfigurations.C:45-77).  dsmcMeshFill (initialiseDsmcParcels/derived/dsmcMeshFill/dsmcMeshFill.C:70-240) is executed by
// dsmcb200_mesh_fill on the device.
void dsmcCloud::initialiseFromDict() {
    const std::string path = caseDir_ + "/system/dsmcInitialiseDict";
    Dict d = foam::readDict(path);
    auto confs = d.dictList("configurations");
    const bool master = rank_ == 0;
    if (master) std::printf("clearing existing field of parcels \n\nCreating dsmc configurations: \n\n");
    if (confs.size() > 1)
        throw FoamError("dsmcInitialiseDict: " + std::to_string(confs.size()) + " configurations; this engine fills the mesh from one configuration\nin: " + path);
    int64_t added = 0;
    for (auto& e : confs) {
        const Dict& cf = *e.second;
        const std::string type = cf.word("type");
        if (type != "dsmcMeshFill")
            throw FoamError("dsmcConfiguration::New(const dictionary&) : \n    unknown dsmcConfiguration type " + type +
                            ", constructor not in hash table\n\n    Valid  types are :\n(dsmcMeshFill)");
        if (master) std::printf("\nInitialising particles\n");
        const double Ttra = cf.scalar("translationalTemperature"), Trot = cf.scalar("rotationalTemperature");
        const double Tvib = cf.scalar("vibrationalTemperature"), Telec = cf.scalar("electronicTemperature");
        const std::vector<double> vel = cf.vector3("velocity");
        const Dict& nd = cf.subDict("numberDensities");
        std::vector<int32_t> ids;
        std::vector<double> dens;
        for (auto& mol : nd.toc()) {
            int t = -1;
            for (size_t k = 0; k < typeIdList_.size(); ++k) if (typeIdList_[k] == mol) t = int(k);
            if (t < 0) throw FoamError("Cannot find typeId: " + mol + "\nin: " + path);   // dsmcMeshFill.C:131-137 shape
            ids.push_back(t);
            dens.push_back(nd.scalar(mol));
        }
        if (ids.empty()) throw FoamError("numberDensities is empty in " + path);
        if (dryRun_) {
            std::printf("configuration dsmcMeshFill: Ttra %g Trot %g Tvib %g Telec %g velocity (%g %g %g)\n", Ttra, Trot, Tvib, Telec, vel[0], vel[1], vel[2]);
            for (size_t k = 0; k < ids.size(); ++k) std::printf("  numberDensity %s %g\n", typeIdList_[ids[k]].c_str(), dens[k]);
            continue;
        }
        check(dsmcb200_mesh_fill(ctx_, int(ids.size()), ids.data(), dens.data(), Ttra, Trot, Tvib, Telec, vel.data()), "dsmcb200_mesh_fill");
        added = nParcels();
    }
    if (dryRun_) return;
    double tot[1] = {double(added)};
    if (nRanks_ > 1) check(dsmcb200_allreduce_sum(ctx_, tot, 1), "dsmcb200_allreduce_sum");
    if (master)
        std::printf("\nInitial no. of parcels: 0 added parcels: %lld, total no. of parcels: %lld\n", (long long)tot[0], (long long)tot[0]);
}

void dsmcCloud::readCloud() {
    const std::string dir = root_ + "/" + timeName_ + "/lagrangian/" + cloudName_ + "/";
    std::vector<double> xyz, U, ERot;
    std::vector<int32_t> cell, typeId, vib, elevel, cls, origId;
    foam::readPositions(dir + "positions", xyz, cell);
    const int64_t n = int64_t(cell.size());
    U = foam::readVectorField(dir + "U");
    typeId = foam::readLabelField(dir + "typeId");
    if (int64_t(U.size()) != 3 * n || int64_t(typeId.size()) != n) throw FoamError("cloud files under " + dir + " have inconsistent sizes");
    dsmcb200_parcels_soa s{};
    s.position = xyz.data(); s.U = U.data(); s.cell = cell.data(); s.typeId = typeId.data();
    s.maxModes = maxModes_;
    if (foam::exists(dir + "ERot")) { ERot = foam::readScalarField(dir + "ERot"); if (int64_t(ERot.size()) == n) s.ERot = ERot.data(); }
    if (foam::exists(dir + "vibLevel")) {
        int w = 0;
        vib = foam::readLabelListList(dir + "vibLevel", w);
        if (w > 0) { s.vibLevel = vib.data(); s.maxModes = w; }
    }
    if (foam::exists(dir + "ELevel")) { elevel = foam::readLabelField(dir + "ELevel"); if (int64_t(elevel.size()) == n) s.ELevel = elevel.data(); }
    if (foam::exists(dir + "classification")) { cls = foam::readLabelField(dir + "classification"); if (int64_t(cls.size()) == n) s.classification = cls.data(); }
    if (foam::exists(dir + "origId")) { origId = foam::readLabelField(dir + "origId"); if (int64_t(origId.size()) == n) s.origId = origId.data(); }
    std::vector<double> radialWeight;   // dsmcParcel::RWF_ (dsmcParcelIO.C); without the file a parcel takes its cell's weight
    if (foam::exists(dir + "radialWeight")) { radialWeight = foam::readScalarField(dir + "radialWeight"); if (int64_t(radialWeight.size()) == n) s.radialWeight = radialWeight.data(); }
    nRead_ = n;
    cloudHash_ = fnv1a(fnv1a(fnv1a(fnv1a(fnv1a(fnv1a(fnv1a(0, xyz), cell), U), typeId), ERot), vib), elevel);
    // <time>/dsmcSigmaTcRMax is MUST_READ (dsmcCloud.C:625-636)
    auto sig = foam::readInternalField(root_ + "/" + timeName_ + "/dsmcSigmaTcRMax", nCells_, 1);
    if (dryRun_) return;
    check(dsmcb200_upload_parcels(ctx_, n, &s), "dsmcb200_upload_parcels");
    check(dsmcb200_upload_cellstate(ctx_, sig.data(), nullptr), "dsmcb200_upload_cellstate");
}

std::string dsmcCloud::summary() const {
    std::ostringstream o;
    o << "case " << caseDir_ << "\n  startTime " << timeName_ << " deltaT " << deltaT_ << " endTime " << endTime_ << " writeControl " << writeControl_
      << " writeInterval " << writeInterval_ << " nTerminalOutputs " << nTerminalOutputs_ << " purgeWrite " << purgeWrite_ << "\n  mesh: " << points_.size() / 3 << " points "
      << nFaces_ << " faces " << nInternal_ << " internal " << nCells_ << " cells " << boundary_.size() << " patches\n";
    for (auto& b : boundary_) o << "    patch " << b.name << " " << b.type << " " << b.nFaces << " @" << b.startFace << "\n";
    o << "  species:";
    for (auto& t : typeIdList_) o << " " << t;
    o << "\n  collisionModel " << models_.collisionModel << " invZv " << models_.invZvFormulation << " nEquivalentParticles " << models_.nEquivalentParticles
      << " seed " << models_.seed << "\n  coordinateSystem " << (models_.coordinateSystem == DSMCB200_COORD_AXISYMMETRIC ? "dsmcAxisymmetric" : (models_.coordinateSystem == DSMCB200_COORD_SPHERICAL ? "dsmcSpherical" : "dsmcCartesian"))
      << " polarAxis " << polarAxis_ << " angularCoordinate " << models_.angularCoordinate << " maxRadialWeightingFactor " << maxRWF_
      << " timeStepModel " << (variableTimeStep_ ? "variable" : "constant") << "\n  patchModels " << patchModels_.size() << " inflows " << inflows_.size() << " fields " << fields_.size() << "\n";
    for (auto& pm : patchModels_)
        if (pm.model == DSMCB200_BND_CLL_WALL)
            o << "    patchModel " << boundary_[pm.patch].name << " dsmcCLLWallPatch temperature " << pm.temperature << " accommodation normal "
              << pm.normalAccommodationCoefficient << " tangential " << pm.tangentialAccommodationCoefficient << " rotational "
              << pm.rotationalEnergyAccommodationCoefficient << "\n";
    for (auto& pm : patchModels_)
        if (pm.linearTemperature)
            o << "    patchModel " << boundary_[pm.patch].name << " temperature " << pm.temperature << " linearTemperature formationLevel "
              << pm.formationLevelTemperature << " depthAxis " << pm.depthAxis << "\n";
    for (auto& f : fields_) {
        o << "    field " << f.fieldName << " typeIds";
        for (int t : f.typeIds) o << " " << t;
        o << " mfp " << f.measureMeanFreePath << " reset " << f.resetAtOutput << " sampleInterval " << f.sampleInterval << "\n";
    }
    o << "  parcels " << nRead_ << "\n";
    o << "  checksum mesh " << std::hex << meshHash_ << " cloud " << cloudHash_ << std::dec << "\n";
    return o.str();
}

bool dsmcCloud::loop() {
    if (!(time_ < endTime_ - 0.5 * deltaT_)) return false;
    ++timeIndex_;
    time_ = startTime_ + double(timeIndex_) * deltaT_;
    timeName_ = foam::timeName(time_, timePrecision_);
    return true;
}

bool dsmcCloud::outputTime() const {
    if (writeControl_ == "timeStep") return timeIndex_ % std::max<int64_t>(1, int64_t(std::llround(writeInterval_))) == 0;
    // runTime / adjustableRunTime: Time::adjustDeltaT-free form of Time::operator++
    const int64_t now = int64_t(((time_ - startTime_) + 0.5 * deltaT_) / writeInterval_);
    const int64_t before = int64_t(((time_ - deltaT_ - startTime_) + 0.5 * deltaT_) / writeInterval_);
    return now > before;
}

int64_t dsmcCloud::nParcels() {
    int64_t n = 0;
    check(dsmcb200_download_parcels(ctx_, 0, &n, nullptr), "dsmcb200_download_parcels");
    return n;
}

void dsmcCloud::evolve() {
    check(dsmcb200_evolve(ctx_, 1), "dsmcb200_evolve");
    ++infoCounter_;
    if (infoCounter_ >= nTerminalOutputs_) {
        dsmcb200_counters c{};
        check(dsmcb200_get_counters(ctx_, &c), "dsmcb200_get_counters");
        double v[2] = {double(c.collisions), double(c.collisionCandidates)};
        if (nRanks_ > 1) check(dsmcb200_allreduce_sum(ctx_, v, 2), "dsmcb200_allreduce_sum");
        if (rank_ == 0) {
            // noTimeCounter.C:320-337
            if (v[1] > 0) std::printf("    Collisions                      = %lld\n\n", (long long)v[0]);
            else std::printf("    No collisions\n");
        }
        infoCounter_ = 0;
    }
}

void dsmcCloud::info() {
    dsmcb200_counters c{};
    check(dsmcb200_get_counters(ctx_, &c), "dsmcb200_get_counters");
    double v[6] = {double(c.nParcels), c.mass, c.linearKineticEnergy, c.rotationalEnergy, c.vibrationalEnergy, c.electronicEnergy};
    if (nRanks_ > 1) check(dsmcb200_allreduce_sum(ctx_, v, 6), "dsmcb200_allreduce_sum");
    if (rank_ != 0) return;
    const double nP = models_.nEquivalentParticles;   // nParticlesOrg (dsmcCloud.C:947)
    const double nMol = c.nMolecules > 0 && nRanks_ == 1 ? c.nMolecules : v[0] * nP;   // infoMeasurements[6]
    // dsmcCloud.C:960-981
    std::printf("    Number of DSMC particles        = %lld\n", (long long)v[0]);
    if (v[0] > 0) {
        std::printf("    Number of stuck particles       = %g\n", 0.0);
        std::printf("    Number of free particles        = %g\n", nMol / nP);
        std::printf("    Average linear kinetic energy   = %g\n", v[2] / nMol);
        std::printf("    Average rotational energy       = %g\n", v[3] / nMol);
        std::printf("    Average vibrational energy      = %g\n", v[4] / nMol);
        std::printf("    Average electronic energy       = %g\n", v[5] / nMol);
        std::printf("    Total energy                    = %g\n", v[2] + v[3] + v[4] + v[5]);
    }
    std::fflush(stdout);
}

void dsmcCloud::selectSet(int set) {
    if (sampleSets_.size() > 1) check(dsmcb200_select_sample_set(ctx_, set), "dsmcb200_select_sample_set");
}

// dsmcVolFields::calculateField reductions (dsmcVolFields.C:1242-1290, 1401-1508, 1663-1790) from the per-species
// moment sums accumulated on the device.
DerivedFields dsmcCloud::calculateField(const FieldSpec& f) {
    selectSet(f.set);
    dsmcb200_accum_info ai{};
    check(dsmcb200_accum_info_get(ctx_, &ai), "dsmcb200_accum_info_get");
    const int S = ai.nSpecies, nQ = ai.nQuantities, nC = ai.nCells;
    std::vector<double> acc(size_t(nC) * S * nQ), coll(size_t(nC) * 2);
    check(dsmcb200_download_accumulators(ctx_, acc.data(), coll.data()), "dsmcb200_download_accumulators");
    // this field's sums since ITS last reset: the shared accumulators minus the field's baseline (dsmcField.C:113-152)
    if (f.baseAcc.size() == acc.size()) for (size_t k = 0; k < acc.size(); ++k) acc[k] -= f.baseAcc[k];
    if (f.baseColl.size() == coll.size()) for (size_t k = 0; k < coll.size(); ++k) coll[k] -= f.baseColl[k];
    const bool internal = ai.nModes >= 0;
    const double nTf = ai.nTimeSteps - f.baseNT;
    const double nT = nTf > 0 ? nTf : 1.0;
    const double kB = models_.kB > 0 ? models_.kB : 1.38065e-23;
    DerivedFields o;
    auto z = [&](std::vector<double>& v, int w = 1) { v.assign(size_t(nC) * w, 0.0); };
    z(o.dsmcNMean); z(o.rhoN); z(o.rhoM); z(o.p); z(o.Ttra); z(o.Trot); z(o.Tvib); z(o.Tov); z(o.Ma); z(o.mfp); z(o.mct); z(o.mctToDt);
    z(o.mfpToDx); z(o.SOF); z(o.measuredCollisionRate); z(o.UMean, 3);
    if (f.measureHeatFluxShearStress) { z(o.pressureTensor, 9); z(o.shearStressTensor, 9); z(o.heatFluxVector, 3); }
    if (f.measureErrors) { z(o.rhoMError); z(o.UError); z(o.TError); z(o.pError); }
    const double NAvo = 6.02214e26;  // OpenFOAM SI physicoChemical::NA is per kmol
    (void)NAvo;
    for (int c = 0; c < nC; ++c) {
        const double V = cellVolumes_[c];
        const double FN = nPtsCell_[c] * rwfCell_[c];   // cloud_.nParticles(cell), dsmcVolFields.C:1098,1128
        double dsmcNCum = 0, mCumP = 0, ErotCum = 0, ZetaRotCum = 0, keP = 0;
        double mom[3] = {0, 0, 0};
        for (int s : f.typeIds) {
            const double* r = &acc[(size_t(c) * S + s) * nQ];
            const double m = species_[s].mass;
            dsmcNCum += r[0]; mCumP += m * r[0];
            for (int k = 0; k < 3; ++k) mom[k] += m * r[1 + k];
            keP += m * r[4];
            if (internal) { ErotCum += r[5]; ZetaRotCum += species_[s].rotationalDegreesOfFreedom * r[0]; }
        }
        const double nCum = FN * dsmcNCum, mCum = FN * mCumP, linearKECum = FN * keP;
        if (dsmcNCum > 1e-3) {
            o.dsmcNMean[c] = dsmcNCum / nT;
            const double rhoNMean = nCum / (nT * V), rhoMMean = mCum / (nT * V);
            o.rhoN[c] = rhoNMean; o.rhoM[c] = rhoMMean;
            double uu = 0;
            for (int k = 0; k < 3; ++k) { o.UMean[3 * c + k] = FN * mom[k] / mCum; uu += o.UMean[3 * c + k] * o.UMean[3 * c + k]; }
            const double linearKEMean = 0.5 * linearKECum / (V * nT);
            o.Ttra[c] = 2.0 / (3.0 * kB * rhoNMean) * (linearKEMean - 0.5 * rhoMMean * uu);
            o.p[c] = rhoNMean * kB * o.Ttra[c];
        } else {
            o.dsmcNMean[c] = 0.001;
        }
        if (f.densityOnly) continue;
        const double zetaRotTot = dsmcNCum > SMALL ? ZetaRotCum / dsmcNCum : 0.0;
        o.Trot[c] = ZetaRotCum > SMALL ? 2.0 * ErotCum / (kB * ZetaRotCum) : 0.0;
        double moleculesRhoN = 0, Tvib = 0, zetaVib = 0;
        if (internal) {
            for (int s : f.typeIds) {
                const double* r = &acc[(size_t(c) * S + s) * nQ];
                double speciesZetaVib = 0, zetaByTvibMod = 0;
                for (int m = 0; m < species_[s].nVibrationalModes; ++m) {
                    const double E = r[7 + m];
                    if (E > VSMALL && r[0] > SMALL) {
                        const double thetaV = species_[s].thetaV[m];
                        const double iMean = E / (kB * thetaV * r[0]);
                        if (iMean > SMALL) {
                            const double logFactor = std::log(1.0 + 1.0 / iMean);
                            const double TvibMod = thetaV / logFactor, zMod = 2.0 * iMean * logFactor;
                            speciesZetaVib += zMod;
                            zetaByTvibMod = zMod * TvibMod;  // assigned, not accumulated (dsmcVolFields.C:1469)
                        }
                    }
                }
                if (speciesZetaVib > SMALL) {
                    const double nS = FN * r[0];
                    moleculesRhoN += nS;
                    Tvib += nS * zetaByTvibMod / speciesZetaVib;
                    zetaVib += nS * speciesZetaVib;
                }
            }
            if (moleculesRhoN > SMALL) { Tvib /= moleculesRhoN; zetaVib /= moleculesRhoN; }
        }
        o.Tvib[c] = Tvib;
        o.Tov[c] = (3.0 * o.Ttra[c] + zetaRotTot * o.Trot[c] + zetaVib * Tvib) / (3.0 + zetaRotTot + zetaVib);
        // pressure tensor, shear-stress tensor and heat-flux vector (dsmcVolFields.C:1509-1622)
        if (f.measureHeatFluxShearStress && models_.measureHeatFluxShearStress && dsmcNCum > SMALL) {
            const int qF = 5 + (internal ? 2 + (ai.nModes > 0 ? ai.nModes : 0) : 0);
            double M[6] = {0, 0, 0, 0, 0, 0}, Mcc[3] = {0, 0, 0}, E[3] = {0, 0, 0}, MccAll = 0, ECum = 0;
            for (int s : f.typeIds) {
                const double* r = &acc[(size_t(c) * S + s) * nQ];
                const double m = species_[s].mass;
                for (int k = 0; k < 6; ++k) M[k] += m * r[qF + k];           // Muu Muv Muw Mvv Mvw Mww
                for (int k = 0; k < 3; ++k) { Mcc[k] += m * r[qF + 6 + k]; E[k] += r[qF + 9 + k]; }
                MccAll += m * r[4];
                if (internal) { ECum += r[5]; for (int md = 0; md < species_[s].nVibrationalModes; ++md) ECum += r[7 + md]; }
            }
            const double* u = &o.UMean[3 * size_t(c)];
            const double k0 = o.rhoN[c] / dsmcNCum;
            double* P = &o.pressureTensor[9 * size_t(c)];
            const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) P[3 * a + b] = k0 * (M[idx[a][b]] - mCumP * u[a] * u[b]);
            const double scalarPressure = (P[0] + P[4] + P[8]) / 3.0;
            double* T = &o.shearStressTensor[9 * size_t(c)];
            for (int k = 0; k < 9; ++k) T[k] = -P[k];
            T[0] += scalarPressure; T[4] += scalarPressure; T[8] += scalarPressure;
            for (int a = 0; a < 3; ++a)
                o.heatFluxVector[3 * size_t(c) + a] = k0 * (0.5 * Mcc[a] - 0.5 * MccAll * u[a] + E[a] - ECum * u[a]) - P[3 * a] * u[0] - P[3 * a + 1] * u[1] -
                                                      P[3 * a + 2] * u[2];
        }
        // Mach number (dsmcVolFields.C:1624-1661)
        double gammaCell = 0.0, particleCv = 0.0;
        if (dsmcNCum > SMALL && o.Ttra[c] > SMALL) {
            double molecularMass = 0, cv = 0, cp = 0;
            for (int s : f.typeIds) {
                const double Xs = acc[(size_t(c) * S + s) * nQ] / dsmcNCum;
                molecularMass += Xs * species_[s].mass;
                cv += Xs * (3.0 + species_[s].rotationalDegreesOfFreedom);
                cp += Xs * (5.0 + species_[s].rotationalDegreesOfFreedom);
            }
            const double gamma = cp / cv;
            gammaCell = gamma; particleCv = cv / 6.02214e26;   // molarCv_trarot / NAvo, OpenFOAM's NA is per kmol (dsmcVolFields.C:1647)
            const double a = std::sqrt(gamma * kB / molecularMass * o.Ttra[c]);
            const double* u = &o.UMean[3 * c];
            o.Ma[c] = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) / a;
        }
        // statistical error estimates (dsmcVolFields.C:1857-1873)
        if (f.measureErrors && o.dsmcNMean[c] > SMALL && o.Ma[c] > SMALL && gammaCell > SMALL && particleCv > SMALL) {
            const double deno = std::sqrt(o.dsmcNMean[c] * nT);
            o.rhoMError[c] = 1.0 / deno;
            o.UError[c] = 1.0 / (deno * o.Ma[c] * std::sqrt(gammaCell));
            o.TError[c] = std::sqrt(kB / particleCv) / deno;
            o.pError[c] = std::sqrt(gammaCell) / deno;
        }
        if (f.measureMeanFreePath && o.Ttra[c] > 1.0) {
            double mfp = 0, mcr = 0;
            for (int sp : f.typeIds) {
                double spMfp = 0, spMcr = 0;
                for (int sq : f.typeIds) {
                    const double Nq = acc[(size_t(c) * S + sq) * nQ];
                    if (!(Nq > SMALL)) continue;
                    const double dPQ = 0.5 * (species_[sp].diameter + species_[sq].diameter);
                    const double omegaPQ = 0.5 * (species_[sp].omega + species_[sq].omega);
                    const double massRatio = species_[sp].mass / species_[sq].mass;
                    const double reducedMass = species_[sp].mass * species_[sq].mass / (species_[sp].mass + species_[sq].mass);
                    const double nDensQ = FN * Nq / (V * nT);
                    spMfp += M_PI * dPQ * dPQ * nDensQ * std::pow(f.mfpReferenceTemperature / o.Ttra[c], omegaPQ - 0.5) * std::sqrt(1.0 + massRatio);
                    spMcr += 2.0 * std::sqrt(M_PI) * dPQ * dPQ * nDensQ * std::pow(o.Ttra[c] / f.mfpReferenceTemperature, 1.0 - omegaPQ) *
                             std::sqrt(2.0 * kB * f.mfpReferenceTemperature / reducedMass);
                }
                if (spMfp > SMALL) spMfp = 1.0 / spMfp;
                if (o.rhoN[c] > SMALL) {
                    const double w = acc[(size_t(c) * S + sp) * nQ] / dsmcNCum;
                    mfp += spMfp * w; mcr += spMcr * w;
                }
            }
            if (mfp < SMALL) mfp = GREAT;
            o.mfp[c] = mfp;
            if (mcr > SMALL) { o.mct[c] = 1.0 / mcr; o.mctToDt[c] = o.mct[c] / dtCell_[c]; } else { o.mct[c] = GREAT; o.mctToDt[c] = GREAT; }
            if (nCum > SMALL) o.measuredCollisionRate[c] = coll[2 * size_t(c)] * FN / (nCum * dtCell_[c]);
            // mfpToDx and the separation-of-free-paths ratio (dsmcVolFields.C:1788-1836)
            if (mfp != GREAT) {
                o.mfpToDx[c] = mfp / cellMaxDx(c);
                const double nColl = coll[2 * size_t(c)];
                const double mcs = nColl > SMALL ? coll[2 * size_t(c) + 1] / nColl : GREAT;   // meanCollisionSeparation_ (:1244-1250)
                o.SOF[c] = mfp > SMALL ? mcs / mfp : 0.0;
            } else {
                o.mfpToDx[c] = GREAT; o.SOF[c] = GREAT;
            }
        }
    }
    return o;
}

double dsmcCloud::cellMaxDx(int c) const {
    if (cellMaxDx_.empty()) {
        std::vector<double> lo(size_t(nCells_) * 3, GREAT), hi(size_t(nCells_) * 3, -GREAT);
        auto take = [&](int cell, int f) {
            for (int q = faceOffsets_[f]; q < faceOffsets_[f + 1]; ++q)
                for (int d = 0; d < 3; ++d) {
                    const double x = points_[3 * size_t(facePoints_[q]) + d];
                    lo[3 * size_t(cell) + d] = std::min(lo[3 * size_t(cell) + d], x);
                    hi[3 * size_t(cell) + d] = std::max(hi[3 * size_t(cell) + d], x);
                }
        };
        for (int f = 0; f < nFaces_; ++f) {
            take(owner_[f], f);
            if (f < nInternal_) take(neighbour_[f], f);
        }
        cellMaxDx_.resize(nCells_);
        for (int k = 0; k < nCells_; ++k)
            cellMaxDx_[k] = std::max(hi[3 * size_t(k)] - lo[3 * size_t(k)], std::max(hi[3 * size_t(k) + 1] - lo[3 * size_t(k) + 1], hi[3 * size_t(k) + 2] - lo[3 * size_t(k) + 2]));
    }
    return cellMaxDx_[c];
}

void dsmcCloud::writeFields(const std::string& timeDir, const std::vector<double>& instN) {
    selectSet(0);
    // Boundary values (dsmcVolFields.C:1878-2206): faces of `wall` patches get every field from the wall accumulators
    // (the *BF_ arrays of boundaryMeasurements), every other non-empty, non-cyclic patch the adjacent cell value.
    int32_t nMeas = 0, nWallQ = 0;
    check(dsmcb200_wall_info(ctx_, &nMeas, &nWallQ), "dsmcb200_wall_info");
    const int S = int(species_.size());
    std::vector<double> wall(size_t(std::max(nMeas, 1)) * S * std::max(nWallQ, 1));
    if (nMeas) check(dsmcb200_download_wall_accumulators(ctx_, wall.data()), "dsmcb200_download_wall_accumulators");
    dsmcb200_accum_info ai{};
    check(dsmcb200_accum_info_get(ctx_, &ai), "dsmcb200_accum_info_get");
    const std::vector<double> wallAll = wall;
    const double kB = models_.kB > 0 ? models_.kB : 1.38065e-23;
    // measured-face index of a boundary face follows the order of the patch models with a wall model
    std::vector<int> measStart(boundary_.size(), -1);
    {
        int k = 0;
        for (auto& pm : patchModels_)
            if (pm.model != DSMCB200_BND_DELETION) { measStart[pm.patch] = k; k += boundary_[pm.patch].nFaces; }
    }
    // WallQ order of the engine (csrc/engine.h): rhoN 0, rhoNInt 1, rhoNElec 2, rhoM 3, linearKE 4, mcc 5, momentum 6-8, Erot 9,
    // zetaRot 10, Evib 11, Eelec 12, q 13, fD 14-16, EvibMod 17+
    struct WallFace { double rhoN, rhoM, U[3], Ttra, Trot, Tvib, Tov, Ma, fD[3], p, tau, q; };
    bool firstField = true;
    for (auto& f : fields_) {
        DerivedFields d = calculateField(f);
        // the wall sums of this field since its last reset
        wall = wallAll;
        if (sampleSets_.size() > 1) {   // ... of this field's sample set (selected by calculateField)
            if (nMeas) check(dsmcb200_download_wall_accumulators(ctx_, wall.data()), "dsmcb200_download_wall_accumulators");
            check(dsmcb200_accum_info_get(ctx_, &ai), "dsmcb200_accum_info_get");
        }
        if (f.baseWall.size() == wall.size()) for (size_t k = 0; k < wall.size(); ++k) wall[k] -= f.baseWall[k];
        const double nT = ai.nTimeSteps - f.baseNT > 0 ? ai.nTimeSteps - f.baseNT : 1.0;
        // fields().overallT(cell) is Tov_ of fields_[0] as written here (dsmcFieldProperties.C:235-240): the "2008" Zv formulation
        // reads it in the collisions of the following steps (dsmcCloud.C:1441-1456)
        if (firstField && models_.invZvFormulation == 1 && (models_.collisionModel == DSMCB200_COLL_LB_VHS || models_.collisionModel == DSMCB200_COLL_LB_VSS))
            check(dsmcb200_upload_overall_temperature(ctx_, d.Tov.data()), "dsmcb200_upload_overall_temperature");
        firstField = false;
        // ---- wall faces of this instance
        std::vector<std::vector<WallFace>> wf(boundary_.size());
        for (size_t j = 0; j < boundary_.size(); ++j) {
            if (boundary_[j].type != "wall" || measStart[j] < 0) continue;
            wf[j].resize(boundary_[j].nFaces);
            for (int k = 0; k < boundary_[j].nFaces; ++k) {
                WallFace& w = wf[j][k];
                std::memset(&w, 0, sizeof(w));
                const int face = boundary_[j].startFace + k;
                auto W = [&](int s, int q) { return wall[(size_t(measStart[j] + k) * S + s) * nWallQ + q]; };
                double rhoNBF = 0, rhoMBF = 0, linearKEBF = 0, momBF[3] = {0, 0, 0}, ErotBF = 0, zetaRotBF = 0, qBF = 0, fDBF[3] = {0, 0, 0};
                for (int s : f.typeIds) {
                    rhoNBF += W(s, 0); rhoMBF += W(s, 3); linearKEBF += W(s, 4); ErotBF += W(s, 9); zetaRotBF += W(s, 10); qBF += W(s, 13);
                    for (int q = 0; q < 3; ++q) { momBF[q] += W(s, 6 + q); fDBF[q] += W(s, 14 + q); }
                }
                // the parcels' RWF is in the sums; what is left is the time-step model's nParticles of the face (dsmcVolFields.C:1905-1913)
                const double FN = nPtsCell_[owner_[face]];
                const double rhoNMean = rhoNBF * FN / nT, rhoMMean = rhoMBF * FN / nT, linearKEMean = linearKEBF * FN / nT;
                w.rhoN = rhoNMean; w.rhoM = rhoMMean;
                if (rhoMMean > VSMALL) {
                    for (int q = 0; q < 3; ++q) w.U[q] = momBF[q] / rhoMBF;
                    w.Ttra = 2.0 / (3.0 * kB * rhoNMean) * (linearKEMean - 0.5 * rhoMMean * (w.U[0] * w.U[0] + w.U[1] * w.U[1] + w.U[2] * w.U[2]));
                }
                const double zetaRotTot = rhoNBF > SMALL ? zetaRotBF / rhoNBF : 0.0;
                w.Trot = zetaRotBF > SMALL ? 2.0 * ErotBF / (kB * zetaRotBF) : 0.0;
                double zetaVibBF = 0.0, moleculesRhoN = 0.0;
                for (int s : f.typeIds) {
                    const double spRhoN = W(s, 0);
                    double spZetaVib = 0.0, zetaByTvibMod = 0.0;
                    if (spRhoN > SMALL) {
                        for (int mod = 0; mod < species_[s].nVibrationalModes; ++mod) {
                            const double thetaV = species_[s].thetaV[mod];
                            const double iMean = (17 + mod < nWallQ ? W(s, 17 + mod) : 0.0) / (kB * thetaV * spRhoN);
                            if (iMean > SMALL) {
                                const double logFactor = std::log(1.0 + 1.0 / iMean);
                                const double Tmod = thetaV / logFactor, zmod = 2.0 * iMean * logFactor;
                                spZetaVib += zmod; zetaByTvibMod += zmod * Tmod;
                            }
                        }
                    }
                    if (spZetaVib > SMALL) {
                        moleculesRhoN += spRhoN;
                        w.Tvib += spRhoN * (zetaByTvibMod / spZetaVib);
                        zetaVibBF += spRhoN * spZetaVib;
                    }
                }
                if (moleculesRhoN > SMALL) { w.Tvib /= moleculesRhoN; zetaVibBF /= moleculesRhoN; }
                w.Tov = (3.0 * w.Ttra + zetaRotTot * w.Trot + zetaVibBF * w.Tvib) / (3.0 + zetaRotTot + zetaVibBF);
                if (rhoNBF > SMALL) {
                    double molecularMassBF = 0, cv = 0, cp = 0;
                    for (int s : f.typeIds) {
                        const double Xs = W(s, 0) / rhoNBF;
                        molecularMassBF += Xs * species_[s].mass;
                        cv += Xs * (3.0 + species_[s].rotationalDegreesOfFreedom);
                        cp += Xs * (5.0 + species_[s].rotationalDegreesOfFreedom);
                    }
                    const double gasConstant = (w.Ttra > SMALL) ? kB / molecularMassBF : 0.0;
                    const double speedOfSound = std::sqrt(cp / cv * gasConstant * w.Ttra);
                    w.Ma = std::sqrt(w.U[0] * w.U[0] + w.U[1] * w.U[1] + w.U[2] * w.U[2]) / speedOfSound;
                }
                // unit vectors of the face (dsmcVolFields::calculateWallUnitVectors, :52-80)
                const double* Sf = &faceAreas_[3 * size_t(face)];
                const double magSf = std::sqrt(Sf[0] * Sf[0] + Sf[1] * Sf[1] + Sf[2] * Sf[2]);
                const double n[3] = {Sf[0] / magSf, Sf[1] / magSf, Sf[2] / magSf};
                const double* p0 = &points_[3 * size_t(facePoints_[faceOffsets_[face]])];
                double t1[3] = {faceCentres_[3 * size_t(face)] - p0[0], faceCentres_[3 * size_t(face) + 1] - p0[1], faceCentres_[3 * size_t(face) + 2] - p0[2]};
                const double m1 = std::sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
                for (double& x : t1) x /= m1;
                double t2[3] = {n[1] * t1[2] - n[2] * t1[1], n[2] * t1[0] - n[0] * t1[2], n[0] * t1[1] - n[1] * t1[0]};
                const double m2 = std::sqrt(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2]);
                for (double& x : t2) x /= m2;
                for (int q = 0; q < 3; ++q) w.fD[q] = fDBF[q] / nT;
                w.p = w.fD[0] * n[0] + w.fD[1] * n[1] + w.fD[2] * n[2];
                const double a1 = w.fD[0] * t1[0] + w.fD[1] * t1[1] + w.fD[2] * t1[2], a2 = w.fD[0] * t2[0] + w.fD[1] * t2[1] + w.fD[2] * t2[2];
                w.tau = std::sqrt(a1 * a1 + a2 * a2);
                w.q = qBF / nT;
            }
        }
        // per-patch values of a scalar field: `pick` selects the wall-face value, other wall / patch faces copy the cell
        auto scalarPatches = [&](const std::vector<double>& cellField, double WallFace::*pick) {
            std::vector<foam::PatchValues> pv;
            for (size_t j = 0; j < boundary_.size(); ++j) {
                foam::PatchValues p;
                p.name = boundary_[j].name; p.type = boundary_[j].type;
                if (p.type == "wall" || p.type == "patch") {
                    p.values.resize(boundary_[j].nFaces);
                    for (int k = 0; k < boundary_[j].nFaces; ++k) {
                        if (pick && !wf[j].empty()) p.values[k] = wf[j][k].*pick;
                        else p.values[k] = cellField.empty() ? 0.0 : cellField[owner_[boundary_[j].startFace + k]];
                    }
                }
                pv.push_back(p);
            }
            return pv;
        };
        auto wr = [&](const std::string& name, const std::string& dims, const std::vector<double>& v, double WallFace::*pick = nullptr) {
            foam::writeVolField(timeDir + "/" + name + "_" + f.fieldName, timeName_, name + "_" + f.fieldName, dims, v.data(), nCells_, 1,
                                scalarPatches(v, pick));
        };
        {
            std::vector<double> dsmcN(size_t(nCells_), 0.0);
            for (int c = 0; c < nCells_; ++c)
                for (int s : f.typeIds) dsmcN[c] += instN[size_t(c) * S + s];
            wr("dsmcN", "[0 0 0 0 0 0 0]", dsmcN);
        }
        wr("dsmcNMean", "[0 0 0 0 0 0 0]", d.dsmcNMean);
        wr("rhoN", "[0 -3 0 0 0 0 0]", d.rhoN, &WallFace::rhoN);
        wr("rhoM", "[1 -3 0 0 0 0 0]", d.rhoM, &WallFace::rhoM);
        if (f.densityOnly) continue;
        wr("p", "[1 -1 -2 0 0 0 0]", d.p, &WallFace::p);
        wr("Ttra", "[0 0 0 1 0 0 0]", d.Ttra, &WallFace::Ttra);
        wr("Trot", "[0 0 0 1 0 0 0]", d.Trot, &WallFace::Trot);
        wr("Tvib", "[0 0 0 1 0 0 0]", d.Tvib, &WallFace::Tvib);
        wr("Tov", "[0 0 0 1 0 0 0]", d.Tov, &WallFace::Tov);
        wr("Ma", "[0 0 0 0 0 0 0]", d.Ma, &WallFace::Ma);
        std::vector<double> zero(size_t(nCells_), 0.0);
        // q_ and tau_ live on the walls only (internal field zero, dsmcVolFields.C:231-257)
        {
            auto wallOnly = [&](double WallFace::*pick) {
                auto pv = scalarPatches(zero, pick);
                return pv;
            };
            foam::writeVolField(timeDir + "/wallHeatFlux_" + f.fieldName, timeName_, "wallHeatFlux_" + f.fieldName, "[1 0 -3 0 0 0 0]", zero.data(), nCells_, 1,
                                wallOnly(&WallFace::q));
            foam::writeVolField(timeDir + "/wallShearStress_" + f.fieldName, timeName_, "wallShearStress_" + f.fieldName, "[1 -1 -2 0 0 0 0]", zero.data(),
                                nCells_, 1, wallOnly(&WallFace::tau));
        }
        if (f.measureMeanFreePath) {
            wr("mfp", "[0 1 0 0 0 0 0]", d.mfp);
            wr("mfpToDx", "[0 0 0 0 0 0 0]", d.mfpToDx);
            wr("mct", "[0 0 1 0 0 0 0]", d.mct);
            wr("mctToDt", "[0 0 0 0 0 0 0]", d.mctToDt);
            wr("SOFP", "[0 0 0 0 0 0 0]", d.SOF);
        }
        if (f.measureErrors) {
            wr("rhoMError", "[0 0 0 0 0 0 0]", d.rhoMError); wr("UError", "[0 0 0 0 0 0 0]", d.UError);
            wr("TError", "[0 0 0 0 0 0 0]", d.TError); wr("pError", "[0 0 0 0 0 0 0]", d.pError);
        }
        if (f.measureHeatFluxShearStress && !d.heatFluxVector.empty()) {
            // zero-gradient boundary values (dsmcVolFields.C:2183-2191)
            auto multi = [&](const std::string& name, const std::string& dims, const std::vector<double>& v, int nc) {
                std::vector<foam::PatchValues> pv;
                for (size_t j = 0; j < boundary_.size(); ++j) {
                    foam::PatchValues p;
                    p.name = boundary_[j].name; p.type = boundary_[j].type;
                    if (p.type == "wall" || p.type == "patch") {
                        p.values.resize(size_t(boundary_[j].nFaces) * nc);
                        for (int k = 0; k < boundary_[j].nFaces; ++k)
                            for (int q = 0; q < nc; ++q) p.values[size_t(k) * nc + q] = v[size_t(owner_[boundary_[j].startFace + k]) * nc + q];
                    }
                    pv.push_back(p);
                }
                foam::writeVolField(timeDir + "/" + name + "_" + f.fieldName, timeName_, name + "_" + f.fieldName, dims, v.data(), nCells_, nc, pv);
            };
            multi("heatFluxVector", "[1 0 -3 0 0 0 0]", d.heatFluxVector, 3);
            multi("pressureTensor", "[1 -1 -2 0 0 0 0]", d.pressureTensor, 9);
            multi("shearStressTensor", "[1 -1 -2 0 0 0 0]", d.shearStressTensor, 9);
        }
        // vectors: U and the wall force density fD
        {
            std::vector<foam::PatchValues> pu, pf;
            std::vector<double> zero3(size_t(nCells_) * 3, 0.0);
            for (size_t j = 0; j < boundary_.size(); ++j) {
                foam::PatchValues a, b;
                a.name = b.name = boundary_[j].name; a.type = b.type = boundary_[j].type;
                if (a.type == "wall" || a.type == "patch") {
                    a.values.resize(size_t(boundary_[j].nFaces) * 3); b.values.assign(size_t(boundary_[j].nFaces) * 3, 0.0);
                    for (int k = 0; k < boundary_[j].nFaces; ++k) {
                        const int c = owner_[boundary_[j].startFace + k];
                        for (int q = 0; q < 3; ++q) {
                            a.values[3 * k + q] = wf[j].empty() ? d.UMean[3 * size_t(c) + q] : wf[j][k].U[q];
                            if (!wf[j].empty()) b.values[3 * k + q] = wf[j][k].fD[q];
                        }
                    }
                }
                pu.push_back(a); pf.push_back(b);
            }
            foam::writeVolField(timeDir + "/U_" + f.fieldName, timeName_, "U_" + f.fieldName, "[0 1 -1 0 0 0 0]", d.UMean.data(), nCells_, 3, pu);
            foam::writeVolField(timeDir + "/fD_" + f.fieldName, timeName_, "fD_" + f.fieldName, "[1 -1 -2 0 0 0 0]", zero3.data(), nCells_, 3, pf);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Sampling restart files.  dsmcVolFields::writeOut / readIn (DSMC/macroscopicProperties/derived/combined/dsmcVolFields/
// dsmcVolFields.C:647-835) keep one dictionary `<time>/uniform/resumeSampling_<fieldName>` per field instance.  The engine
// samples per SPECIES and derives every instance from those rows, so its own restart state is the per-species rows:
// they go, losslessly, to `<time>/uniform/resumeSampling_dsmcb200` (read back here), and each instance's dictionary is
// written with the reference's key set computed from them (accumulation rules of dsmcVolFields.C:1115-1237), so that
// the reference's tools and dsmcFoam+ itself can pick the run up.  Not tracked by the engine and written as zeros:
// dsmcNGrndElecLvlSpeciesCum / dsmcN1stElecLvlSpeciesCum (only feed Telec, which the reference forces to 0, :1497).
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct ListOut {
    FILE* f;
    int prec;
    void num(double v) const { std::fprintf(f, "%.*g", prec, v); }
    void scalars(const double* a, int64_t n) const {
        if (n == 0) { std::fputs("0 ( )", f); return; }
        bool uniform = true;
        for (int64_t i = 1; i < n && uniform; ++i) uniform = a[i] == a[0];
        if (uniform) { std::fprintf(f, "%lld { ", (long long)n); num(a[0]); std::fputs(" }", f); return; }
        std::fprintf(f, "%lld ( ", (long long)n);
        for (int64_t i = 0; i < n; ++i) { num(a[i]); std::fputc(' ', f); }
        std::fputc(')', f);
    }
    void vectors(const double* a, int64_t n) const {
        if (n == 0) { std::fputs("0 ( )", f); return; }
        bool uniform = true;
        for (int64_t i = 1; i < n && uniform; ++i) uniform = a[3 * i] == a[0] && a[3 * i + 1] == a[1] && a[3 * i + 2] == a[2];
        auto one = [&](const double* v) { std::fputs("( ", f); num(v[0]); std::fputc(' ', f); num(v[1]); std::fputc(' ', f); num(v[2]); std::fputs(" )", f); };
        if (uniform) { std::fprintf(f, "%lld { ", (long long)n); one(a); std::fputs(" }", f); return; }
        std::fprintf(f, "%lld ( ", (long long)n);
        for (int64_t i = 0; i < n; ++i) { one(a + 3 * i); std::fputc(' ', f); }
        std::fputc(')', f);
    }
    void key(const char* k) const { std::fprintf(f, "%-15s ", k); }
    void end() const { std::fputs(";\n\n", f); }
    void entry(const char* k, const std::vector<double>& a) const { key(k); scalars(a.data(), int64_t(a.size())); end(); }
    void entryV(const char* k, const std::vector<double>& a) const { key(k); vectors(a.data(), int64_t(a.size() / 3)); end(); }
    // List<scalarField>: one list per species / per patch
    void entryLL(const char* k, const std::vector<std::vector<double>>& a) const {
        key(k);
        std::fprintf(f, "%zu ( ", a.size());
        for (auto& v : a) { scalars(v.data(), int64_t(v.size())); std::fputc(' ', f); }
        std::fputc(')', f);
        end();
    }
};
}  // namespace

void dsmcCloud::writeResumeSampling(const std::string& timeDir) {
    for (size_t set = 0; set < std::max<size_t>(1, sampleSets_.size()); ++set) writeResumeSamplingOf(timeDir, int(set));
    selectSet(0);
}

// the fields of one sample set: uniform/resumeSampling_dsmcb200 (set 0), resumeSampling_dsmcb200_set<k> (the others)
void dsmcCloud::writeResumeSamplingOf(const std::string& timeDir, int set) {
    bool any = false;
    for (auto& f : fields_) any = any || (f.set == set && f.averagingAcrossManyRuns);
    if (!any) return;
    selectSet(set);
    const std::string own = set == 0 ? std::string("resumeSampling_dsmcb200") : "resumeSampling_dsmcb200_set" + std::to_string(set);
    dsmcb200_accum_info ai{};
    check(dsmcb200_accum_info_get(ctx_, &ai), "dsmcb200_accum_info_get");
    const int S = ai.nSpecies, nQ = ai.nQuantities, nC = ai.nCells;
    std::vector<double> acc(size_t(nC) * S * nQ), coll(size_t(nC) * 2), sig(nC), rem(nC);
    check(dsmcb200_download_accumulators(ctx_, acc.data(), coll.data()), "dsmcb200_download_accumulators");
    check(dsmcb200_download_cellstate(ctx_, sig.data(), rem.data()), "dsmcb200_download_cellstate");
    int32_t nMeas = 0, nWallQ = 0;
    check(dsmcb200_wall_info(ctx_, &nMeas, &nWallQ), "dsmcb200_wall_info");
    std::vector<double> wall(size_t(nMeas) * S * nWallQ);
    if (nMeas) check(dsmcb200_download_wall_accumulators(ctx_, wall.data()), "dsmcb200_download_wall_accumulators");
    const std::string ud = timeDir + "/uniform";
    foam::makeDirs(ud);
    const bool internal = ai.nModes >= 0;
    const int nModes = internal ? ai.nModes : 0;
    const int qFlux = 5 + (internal ? 2 + nModes : 0);
    const bool hasFlux = models_.measureHeatFluxShearStress != 0;
    const int qClass = qFlux + (hasFlux ? 12 : 0);
    const bool hasClass = models_.measureClassifications != 0;

    // ---- the engine's own state: per-species rows, lossless
    {
        FILE* f = std::fopen((ud + "/" + own).c_str(), "w");
        if (!f) throw FoamError("cannot write " + ud + "/" + own);
        std::fputs(foam::asciiHeader("dictionary", timeName_ + "/uniform", own).c_str(), f);
        ListOut o{f, 17};
        std::fprintf(f, "nTimeSteps      %.17g;\n\nnCells          %d;\n\nnSpecies        %d;\n\nnQuantities     %d;\n\nnMeasuredFaces  %d;\n\nnWallQuantities %d;\n\n",
                     ai.nTimeSteps, nC, S, nQ, nMeas, nWallQ);
        o.entry("accumulators", acc);
        o.entry("collisionCumulative", coll);
        o.entry("wallAccumulators", wall);
        o.entry("collisionSelectionRemainder", rem);
        std::fclose(f);
    }

    // ---- one dictionary per field instance with the reference's keys
    std::vector<int> measStart(boundary_.size(), -1);
    {
        int k = 0;
        for (auto& pm : patchModels_)
            if (pm.model != DSMCB200_BND_DELETION) { measStart[pm.patch] = k; k += boundary_[pm.patch].nFaces; }
    }
    for (auto& fs : fields_) {
        if (!fs.averagingAcrossManyRuns || fs.set != set) continue;
        // the sums of this field since its own last reset (dsmcCloud::write)
        std::vector<double> accL = acc, collL = coll, wallL = wall;
        if (fs.baseAcc.size() == accL.size()) for (size_t k = 0; k < accL.size(); ++k) accL[k] -= fs.baseAcc[k];
        if (fs.baseColl.size() == collL.size()) for (size_t k = 0; k < collL.size(); ++k) collL[k] -= fs.baseColl[k];
        if (fs.baseWall.size() == wallL.size()) for (size_t k = 0; k < wallL.size(); ++k) wallL[k] -= fs.baseWall[k];
        const int nT = int(fs.typeIds.size());
        auto cells = [&]() { return std::vector<double>(size_t(nC), 0.0); };
        std::vector<double> dsmcNCum = cells(), dsmcMCum = cells(), dsmcLinearKECum = cells(), dsmcErotCum = cells(), dsmcZetaRotCum = cells(),
                            dsmcMom(size_t(nC) * 3, 0.0), dsmcECum = cells(), dsmcNElecLvlCum = cells(), nColls = cells(), collSep = cells();
        std::vector<double> Muu = cells(), Muv = cells(), Muw = cells(), Mvv = cells(), Mvw = cells(), Mww = cells(), Mccu = cells(), Mccv = cells(),
                            Mccw = cells(), Eu = cells(), Ev = cells(), Ew = cells(), cI = cells(), cII = cells(), cIII = cells();
        std::vector<std::vector<double>> spN(nT, cells()), spMcc(nT, cells()), spEelec(nT, cells()), spZero(nT, cells());
        std::vector<std::vector<std::vector<double>>> spEvibMod(nT);
        for (int t = 0; t < nT; ++t) spEvibMod[t].assign(species_[fs.typeIds[t]].nVibrationalModes, cells());
        for (int c = 0; c < nC; ++c) {
            for (int t = 0; t < nT; ++t) {
                const int s = fs.typeIds[t];
                const double* r = &accL[(size_t(c) * S + s) * nQ];
                const double m = species_[s].mass;
                dsmcNCum[c] += r[0]; dsmcMCum[c] += m * r[0]; dsmcLinearKECum[c] += m * r[4];
                for (int k = 0; k < 3; ++k) dsmcMom[3 * size_t(c) + k] += m * r[1 + k];
                spN[t][c] = r[0]; spMcc[t][c] = m * r[4];
                if (internal) {
                    dsmcErotCum[c] += r[5]; dsmcZetaRotCum[c] += species_[s].rotationalDegreesOfFreedom * r[0];
                    spEelec[t][c] = r[6];
                    double eint = r[5];
                    for (int md = 0; md < species_[s].nVibrationalModes; ++md) { spEvibMod[t][md][c] = r[7 + md]; eint += r[7 + md]; }
                    dsmcECum[c] += eint;
                    if (species_[s].nElectronicLevels > 1) dsmcNElecLvlCum[c] += r[0];
                }
                if (hasFlux) {
                    const double* q = r + qFlux;
                    Muu[c] += m * q[0]; Muv[c] += m * q[1]; Muw[c] += m * q[2]; Mvv[c] += m * q[3]; Mvw[c] += m * q[4]; Mww[c] += m * q[5];
                    Mccu[c] += m * q[6]; Mccv[c] += m * q[7]; Mccw[c] += m * q[8];
                    Eu[c] += q[9]; Ev[c] += q[10]; Ew[c] += q[11];
                }
                if (hasClass) { cI[c] += r[qClass]; cII[c] += r[qClass + 1]; cIII[c] += r[qClass + 2]; }
            }
            nColls[c] = collL[2 * size_t(c)]; collSep[c] = collL[2 * size_t(c) + 1];
        }
        auto scaled = [&](const std::vector<double>& v) {   // per cell, nCmpt values each, times cloud_.nParticles(cell)
            std::vector<double> o(v);
            const size_t w = o.size() / size_t(std::max(nC, 1));
            for (size_t k = 0; k < o.size(); ++k) o[k] *= nPtsCell_[k / w] * rwfCell_[k / w];
            return o;
        };
        const std::string name = "resumeSampling_" + fs.fieldName;
        FILE* f = std::fopen((ud + "/" + name).c_str(), "w");
        if (!f) throw FoamError("cannot write " + ud + "/" + name);
        std::fputs(foam::asciiHeader("dictionary", timeName_ + "/uniform", name).c_str(), f);
        ListOut o{f, 10};
        std::fprintf(f, "nTimeSteps      %.10g;\n\n", ai.nTimeSteps - fs.baseNT);
        o.entry("dsmcNCum", dsmcNCum); o.entry("dsmcMCum", dsmcMCum); o.entry("dsmcLinearKECum", dsmcLinearKECum);
        o.entryV("dsmcMomentumCum", dsmcMom); o.entry("dsmcErotCum", dsmcErotCum); o.entry("dsmcZetaRotCum", dsmcZetaRotCum);
        o.entryLL("dsmcSpeciesEelecCum", spEelec); o.entryLL("dsmcNSpeciesCum", spN); o.entryLL("dsmcMccSpeciesCum", spMcc);
        o.entry("dsmcMuuCum", Muu); o.entry("dsmcMuvCum", Muv); o.entry("dsmcMuwCum", Muw); o.entry("dsmcMvvCum", Mvv); o.entry("dsmcMvwCum", Mvw);
        o.entry("dsmcMwwCum", Mww); o.entry("dsmcMccCum", dsmcLinearKECum); o.entry("dsmcMccuCum", Mccu); o.entry("dsmcMccvCum", Mccv);
        o.entry("dsmcMccwCum", Mccw); o.entry("dsmcEuCum", Eu); o.entry("dsmcEvCum", Ev); o.entry("dsmcEwCum", Ew); o.entry("dsmcECum", dsmcECum);
        o.entry("dsmcNElecLvlCum", dsmcNElecLvlCum); o.entryLL("dsmcNGrndElecLvlSpeciesCum", spZero); o.entryLL("dsmcN1stElecLvlSpeciesCum", spZero);
        if (fs.measureClassifications) { o.entry("dsmcNClassICum", cI); o.entry("dsmcNClassIICum", cII); o.entry("dsmcNClassIIICum", cIII); }
        {   // List<List<scalarField>> [species][mode]
            o.key("dsmcSpeciesEvibModCum");
            std::fprintf(f, "%d ( ", nT);
            for (int t = 0; t < nT; ++t) {
                std::fprintf(f, "%zu ( ", spEvibMod[t].size());
                for (auto& v : spEvibMod[t]) { o.scalars(v.data(), int64_t(v.size())); std::fputc(' ', f); }
                std::fputs(") ", f);
            }
            std::fputc(')', f);
            o.end();
        }
        o.entry("dsmcNCollsCum", nColls);
        o.entry("nCum", scaled(dsmcNCum)); o.entry("mCum", scaled(dsmcMCum));
        {
            std::vector<std::vector<double>> nSp;
            for (auto& v : spN) nSp.push_back(scaled(v));
            o.entryLL("nSpeciesCum", nSp);
        }
        o.entryV("momentumCum", scaled(dsmcMom)); o.entry("linearKECum", scaled(dsmcLinearKECum)); o.entry("collisionSeparation", collSep);
        // ---- boundary measurements: List<scalarField> over ALL patches (zero where no wall model samples)
        auto patchSum = [&](int wq, int nCmpt) {
            std::vector<std::vector<double>> out(boundary_.size());
            for (size_t j = 0; j < boundary_.size(); ++j) {
                out[j].assign(size_t(boundary_[j].nFaces) * nCmpt, 0.0);
                if (measStart[j] < 0) continue;
                for (int k = 0; k < boundary_[j].nFaces; ++k)
                    for (int s : fs.typeIds)
                        for (int q = 0; q < nCmpt; ++q) out[j][size_t(k) * nCmpt + q] += wallL[(size_t(measStart[j] + k) * S + s) * nWallQ + wq + q];
            }
            return out;
        };
        auto patchSpecies = [&](int wq) {
            std::vector<std::vector<std::vector<double>>> out(nT, std::vector<std::vector<double>>(boundary_.size()));
            for (int t = 0; t < nT; ++t)
                for (size_t j = 0; j < boundary_.size(); ++j) {
                    out[t][j].assign(size_t(boundary_[j].nFaces), 0.0);
                    if (measStart[j] < 0 || wq < 0) continue;
                    for (int k = 0; k < boundary_[j].nFaces; ++k) out[t][j][k] = wallL[(size_t(measStart[j] + k) * S + fs.typeIds[t]) * nWallQ + wq];
                }
            return out;
        };
        auto writeVecPatches = [&](const char* k, const std::vector<std::vector<double>>& a) {
            o.key(k);
            std::fprintf(f, "%zu ( ", a.size());
            for (auto& v : a) { o.vectors(v.data(), int64_t(v.size() / 3)); std::fputc(' ', f); }
            std::fputc(')', f);
            o.end();
        };
        auto writeSpeciesPatches = [&](const char* k, const std::vector<std::vector<std::vector<double>>>& a) {
            o.key(k);
            std::fprintf(f, "%zu ( ", a.size());
            for (auto& sp : a) {
                std::fprintf(f, "%zu ( ", sp.size());
                for (auto& v : sp) { o.scalars(v.data(), int64_t(v.size())); std::fputc(' ', f); }
                std::fputs(") ", f);
            }
            std::fputc(')', f);
            o.end();
        };
        // WallQ order of the engine (csrc/engine.h): rhoN 0, rhoNInt 1, rhoNElec 2, rhoM 3, linearKE 4, mcc 5, momentum 6-8, Erot 9,
        // zetaRot 10, Evib 11, Eelec 12, q 13, fD 14-16, EvibMod 17+
        if (nWallQ >= 17) {
            o.entryLL("rhoNBF", patchSum(0, 1)); o.entryLL("rhoMBF", patchSum(3, 1)); o.entryLL("linearKEBF", patchSum(4, 1));
            writeVecPatches("momentumBF", patchSum(6, 3));
            o.entryLL("ErotBF", patchSum(9, 1)); o.entryLL("zetaRotBF", patchSum(10, 1)); o.entryLL("rhoNIntBF", patchSum(1, 1));
            o.entryLL("rhoNElecBF", patchSum(2, 1)); o.entryLL("qBF", patchSum(13, 1));
            writeVecPatches("fDBF", patchSum(14, 3));
            writeSpeciesPatches("speciesRhoNBF", patchSpecies(0)); writeSpeciesPatches("speciesEvibBF", patchSpecies(11));
            writeSpeciesPatches("speciesEelecBF", patchSpecies(12)); writeSpeciesPatches("speciesMccBF", patchSpecies(5));
            {   // [species][mode][patch]
                o.key("speciesEvibModBF");
                std::fprintf(f, "%d ( ", nT);
                for (int t = 0; t < nT; ++t) {
                    const int nM = species_[fs.typeIds[t]].nVibrationalModes;
                    std::fprintf(f, "%d ( ", nM);
                    for (int md = 0; md < nM; ++md) {
                        std::fprintf(f, "%zu ( ", boundary_.size());
                        for (size_t j = 0; j < boundary_.size(); ++j) {
                            std::vector<double> v(size_t(boundary_[j].nFaces), 0.0);
                            if (measStart[j] >= 0 && 17 + md < nWallQ)
                                for (int k = 0; k < boundary_[j].nFaces; ++k) v[k] = wallL[(size_t(measStart[j] + k) * S + fs.typeIds[t]) * nWallQ + 17 + md];
                            o.scalars(v.data(), int64_t(v.size())); std::fputc(' ', f);
                        }
                        std::fputs(") ", f);
                    }
                    std::fputs(") ", f);
                }
                std::fputc(')', f);
                o.end();
            }
        }
        std::fclose(f);
    }
}

void dsmcCloud::readResumeSampling() {
    for (size_t set = 0; set < std::max<size_t>(1, sampleSets_.size()); ++set) readResumeSamplingOf(int(set));
    selectSet(0);
}

void dsmcCloud::readResumeSamplingOf(int set) {
    // dsmcVolFields.C:1052-1066: read only with averagingAcrossManyRuns and resetAtOutput off
    bool any = false, reset = false, members = false;
    for (auto& f : fields_)
        if (f.set == set) {
            any = any || f.averagingAcrossManyRuns;
            const bool r = f.resetAtOutput && !(time_ + deltaT_ > f.resetAtOutputUntilTime);
            reset = members ? (reset && r) : r;
            members = true;
        }
    if (!any) return;
    selectSet(set);
    if (reset) {
        std::printf("Averaging across many runs will be enabled as soon as resetAtOutput is turned off.\n");
        return;
    }
    const std::string path = root_ + "/" + timeName_ + "/uniform/resumeSampling_dsmcb200" + (set == 0 ? std::string() : "_set" + std::to_string(set));
    if (!foam::exists(path)) {
        // a case checkpointed by dsmcFoam+ itself has only the per-field dictionaries (dsmcVolFields.C:1052-1066): say that they are not read
        for (auto& f : fields_)
            if (f.set == set && f.averagingAcrossManyRuns && foam::exists(root_ + "/" + timeName_ + "/uniform/resumeSampling_" + f.fieldName))
                std::printf("WARNING: uniform/resumeSampling_%s (written by dsmcFoam+) is not read; the averages of this run start at zero. "
                            "Only resumeSampling_dsmcb200, written by this engine, restores the sampling.\n", f.fieldName.c_str());
        return;
    }
    foam::Dict d = foam::readDict(path);
    dsmcb200_accum_info ai{};
    check(dsmcb200_accum_info_get(ctx_, &ai), "dsmcb200_accum_info_get");
    int32_t nMeas = 0, nWallQ = 0;
    check(dsmcb200_wall_info(ctx_, &nMeas, &nWallQ), "dsmcb200_wall_info");
    if (d.labelOr("nCells", -1) != ai.nCells || d.labelOr("nSpecies", -1) != ai.nSpecies || d.labelOr("nQuantities", -1) != ai.nQuantities ||
        d.labelOr("nMeasuredFaces", -1) != nMeas || d.labelOr("nWallQuantities", -1) != nWallQ) {
        std::printf("resumeSampling_dsmcb200 does not match the current mesh / species / field set: sampling starts afresh\n");
        return;
    }
    std::vector<double> acc = d.scalarList("accumulators"), coll = d.scalarList("collisionCumulative");
    if (acc.size() != size_t(ai.nCells) * ai.nSpecies * ai.nQuantities || coll.size() != size_t(ai.nCells) * 2)
        throw FoamError("resumeSampling_dsmcb200: list sizes do not match the header in " + path);
    check(dsmcb200_upload_accumulators(ctx_, acc.data(), coll.data(), d.scalar("nTimeSteps")), "dsmcb200_upload_accumulators");
    if (nMeas) {
        std::vector<double> wall = d.scalarList("wallAccumulators");
        if (wall.size() != size_t(nMeas) * ai.nSpecies * nWallQ) throw FoamError("resumeSampling_dsmcb200: wallAccumulators size mismatch in " + path);
        check(dsmcb200_upload_wall_accumulators(ctx_, wall.data()), "dsmcb200_upload_wall_accumulators");
    }
    if (d.found("collisionSelectionRemainder")) {
        std::vector<double> rem = d.scalarList("collisionSelectionRemainder");
        if (rem.size() == size_t(ai.nCells)) check(dsmcb200_upload_cellstate(ctx_, nullptr, rem.data()), "dsmcb200_upload_cellstate");
    }
    std::printf("Resuming sampling from %s (nTimeSteps = %g)\n", path.c_str(), d.scalar("nTimeSteps"));
}

void dsmcCloud::write() {
    const std::string timeDir = root_ + "/" + timeName_;
    // Time::writeObject (OpenFOAM v1706 Time/TimeIO.C): with purgeWrite N only the N most recent time directories written by this run
    // are kept; each rank purges its own processorN tree
    if (purgeWrite_ > 0) {
        previousWriteTimes_.push_back(timeDir);
        while (int(previousWriteTimes_.size()) > purgeWrite_) {
            std::error_code ec;
            std::filesystem::remove_all(previousWriteTimes_.front(), ec);
            previousWriteTimes_.pop_front();
        }
    }
    const std::string cdir = timeDir + "/lagrangian/" + cloudName_;
    foam::makeDirs(cdir);
    const int64_t n = nParcels();
    std::vector<double> xyz(size_t(n) * 3), U(size_t(n) * 3), ERot(static_cast<size_t>(n));
    std::vector<int32_t> cell(n), typeId(n), vib(size_t(n) * maxModes_), elevel(n), cls(n), origId(n), newParcel(size_t(n), -1);
    dsmcb200_parcels_soa s{};
    s.position = xyz.data(); s.U = U.data(); s.ERot = ERot.data(); s.cell = cell.data(); s.typeId = typeId.data(); s.vibLevel = vib.data();
    s.ELevel = elevel.data(); s.classification = cls.data(); s.origId = origId.data(); s.maxModes = maxModes_;
    std::vector<double> radialWeight(static_cast<size_t>(n), 1.0);
    s.radialWeight = radialWeight.data();
    int64_t got = 0;
    check(dsmcb200_download_parcels(ctx_, n, &got, &s), "dsmcb200_download_parcels");
    const std::string loc = timeName_ + "/lagrangian/" + cloudName_;
    // the file set of dsmcParcel::writeFields (DSMC/parcels/dsmcParcelIO.C:338-450): optional files only if non-trivial
    foam::writePositions(cdir + "/positions", loc, xyz.data(), cell.data(), got);
    foam::writeVectorField(cdir + "/U", "vectorField", loc, "U", U.data(), got);
    foam::writeLabelField(cdir + "/typeId", "labelField", loc, "typeId", typeId.data(), got);
    foam::writeLabelField(cdir + "/newParcel", "labelField", loc, "newParcel", newParcel.data(), got);
    foam::writeLabelField(cdir + "/classification", "labelField", loc, "classification", cls.data(), got);
    foam::writeLabelField(cdir + "/origId", "labelField", loc, "origId", origId.data(), got);
    std::vector<int32_t> origProc(size_t(got), rank_);
    foam::writeLabelField(cdir + "/origProcId", "labelField", loc, "origProcId", origProc.data(), got);
    bool anyRot = false, anyEl = false;
    for (int64_t i = 0; i < got; ++i) { anyRot |= ERot[i] > 0; anyEl |= elevel[i] > 0; }
    if (anyRot) foam::writeScalarField(cdir + "/ERot", "scalarField", loc, "ERot", ERot.data(), got);
    if (anyEl) foam::writeLabelField(cdir + "/ELevel", "labelField", loc, "ELevel", elevel.data(), got);
    foam::writeLabelListList(cdir + "/vibLevel", "labelFieldField", loc, "vibLevel", vib.data(), got, maxModes_);
    foam::writeScalarField(cdir + "/radialWeight", "scalarField", loc, "radialWeight", radialWeight.data(), got);
    {
        const std::string ud = timeDir + "/uniform/lagrangian/" + cloudName_;
        foam::makeDirs(ud);
        FILE* f = std::fopen((ud + "/cloudProperties").c_str(), "w");
        if (f) {
            std::fputs(foam::asciiHeader("dictionary", timeName_ + "/uniform/lagrangian/" + cloudName_, "cloudProperties").c_str(), f);
            std::fprintf(f, "processor%d\n{\n    particleCount   %lld;\n}\n", rank_, (long long)got);
            std::fclose(f);
        }
    }
    std::vector<double> sig(nCells_), rem(nCells_);
    check(dsmcb200_download_cellstate(ctx_, sig.data(), rem.data()), "dsmcb200_download_cellstate");
    std::vector<foam::PatchValues> pv;
    for (auto& b : boundary_) {
        foam::PatchValues p;
        p.name = b.name; p.type = b.type;
        if (b.type == "wall" || b.type == "patch") { p.values.resize(b.nFaces); for (int k = 0; k < b.nFaces; ++k) p.values[k] = sig[owner_[b.startFace + k]]; }
        pv.push_back(p);
    }
    foam::writeVolField(timeDir + "/dsmcSigmaTcRMax", timeName_, "dsmcSigmaTcRMax", "[0 3 -1 0 0 0 0]", sig.data(), nCells_, 1, pv);
    // AUTO_WRITE fields of the coordinate system / time-step model: RWF (dsmcAxisymmetric.C:316-328), nParticles and deltaT
    // (dsmcVariableTimeStepModel.C:103-121); boundary values are the face cells'
    auto cellField = [&](const std::string& name, const std::string& dims, const std::vector<double>& v) {
        std::vector<foam::PatchValues> bv;
        for (auto& b : boundary_) {
            foam::PatchValues p;
            p.name = b.name; p.type = b.type;
            if (b.type == "wall" || b.type == "patch") { p.values.resize(b.nFaces); for (int k = 0; k < b.nFaces; ++k) p.values[k] = v[owner_[b.startFace + k]]; }
            bv.push_back(p);
        }
        foam::writeVolField(timeDir + "/" + name, timeName_, name, dims, v.data(), nCells_, 1, bv);
    };
    if (models_.coordinateSystem != DSMCB200_COORD_CARTESIAN) cellField("RWF", "[0 0 0 0 0 0 0]", rwfCell_);
    if (variableTimeStep_) { cellField("nParticles", "[0 0 0 0 0 0 0]", nPtsCell_); cellField("deltaT", "[0 0 1 0 0 0 0]", dtCell_); }
    if (initialise_) return;  // dsmcInitialise+ writes the cloud and dsmcSigmaTcRMax only
    {
        const int S = int(species_.size());
        std::vector<double> instN(size_t(nCells_) * S, 0.0);
        for (int64_t i = 0; i < got; ++i)
            if (cell[i] >= 0 && cell[i] < nCells_ && typeId[i] >= 0 && typeId[i] < S) instN[size_t(cell[i]) * S + typeId[i]] += 1.0;
        writeFields(timeDir, instN);
    }
    // resetAtOutput / resetAtOutputUntilTime (dsmcField.C:113-152) are per field, the accumulators are shared: when every field resets the
    // engine's sums are cleared; otherwise a field that resets takes the present sums as its new baseline
    bool reset = !fields_.empty();
    for (size_t set = 0; set < std::max<size_t>(1, sampleSets_.size()); ++set) {   // the sums are shared by the fields of one sample set
        auto resets = [&](const FieldSpec& f) { return f.resetAtOutput && !(time_ + deltaT_ > f.resetAtOutputUntilTime); };
        bool all = true, some = false, members = false;
        for (auto& f : fields_) if (f.set == int(set)) { members = true; all = all && resets(f); some = some || resets(f); }
        if (!members) continue;
        reset = reset && all;
        selectSet(int(set));
        if (all) {
            check(dsmcb200_reset_accumulators(ctx_), "dsmcb200_reset_accumulators");
            for (auto& f : fields_) if (f.set == int(set)) { f.baseAcc.clear(); f.baseColl.clear(); f.baseWall.clear(); f.baseNT = 0; }
        } else if (some) {
            dsmcb200_accum_info ai{};
            check(dsmcb200_accum_info_get(ctx_, &ai), "dsmcb200_accum_info_get");
            std::vector<double> acc(size_t(ai.nCells) * ai.nSpecies * ai.nQuantities), coll(size_t(ai.nCells) * 2);
            check(dsmcb200_download_accumulators(ctx_, acc.data(), coll.data()), "dsmcb200_download_accumulators");
            int32_t nMeas = 0, nWallQ = 0;
            check(dsmcb200_wall_info(ctx_, &nMeas, &nWallQ), "dsmcb200_wall_info");
            std::vector<double> wall(size_t(std::max(nMeas, 1)) * ai.nSpecies * std::max(nWallQ, 1), 0.0);
            if (nMeas) check(dsmcb200_download_wall_accumulators(ctx_, wall.data()), "dsmcb200_download_wall_accumulators");
            for (auto& f : fields_)
                if (f.set == int(set) && resets(f)) { f.baseAcc = acc; f.baseColl = coll; f.baseWall = wall; f.baseNT = ai.nTimeSteps; }
        }
    }
    selectSet(0);
    // dsmcVolFields.C:2375-2378: only with averagingAcrossManyRuns and resetAtOutput off
    if (!reset) writeResumeSampling(timeDir);
}

}  // namespace dsmcb200
