// dsmc_cloud.h -- host-side mirror of the reference's dsmcCloud and of the run-time-selection surface
// it exposes (level 1), for callers that are not OpenFOAM: the standalone driver
// dsmcb200_run reads an unchanged dsmcFoam+ case directory with it.
//
// Mirrors  Foam::dsmcCloud            DSMC/clouds/dsmcCloud.H:76,257-263,574-681  (ctor, evolve, info, nTerminalOutputs)
//          BinaryCollisionModel::New  DSMC/collisions/basic/BinaryCollisionModel/BinaryCollisionModel.C:59-98
//          collisionPartnerSelection::New  DSMC/collisionPartnerSelection/basic/collisionPartnerSelection.C:60-95
//          dsmcBoundaries / dsmcPatchBoundary::New / dsmcGeneralBoundary::New
//                                     DSMC/boundaries/basic/dsmcBoundaries/dsmcBoundaries.C:82-520
//          dsmcFieldProperties / dsmcVolFields  DSMC/macroscopicProperties/...  (createField, calculateField, writeField)
// All compute goes through the C ABI (include/dsmcb200.h); nothing here touches CUDA directly.
#pragma once
#include <cstdint>
#include <map>
#include <deque>
#include <string>
#include <vector>

#include "../../include/dsmcb200.h"
#include "foam_io.h"

namespace dsmcb200 {

// name -> enum tables with the reference's "unknown ... type / Valid ... types are" failure
int selectBinaryCollisionModel(const std::string& name);
int selectCollisionPartnerSelection(const std::string& name);
int selectPatchBoundaryModel(const std::string& name);
void selectGeneralBoundaryModel(const std::string& name);
void selectFieldModel(const std::string& name);
int selectCoordinateSystem(const std::string& name);   // dsmcb200_coordinate_system
bool selectTimeStepModel(const std::string& name);     // true: variable
int patchTypeFromWord(const std::string& type);

struct FieldSpec {  // one dsmcVolFields entry of system/fieldPropertiesDict
    std::string fieldName;
    std::vector<int> typeIds;
    bool measureMeanFreePath = false, densityOnly = false, measureHeatFluxShearStress = false, measureClassifications = false, measureErrors = false;
    double mfpReferenceTemperature = 273.0;
    bool resetAtOutput = true;
    double resetAtOutputUntilTime = 1e300;
    int sampleInterval = 1;
    int set = 0;                           // the engine's sample set of this field's sampleInterval (dsmcb200_set_sample_sets)
    bool averagingAcrossManyRuns = false;  // dsmcVolFields.C:1048: keep / restore uniform/resumeSampling_<fieldName>
    // the shared accumulators as of this field's last reset (empty: zero), see dsmcCloud::write
    std::vector<double> baseAcc, baseColl, baseWall;
    double baseNT = 0;
};

struct DerivedFields {  // per-cell results of dsmcVolFields::calculateField for one instance
    std::vector<double> dsmcNMean, rhoN, rhoM, p, Ttra, Trot, Tvib, Tov, Ma, mfp, mct, mctToDt, mfpToDx, SOF, measuredCollisionRate;
    std::vector<double> UMean;  // [3n]
    std::vector<double> pressureTensor, shearStressTensor;  // [9n] (measureHeatFluxShearStress)
    std::vector<double> heatFluxVector;                     // [3n]
    std::vector<double> rhoMError, UError, TError, pError;  // measureErrors
};

class dsmcCloud {
   public:
    dsmcCloud(const std::string& caseDir, const std::string& cloudName = "dsmc", int rank = 0, int nRanks = 1, int device = 0,
              const void* ncclId128 = nullptr, bool dryRun = false, bool initialise = false);
    ~dsmcCloud();
    // cell labels inside the engine for clouds constructed from now on (dsmcb200_set_cell_order; the case files keep their labels)
    static void cellOrder(int mode);
    dsmcCloud(const dsmcCloud&) = delete;

    void evolve();                 // dsmcCloud::evolve(): one time step
    void info();                   // dsmcCloud::info(): the 8-line summary (same strings as the reference)
    int nTerminalOutputs() const { return nTerminalOutputs_; }
    bool loop();                   // runTime.loop()
    bool outputTime() const;       // runTime.outputTime()
    void write();                  // runTime.write(): cloud, dsmcSigmaTcRMax, fields
    const std::string& timeName() const { return timeName_; }
    double time() const { return time_; }
    int64_t nParcels();
    int nCells() const { return nCells_; }
    const std::vector<FieldSpec>& fields() const { return fields_; }
    DerivedFields calculateField(const FieldSpec& f);   // from the current accumulators
    dsmcb200_ctx* ctx() { return ctx_; }
    const std::vector<std::string>& typeIdList() const { return typeIdList_; }

   private:
    void readControl();
    void readMesh();
    void readProperties();
    void setCellFields();
    void readReactions();
    void readBoundaries();
    void readFieldProperties();
    void readCloud();
    // instN: [nCells][nSpecies] parcels in the cells right now (dsmcN_, the AUTO_WRITE instantaneous count of dsmcVolFields.C:106-117)
    void writeFields(const std::string& timeDir, const std::vector<double>& instN);
    double cellMaxDx(int c) const;  // largest extent of the cell's points along x, y, z (dsmcVolFields.C:1795-1821)
    void writeResumeSampling(const std::string& timeDir);  // dsmcVolFields::writeOut + the engine's own lossless checkpoint
    void writeResumeSamplingOf(const std::string& timeDir, int set);
    void readResumeSampling();                              // dsmcVolFields::readIn
    void readResumeSamplingOf(int set);
    void selectSet(int set);                                // the sample set the accumulator calls act on
    std::vector<int32_t> sampleSets_;                       // distinct sampleIntervals of the field{} entries, in order of appearance
    void check(int rc, const char* what);

    std::string caseDir_, root_, cloudName_, timeName_;
    int rank_, nRanks_;
    dsmcb200_ctx* ctx_ = nullptr;
    // time
    double time_ = 0, startTime_ = 0, endTime_ = 0, deltaT_ = 0, writeInterval_ = 1;
    std::string writeControl_ = "timeStep";
    int timePrecision_ = 6, nTerminalOutputs_ = 1, infoCounter_ = 0;
    int purgeWrite_ = 0;                          // controlDict purgeWrite: time directories kept (0 = all)
    std::deque<std::string> previousWriteTimes_;  // FIFO of the directories this run wrote
    int64_t timeIndex_ = 0, startIndex_ = 0;
    // mesh
    int nCells_ = 0, nFaces_ = 0, nInternal_ = 0;
    std::vector<double> points_;
    std::vector<int32_t> faceOffsets_, facePoints_, owner_, neighbour_;
    std::vector<foam::BoundaryPatch> boundary_;
    std::vector<dsmcb200_patch> patches_;
    std::vector<double> cellVolumes_, cellCentres_, faceAreas_, faceCentres_;
    mutable std::vector<double> cellMaxDx_;
    // models
    std::vector<std::string> typeIdList_;
    std::vector<dsmcb200_species> species_;
    std::vector<dsmcb200_reaction> reactions_;     // system/chemReactDict
    // coordinate system / time-step model: per-cell nParticles (time-step model), deltaT and radial weighting factor
    bool variableTimeStep_ = false;
    int polarAxis_ = 1;
    double maxRWF_ = 1.0, radialExtent_ = 0.0;
    std::vector<double> nPtsCell_, dtCell_, rwfCell_;
    std::vector<std::string> reactionNames_;
    dsmcb200_models models_{};
    std::vector<dsmcb200_patch_model> patchModels_;
    std::vector<dsmcb200_inflow> inflows_;
    std::vector<FieldSpec> fields_;
    int maxModes_ = 1;
    int64_t lastCollisions_ = 0;
    bool dryRun_ = false;
    bool initialise_ = false;   // dsmcInitialise+: no cloud is read, system/dsmcInitialiseDict fills the mesh
    void initialiseFromDict();  // dsmcAllConfigurations::setInitialConfig (dsmcCloud.C:795-796)
    int64_t nRead_ = 0;
    double origin_[3] = {0, 0, 0};   // sphericalProperties origin (dsmcSpherical)
    uint64_t meshHash_ = 0, cloudHash_ = 0;   // FNV-1a over the bytes read (printed by summary(): the same case in ASCII and binary reads the same)

   public:
    // -dryRun: what was parsed from the case directory (no GPU context is created)
    std::string summary() const;
};

}  // namespace dsmcb200
