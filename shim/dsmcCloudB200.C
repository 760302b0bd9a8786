/*---------------------------------------------------------------------------*\
  dsmcCloudB200.C -- see dsmcCloudB200.H.  Every compute step is a call of the C ABI (include/dsmcb200.h);
  this file only translates OpenFOAM objects (mesh, dictionaries, lagrangian files, volFields) to and from it.
\*---------------------------------------------------------------------------*/
#include "dsmcCloudB200.H"

#include "passiveParticleCloud.H"
#include "IOField.H"
#include "processorPolyPatch.H"
#include "processorCyclicPolyPatch.H"
#include "cyclicPolyPatch.H"
#include "wallPolyPatch.H"
#include "emptyPolyPatch.H"
#include "symmetryPolyPatch.H"
#include "symmetryPlanePolyPatch.H"
#include "wedgePolyPatch.H"
#include "zeroGradientFvPatchFields.H"
#include "calculatedFvPatchFields.H"

#include <cstring>
#include <cuda_runtime_api.h>   // cudaGetDeviceCount only: which device this rank uses

namespace Foam
{
    defineTypeNameAndDebug(dsmcCloud, 0);
}

// * * * * * * * * * * * * * * * * helpers  * * * * * * * * * * * * * * * * * //

void Foam::dsmcCloud::ck(const int rc, const char* what) const
{
    // nothing throws across the ABI: status codes become FatalErrors here
    if (rc)
    {
        FatalErrorIn("dsmcCloud (dsmcb200)")
            << what << ": " << dsmcb200_last_error(ctx_) << " (status " << rc << ")" << nl
            << exit(FatalError);
    }
}


Foam::label Foam::dsmcCloud::lookupOrFail(const HashTable<label>& table, const word& name, const char* family)
{
    // the failure text of the reference's New() selectors (e.g. BinaryCollisionModel.C:77-87)
    HashTable<label>::const_iterator it = table.find(name);
    if (it == table.end())
    {
        FatalErrorIn("dsmcCloud (dsmcb200)")
            << "Unknown " << family << " type " << name << nl << nl
            << "Valid " << family << " types are:" << nl
            << table.sortedToc() << nl
            << exit(FatalError);
    }
    return it();
}


Foam::label Foam::dsmcCloud::typeIdOf(const word& name) const
{
    const label id = findIndex(typeIdList_, name);
    if (id == -1)
    {
        FatalErrorIn("dsmcCloud (dsmcb200)")
            << "Cannot find typeId: " << name << " in typeIdList " << typeIdList_ << nl
            << exit(FatalError);
    }
    return id;
}


void Foam::dsmcCloud::fillPatch(dsmcb200_patch& out, const polyPatch& pp, const polyMesh& mesh)
{
    std::memset(&out, 0, sizeof(out));
    std::strncpy(out.name, pp.name().c_str(), DSMCB200_NAME_LEN - 1);
    out.start = pp.start();
    out.size = pp.size();
    out.neighbPatch = -1; out.myProcNo = -1; out.neighbProcNo = -1; out.referPatch = -1;

    // most derived types first: processorCyclic is a processor, symmetryPlane is not a symmetry
    if (isA<processorCyclicPolyPatch>(pp))
    {
        const processorCyclicPolyPatch& p = refCast<const processorCyclicPolyPatch>(pp);
        out.type = DSMCB200_PATCH_PROCESSORCYCLIC;
        out.myProcNo = p.myProcNo(); out.neighbProcNo = p.neighbProcNo(); out.referPatch = p.referPatchID();
        // the receiving side subtracts the separation (particle::correctAfterParallelTransfer, particleTemplates.C:97-121)
        if (p.separated())
        {
            const vector& s = (p.separation().size() == 1) ? p.separation()[0] : p.separation()[0];
            out.separation[0] = s.x(); out.separation[1] = s.y(); out.separation[2] = s.z();
            out.hasSeparation = 1;
        }
        if (!p.parallel())
        {
            FatalErrorIn("dsmcCloud (dsmcb200)") << "rotational processorCyclic patch " << pp.name() << " is not supported" << exit(FatalError);
        }
    }
    else if (isA<processorPolyPatch>(pp))
    {
        const processorPolyPatch& p = refCast<const processorPolyPatch>(pp);
        out.type = DSMCB200_PATCH_PROCESSOR;
        out.myProcNo = p.myProcNo(); out.neighbProcNo = p.neighbProcNo();
    }
    else if (isA<cyclicPolyPatch>(pp))
    {
        const cyclicPolyPatch& p = refCast<const cyclicPolyPatch>(pp);
        out.type = DSMCB200_PATCH_CYCLIC;
        out.neighbPatch = p.neighbPatchID();
        if (!p.parallel())
        {
            FatalErrorIn("dsmcCloud (dsmcb200)") << "rotational cyclic patch " << pp.name() << " is not supported" << exit(FatalError);
        }
        // particle::hitCyclicPatch: position -= receiving patch's separation (particleTemplates.C:1553-1566)
        if (p.separated())
        {
            const vector& s = p.separation()[0];
            out.separation[0] = s.x(); out.separation[1] = s.y(); out.separation[2] = s.z();
            out.hasSeparation = 1;
        }
    }
    else if (isA<wallPolyPatch>(pp))            { out.type = DSMCB200_PATCH_WALL; }
    else if (isA<emptyPolyPatch>(pp))           { out.type = DSMCB200_PATCH_EMPTY; }
    else if (isA<symmetryPlanePolyPatch>(pp))   { out.type = DSMCB200_PATCH_SYMMETRYPLANE; }
    else if (isA<symmetryPolyPatch>(pp))        { out.type = DSMCB200_PATCH_SYMMETRY; }
    else if (isA<wedgePolyPatch>(pp))           { out.type = DSMCB200_PATCH_WEDGE; }
    else if (pp.type() == polyPatch::typeName)  { out.type = DSMCB200_PATCH_PATCH; }
    else
    {
        FatalErrorIn("dsmcCloud (dsmcb200)")
            << "patch " << pp.name() << " of type " << pp.type() << " is not handled by the tracker" << nl
            << "Valid patch types are: patch wall cyclic empty symmetry symmetryPlane wedge processor processorCyclic" << nl
            << exit(FatalError);
    }
}


void Foam::dsmcCloud::sendMesh()
{
    const polyMesh& mesh = mesh_;
    // OpenFOAM's own geometry is handed over, so centres / volumes / tet base points are the reference's by construction
    labelList faceOffsets(mesh.nFaces() + 1, 0);
    forAll(mesh.faces(), f) { faceOffsets[f + 1] = faceOffsets[f] + mesh.faces()[f].size(); }
    labelList facePoints(faceOffsets[mesh.nFaces()]);
    forAll(mesh.faces(), f)
    {
        const face& fc = mesh.faces()[f];
        forAll(fc, i) { facePoints[faceOffsets[f] + i] = fc[i]; }
    }
    List<dsmcb200_patch> patches(mesh.boundaryMesh().size());
    forAll(mesh.boundaryMesh(), p) { fillPatch(patches[p], mesh.boundaryMesh()[p], mesh); }

    dsmcb200_mesh m;
    std::memset(&m, 0, sizeof(m));
    m.nPoints = mesh.nPoints(); m.nFaces = mesh.nFaces(); m.nInternalFaces = mesh.nInternalFaces(); m.nCells = mesh.nCells();
    m.nPatches = patches.size();
    // vector is three contiguous scalars, label is int32 in the builds the ABI is declared for (WM_LABEL_SIZE=32, WM_PRECISION_OPTION=DP)
    m.points = reinterpret_cast<const double*>(mesh.points().begin());
    m.faceOffsets = faceOffsets.begin(); m.facePoints = facePoints.begin();
    m.owner = mesh.faceOwner().begin(); m.neighbour = mesh.faceNeighbour().begin();
    m.patches = patches.begin();
    m.cellCentres = reinterpret_cast<const double*>(mesh.cellCentres().begin());
    m.cellVolumes = mesh.cellVolumes().begin();
    m.faceCentres = reinterpret_cast<const double*>(mesh.faceCentres().begin());
    m.faceAreas = reinterpret_cast<const double*>(mesh.faceAreas().begin());
    m.tetBasePtIs = mesh.tetBasePtIs().begin();
    // optional, constant/dsmcProperties: `dsmcb200CellOrder zCurve;` relabels the cells inside the library (renumberMesh in memory; the
    // fields and the cloud this class writes keep the mesh's labels)
    const word order(particleProperties_.lookupOrDefault<word>("dsmcb200CellOrder", "asGiven"));
    if (order == "zCurve") { ck(dsmcb200_set_cell_order(ctx_, DSMCB200_CELL_ORDER_Z_CURVE, NULL, 0), "dsmcb200_set_cell_order"); }
    else if (order != "asGiven")
    {
        FatalErrorIn("dsmcCloud::sendMesh()") << "dsmcb200CellOrder " << order << " is not in enumeration: 2(asGiven zCurve)" << exit(FatalError);
    }
    ck(dsmcb200_set_mesh(ctx_, &m), "dsmcb200_set_mesh");
}


void Foam::dsmcCloud::readSpecies()
{
    // dsmcCloud::buildConstProps (dsmcCloud.C:77-108) + dsmcParcel::constantProperties (dsmcParcelI.H:37-200)
    typeIdList_ = wordList(particleProperties_.lookup("typeIdList"));
    const dictionary& moleculeProperties = particleProperties_.subDict("moleculeProperties");
    if (typeIdList_.size() > DSMCB200_MAX_SPECIES)
    {
        FatalErrorIn("dsmcCloud (dsmcb200)") << "at most " << DSMCB200_MAX_SPECIES << " species" << exit(FatalError);
    }
    species_.setSize(typeIdList_.size());
    maxModes_ = 1;
    forAll(typeIdList_, i)
    {
        const word& id = typeIdList_[i];
        Info<< "    " << id << endl;
        const dictionary& d = moleculeProperties.subDict(id);
        dsmcb200_species& s = species_[i];
        std::memset(&s, 0, sizeof(s));
        std::strncpy(s.name, id.c_str(), DSMCB200_NAME_LEN - 1);
        s.mass = readScalar(d.lookup("mass"));
        s.diameter = readScalar(d.lookup("diameter"));
        s.omega = readScalar(d.lookup("omega"));
        s.alpha = d.lookupOrDefault<scalar>("alpha", 1.0);
        s.rotationalDegreesOfFreedom = d.lookupOrDefault<scalar>("rotationalDegreesOfFreedom", 0);
        s.nVibrationalModes = label(d.lookupOrDefault<scalar>("nVibrationalModes", 0));
        const scalarList thetaV(d.lookupOrDefault<scalarList>("characteristicVibrationalTemperature", scalarList()));
        const scalarList Zref(d.lookupOrDefault<scalarList>("Zref", scalarList()));
        const scalarList TrefZv(d.lookupOrDefault<scalarList>("referenceTempForZref", scalarList()));
        if (thetaV.size() != s.nVibrationalModes)
        {
            FatalErrorIn("dsmcParcel::constantProperties::constantProperties")
                << "Number of characteristic vibrational temperatures is " << thetaV.size() << ", instead of " << s.nVibrationalModes << nl
                << exit(FatalError);
        }
        if (Zref.size() != s.nVibrationalModes || TrefZv.size() != s.nVibrationalModes)
        {
            FatalErrorIn("dsmcParcel::constantProperties::constantProperties")
                << "Number of reference vibrational relaxation numbers / temperatures is " << Zref.size() << " / " << TrefZv.size()
                << ", instead of " << s.nVibrationalModes << nl << exit(FatalError);
        }
        if (s.nVibrationalModes > DSMCB200_MAX_VIB_MODES)
        {
            FatalErrorIn("dsmcCloud (dsmcb200)") << "at most " << DSMCB200_MAX_VIB_MODES << " vibrational modes per species" << exit(FatalError);
        }
        for (label m = 0; m < s.nVibrationalModes; ++m) { s.thetaV[m] = thetaV[m]; s.Zref[m] = Zref[m]; s.TrefZv[m] = TrefZv[m]; }
        maxModes_ = max(maxModes_, label(s.nVibrationalModes));
        s.thetaD = d.lookupOrDefault<scalar>("dissociationTemperature", 0.0);
        s.charge = d.lookupOrDefault<label>("charge", 0);
        s.nElectronicLevels = d.lookupOrDefault<label>("nElectronicLevels", 1);
        const scalarList eList(d.lookupOrDefault<scalarList>("electronicEnergyList", scalarList(1, 0.0)));
        const labelList gList(d.lookupOrDefault<labelList>("electronicDegeneracyList", labelList(label(1), 1)));
        if (eList.size() != s.nElectronicLevels || gList.size() != s.nElectronicLevels)
        {
            FatalErrorIn("dsmcParcel::constantProperties::constantProperties")
                << "Number of energy / degeneracy levels should be " << s.nElectronicLevels << ", instead of "
                << eList.size() << " / " << gList.size() << nl << exit(FatalError);
        }
        if (s.nElectronicLevels > DSMCB200_MAX_ELEC_LEVELS)
        {
            FatalErrorIn("dsmcCloud (dsmcb200)") << "at most " << DSMCB200_MAX_ELEC_LEVELS << " electronic levels per species" << exit(FatalError);
        }
        forAll(eList, l) { s.electronicEnergyList[l] = eList[l]; s.electronicDegeneracyList[l] = gList[l]; }
    }
    ck(dsmcb200_set_species(ctx_, species_.size(), species_.begin()), "dsmcb200_set_species");
    readReactions();
}


void Foam::dsmcCloud::readReactions()
{
    // dsmcReactions ctor (dsmcReactions.C:69-118): system/chemReactDict `reactions ( name { reactionModel M; reactants (A B); ... } )`
    // for the three quantum-kinetic models of the engine (dissociationQK.C:44-195, exchangeQK.C:44-176)
    IOobject header("chemReactDict", time_.system(), mesh_, IOobject::READ_IF_PRESENT, IOobject::NO_WRITE);
    if (!header.headerOk()) return;
    const IOdictionary chem(header);
    if (!chem.found("reactions")) return;
    Info<< nl << "Creating dsmcReactions" << nl << endl;
    const PtrList<entry> reactionList(chem.lookup("reactions"));
    HashTable<label> reactionModels;
    reactionModels.insert("dissociationQK", DSMCB200_REACT_DISSOCIATION_QK);
    reactionModels.insert("exchangeQK", DSMCB200_REACT_EXCHANGE_QK);
    reactionModels.insert("dissociationExchangeQK", DSMCB200_REACT_DISSOCIATION_EXCHANGE_QK);
    reactions_.setSize(reactionList.size());
    forAll(reactionList, r)
    {
        const dictionary& dict = reactionList[r].dict();
        const word name(reactionList[r].keyword());
        const word model(dict.lookup("reactionModel"));
        Info<< "Selecting the reaction model " << model << endl;
        dsmcb200_reaction& R = reactions_[r];
        std::memset(&R, 0, sizeof(R));
        R.model = lookupOrFail(reactionModels, model, "dsmc reaction model");
        const wordList reactants(dict.lookup("reactants"));
        if (reactants.size() != 2)
        {
            FatalErrorIn("dsmcReaction::setProperties()") << "For reaction named " << name << nl
                << "There should be two reactants, instead of " << reactants.size() << nl << exit(FatalError);
        }
        forAll(reactants, k)
        {
            R.reactants[k] = findIndex(typeIdList_, reactants[k]);
            if (R.reactants[k] == -1)
            {
                FatalErrorIn("dsmcReaction::setProperties()") << "For reaction named " << name << nl
                    << "Cannot find type id: " << reactants[k] << nl << exit(FatalError);
            }
        }
        R.allowSplitting = Switch(dict.lookupOrDefault<Switch>("allowSplitting", true)) ? 1 : 0;
        for (label k = 0; k < 2; ++k) { R.dissociationProducts[k][0] = R.dissociationProducts[k][1] = -1; R.exchangeProducts[k] = -1; }
        if (R.model != DSMCB200_REACT_EXCHANGE_QK)
        {
            const List<wordList> products(dict.subDict("dissociationQKProperties").lookup("dissociationProducts"));
            if (products.size() != 2)
            {
                FatalErrorIn("dissociationQK::setProperties()") << "For reaction named " << name << nl
                    << "There should be two lists of products, instead of " << products.size() << nl
                    << "NB: a list can be left empty" << nl << exit(FatalError);
            }
            forAll(products, k)
            {
                if (products[k].size() != 0 && products[k].size() != 2)
                {
                    FatalErrorIn("dissociationQK::setProperties()") << "For reaction named " << name << nl
                        << "There should be 2 dissociation products for molecule " << reactants[k] << " instead of "
                        << products[k].size() << ", that is " << products[k] << exit(FatalError);
                }
                forAll(products[k], q)
                {
                    R.dissociationProducts[k][q] = findIndex(typeIdList_, products[k][q]);
                    if (R.dissociationProducts[k][q] == -1)
                    {
                        FatalErrorIn("dissociationQK::setProperties()") << "For reaction named " << name << nl
                            << "Cannot find type id: " << products[k][q] << nl << exit(FatalError);
                    }
                }
            }
        }
        if (R.model != DSMCB200_REACT_DISSOCIATION_QK)
        {
            const dictionary& x = dict.subDict("exchangeQKProperties");
            const wordList products(x.lookup("exchangeProducts"));
            if (products.size() != 2)
            {
                FatalErrorIn("exchangeQK::setProperties()") << "For reaction named " << name << nl
                    << "There should be two products, instead of " << products.size() << nl << exit(FatalError);
            }
            forAll(products, k)
            {
                R.exchangeProducts[k] = findIndex(typeIdList_, products[k]);
                if (R.exchangeProducts[k] == -1)
                {
                    FatalErrorIn("exchangeQK::setProperties()") << "For reaction named " << name << nl
                        << "Cannot find type id: " << products[k] << nl << exit(FatalError);
                }
            }
            R.heatOfReactionExchange = readScalar(x.lookup("heatOfReactionExchange"));
            R.aCoeff = readScalar(x.lookup("aCoeff"));
            R.bCoeff = readScalar(x.lookup("bCoeff"));
        }
    }
    if (reactions_.size())
    {
        Info<< "Number of reactions created: " << reactions_.size() << endl;
        // the remaining checks of <model>::setProperties (molecule / atom roles, product types) and the typeId-pair addressing of
        // dsmcReactions::initialConfiguration are applied by the engine with the reference's messages
        ck(dsmcb200_set_reactions(ctx_, reactions_.size(), reactions_.begin()), "dsmcb200_set_reactions");
    }
    else
    {
        Info<< "There are no chemical reactions defined." << endl;
    }
}


void Foam::dsmcCloud::readBoundaries()
{
    // dsmcBoundaries (dsmcBoundaries.C:82-520): three lists of `boundary { ...Properties { patchName } boundaryModel M; MProperties {} }`
    HashTable<label> patchModels;
    patchModels.insert("dsmcDiffuseWallPatch", DSMCB200_BND_DIFFUSE_WALL);
    patchModels.insert("dsmcSpecularWallPatch", DSMCB200_BND_SPECULAR_WALL);
    patchModels.insert("dsmcDiffuseSpecularWallPatch", DSMCB200_BND_DIFFUSE_SPECULAR_WALL);
    patchModels.insert("dsmcDeletionPatch", DSMCB200_BND_DELETION);
    patchModels.insert("dsmcCLLWallPatch", DSMCB200_BND_CLL_WALL);
    HashTable<label> generalModels;
    generalModels.insert("dsmcFreeStreamInflowPatch", 1);

    const PtrList<entry> pList(boundariesDict_.lookup("dsmcPatchBoundaries"));
    const PtrList<entry> cList(boundariesDict_.lookup("dsmcCyclicBoundaries"));
    const PtrList<entry> gList(boundariesDict_.lookup("dsmcGeneralBoundaries"));
    if (cList.size())
    {
        // dsmcCyclicBoundary models (dsmcReflectiveParticleMembranePatch) act on cyclic crossings: not part of the hot path
        FatalErrorIn("dsmcCloud (dsmcb200)") << "dsmcCyclicBoundaries models are not supported; leave the list empty" << exit(FatalError);
    }
    patchModels_.setSize(pList.size());
    forAll(pList, i)
    {
        const dictionary& d = pList[i].dict();
        const word patchName(d.subDict("patchBoundaryProperties").lookup("patchName"));
        const word model(d.lookup("boundaryModel"));
        const label patchId = mesh_.boundaryMesh().findPatchID(patchName);
        if (patchId == -1)
        {
            FatalErrorIn("dsmcPatchBoundary::dsmcPatchBoundary") << "Cannot find patch: " << patchName << nl << "in: " << boundariesDict_.name() << exit(FatalError);
        }
        dsmcb200_patch_model& pm = patchModels_[i];
        std::memset(&pm, 0, sizeof(pm));
        pm.patch = patchId;
        pm.model = lookupOrFail(patchModels, model, "dsmcPatchBoundary");
        pm.diffuseFraction = 1.0;
        pm.depthAxis = 1;
        if (pm.model == DSMCB200_BND_DIFFUSE_WALL || pm.model == DSMCB200_BND_DIFFUSE_SPECULAR_WALL)
        {
            const dictionary& p = d.subDict(model + "Properties");
            const vector v(p.lookup("velocity"));
            pm.velocity[0] = v.x(); pm.velocity[1] = v.y(); pm.velocity[2] = v.z();
            // dsmcDiffuseWallPatch.C:49-64,169-187: temperature | groundLevelTemperature (+ formationLevelTemperature, depthAxis)
            if (p.found("groundLevelTemperature"))
            {
                pm.temperature = readScalar(p.lookup("groundLevelTemperature"));
                pm.formationLevelTemperature = p.lookupOrDefault<scalar>("formationLevelTemperature", pm.temperature);
                pm.linearTemperature = pm.formationLevelTemperature != pm.temperature;
                const word axis(p.lookupOrDefault<word>("depthAxis", "y"));
                pm.depthAxis = (axis == "x") ? 0 : (axis == "z") ? 2 : 1;
            }
            else
            {
                pm.temperature = readScalar(p.lookup("temperature"));
                pm.formationLevelTemperature = pm.temperature;
            }
            if (pm.model == DSMCB200_BND_DIFFUSE_SPECULAR_WALL) { pm.diffuseFraction = readScalar(p.lookup("diffuseFraction")); }
        }
        if (pm.model == DSMCB200_BND_CLL_WALL)
        {
            // dsmcCLLWallPatch.C:45-75,330-334
            const dictionary& p = d.subDict(model + "Properties");
            const vector v(p.lookup("velocity"));
            pm.velocity[0] = v.x(); pm.velocity[1] = v.y(); pm.velocity[2] = v.z();
            pm.temperature = readScalar(p.lookup("temperature"));
            pm.normalAccommodationCoefficient = readScalar(p.lookup("normalAccommodationCoefficient"));
            pm.tangentialAccommodationCoefficient = readScalar(p.lookup("tangentialAccommodationCoefficient"));
            pm.rotationalEnergyAccommodationCoefficient = readScalar(p.lookup("rotationalEnergyAccommodationCoefficient"));
            readScalar(p.lookup("vibrationalEnergyAccommodationCoefficient"));   // mandatory in the reference, used nowhere
        }
    }
    inflows_.setSize(gList.size());
    forAll(gList, i)
    {
        const dictionary& d = gList[i].dict();
        const word patchName(d.subDict("generalBoundaryProperties").lookup("patchName"));
        const word model(d.lookup("boundaryModel"));
        lookupOrFail(generalModels, model, "dsmcGeneralBoundary");
        const label patchId = mesh_.boundaryMesh().findPatchID(patchName);
        if (patchId == -1)
        {
            FatalErrorIn("dsmcGeneralBoundary::dsmcGeneralBoundary") << "Cannot find patch: " << patchName << exit(FatalError);
        }
        // dsmcFreeStreamInflowPatch::setProperties (dsmcFreeStreamInflowPatch.C:393-470)
        const dictionary& p = d.subDict(model + "Properties");
        dsmcb200_inflow& in = inflows_[i];
        std::memset(&in, 0, sizeof(in));
        in.patch = patchId;
        const vector v(p.lookup("velocity"));
        in.velocity[0] = v.x(); in.velocity[1] = v.y(); in.velocity[2] = v.z();
        in.translationalTemperature = readScalar(p.lookup("translationalTemperature"));
        in.rotationalTemperature = p.lookupOrDefault<scalar>("rotationalTemperature", 0.0);
        in.vibrationalTemperature = p.lookupOrDefault<scalar>("vibrationalTemperature", 0.0);
        in.electronicTemperature = p.lookupOrDefault<scalar>("electronicTemperature", 0.0);
        const wordList molecules(p.lookup("typeIds"));
        if (molecules.size() == 0)
        {
            FatalErrorIn("dsmcFreeStreamInflowPatch::setProperties()") << "Cannot have zero typeIds being inserted." << exit(FatalError);
        }
        DynamicList<word> reduced(0);
        forAll(molecules, k) { if (findIndex(reduced, molecules[k]) == -1) { reduced.append(molecules[k]); } }
        const dictionary& nd = p.subDict("numberDensities");
        in.nTypes = reduced.size();
        forAll(reduced, k)
        {
            in.typeIds[k] = typeIdOf(reduced[k]);
            in.numberDensities[k] = readScalar(nd.lookup(reduced[k]));
        }
    }
}


void Foam::dsmcCloud::readFieldProperties()
{
    // dsmcFieldProperties (dsmcFieldProperties.C:60-130): dsmcFields ( field { fieldModel dsmcVolFields; timeProperties{} dsmcVolFieldsProperties{} } )
    HashTable<label> fieldModels;
    fieldModels.insert("dsmcVolFields", 1);
    const PtrList<entry> fList(fieldPropertiesDict_.lookup("dsmcFields"));
    fields_.setSize(fList.size());
    models_.sampleInterval = 1;
    sampleIntervals_.clear();
    forAll(fList, i)
    {
        const dictionary& d = fList[i].dict();
        lookupOrFail(fieldModels, word(d.lookup("fieldModel")), "dsmcField");
        const dictionary& p = d.subDict("dsmcVolFieldsProperties");
        fieldSpec& f = fields_[i];
        f.fieldName = word(p.lookup("fieldName"));
        const wordList ids(p.lookup("typeIds"));
        f.typeIds.setSize(ids.size());
        forAll(ids, k) { f.typeIds[k] = typeIdOf(ids[k]); }
        const dictionary& tp = d.subDict("timeProperties");
        f.resetAtOutput = Switch(tp.lookupOrDefault<Switch>("resetAtOutput", true));
        f.resetAtOutputUntilTime = tp.lookupOrDefault<scalar>("resetAtOutputUntilTime", VGREAT);
        // every field samples on its own cadence (dsmcVolFields.C:1073-1081): fields with the same sampleInterval share one set of the
        // library's sums, resetAtOutput / resetAtOutputUntilTime stay per field through the baselines below (dsmcField.C:113-152)
        const label si = max(label(1), p.lookupOrDefault<label>("sampleInterval", 1));
        label set = -1;
        forAll(sampleIntervals_, k) { if (sampleIntervals_[k] == si) { set = k; } }
        if (set < 0)
        {
            if (sampleIntervals_.size() == 8)
            {
                FatalErrorIn("dsmcCloud (dsmcb200)") << "field " << f.fieldName << ": more than 8 different sampleIntervals" << exit(FatalError);
            }
            set = sampleIntervals_.size();
            sampleIntervals_.append(si);
        }
        f.sampleSet = set;
        f.baseNT = 0;
        if (Switch(p.lookupOrDefault<Switch>("measureHeatFluxShearStress", false))) { models_.measureHeatFluxShearStress = 1; }
        if (Switch(p.lookupOrDefault<Switch>("measureClassifications", false))) { models_.measureClassifications = 1; }
    }
    if (sampleIntervals_.size()) { models_.sampleInterval = sampleIntervals_[0]; }
    if (sampleIntervals_.size() > 1)
    {
        ck(dsmcb200_set_sample_sets(ctx_, sampleIntervals_.size(), sampleIntervals_.begin()), "dsmcb200_set_sample_sets");
    }
}


void Foam::dsmcCloud::selectSampleSet(const label set) const
{
    if (sampleIntervals_.size() > 1) { ck(dsmcb200_select_sample_set(ctx_, set), "dsmcb200_select_sample_set"); }
}


void Foam::dsmcCloud::readModels()
{
    std::memset(&models_, 0, sizeof(models_));
    HashTable<label> collisionModels;
    collisionModels.insert("NoBinaryCollision", DSMCB200_COLL_NONE);
    collisionModels.insert("VariableHardSphere", DSMCB200_COLL_VHS);
    collisionModels.insert("LarsenBorgnakkeVariableHardSphere", DSMCB200_COLL_LB_VHS);
    collisionModels.insert("VariableSoftSphere", DSMCB200_COLL_VSS);
    collisionModels.insert("LarsenBorgnakkeVariableSoftSphere", DSMCB200_COLL_LB_VSS);
    HashTable<label> partnerModels;
    partnerModels.insert("noTimeCounter", 1);
    HashTable<label> coordinateSystems;
    coordinateSystems.insert("dsmcCartesian", DSMCB200_COORD_CARTESIAN);
    coordinateSystems.insert("dsmcAxisymmetric", DSMCB200_COORD_AXISYMMETRIC);
    coordinateSystems.insert("dsmcSpherical", DSMCB200_COORD_SPHERICAL);
    HashTable<label> timeStepModels;
    timeStepModels.insert("constant", 0);
    timeStepModels.insert("variable", 1);

    const word collisionModel(particleProperties_.lookup("BinaryCollisionModel"));
    models_.collisionModel = lookupOrFail(collisionModels, collisionModel, "BinaryCollisionModel");
    lookupOrFail(partnerModels, word(particleProperties_.lookup("collisionPartnerSelectionModel")), "collisionPartnerSelection");
    models_.coordinateSystem =
        lookupOrFail(coordinateSystems, particleProperties_.lookupOrDefault<word>("coordinateSystem", "dsmcCartesian"), "dsmcCoordinateSystem");
    variableTimeStep_ = lookupOrFail(timeStepModels, particleProperties_.lookupOrDefault<word>("timeStepModel", "constant"), "dsmcTimeStepModel") == 1;
    polarAxis_ = 1; models_.angularCoordinate = 2;
    if (models_.coordinateSystem == DSMCB200_COORD_AXISYMMETRIC)
    {
        // dsmcAxisymmetric::checkCoordinateSystemInputs (dsmcAxisymmetric.C:337-420)
        const dictionary& ax = particleProperties_.subDict("axisymmetricProperties");
        const word rev(ax.lookupOrDefault<word>("revolutionAxis", word::null)), pol(ax.lookupOrDefault<word>("polarAxis", word::null));
        label polarAxis = 1, angular = 2;
        bool bad = false;
        if (rev == "z")
        {
            if (pol == word::null || pol == "x") { polarAxis = 0; angular = 1; } else if (pol == "y") { polarAxis = 1; angular = 0; } else { bad = true; }
        }
        else if (rev == "y")
        {
            if (pol == word::null || pol == "z") { polarAxis = 2; angular = 0; } else if (pol == "x") { polarAxis = 0; angular = 2; } else { bad = true; }
        }
        else if (rev == "x")
        {
            if (pol == "z") { polarAxis = 2; angular = 1; } else if (pol != "y") { bad = true; }
        }
        if (bad)
        {
            FatalErrorIn("dsmcAxisymmetric::checkCoordinateSystemInputs(const bool init)")
                << "Revolution and polar axes are badly defined in constant/dsmcProperties axisymmetricProperties{}" << exit(FatalError);
        }
        polarAxis_ = polarAxis; models_.angularCoordinate = angular;
    }
    models_.nEquivalentParticles = readScalar(particleProperties_.lookup("nEquivalentParticles"));
    models_.seed = uint64_t(particleProperties_.lookupOrDefault<label>("seedNumber", 1));
    models_.deltaT = mesh_.time().deltaTValue();
    models_.kB = physicoChemical::k.value();
    models_.Tref = 273.0;
    models_.invZvFormulation = 2;
    if (models_.collisionModel != DSMCB200_COLL_NONE)
    {
        const dictionary& coeffs = particleProperties_.subDict(collisionModel + "Coeffs");
        models_.Tref = coeffs.lookupOrDefault<scalar>("Tref", 273.0);          // VariableHardSphere.C:62
        if (models_.collisionModel == DSMCB200_COLL_LB_VHS || models_.collisionModel == DSMCB200_COLL_LB_VSS)
        {
            // LarsenBorgnakkeVariableHardSphere.C:58-99
            models_.rotationalRelaxationCollisionNumber = coeffs.lookupOrDefault<scalar>("rotationalRelaxationCollisionNumber", 5.0);
            models_.vibrationalRelaxationCollisionNumber = coeffs.lookupOrDefault<scalar>("vibrationalRelaxationCollisionNumber", 0.0);
            models_.electronicRelaxationCollisionNumber = coeffs.lookupOrDefault<scalar>("electronicRelaxationCollisionNumber", 500.0);
            const word zv(coeffs.lookupOrDefault<word>("inverseZvFormulation", "default"));
            models_.invZvFormulation = (zv == "pre-2008") ? 0 : (zv == "2008") ? 1 : 2;
        }
    }
    readBoundaries();
    readFieldProperties();
    models_.nPatchModels = patchModels_.size(); models_.patchModels = patchModels_.begin();
    models_.nInflows = inflows_.size(); models_.inflows = inflows_.begin();
    ck(dsmcb200_set_models(ctx_, &models_), "dsmcb200_set_models");
}


void Foam::dsmcCloud::setCellFields()
{
    // the volScalarFields behind nParticles(cell) / deltaTValue(cell): dsmcVariableTimeStepModel::updatenParticles / updateTimeStep
    // (dsmcVariableTimeStepModel.C:48-100) and dsmcAxisymmetric::checkCoordinateSystemInputs + recalculateRWF, method "cell"
    // (dsmcAxisymmetric.C:236-275, 337-470)
    const label nC = mesh_.nCells();
    nParticlesCell_.setSize(nC); nParticlesCell_ = models_.nEquivalentParticles;
    deltaTCell_.setSize(nC); deltaTCell_ = models_.deltaT;
    RWFCell_.setSize(nC); RWFCell_ = 1.0;
    if (variableTimeStep_)
    {
        const scalarField& V = mesh_.cellVolumes();
        scalar minVolume = gMin(V);
        label refCell = -1;
        forAll(V, c) { if (mag(V[c] - minVolume) < SMALL) { refCell = c; break; } }
        if (refCell == -1)
        {
            FatalErrorIn("dsmcCloud (dsmcb200)") << "variable time-step model: the smallest cell of the mesh is on another processor; "
                << "the reference takes its reference cell from the local mesh (dsmcVariableTimeStepModel.C:48-68)" << exit(FatalError);
        }
        const scalar nParticleRef = nParticlesCell_[refCell];
        forAll(V, c) { nParticlesCell_[c] = nParticleRef*V[c]/minVolume; }
        const scalar nParticleTimeStepRatio = nParticlesCell_[refCell]/deltaTCell_[refCell];
        forAll(V, c) { deltaTCell_[c] = nParticlesCell_[c]/nParticleTimeStepRatio; }
        Info<< "Variable time-step model:" << nl << "- Initial time-step [sec]" << tab << deltaTCell_[0] << nl << endl;
    }
    const bool axisymmetric = models_.coordinateSystem == DSMCB200_COORD_AXISYMMETRIC;
    if (axisymmetric)
    {
        const dictionary& ax = particleProperties_.subDict("axisymmetricProperties");
        const word method(ax.lookupOrDefault<word>("radialWeightingMethod", "cell"));
        if (method != "cell")
        {
            FatalErrorIn("dsmcCloud (dsmcb200)") << "radialWeightingMethod " << method
                << ": only the cell-based radial weighting is part of this engine" << exit(FatalError);
        }
        const label polarAxis = polarAxis_, angular = models_.angularCoordinate;
        scalar radialExtent = gMax(mesh_.faceCentres().component(polarAxis));
        if (!(radialExtent > 0)) { radialExtent = -gMin(mesh_.faceCentres().component(polarAxis)); }
        const scalar maxRWF = readScalar(ax.lookup("maxRadialWeightingFactor"));
        forAll(RWFCell_, c) { RWFCell_[c] = 1.0 + (maxRWF - 1.0)*mag(mesh_.cellCentres()[c].component(polarAxis))/radialExtent; }
        Info<< nl << "Axisymmetric simulation:" << nl << "- polar axis label" << tab << polarAxis << nl << "- angular coordinate label" << tab
            << angular << nl << "- radial weighting method" << tab << "cell-based" << nl << "- radial extent" << tab << radialExtent << nl
            << "- maximum radial weighting factor" << tab << maxRWF << nl << endl;
    }
    const bool spherical = models_.coordinateSystem == DSMCB200_COORD_SPHERICAL;
    if (spherical)
    {
        // dsmcSpherical::checkCoordinateSystemInputs / recalculateRWF (dsmcSpherical.C:232-275,325-386), radial weighting method "cell"
        const dictionary& sph = particleProperties_.subDict("sphericalProperties");
        const word method(sph.lookupOrDefault<word>("radialWeightingMethod", "cell"));
        if (method != "cell")
        {
            FatalErrorIn("dsmcCloud (dsmcb200)") << "radialWeightingMethod " << method
                << ": only the cell-based radial weighting is part of this engine" << exit(FatalError);
        }
        const scalar maxRWF = readScalar(sph.lookup("maxRadialWeightingFactor"));
        const vector origin(sph.lookupOrDefault<vector>("origin", vector::zero));
        const scalar radialExtent = gMax(mag(mesh_.faceCentres() - origin));
        forAll(RWFCell_, c) { RWFCell_[c] = 1.0 + (maxRWF - 1.0)*sqr(mag(mesh_.cellCentres()[c] - origin)/radialExtent); }
        Info<< nl << "Spherical simulation:" << nl << "- coordinate system origin" << tab << origin << nl << "- radial weighting method" << tab
            << "cell-based" << nl << "- radial extent" << tab << radialExtent << nl << "- maximum radial weighting factor" << tab << maxRWF << nl << endl;
    }
    const bool weighted = axisymmetric || spherical;
    if (variableTimeStep_ || weighted)
    {
        ck(dsmcb200_set_cell_fields(ctx_, variableTimeStep_ ? nParticlesCell_.begin() : NULL, variableTimeStep_ ? deltaTCell_.begin() : NULL,
                                    weighted ? RWFCell_.begin() : NULL), "dsmcb200_set_cell_fields");
    }
}


void Foam::dsmcCloud::readCloud()
{
    // Cloud<dsmcParcel>::initCloud + dsmcParcel::readFields (dsmcParcelIO.C:133-335) with the stock lagrangian readers
    passiveParticleCloud positions(mesh_, cloudName_);    // positions + cell; OpenFOAM locates tetFace / tetPt (particleI.H:851-996)
    const label n = positions.size();
    IOField<vector> U(positions.fieldIOobject("U", IOobject::MUST_READ));
    IOField<label> typeId(positions.fieldIOobject("typeId", IOobject::MUST_READ));
    IOField<scalar> ERot(positions.fieldIOobject("ERot", IOobject::READ_IF_PRESENT));
    IOField<label> ELevel(positions.fieldIOobject("ELevel", IOobject::READ_IF_PRESENT));
    IOField<label> classification(positions.fieldIOobject("classification", IOobject::READ_IF_PRESENT));
    IOField<labelField> vibLevel(positions.fieldIOobject("vibLevel", IOobject::READ_IF_PRESENT));
    IOField<scalar> radialWeight(positions.fieldIOobject("radialWeight", IOobject::READ_IF_PRESENT));

    Field<vector> pos(n);
    labelList cell(n), tetFace(n), tetPt(n), origId(n), origProc(n), vib(n*maxModes_, 0);
    label i = 0;
    forAllConstIter(passiveParticleCloud, positions, iter)
    {
        const passiveParticle& p = iter();
        pos[i] = p.position(); cell[i] = p.cell(); tetFace[i] = p.tetFace(); tetPt[i] = p.tetPt();
        origId[i] = p.origId(); origProc[i] = p.origProc();
        if (vibLevel.size() == n) { forAll(vibLevel[i], m) { if (m < maxModes_) { vib[i*maxModes_ + m] = vibLevel[i][m]; } } }
        ++i;
    }
    dsmcb200_parcels_soa soa;
    std::memset(&soa, 0, sizeof(soa));
    soa.position = reinterpret_cast<double*>(pos.begin());
    soa.U = reinterpret_cast<double*>(U.begin());
    soa.cell = cell.begin(); soa.tetFace = tetFace.begin(); soa.tetPt = tetPt.begin(); soa.typeId = typeId.begin();
    soa.origId = origId.begin(); soa.origProc = origProc.begin();
    if (ERot.size() == n) { soa.ERot = ERot.begin(); }
    if (ELevel.size() == n) { soa.ELevel = ELevel.begin(); }
    if (classification.size() == n) { soa.classification = classification.begin(); }
    if (radialWeight.size() == n) { soa.radialWeight = radialWeight.begin(); }
    soa.vibLevel = vib.begin(); soa.maxModes = maxModes_;
    ck(dsmcb200_upload_parcels(ctx_, n, &soa), "dsmcb200_upload_parcels");

    // sigmaTcRMax (dsmcCloud.C:612-628)
    volScalarField sigmaTcRMax
    (
        IOobject("dsmcSigmaTcRMax", mesh_.time().timeName(), mesh_, IOobject::MUST_READ, IOobject::NO_WRITE),
        mesh_
    );
    ck(dsmcb200_upload_cellstate(ctx_, sigmaTcRMax.primitiveField().begin(), NULL), "dsmcb200_upload_cellstate");
}


// * * * * * * * * * * * * * * * * Constructors  * * * * * * * * * * * * * * //

Foam::dsmcCloud::dsmcCloud(Time& t, const word& cloudName, const dynamicFvMesh& mesh, bool readFields)
:
    regIOobject(IOobject(cloudName + "B200", t.timeName(), mesh, IOobject::NO_READ, IOobject::AUTO_WRITE)),
    mesh_(mesh),
    time_(t),
    cloudName_(cloudName),
    particleProperties_(IOobject(cloudName + "Properties", t.constant(), mesh, IOobject::MUST_READ, IOobject::NO_WRITE)),
    controlDict_(IOobject("controlDict", t.system(), mesh, IOobject::MUST_READ, IOobject::NO_WRITE)),
    boundariesDict_(IOobject("boundariesDict", t.system(), mesh, IOobject::MUST_READ, IOobject::NO_WRITE)),
    fieldPropertiesDict_(IOobject("fieldPropertiesDict", t.system(), mesh, IOobject::MUST_READ, IOobject::NO_WRITE)),
    ctx_(NULL),
    nTerminalOutputs_(controlDict_.lookupOrDefault<label>("nTerminalOutputs", 1)),
    maxModes_(1)
{
    int nDevices = 0;
    if (cudaGetDeviceCount(&nDevices) != cudaSuccess || nDevices == 0)
    {
        FatalErrorIn("dsmcCloud (dsmcb200)") << "no CUDA device: this library has no CPU path" << exit(FatalError);
    }
    const int rc = dsmcb200_create(&ctx_, Pstream::myProcNo() % nDevices, Pstream::myProcNo(), Pstream::nProcs());
    if (rc) { FatalErrorIn("dsmcCloud (dsmcb200)") << "dsmcb200_create failed with status " << rc << exit(FatalError); }
    if (Pstream::parRun())
    {
        // the ncclUniqueId travels over Pstream once; every later exchange is NCCL on device buffers
        List<char> id(128, 0);
        if (Pstream::master()) { if (dsmcb200_nccl_unique_id(id.begin())) { FatalErrorIn("dsmcCloud (dsmcb200)") << "ncclGetUniqueId failed" << exit(FatalError); } }
        Pstream::scatter(id);
        ck(dsmcb200_init_comm(ctx_, id.begin()), "dsmcb200_init_comm");
    }
    Info<< "Reading the species of typeIdList:" << endl;
    sendMesh();
    readSpecies();
    readModels();
    setCellFields();
    if (readFields) { readCloud(); }
}


// * * * * * * * * * * * * * * * * Destructor  * * * * * * * * * * * * * * * //

Foam::dsmcCloud::~dsmcCloud()
{
    dsmcb200_destroy(ctx_);
}


// * * * * * * * * * * * * * * * Member Functions  * * * * * * * * * * * * * //

Foam::label Foam::dsmcCloud::size() const
{
    int64_t n = 0;
    ck(dsmcb200_download_parcels(ctx_, 0, &n, NULL), "dsmcb200_download_parcels");
    return label(n);
}


void Foam::dsmcCloud::evolve()
{
    ck(dsmcb200_evolve(ctx_, 1), "dsmcb200_evolve");
    // dsmcField::updateTime: sampling restarts at output times while resetAtOutput is on (dsmcField.C:113-152); handled after the write
}


void Foam::dsmcCloud::info()
{
    // dsmcCloud::info (dsmcCloud.C:935-985) and noTimeCounter's collision line (noTimeCounter.C:320-337): global sums over the ranks
    dsmcb200_counters c;
    ck(dsmcb200_get_counters(ctx_, &c), "dsmcb200_get_counters");
    double v[8] = {double(c.nParcels), c.mass, c.linearKineticEnergy, c.rotationalEnergy, c.vibrationalEnergy, c.electronicEnergy,
                   double(c.collisions), 0.0};
    ck(dsmcb200_allreduce_sum(ctx_, v, 7), "dsmcb200_allreduce_sum");
    const scalar nMol = v[0];
    Info<< "    Collisions                      = " << label(v[6]) << nl
        << "    Number of DSMC particles        = " << label(nMol) << nl;
    if (nMol > VSMALL)
    {
        const scalar nP = 1.0;   // mass and energies carry nParticles(cell) per parcel (dsmcCloudI.H:268-297)
        Info<< "    Number of molecules             = " << c.nMolecules << nl
            << "    Mass in system                  = " << v[1]*nP << nl
            << "    Average linear kinetic energy   = " << v[2]/nMol << nl
            << "    Average rotational energy       = " << v[3]/nMol << nl
            << "    Average vibrational energy      = " << v[4]/nMol << nl
            << "    Average electronic energy       = " << v[5]/nMol << nl
            << "    Total energy                    = " << (v[2] + v[3] + v[4] + v[5])*nP << endl;
    }
}


void Foam::dsmcCloud::loadBalanceCheck()
{
    // dsmcDynamicLoadBalancing::update (dsmcDynamicLoadBalancing.C:100-149): the imbalance figure, at output times
    if (!time_.outputTime() || !Pstream::parRun()) { return; }
    scalar nLocal = size();
    scalar nGlobal = nLocal;
    reduce(nGlobal, sumOp<scalar>());
    const scalar ideal = nGlobal/scalar(Pstream::nProcs());
    scalar imbalance = mag(nLocal - ideal);
    reduce(imbalance, maxOp<scalar>());
    balancer_.maxImbalance() = imbalance/ideal;
    Info<< "    Maximum imbalance = " << 100*balancer_.maxImbalance() << "%" << nl << endl;
}


void Foam::dsmcCloud::writeCloud() const
{
    // dsmcParcel::writeFields (dsmcParcelIO.C:338-450): positions + one IOField per member, in cloud-list order
    int64_t n64 = 0;
    ck(dsmcb200_download_parcels(ctx_, 0, &n64, NULL), "dsmcb200_download_parcels");
    const label n = label(n64);
    Field<vector> pos(n), U(n);
    scalarField ERot(n, 0.0), radialWeight(n, 1.0);
    labelList cell(n), tetFace(n), tetPt(n), typeId(n), ELevel(n, 0), newParcel(n, -1), classification(n, 0), origId(n), origProc(n), vib(n*maxModes_, 0);
    dsmcb200_parcels_soa soa;
    std::memset(&soa, 0, sizeof(soa));
    soa.position = reinterpret_cast<double*>(pos.begin()); soa.U = reinterpret_cast<double*>(U.begin()); soa.ERot = ERot.begin();
    soa.cell = cell.begin(); soa.tetFace = tetFace.begin(); soa.tetPt = tetPt.begin(); soa.typeId = typeId.begin();
    soa.vibLevel = vib.begin(); soa.maxModes = maxModes_; soa.ELevel = ELevel.begin(); soa.newParcel = newParcel.begin();
    soa.classification = classification.begin(); soa.origId = origId.begin(); soa.origProc = origProc.begin();
    soa.radialWeight = radialWeight.begin();
    ck(dsmcb200_download_parcels(ctx_, n, &n64, &soa), "dsmcb200_download_parcels");

    passiveParticleCloud outCloud(mesh_, cloudName_, IDLList<passiveParticle>());
    for (label i = 0; i < n; ++i)
    {
        passiveParticle* p = new passiveParticle(mesh_, pos[i], cell[i], tetFace[i], tetPt[i]);
        p->origId() = origId[i]; p->origProc() = origProc[i];
        outCloud.addParticle(p);
    }
    outCloud.write();                                               // positions (+ origProcId, origId)
    IOField<vector> fU(outCloud.fieldIOobject("U", IOobject::NO_READ), U);
    IOField<scalar> fERot(outCloud.fieldIOobject("ERot", IOobject::NO_READ), ERot);
    IOField<label> fELevel(outCloud.fieldIOobject("ELevel", IOobject::NO_READ), ELevel);
    IOField<label> fTypeId(outCloud.fieldIOobject("typeId", IOobject::NO_READ), typeId);
    IOField<label> fNewParcel(outCloud.fieldIOobject("newParcel", IOobject::NO_READ), newParcel);
    IOField<label> fClass(outCloud.fieldIOobject("classification", IOobject::NO_READ), classification);
    IOField<labelField> fVib(outCloud.fieldIOobject("vibLevel", IOobject::NO_READ), n);
    for (label i = 0; i < n; ++i)
    {
        const label nM = species_[typeId[i]].nVibrationalModes;
        fVib[i].setSize(nM);
        for (label m = 0; m < nM; ++m) { fVib[i][m] = vib[i*maxModes_ + m]; }
    }
    IOField<scalar> fRadialWeight(outCloud.fieldIOobject("radialWeight", IOobject::NO_READ), radialWeight);
    fU.write(); fERot.write(); fELevel.write(); fTypeId.write(); fNewParcel.write(); fClass.write(); fVib.write(); fRadialWeight.write();

    volScalarField sigmaTcRMax
    (
        IOobject("dsmcSigmaTcRMax", mesh_.time().timeName(), mesh_, IOobject::NO_READ, IOobject::NO_WRITE),
        mesh_, dimensionedScalar("zero", dimensionSet(0, 3, -1, 0, 0), 0.0), zeroGradientFvPatchScalarField::typeName
    );
    ck(dsmcb200_download_cellstate(ctx_, sigmaTcRMax.primitiveFieldRef().begin(), NULL), "dsmcb200_download_cellstate");
    sigmaTcRMax.correctBoundaryConditions();
    sigmaTcRMax.write();
}


void Foam::dsmcCloud::writeFields() const
{
    // dsmcVolFields::calculateField / writeField (dsmcVolFields.C:1242-1290 every step, :1401-1508 at output time, :2208-2266):
    // every instance is a sum over its typeIds of the per-species moment sums the engine sampled.
    dsmcb200_accum_info ai;
    ck(dsmcb200_accum_info_get(ctx_, &ai), "dsmcb200_accum_info_get");
    const label nC = ai.nCells, nS = ai.nSpecies, nQ = ai.nQuantities;
    scalarField acc(nC*nS*nQ), coll(2*nC), accSet(nC*nS*nQ), collSet(2*nC);
    label loadedSet = -1;
    scalar nTSet = 0;
    const scalar kB = models_.kB;
    const scalarField& V = mesh_.cellVolumes();
    const bool internal = ai.nModes >= 0;

    forAll(fields_, fi)
    {
        const fieldSpec& f = fields_[fi];
        const word& nm = f.fieldName;
        if (f.sampleSet != loadedSet)      // the sums of this field's sample set ...
        {
            selectSampleSet(f.sampleSet);
            ck(dsmcb200_accum_info_get(ctx_, &ai), "dsmcb200_accum_info_get");
            ck(dsmcb200_download_accumulators(ctx_, accSet.begin(), collSet.begin()), "dsmcb200_download_accumulators");
            nTSet = ai.nTimeSteps;
            loadedSet = f.sampleSet;
        }
        acc = accSet; coll = collSet;      // ... since the field's own last reset
        if (f.baseAcc.size() == acc.size()) { acc -= f.baseAcc; }
        if (f.baseColl.size() == coll.size()) { coll -= f.baseColl; }
        const scalar nT = max(nTSet - f.baseNT, 1.0);
        #define DSMCB200_SCALAR_FIELD(var, prefix, dims) \
            volScalarField var(IOobject(word(prefix) + "_" + nm, mesh_.time().timeName(), mesh_, IOobject::NO_READ, IOobject::NO_WRITE), \
                               mesh_, dimensionedScalar("zero", dims, 0.0), calculatedFvPatchScalarField::typeName)
        DSMCB200_SCALAR_FIELD(dsmcNMean, "dsmcNMean", dimless);
        DSMCB200_SCALAR_FIELD(rhoN, "rhoN", dimensionSet(0, -3, 0, 0, 0));
        DSMCB200_SCALAR_FIELD(rhoM, "rhoM", dimensionSet(1, -3, 0, 0, 0));
        DSMCB200_SCALAR_FIELD(p, "p", dimPressure);
        DSMCB200_SCALAR_FIELD(Ttra, "Ttra", dimTemperature);
        DSMCB200_SCALAR_FIELD(Trot, "Trot", dimTemperature);
        DSMCB200_SCALAR_FIELD(Tvib, "Tvib", dimTemperature);
        DSMCB200_SCALAR_FIELD(Tov, "Tov", dimTemperature);
        #undef DSMCB200_SCALAR_FIELD
        volVectorField UMean(IOobject("U_" + nm, mesh_.time().timeName(), mesh_, IOobject::NO_READ, IOobject::NO_WRITE),
                             mesh_, dimensionedVector("zero", dimVelocity, vector::zero), calculatedFvPatchVectorField::typeName);

        for (label c = 0; c < nC; ++c)
        {
            const scalar nP = nParticlesCell_[c]*RWFCell_[c];   // cloud_.nParticles(cell), dsmcVolFields.C:1098,1128
            scalar dsmcNCum = 0, mCum = 0, linearKECum = 0, ErotCum = 0, zetaRotCum = 0;
            vector momentumCum(vector::zero);
            scalar TvibSum = 0, zetaVibSum = 0, moleculesRhoN = 0;
            forAll(f.typeIds, k)
            {
                const label s = f.typeIds[k];
                const scalar* a = &acc[(c*nS + s)*nQ];
                const dsmcb200_species& sp = species_[s];
                const scalar Ns = a[DSMCB200_Q_N];
                dsmcNCum += Ns;
                mCum += nP*sp.mass*Ns;
                momentumCum += nP*sp.mass*vector(a[DSMCB200_Q_PX], a[DSMCB200_Q_PY], a[DSMCB200_Q_PZ]);
                linearKECum += nP*sp.mass*a[DSMCB200_Q_CC];
                if (internal)
                {
                    ErotCum += a[DSMCB200_Q_EROT];
                    zetaRotCum += sp.rotationalDegreesOfFreedom*Ns;
                    // vibrational temperature and degrees of freedom of the species (dsmcVolFields.C:1433-1494)
                    scalar spZeta = 0, zetaByT = 0;
                    for (label m = 0; m < sp.nVibrationalModes; ++m)
                    {
                        const scalar E = a[DSMCB200_Q_EVIB0 + m];
                        if (E > VSMALL && Ns > SMALL)
                        {
                            const scalar iMean = E/(kB*sp.thetaV[m]*Ns);
                            if (iMean > SMALL)
                            {
                                const scalar TvibMod = sp.thetaV[m]/log(1.0 + 1.0/iMean);
                                const scalar zMod = 2.0*iMean*log(1.0 + 1.0/iMean);
                                spZeta += zMod;
                                zetaByT = zMod*TvibMod;             // assigned, not accumulated (:1469)
                            }
                        }
                    }
                    if (spZeta > SMALL)
                    {
                        moleculesRhoN += nP*Ns;
                        TvibSum += nP*Ns*zetaByT/spZeta;
                        zetaVibSum += nP*Ns*spZeta;
                    }
                }
            }
            if (dsmcNCum > 1e-3)
            {
                const scalar nCum = nP*dsmcNCum;
                dsmcNMean[c] = dsmcNCum/nT;
                rhoN[c] = nCum/(nT*V[c]);
                rhoM[c] = mCum/(nT*V[c]);
                UMean[c] = momentumCum/mCum;
                const scalar linearKEMean = 0.5*linearKECum/(V[c]*nT);
                Ttra[c] = 2.0/(3.0*kB*rhoN[c])*(linearKEMean - 0.5*rhoM[c]*(UMean[c] & UMean[c]));
                p[c] = rhoN[c]*kB*Ttra[c];
                if (internal)
                {
                    const scalar zetaRotTot = zetaRotCum/dsmcNCum;
                    Trot[c] = (zetaRotCum > SMALL) ? 2.0*ErotCum/(kB*zetaRotCum) : 0.0;
                    scalar zetaVib = zetaVibSum;
                    Tvib[c] = TvibSum;
                    if (moleculesRhoN > SMALL) { Tvib[c] /= moleculesRhoN; zetaVib /= moleculesRhoN; }
                    Tov[c] = (3.0*Ttra[c] + zetaRotTot*Trot[c] + zetaVib*Tvib[c])/(3.0 + zetaRotTot + zetaVib);
                }
                else
                {
                    Tov[c] = Ttra[c];
                }
            }
            else
            {
                dsmcNMean[c] = 0.001;
            }
        }
        dsmcNMean.write(); rhoN.write(); rhoM.write(); p.write(); Ttra.write(); Trot.write(); Tvib.write(); Tov.write(); UMean.write();
    }

    // dsmcField::updateTime (dsmcField.C:113-152): while resetAtOutput is on (and until resetAtOutputUntilTime) a field's sampling restarts
    // here.  The sums belong to a sample set: when every field of the set resets they are cleared, otherwise a field that resets takes
    // the present sums as its baseline.
    forAll(sampleIntervals_, set)
    {
        bool all = true, some = false;
        forAll(fields_, fi)
        {
            const fieldSpec& f = fields_[fi];
            if (f.sampleSet != set) { continue; }
            const bool resets = f.resetAtOutput && time_.value() <= f.resetAtOutputUntilTime;
            all = all && resets;
            some = some || resets;
        }
        if (!some) { continue; }
        selectSampleSet(set);
        if (all)
        {
            ck(dsmcb200_reset_accumulators(ctx_), "dsmcb200_reset_accumulators");
            forAll(fields_, fi)
            {
                const fieldSpec& f = fields_[fi];
                if (f.sampleSet == set) { f.baseAcc.clear(); f.baseColl.clear(); f.baseNT = 0; }
            }
        }
        else
        {
            ck(dsmcb200_accum_info_get(ctx_, &ai), "dsmcb200_accum_info_get");
            ck(dsmcb200_download_accumulators(ctx_, accSet.begin(), collSet.begin()), "dsmcb200_download_accumulators");
            forAll(fields_, fi)
            {
                const fieldSpec& f = fields_[fi];
                if (f.sampleSet == set && f.resetAtOutput && time_.value() <= f.resetAtOutputUntilTime)
                {
                    f.baseAcc = accSet; f.baseColl = collSet; f.baseNT = ai.nTimeSteps;
                }
            }
        }
    }
    selectSampleSet(0);
}


bool Foam::dsmcCloud::writeObject(IOstream::streamFormat fmt, IOstream::versionNumber, IOstream::compressionType) const
{
    if (fmt != IOstream::ASCII)
    {
        WarningIn("dsmcCloud (dsmcb200)") << "the lagrangian fields are written through the stock IOField writers in the format of controlDict" << endl;
    }
    writeCloud();
    writeFields();
    return true;
}

// ************************************************************************* //
