// foam_io.h -- reader/writer for the ASCII and binary OpenFOAM files of an unchanged dsmcFoam+ case directory.
//
// The reference reads its configuration through OpenFOAM's IOdictionary / IOField machinery
// (DSMC/clouds/dsmcCloud.C:597-636, DSMC/parcels/dsmcParcelIO.C:133-450,
// BASIC/IOPosition/IOPosition.C:65-150).  OpenFOAM is not available outside an OpenFOAM install,
// so the standalone driver parses the same files itself: FoamFile header, `{}` dictionaries,
// `( )` lists, `//` and `/* */` comments, the `N ( ... )` and `N{v}` list forms.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace foam {

struct FoamError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Dict;

// One token of a primitive entry: a word / number, a ( ... ) list, or a { ... } dictionary.
struct Node {
    enum Kind { WORD, LIST, DICT } kind = WORD;
    std::string word;
    std::vector<Node> list;
    std::shared_ptr<Dict> dict;
    bool isNumber() const;
    double number() const;
};

struct Dict {
    std::string name;  // path-like name for error messages
    std::vector<std::pair<std::string, std::vector<Node>>> entries;

    bool found(const std::string& key) const;
    const std::vector<Node>& stream(const std::string& key) const;  // throws "keyword ... is undefined in dictionary ..."
    const Dict& subDict(const std::string& key) const;
    bool isDict(const std::string& key) const;
    double scalar(const std::string& key) const;
    double scalarOr(const std::string& key, double dflt) const;
    int64_t label(const std::string& key) const;
    int64_t labelOr(const std::string& key, int64_t dflt) const;
    std::string word(const std::string& key) const;
    std::string wordOr(const std::string& key, const std::string& dflt) const;
    bool boolOr(const std::string& key, bool dflt) const;  // Switch: on/off, yes/no, true/false
    std::vector<double> scalarList(const std::string& key) const;
    std::vector<double> scalarListOr(const std::string& key, const std::vector<double>& dflt) const;
    std::vector<int64_t> labelListOr(const std::string& key, const std::vector<int64_t>& dflt) const;
    std::vector<std::string> wordList(const std::string& key) const;
    std::vector<double> vector3(const std::string& key) const;
    // `key ( name { ... } name { ... } )`  (boundariesDict / fieldPropertiesDict style)
    std::vector<std::pair<std::string, const Dict*>> dictList(const std::string& key) const;
    std::vector<std::string> toc() const;
};

Dict parseDict(const std::string& text, const std::string& name);
Dict readDict(const std::string& path);
bool exists(const std::string& path);
std::vector<std::string> listDir(const std::string& path);

// ---- bulk data ----
std::vector<double> readVectorField(const std::string& path);                 // [3n]
std::vector<double> readScalarField(const std::string& path);                 // [n]
std::vector<int32_t> readLabelField(const std::string& path);                 // [n]
void readFaces(const std::string& path, std::vector<int32_t>& offsets, std::vector<int32_t>& labels);
void readPositions(const std::string& path, std::vector<double>& xyz, std::vector<int32_t>& cell);
std::vector<int32_t> readLabelListList(const std::string& path, int& width);  // [n*width], ragged rows zero padded
std::vector<double> readInternalField(const std::string& path, int64_t nCells, int nCmpt);

struct BoundaryPatch {
    std::string name, type, neighbourPatch, referPatch;
    int32_t nFaces = 0, startFace = 0, myProcNo = -1, neighbProcNo = -1;
    bool hasSeparation = false;
    double separation[3] = {0, 0, 0};
};
std::vector<BoundaryPatch> readBoundary(const std::string& path);

// ---- writers ----
std::string header(const std::string& cls, const std::string& location, const std::string& object);       // in the write format
std::string asciiHeader(const std::string& cls, const std::string& location, const std::string& object);  // for files this driver writes as text in any case
void writeScalarField(const std::string& path, const std::string& cls, const std::string& location, const std::string& object,
                      const double* a, int64_t n);
void writeLabelField(const std::string& path, const std::string& cls, const std::string& location, const std::string& object,
                     const int32_t* a, int64_t n);
void writeVectorField(const std::string& path, const std::string& cls, const std::string& location, const std::string& object,
                      const double* a, int64_t n);
void writePositions(const std::string& path, const std::string& location, const double* xyz, const int32_t* cell, int64_t n);
void writeLabelListList(const std::string& path, const std::string& cls, const std::string& location, const std::string& object,
                        const int32_t* a, int64_t n, int width);
// volScalarField / volVectorField with per-patch boundary values (size 0 -> `calculated; value uniform 0`)
struct PatchValues {
    std::string name, type;     // patch name, polyPatch type (empty / cyclic / processor get their constraint type)
    std::vector<double> values; // nFaces*nCmpt or empty
};
void writeVolField(const std::string& path, const std::string& location, const std::string& object, const std::string& dimensions,
                   const double* internal, int64_t nCells, int nCmpt, const std::vector<PatchValues>& patches);
void makeDirs(const std::string& path);
// controlDict writeFormat binary: every writer above emits `format binary;` files (contiguous lists as their size and the raw bytes in
// round brackets, BASIC/particle/particleIO.C:121-143, BASIC/IOPosition/IOPosition.C:65-83); the readers take either format from the header
void setWriteBinary(bool binary);
bool writeBinary();
// controlDict writePrecision (IOstream::defaultPrecision): significant digits of every ASCII writer above
void setWritePrecision(int p);
int writePrecision();
std::string timeName(double t, int precision);

}  // namespace foam
