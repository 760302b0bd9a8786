// foam_io.cpp -- see foam_io.h
#include "foam_io.h"

#include <dirent.h>
#include <sys/stat.h>

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace foam {

namespace {

std::string slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw FoamError("cannot open file " + path);
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

std::string stripComments(const std::string& s) {
    std::string o;
    o.reserve(s.size());
    size_t i = 0, n = s.size();
    while (i < n) {
        if (s[i] == '/' && i + 1 < n && s[i + 1] == '/') {
            while (i < n && s[i] != '\n') ++i;
        } else if (s[i] == '/' && i + 1 < n && s[i + 1] == '*') {
            i += 2;
            while (i + 1 < n && !(s[i] == '*' && s[i + 1] == '/')) ++i;
            i += 2;
            o.push_back(' ');
        } else if (s[i] == '"') {
            o.push_back(s[i++]);
            while (i < n && s[i] != '"') o.push_back(s[i++]);
            if (i < n) o.push_back(s[i++]);
        } else {
            o.push_back(s[i++]);
        }
    }
    return o;
}

struct Tokens {
    std::vector<std::string> t;
    size_t i = 0;
    bool end() const { return i >= t.size(); }
    const std::string& peek() const { return t[i]; }
    std::string next() { return t[i++]; }
};

Tokens tokenise(const std::string& s) {
    Tokens k;
    size_t i = 0, n = s.size();
    while (i < n) {
        char c = s[i];
        if (std::isspace(static_cast<unsigned char>(c))) { ++i; continue; }
        if (c == '{' || c == '}' || c == '(' || c == ')' || c == ';') { k.t.emplace_back(1, c); ++i; continue; }
        if (c == '"') {
            size_t j = s.find('"', i + 1);
            if (j == std::string::npos) j = n - 1;
            k.t.push_back(s.substr(i + 1, j - i - 1));
            i = j + 1;
            continue;
        }
        size_t j = i;
        while (j < n && !std::isspace(static_cast<unsigned char>(s[j])) && s[j] != '{' && s[j] != '}' && s[j] != '(' && s[j] != ')' && s[j] != ';') ++j;
        k.t.push_back(s.substr(i, j - i));
        i = j;
    }
    return k;
}

bool looksNumber(const std::string& w) {
    if (w.empty()) return false;
    char* e = nullptr;
    std::strtod(w.c_str(), &e);
    return e && *e == '\0';
}

Dict parseBody(Tokens& k, const std::string& name);

Node parseList(Tokens& k, const std::string& name) {
    Node nd;
    nd.kind = Node::LIST;
    k.next();  // (
    while (!k.end() && k.peek() != ")") {
        if (k.peek() == "(") {
            nd.list.push_back(parseList(k, name));
        } else if (k.peek() == "{") {
            k.next();
            Node d;
            d.kind = Node::DICT;
            d.dict = std::make_shared<Dict>(parseBody(k, name));
            nd.list.push_back(d);
        } else {
            std::string w = k.next();
            if (!k.end() && k.peek() == "{" && !looksNumber(w)) {  // named dictionary inside a list
                k.next();
                Node d;
                d.kind = Node::DICT;
                d.word = w;
                d.dict = std::make_shared<Dict>(parseBody(k, name + "/" + w));
                nd.list.push_back(d);
            } else if (!k.end() && k.peek() == "(" && looksNumber(w)) {  // sized list N ( ... )
                nd.list.push_back(parseList(k, name));
            } else {
                Node x;
                x.word = w;
                nd.list.push_back(x);
            }
        }
    }
    if (!k.end()) k.next();  // )
    return nd;
}

Dict parseBody(Tokens& k, const std::string& name) {
    Dict d;
    d.name = name;
    while (!k.end() && k.peek() != "}") {
        std::string key = k.next();
        if (key == ";") continue;  // a stray ';' after a sub-dictionary's '}' (common in the shipped dsmcInitialiseDict files)
        if (k.end()) break;
        if (k.peek() == "{") {
            k.next();
            Node nd;
            nd.kind = Node::DICT;
            nd.dict = std::make_shared<Dict>(parseBody(k, name + "/" + key));
            d.entries.push_back({key, {nd}});
            continue;
        }
        std::vector<Node> vals;
        while (!k.end() && k.peek() != ";") {
            if (k.peek() == "(") vals.push_back(parseList(k, name));
            else if (k.peek() == "{") {
                k.next();
                Node nd;
                nd.kind = Node::DICT;
                nd.dict = std::make_shared<Dict>(parseBody(k, name + "/" + key));
                vals.push_back(nd);
            } else if (k.peek() == "}") break;  // tolerate a missing ';'
            else {
                Node x;
                x.word = k.next();
                vals.push_back(x);
            }
        }
        if (!k.end() && k.peek() == ";") k.next();
        if (vals.size() == 2 && vals[0].kind == Node::WORD && vals[0].isNumber() && vals[1].kind == Node::LIST) vals.erase(vals.begin());
        d.entries.push_back({key, vals});
    }
    if (!k.end()) k.next();  // }
    return d;
}

// body of a data file after the FoamFile header.  `format binary;` files (IOstream::BINARY: every contiguous list is its size followed by
// the raw bytes in round brackets, OSstream::write(const char*, streamsize)) keep their bytes: comments are skipped by the scanner, not
// stripped from the text.  The header's arch entry ("LSB;label=32;scalar=64") gives the label width; scalars are doubles.
struct DataFile {
    std::string s;
    bool binary = false;
    int labelBytes = 4;
};
DataFile openData(const std::string& path) {
    DataFile f;
    std::string raw = slurp(path);
    size_t p = raw.find("FoamFile");
    size_t e = p == std::string::npos ? std::string::npos : raw.find('}', p);
    if (e != std::string::npos) {
        const std::string hdr = raw.substr(p, e - p);
        size_t q = hdr.find("format");
        if (q != std::string::npos) {
            q += 6;
            while (q < hdr.size() && std::isspace(static_cast<unsigned char>(hdr[q]))) ++q;
            f.binary = hdr.compare(q, 6, "binary") == 0;
        }
        if (hdr.find("label=64") != std::string::npos) f.labelBytes = 8;
        if (f.binary && hdr.find("scalar=32") != std::string::npos) throw FoamError(path + ": binary files of a single-precision build (scalar=32) are not supported");
    }
    if (f.binary) {
        f.s = raw.substr(e + 1);
    } else {
        f.s = stripComments(raw);
        p = f.s.find("FoamFile");
        if (p != std::string::npos) {
            e = f.s.find('}', p);
            if (e != std::string::npos) f.s = f.s.substr(e + 1);
        }
    }
    return f;
}
std::string dataBody(const std::string& path) {
    return openData(path).s;
}

struct Scanner {
    const char* p;
    const char* e;
    void ws() {   // white space and, in files whose comments were not stripped, comments
        for (;;) {
            while (p < e && (std::isspace(static_cast<unsigned char>(*p)))) ++p;
            if (p + 1 < e && p[0] == '/' && p[1] == '/') { while (p < e && *p != '\n') ++p; continue; }
            if (p + 1 < e && p[0] == '/' && p[1] == '*') { p += 2; while (p + 1 < e && !(p[0] == '*' && p[1] == '/')) ++p; p += 2; continue; }
            break;
        }
    }
    // one binary block: '(' count raw bytes ')'
    void raw(void* dst, size_t bytes) {
        ws();
        if (p >= e || *p != '(') throw FoamError("'(' expected in front of a binary block");
        ++p;
        if (size_t(e - p) < bytes + 1) throw FoamError("binary block is truncated");
        std::memcpy(dst, p, bytes);
        p += bytes;
        if (*p != ')') throw FoamError("')' expected after a binary block");
        ++p;
    }
    // a binary List<label>: labels of the file's width narrowed to 32 bits
    void rawLabels(int32_t* dst, size_t n, int labelBytes) {
        if (n == 0) return;
        if (labelBytes == 4) { raw(dst, n * 4); return; }
        std::vector<int64_t> w(n);
        raw(w.data(), n * 8);
        for (size_t i = 0; i < n; ++i) dst[i] = int32_t(w[i]);
    }
    bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
    double num() {
        ws();
        char* q = nullptr;
        double v = std::strtod(p, &q);
        if (q == p) throw FoamError("number expected");
        p = q;
        return v;
    }
    long integer() {
        ws();
        char* q = nullptr;
        long v = std::strtol(p, &q, 10);
        if (q == p) throw FoamError("label expected");
        p = q;
        return v;
    }
};

// binary lists: the size only (the block follows)
int64_t binarySize(Scanner& sc) {
    sc.ws();
    while (sc.p < sc.e && !std::isdigit(static_cast<unsigned char>(*sc.p))) { ++sc.p; sc.ws(); }
    return sc.integer();
}

// positions the scanner after "N (" and returns N; uniform form N{v} returns uniformText
int64_t openSized(Scanner& sc, bool& uniform) {
    sc.ws();
    while (sc.p < sc.e && !std::isdigit(static_cast<unsigned char>(*sc.p))) { ++sc.p; sc.ws(); }
    int64_t n = sc.integer();
    sc.ws();
    uniform = false;
    if (sc.p < sc.e && *sc.p == '{') { ++sc.p; uniform = true; return n; }
    if (!sc.eat('(')) throw FoamError("'(' expected after list size");
    return n;
}

}  // namespace

bool Node::isNumber() const { return kind == WORD && looksNumber(word); }
double Node::number() const {
    if (!isNumber()) throw FoamError("number expected, found '" + word + "'");
    return std::strtod(word.c_str(), nullptr);
}

bool Dict::found(const std::string& key) const {
    for (auto& e : entries) if (e.first == key) return true;
    return false;
}
const std::vector<Node>& Dict::stream(const std::string& key) const {
    for (auto& e : entries) if (e.first == key) return e.second;
    throw FoamError("keyword " + key + " is undefined in dictionary \"" + name + "\"");
}
bool Dict::isDict(const std::string& key) const {
    if (!found(key)) return false;
    auto& s = stream(key);
    return s.size() == 1 && s[0].kind == Node::DICT;
}
const Dict& Dict::subDict(const std::string& key) const {
    auto& s = stream(key);
    if (s.size() != 1 || s[0].kind != Node::DICT) throw FoamError("keyword " + key + " is not a dictionary in \"" + name + "\"");
    return *s[0].dict;
}
double Dict::scalar(const std::string& key) const {
    auto& s = stream(key);
    if (s.empty()) throw FoamError("empty entry " + key + " in \"" + name + "\"");
    return s[0].number();
}
double Dict::scalarOr(const std::string& key, double d) const { return found(key) ? scalar(key) : d; }
int64_t Dict::label(const std::string& key) const { return int64_t(std::llround(scalar(key))); }
int64_t Dict::labelOr(const std::string& key, int64_t d) const { return found(key) ? label(key) : d; }
std::string Dict::word(const std::string& key) const {
    auto& s = stream(key);
    if (s.empty() || s[0].kind != Node::WORD) throw FoamError("word expected for " + key + " in \"" + name + "\"");
    return s[0].word;
}
std::string Dict::wordOr(const std::string& key, const std::string& d) const { return found(key) ? word(key) : d; }
bool Dict::boolOr(const std::string& key, bool d) const {
    if (!found(key)) return d;
    std::string w = word(key);
    return w == "on" || w == "yes" || w == "true" || w == "y" || w == "t";
}
std::vector<double> Dict::scalarList(const std::string& key) const {
    auto& s = stream(key);
    std::vector<double> o;
    if (s.empty()) return o;
    if (s[0].kind != Node::LIST) { o.push_back(s[0].number()); return o; }
    for (auto& n : s[0].list) o.push_back(n.number());
    return o;
}
std::vector<double> Dict::scalarListOr(const std::string& key, const std::vector<double>& d) const { return found(key) ? scalarList(key) : d; }
std::vector<int64_t> Dict::labelListOr(const std::string& key, const std::vector<int64_t>& d) const {
    if (!found(key)) return d;
    std::vector<int64_t> o;
    for (double v : scalarList(key)) o.push_back(int64_t(std::llround(v)));
    return o;
}
std::vector<std::string> Dict::wordList(const std::string& key) const {
    auto& s = stream(key);
    std::vector<std::string> o;
    if (s.empty()) return o;
    if (s[0].kind != Node::LIST) { o.push_back(s[0].word); return o; }
    for (auto& n : s[0].list) o.push_back(n.word);
    return o;
}
std::vector<double> Dict::vector3(const std::string& key) const {
    auto v = scalarList(key);
    if (v.size() != 3) throw FoamError("vector (x y z) expected for " + key + " in \"" + name + "\"");
    return v;
}
std::vector<std::pair<std::string, const Dict*>> Dict::dictList(const std::string& key) const {
    std::vector<std::pair<std::string, const Dict*>> o;
    if (!found(key)) return o;
    auto& s = stream(key);
    if (s.empty()) return o;
    if (s[0].kind != Node::LIST) throw FoamError("list expected for " + key + " in \"" + name + "\"");
    for (auto& n : s[0].list)
        if (n.kind == Node::DICT) o.push_back({n.word, n.dict.get()});
    return o;
}
std::vector<std::string> Dict::toc() const {
    std::vector<std::string> o;
    for (auto& e : entries) o.push_back(e.first);
    return o;
}

Dict parseDict(const std::string& text, const std::string& name) {
    Tokens k = tokenise(stripComments(text));
    Dict d = parseBody(k, name);
    // top-level bare lists:  name ( ... );  are parsed by parseBody as entries already
    return d;
}
Dict readDict(const std::string& path) { return parseDict(slurp(path), path); }

bool exists(const std::string& path) {
    struct stat st;
    return stat(path.c_str(), &st) == 0;
}
std::vector<std::string> listDir(const std::string& path) {
    std::vector<std::string> o;
    DIR* d = opendir(path.c_str());
    if (!d) return o;
    while (dirent* e = readdir(d)) {
        std::string n = e->d_name;
        if (n != "." && n != "..") o.push_back(n);
    }
    closedir(d);
    return o;
}

std::vector<double> readVectorField(const std::string& path) {
    const DataFile df = openData(path);
    const std::string& s = df.s;
    Scanner sc{s.data(), s.data() + s.size()};
    if (df.binary) {
        const int64_t n = binarySize(sc);
        std::vector<double> o(size_t(n) * 3);
        if (n) sc.raw(o.data(), o.size() * 8);
        return o;
    }
    bool uni;
    int64_t n = openSized(sc, uni);
    std::vector<double> o(size_t(n) * 3);
    if (uni) {
        sc.eat('(');
        double v[3] = {sc.num(), sc.num(), sc.num()};
        for (int64_t i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) o[3 * i + d] = v[d];
        return o;
    }
    for (int64_t i = 0; i < n; ++i) {
        if (!sc.eat('(')) throw FoamError(path + ": '(' expected in vector list");
        o[3 * i] = sc.num(); o[3 * i + 1] = sc.num(); o[3 * i + 2] = sc.num();
        sc.eat(')');
    }
    return o;
}
namespace {
std::vector<double> scalarFieldOf(const DataFile& df) {
    const std::string& s = df.s;
    Scanner sc{s.data(), s.data() + s.size()};
    if (df.binary) {
        const int64_t n = binarySize(sc);
        std::vector<double> o(static_cast<size_t>(n));
        if (n) sc.raw(o.data(), o.size() * 8);
        return o;
    }
    bool uni;
    int64_t n = openSized(sc, uni);
    std::vector<double> o(static_cast<size_t>(n));
    if (uni) { double v = sc.num(); for (auto& x : o) x = v; return o; }
    for (int64_t i = 0; i < n; ++i) o[i] = sc.num();
    return o;
}
}  // namespace
std::vector<double> readScalarField(const std::string& path) { return scalarFieldOf(openData(path)); }
std::vector<int32_t> readLabelField(const std::string& path) {
    const DataFile df = openData(path);
    if (df.binary) {
        Scanner sc{df.s.data(), df.s.data() + df.s.size()};
        const int64_t n = binarySize(sc);
        std::vector<int32_t> o(static_cast<size_t>(n));
        sc.rawLabels(o.data(), o.size(), df.labelBytes);
        return o;
    }
    auto v = scalarFieldOf(df);
    std::vector<int32_t> o(v.size());
    for (size_t i = 0; i < v.size(); ++i) o[i] = int32_t(std::llround(v[i]));
    return o;
}
void readFaces(const std::string& path, std::vector<int32_t>& offsets, std::vector<int32_t>& labels) {
    const DataFile df = openData(path);
    const std::string& s = df.s;
    Scanner sc{s.data(), s.data() + s.size()};
    if (df.binary) {   // faceCompactList: the offsets (nFaces + 1) and the point labels of all faces, two contiguous lists
        const int64_t n1 = binarySize(sc);
        offsets.assign(size_t(n1), 0);
        sc.rawLabels(offsets.data(), offsets.size(), df.labelBytes);
        const int64_t m = binarySize(sc);
        labels.assign(size_t(m), 0);
        sc.rawLabels(labels.data(), labels.size(), df.labelBytes);
        if (offsets.empty() || offsets.back() != int32_t(m)) throw FoamError(path + ": face offsets do not match the label list");
        return;
    }
    bool uni;
    int64_t n = openSized(sc, uni);
    offsets.assign(size_t(n) + 1, 0);
    labels.clear();
    for (int64_t i = 0; i < n; ++i) {
        long m = sc.integer();
        if (!sc.eat('(')) throw FoamError(path + ": '(' expected in face");
        for (long k = 0; k < m; ++k) labels.push_back(int32_t(sc.integer()));
        sc.eat(')');
        offsets[i + 1] = int32_t(labels.size());
    }
}
void readPositions(const std::string& path, std::vector<double>& xyz, std::vector<int32_t>& cell) {
    const DataFile df = openData(path);
    const std::string& s = df.s;
    Scanner sc{s.data(), s.data() + s.size()};
    bool uni;
    int64_t n = openSized(sc, uni);
    xyz.assign(size_t(n) * 3, 0.0);
    cell.assign(size_t(n), 0);
    if (df.binary) {
        // particle::write(os, false) in binary (particleIO.C:121-143): position, cellI, faceI, stepFraction as one block per particle
        const size_t bytes = 24 + 2 * size_t(df.labelBytes) + 8;
        unsigned char rec[48];
        for (int64_t i = 0; i < n; ++i) {
            sc.raw(rec, bytes);
            std::memcpy(&xyz[3 * i], rec, 24);
            if (df.labelBytes == 4) { int32_t c; std::memcpy(&c, rec + 24, 4); cell[i] = c; }
            else { int64_t c; std::memcpy(&c, rec + 24, 8); cell[i] = int32_t(c); }
        }
        return;
    }
    for (int64_t i = 0; i < n; ++i) {
        if (!sc.eat('(')) throw FoamError(path + ": '(' expected in positions");
        xyz[3 * i] = sc.num(); xyz[3 * i + 1] = sc.num(); xyz[3 * i + 2] = sc.num();
        sc.eat(')');
        cell[i] = int32_t(sc.integer());
    }
}
std::vector<int32_t> readLabelListList(const std::string& path, int& width) {
    const DataFile df = openData(path);
    const std::string& s = df.s;
    Scanner sc{s.data(), s.data() + s.size()};
    bool uni;
    int64_t n = openSized(sc, uni);
    std::vector<std::vector<int32_t>> rows(static_cast<size_t>(n));
    if (df.binary) {   // a list of lists is not contiguous: the rows follow one another, each a size and (if not empty) a block
        for (int64_t i = 0; i < n; ++i) {
            const long m = sc.integer();
            rows[i].assign(size_t(m), 0);
            sc.rawLabels(rows[i].data(), size_t(m), df.labelBytes);
        }
    } else if (uni) {
        long m = sc.integer();
        sc.eat('(');
        std::vector<int32_t> r;
        for (long k = 0; k < m; ++k) r.push_back(int32_t(sc.integer()));
        for (auto& x : rows) x = r;
    } else {
        for (int64_t i = 0; i < n; ++i) {
            long m = sc.integer();
            if (m > 0) {
                sc.ws();
                if (sc.p < sc.e && *sc.p == '{') {  // m{v}
                    ++sc.p;
                    int32_t v = int32_t(sc.integer());
                    sc.eat('}');
                    rows[i].assign(size_t(m), v);
                } else {
                    sc.eat('(');
                    for (long k = 0; k < m; ++k) rows[i].push_back(int32_t(sc.integer()));
                    sc.eat(')');
                }
            } else {
                sc.ws();
                if (sc.p < sc.e && *sc.p == '(') { ++sc.p; sc.eat(')'); }
            }
        }
    }
    width = 0;
    for (auto& r : rows) width = std::max<int>(width, int(r.size()));
    std::vector<int32_t> o(size_t(n) * size_t(std::max(width, 1)), 0);
    for (int64_t i = 0; i < n; ++i)
        for (size_t k = 0; k < rows[i].size(); ++k) o[size_t(i) * std::max(width, 1) + k] = rows[i][k];
    return o;
}
std::vector<double> readInternalField(const std::string& path, int64_t nCells, int nCmpt) {
    const DataFile df = openData(path);
    const std::string& s = df.s;
    size_t p = s.find("internalField");
    if (p == std::string::npos) throw FoamError(path + ": no internalField");
    Scanner sc{s.data() + p + 13, s.data() + s.size()};
    sc.ws();
    std::vector<double> o(size_t(nCells) * nCmpt);
    if (std::strncmp(sc.p, "uniform", 7) == 0) {
        sc.p += 7;
        sc.eat('(');
        std::vector<double> v(nCmpt);
        for (int d = 0; d < nCmpt; ++d) v[d] = sc.num();
        for (int64_t i = 0; i < nCells; ++i) for (int d = 0; d < nCmpt; ++d) o[i * nCmpt + d] = v[d];
        return o;
    }
    const char* q = std::strchr(sc.p, '>');
    if (!q) throw FoamError(path + ": malformed internalField");
    sc.p = q + 1;
    if (df.binary) {
        if (binarySize(sc) != nCells) throw FoamError(path + ": internalField size does not match the mesh");
        if (nCells) sc.raw(o.data(), o.size() * 8);
        return o;
    }
    bool uni;
    int64_t n = openSized(sc, uni);
    if (n != nCells) throw FoamError(path + ": internalField size does not match the mesh");
    for (int64_t i = 0; i < n; ++i) {
        if (nCmpt > 1) sc.eat('(');
        for (int d = 0; d < nCmpt; ++d) o[i * nCmpt + d] = sc.num();
        if (nCmpt > 1) sc.eat(')');
    }
    return o;
}

std::vector<BoundaryPatch> readBoundary(const std::string& path) {
    std::string s = stripComments(dataBody(path));   // entries only, whatever the header's format says
    size_t p = 0;
    while (p < s.size() && !std::isdigit(static_cast<unsigned char>(s[p]))) ++p;
    size_t q = s.find('(', p);
    size_t r = s.rfind(')');
    if (q == std::string::npos || r == std::string::npos) throw FoamError(path + ": malformed boundary file");
    Dict d = parseDict(s.substr(q + 1, r - q - 1), path);
    std::vector<BoundaryPatch> o;
    for (auto& e : d.entries) {
        if (e.second.size() != 1 || e.second[0].kind != Node::DICT) continue;
        const Dict& pd = *e.second[0].dict;
        BoundaryPatch b;
        b.name = e.first;
        b.type = pd.word("type");
        b.nFaces = int32_t(pd.label("nFaces"));
        b.startFace = int32_t(pd.label("startFace"));
        b.neighbourPatch = pd.wordOr("neighbourPatch", "");
        b.referPatch = pd.wordOr("referPatch", "");
        b.myProcNo = int32_t(pd.labelOr("myProcNo", -1));
        b.neighbProcNo = int32_t(pd.labelOr("neighbProcNo", -1));
        if (pd.found("separationVector")) {
            auto v = pd.vector3("separationVector");
            b.hasSeparation = true;
            for (int k = 0; k < 3; ++k) b.separation[k] = v[k];
        }
        o.push_back(b);
    }
    return o;
}

// ---------------------------------------------------------------------------------------------
namespace {
bool gWriteBinary = false;
}
void setWriteBinary(bool b) { gWriteBinary = b; }
bool writeBinary() { return gWriteBinary; }

namespace {
std::string headerOf(const std::string& cls, const std::string& location, const std::string& object, bool binary);
}
std::string header(const std::string& cls, const std::string& location, const std::string& object) { return headerOf(cls, location, object, gWriteBinary); }
std::string asciiHeader(const std::string& cls, const std::string& location, const std::string& object) { return headerOf(cls, location, object, false); }
namespace {
std::string headerOf(const std::string& cls, const std::string& location, const std::string& object, bool binary) {
    std::ostringstream s;
    s << "/*--------------------------------*- C++ -*----------------------------------*\\\n"
         "| =========                 |                                                 |\n"
         "| \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox           |\n"
         "|  \\\\    /   O peration     | Version:  v1706                                 |\n"
         "|   \\\\  /    A nd           | Web:      www.OpenFOAM.com                      |\n"
         "|    \\\\/     M anipulation  |                                                 |\n"
         "\\*---------------------------------------------------------------------------*/\n"
         "FoamFile\n{\n    version     2.0;\n    format      " << (binary ? "binary" : "ascii") << ";\n"
      << (binary ? "    arch        \"LSB;label=32;scalar=64\";\n" : "") << "    class       "
      << cls << ";\n    location    \"" << location << "\";\n    object      " << object
      << ";\n}\n// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n";
    return s.str();
}
}  // namespace

namespace {
int gWritePrecision = 10;
}
void setWritePrecision(int p) { gWritePrecision = p < 1 ? 1 : (p > 17 ? 17 : p); }
int writePrecision() { return gWritePrecision; }
namespace {
void fmt(FILE* f, double v) { std::fprintf(f, "%.*g", gWritePrecision, v); }
FILE* openw(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) throw FoamError("cannot write " + path);
    return f;
}
// OSstream::write(const char*, streamsize): the bytes in round brackets
void block(FILE* f, const void* a, size_t bytes) {
    std::fputc('(', f);
    if (bytes && std::fwrite(a, 1, bytes, f) != bytes) throw FoamError("short write");
    std::fputc(')', f);
}
// Ostream << List<T> of a contiguous T in binary: nl, size, nl and, unless empty, the block
void binaryList(FILE* f, const void* a, int64_t n, size_t elemBytes) {
    std::fprintf(f, "\n%lld\n", (long long)n);
    if (n) block(f, a, size_t(n) * elemBytes);
}
}  // namespace

void makeDirs(const std::string& path) {
    std::string cur;
    for (size_t i = 0; i <= path.size(); ++i) {
        if (i == path.size() || path[i] == '/') {
            if (!cur.empty()) mkdir(cur.c_str(), 0755);
        }
        if (i < path.size()) cur.push_back(path[i]);
    }
}

std::string timeName(double t, int precision) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.*g", precision, t);
    return buf;
}

void writeScalarField(const std::string& path, const std::string& cls, const std::string& location, const std::string& object,
                      const double* a, int64_t n) {
    FILE* f = openw(path);
    std::fputs(header(cls, location, object).c_str(), f);
    if (gWriteBinary) { binaryList(f, a, n, 8); std::fputc('\n', f); std::fclose(f); return; }
    bool uni = n > 0;
    for (int64_t i = 1; i < n && uni; ++i) uni = a[i] == a[0];
    if (uni) { std::fprintf(f, "%lld{", (long long)n); fmt(f, a[0]); std::fputs("}\n", f); }
    else {
        std::fprintf(f, "%lld\n(\n", (long long)n);
        for (int64_t i = 0; i < n; ++i) { fmt(f, a[i]); std::fputc('\n', f); }
        std::fputs(")\n", f);
    }
    std::fclose(f);
}
void writeLabelField(const std::string& path, const std::string& cls, const std::string& location, const std::string& object,
                     const int32_t* a, int64_t n) {
    FILE* f = openw(path);
    std::fputs(header(cls, location, object).c_str(), f);
    if (gWriteBinary) { binaryList(f, a, n, 4); std::fputc('\n', f); std::fclose(f); return; }
    bool uni = n > 0;
    for (int64_t i = 1; i < n && uni; ++i) uni = a[i] == a[0];
    if (uni) std::fprintf(f, "%lld{%d}\n", (long long)n, a[0]);
    else {
        std::fprintf(f, "%lld\n(\n", (long long)n);
        for (int64_t i = 0; i < n; ++i) std::fprintf(f, "%d\n", a[i]);
        std::fputs(")\n", f);
    }
    std::fclose(f);
}
void writeVectorField(const std::string& path, const std::string& cls, const std::string& location, const std::string& object,
                      const double* a, int64_t n) {
    FILE* f = openw(path);
    std::fputs(header(cls, location, object).c_str(), f);
    if (gWriteBinary) { binaryList(f, a, n, 24); std::fputc('\n', f); std::fclose(f); return; }
    std::fprintf(f, "%lld\n(\n", (long long)n);
    for (int64_t i = 0; i < n; ++i) std::fprintf(f, "(%.*g %.*g %.*g)\n", gWritePrecision, a[3 * i], gWritePrecision, a[3 * i + 1], gWritePrecision, a[3 * i + 2]);
    std::fputs(")\n", f);
    std::fclose(f);
}
void writePositions(const std::string& path, const std::string& location, const double* xyz, const int32_t* cell, int64_t n) {
    FILE* f = openw(path);
    std::fputs(header("Cloud<dsmcParcel>", location, "positions").c_str(), f);
    std::fprintf(f, "%lld\n(\n", (long long)n);
    if (gWriteBinary) {
        // IOPosition::writeData + particle::write(os, false): position, cellI, faceI (-1 between steps), stepFraction (0) per particle
        for (int64_t i = 0; i < n; ++i) {
            unsigned char rec[40];
            const int32_t face = -1;
            const double stepFraction = 0.0;
            std::memcpy(rec, xyz + 3 * i, 24); std::memcpy(rec + 24, cell + i, 4); std::memcpy(rec + 28, &face, 4); std::memcpy(rec + 32, &stepFraction, 8);
            block(f, rec, 40);
            std::fputc('\n', f);
        }
        std::fputs(")\n", f);
        std::fclose(f);
        return;
    }
    for (int64_t i = 0; i < n; ++i) std::fprintf(f, "(%.*g %.*g %.*g) %d\n", gWritePrecision, xyz[3 * i], gWritePrecision, xyz[3 * i + 1], gWritePrecision, xyz[3 * i + 2], cell[i]);
    std::fputs(")\n", f);
    std::fclose(f);
}
void writeLabelListList(const std::string& path, const std::string& cls, const std::string& location, const std::string& object,
                        const int32_t* a, int64_t n, int width) {
    FILE* f = openw(path);
    std::fputs(header(cls, location, object).c_str(), f);
    std::fprintf(f, "%lld\n(\n", (long long)n);
    if (gWriteBinary) {
        for (int64_t i = 0; i < n; ++i) binaryList(f, a + i * width, width, 4);
        std::fputs("\n)\n", f);
        std::fclose(f);
        return;
    }
    for (int64_t i = 0; i < n; ++i) {
        std::fprintf(f, "%d(", width);
        for (int k = 0; k < width; ++k) std::fprintf(f, k ? " %d" : "%d", a[i * width + k]);
        std::fputs(")\n", f);
    }
    std::fputs(")\n", f);
    std::fclose(f);
}

void writeVolField(const std::string& path, const std::string& location, const std::string& object, const std::string& dimensions,
                   const double* internal, int64_t nCells, int nCmpt, const std::vector<PatchValues>& patches) {
    FILE* f = openw(path);
    std::fputs(header(nCmpt == 1 ? "volScalarField" : (nCmpt == 3 ? "volVectorField" : "volTensorField"), location, object).c_str(), f);
    std::fprintf(f, "dimensions      %s;\n\n", dimensions.c_str());
    auto put = [&](const double* v) {
        if (nCmpt == 1) fmt(f, v[0]);
        else { std::fputc('(', f); for (int d = 0; d < nCmpt; ++d) { if (d) std::fputc(' ', f); fmt(f, v[d]); } std::fputc(')', f); }
    };
    if (gWriteBinary) {
        std::fprintf(f, "internalField   nonuniform List<%s> ", nCmpt == 1 ? "scalar" : (nCmpt == 3 ? "vector" : "tensor"));
        binaryList(f, internal, nCells, size_t(nCmpt) * 8);
        std::fputs(";\n\nboundaryField\n{\n", f);
    } else {
        std::fprintf(f, "internalField   nonuniform List<%s> \n%lld\n(\n", nCmpt == 1 ? "scalar" : (nCmpt == 3 ? "vector" : "tensor"), (long long)nCells);
        for (int64_t i = 0; i < nCells; ++i) { put(internal + i * nCmpt); std::fputc('\n', f); }
        std::fputs(")\n;\n\nboundaryField\n{\n", f);
    }
    for (auto& p : patches) {
        std::fprintf(f, "    %s\n    {\n", p.name.c_str());
        if (p.type == "empty" || p.type == "cyclic" || p.type == "processor" || p.type == "processorCyclic" || p.type == "symmetryPlane" ||
            p.type == "symmetry" || p.type == "wedge") {
            std::fprintf(f, "        type            %s;\n", p.type.c_str());
        } else {
            std::fputs("        type            calculated;\n", f);
            const int64_t nf = int64_t(p.values.size()) / nCmpt;
            if (nf == 0) {
                std::fputs(nCmpt == 1 ? "        value           uniform 0;\n" : (nCmpt == 3 ? "        value           uniform (0 0 0);\n" : "        value           uniform (0 0 0 0 0 0 0 0 0);\n"), f);
            } else if (gWriteBinary) {
                std::fprintf(f, "        value           nonuniform List<%s> ", nCmpt == 1 ? "scalar" : (nCmpt == 3 ? "vector" : "tensor"));
                binaryList(f, p.values.data(), nf, size_t(nCmpt) * 8);
                std::fputs(";\n", f);
            } else {
                std::fprintf(f, "        value           nonuniform List<%s> \n%lld\n(\n", nCmpt == 1 ? "scalar" : (nCmpt == 3 ? "vector" : "tensor"), (long long)nf);
                for (int64_t i = 0; i < nf; ++i) { put(p.values.data() + i * nCmpt); std::fputc('\n', f); }
                std::fputs(")\n;\n", f);
            }
        }
        std::fputs("    }\n", f);
    }
    std::fputs("}\n\n\n// ************************************************************************* //\n", f);
    std::fclose(f);
}

}  // namespace foam
