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

// body of a data file after the FoamFile header
std::string dataBody(const std::string& path) {
    std::string s = stripComments(slurp(path));
    size_t p = s.find("FoamFile");
    if (p != std::string::npos) {
        size_t e = s.find('}', p);
        if (e != std::string::npos) s = s.substr(e + 1);
    }
    return s;
}

struct Scanner {
    const char* p;
    const char* e;
    void ws() { while (p < e && (std::isspace(static_cast<unsigned char>(*p)))) ++p; }
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

// positions the scanner after "N (" and returns N; uniform form N{v} returns uniformText
int64_t openSized(Scanner& sc, bool& uniform) {
    sc.ws();
    while (sc.p < sc.e && !std::isdigit(static_cast<unsigned char>(*sc.p))) ++sc.p;
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
    std::string s = dataBody(path);
    Scanner sc{s.data(), s.data() + s.size()};
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
std::vector<double> readScalarField(const std::string& path) {
    std::string s = dataBody(path);
    Scanner sc{s.data(), s.data() + s.size()};
    bool uni;
    int64_t n = openSized(sc, uni);
    std::vector<double> o(static_cast<size_t>(n));
    if (uni) { double v = sc.num(); for (auto& x : o) x = v; return o; }
    for (int64_t i = 0; i < n; ++i) o[i] = sc.num();
    return o;
}
std::vector<int32_t> readLabelField(const std::string& path) {
    auto v = readScalarField(path);
    std::vector<int32_t> o(v.size());
    for (size_t i = 0; i < v.size(); ++i) o[i] = int32_t(std::llround(v[i]));
    return o;
}
void readFaces(const std::string& path, std::vector<int32_t>& offsets, std::vector<int32_t>& labels) {
    std::string s = dataBody(path);
    Scanner sc{s.data(), s.data() + s.size()};
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
    std::string s = dataBody(path);
    Scanner sc{s.data(), s.data() + s.size()};
    bool uni;
    int64_t n = openSized(sc, uni);
    xyz.assign(size_t(n) * 3, 0.0);
    cell.assign(size_t(n), 0);
    for (int64_t i = 0; i < n; ++i) {
        if (!sc.eat('(')) throw FoamError(path + ": '(' expected in positions (binary clouds are not supported)");
        xyz[3 * i] = sc.num(); xyz[3 * i + 1] = sc.num(); xyz[3 * i + 2] = sc.num();
        sc.eat(')');
        cell[i] = int32_t(sc.integer());
    }
}
std::vector<int32_t> readLabelListList(const std::string& path, int& width) {
    std::string s = dataBody(path);
    Scanner sc{s.data(), s.data() + s.size()};
    bool uni;
    int64_t n = openSized(sc, uni);
    std::vector<std::vector<int32_t>> rows(static_cast<size_t>(n));
    if (uni) {
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
    std::string s = dataBody(path);
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
    std::string s = dataBody(path);
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
std::string header(const std::string& cls, const std::string& location, const std::string& object) {
    std::ostringstream s;
    s << "/*--------------------------------*- C++ -*----------------------------------*\\\n"
         "| =========                 |                                                 |\n"
         "| \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox           |\n"
         "|  \\\\    /   O peration     | Version:  v1706                                 |\n"
         "|   \\\\  /    A nd           | Web:      www.OpenFOAM.com                      |\n"
         "|    \\\\/     M anipulation  |                                                 |\n"
         "\\*---------------------------------------------------------------------------*/\n"
         "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       "
      << cls << ";\n    location    \"" << location << "\";\n    object      " << object
      << ";\n}\n// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n";
    return s.str();
}

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
    std::fprintf(f, "%lld\n(\n", (long long)n);
    for (int64_t i = 0; i < n; ++i) std::fprintf(f, "(%.*g %.*g %.*g)\n", gWritePrecision, a[3 * i], gWritePrecision, a[3 * i + 1], gWritePrecision, a[3 * i + 2]);
    std::fputs(")\n", f);
    std::fclose(f);
}
void writePositions(const std::string& path, const std::string& location, const double* xyz, const int32_t* cell, int64_t n) {
    FILE* f = openw(path);
    std::fputs(header("Cloud<dsmcParcel>", location, "positions").c_str(), f);
    std::fprintf(f, "%lld\n(\n", (long long)n);
    for (int64_t i = 0; i < n; ++i) std::fprintf(f, "(%.*g %.*g %.*g) %d\n", gWritePrecision, xyz[3 * i], gWritePrecision, xyz[3 * i + 1], gWritePrecision, xyz[3 * i + 2], cell[i]);
    std::fputs(")\n", f);
    std::fclose(f);
}
void writeLabelListList(const std::string& path, const std::string& cls, const std::string& location, const std::string& object,
                        const int32_t* a, int64_t n, int width) {
    FILE* f = openw(path);
    std::fputs(header(cls, location, object).c_str(), f);
    std::fprintf(f, "%lld\n(\n", (long long)n);
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
    std::fprintf(f, "internalField   nonuniform List<%s> \n%lld\n(\n", nCmpt == 1 ? "scalar" : (nCmpt == 3 ? "vector" : "tensor"), (long long)nCells);
    for (int64_t i = 0; i < nCells; ++i) { put(internal + i * nCmpt); std::fputc('\n', f); }
    std::fputs(")\n;\n\nboundaryField\n{\n", f);
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
