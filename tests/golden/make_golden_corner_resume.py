#!/usr/bin/env python
"""Generates tests/golden/hypersonicCorner_resume.npz: RAW ACCUMULATORS written by dsmcFoam+ next to the FIELDS it derived from them.

The reference ships, for its hypersonicCorner tutorial, the per-processor sampling state of a 4-way scotch run
(backup-processors/processorN/0.003/uniform/resumeSampling_Ar: dsmcNCum, dsmcMomentumCum, dsmcLinearKECum, dsmcNCollsCum,
collisionSeparation, ... per processor cell, nTimeSteps = 1500, written by dsmcVolFields::writeOut, dsmcVolFields.C:745-835) and the
reconstructed fields of the same instant (backup-0.003/*_Ar, written by dsmcVolFields::calculateField / writeField, :1242-1290,
:1624-1790).  The processor meshes are not shipped, so a processor cell is tied to its global cell through dsmcNMean = dsmcNCum /
nTimeSteps (an integer count: exact); counts that occur more than once are dropped.  All cells of this mesh are 1 cm cubes.

The fixture pins oracle/fields_ref.derive (and with it the driver's calculateField) on data the reference itself produced.
Run in the build container (needs /root/reference):  python tests/golden/make_golden_corner_resume.py"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hystrath_b200 import foamfile as ff  # noqa: E402

CASE = "/root/reference/run/hyStrath/dsmcFoam+/hypersonicCorner"
KEEP = 3000


def entry(text, key, width=1):
    m = re.search(r"^%s\s+(\d+)\s*\(" % re.escape(key), text, re.M)
    n = int(m.group(1))
    end = text.index(";", m.end())
    body = text[m.end():end].replace("(", " ").replace(")", " ")
    a = np.array(body.split(), dtype=np.float64)
    assert a.size == n * width, (key, a.size, n, width)
    return a.reshape(n, width) if width > 1 else a


def patch_lists(text, key, width=1, species=False):
    """`key nPatches ( N ( v ... ) | N { v } ... );` (a List<scalarField> / List<vectorField> over the patches; `species` = one more
    list level with a single species) -> list of [N] or [N, 3] arrays."""
    m = re.search(r"^%s\s+(\d+)\s*\(" % re.escape(key), text, re.M)
    toks = text[m.end():text.index(";", m.end())].replace("(", " ( ").replace(")", " ) ").replace("{", " { ").replace("}", " } ").split()
    i = 0
    if species:
        assert int(m.group(1)) == 1
        n_patches = int(toks[0]); assert toks[1] == "("
        i = 2
    else:
        n_patches = int(m.group(1))
    out = []
    for _ in range(n_patches):
        n = int(toks[i]); opener = toks[i + 1]; i += 2
        if opener == "{":                                   # uniform list
            if width == 1:
                v = [float(toks[i])]; i += 1
            else:
                assert toks[i] == "("; v = [float(t) for t in toks[i + 1:i + 1 + width]]; i += width + 2
            assert toks[i] == "}"; i += 1
            out.append(np.tile(np.array(v), (n, 1)).reshape(n, width) if width > 1 else np.full(n, v[0]))
        else:
            vals = []
            for _k in range(n):
                if width == 1:
                    vals.append(float(toks[i])); i += 1
                else:
                    assert toks[i] == "("; vals.append([float(t) for t in toks[i + 1:i + 1 + width]]); i += width + 2
            assert toks[i] == ")"; i += 1
            out.append(np.array(vals, dtype=np.float64).reshape(n, width) if width > 1 else np.array(vals, dtype=np.float64))
    return out


WALL_KEYS = [("rhoNBF", 1, False), ("rhoNIntBF", 1, False), ("rhoNElecBF", 1, False), ("rhoMBF", 1, False), ("linearKEBF", 1, False),
             ("speciesMccBF", 1, True), ("momentumBF", 3, False), ("ErotBF", 1, False), ("zetaRotBF", 1, False), ("speciesEvibBF", 1, True),
             ("speciesEelecBF", 1, True), ("qBF", 1, False), ("fDBF", 3, False)]   # the engine's WallQ order 0..16


def walls(out, nT, fnum):
    """Wall faces: the `walls` patch is the third of every processor's boundary list (flow, entrance, walls, processor...)."""
    rows = []
    for proc in range(4):
        text = open(os.path.join(CASE, "backup-processors", f"processor{proc}", "0.003", "uniform", "resumeSampling_Ar")).read()
        cols = []
        for key, width, species in WALL_KEYS:
            a = patch_lists(text, key, width, species)[2]
            cols.append(a.reshape(len(a), width))
        rows.append(np.concatenate(cols, axis=1))
    w = np.concatenate(rows)                                    # [nLocalWallFaces, 17]
    d = os.path.join(CASE, "backup-0.003")
    gf = {n: ff.read_patch_field(os.path.join(d, f"{n}_Ar"), "walls") for n in ("rhoN", "rhoM", "Ttra", "p", "wallHeatFlux", "wallShearStress", "fD", "U", "Ma")}
    # tie a processor face to its global face through rhoN = rhoNBF F_N / nTimeSteps (printed with 10 digits); unique values only
    key_local = np.round(np.log(w[:, 0] * fnum / nT) * 1e8).astype(np.int64)
    key_glob = np.round(np.log(gf["rhoN"]) * 1e8).astype(np.int64)
    vg, ig, cg = np.unique(key_glob, return_index=True, return_counts=True)
    vl, cl = np.unique(key_local, return_counts=True)
    glob = dict(zip(vg[cg == 1].tolist(), ig[cg == 1].tolist()))
    single = set(vl[cl == 1].tolist())
    pairs = [(i, glob[k]) for i, k in enumerate(key_local.tolist()) if k in single and k in glob]
    li = np.array([p[0] for p in pairs]); gi = np.array([p[1] for p in pairs])
    out["wall_acc"] = w[li]
    for n, v in gf.items():
        out["wall_field_" + n] = v[gi]
    # the plate a face lies on: p = fD . n with n = -e_y (y = 0 plate) or -e_z (z = 0 plate)
    fD, p = gf["fD"][gi], gf["p"][gi]
    on_y = np.abs(-fD[:, 1] - p) < np.abs(-fD[:, 2] - p)
    out["wall_normal"] = np.where(on_y[:, None], np.array([0.0, -1.0, 0.0]), np.array([0.0, 0.0, -1.0]))
    return len(pairs), len(w)


def main():
    acc = {k: [] for k in ("dsmcNCum", "dsmcMCum", "dsmcLinearKECum", "dsmcMomentumCum", "dsmcNCollsCum", "collisionSeparation", "nCum", "mCum")}
    nT = None
    for proc in range(4):
        text = open(os.path.join(CASE, "backup-processors", f"processor{proc}", "0.003", "uniform", "resumeSampling_Ar")).read()
        nT = float(re.search(r"^nTimeSteps\s+([0-9.eE+-]+);", text, re.M).group(1))
        for k in acc:
            acc[k].append(entry(text, k, 3 if k == "dsmcMomentumCum" else 1))
    acc = {k: np.concatenate(v) for k, v in acc.items()}
    d = os.path.join(CASE, "backup-0.003")
    fields = {n: ff.read_internal_field(os.path.join(d, f"{n}_Ar")) for n in ("dsmcNMean", "rhoN", "rhoM", "Ttra", "p", "Ma", "mfp", "mct", "SOFP", "U")}
    count = np.rint(fields["dsmcNMean"] * nT).astype(np.int64)
    assert np.abs(fields["dsmcNMean"] * nT - count).max() < 1e-3
    uniq, first, n_occ = np.unique(count, return_index=True, return_counts=True)
    glob = dict(zip(uniq[n_occ == 1].tolist(), first[n_occ == 1].tolist()))
    pc = acc["dsmcNCum"].astype(np.int64)
    _, _, occ_p = np.unique(pc, return_inverse=True, return_counts=True)[0:3]
    vals, cnts = np.unique(pc, return_counts=True)
    single = set(vals[cnts == 1].tolist())
    rows = [(i, glob[c]) for i, c in enumerate(pc.tolist()) if c in single and c in glob]
    rows = rows[:: max(1, len(rows) // KEEP)][:KEEP]
    pi = np.array([r[0] for r in rows]); gi = np.array([r[1] for r in rows])
    out = {"nTimeSteps": np.float64(nT), "globalCell": gi.astype(np.int32)}
    for k, v in acc.items():
        out["acc_" + k] = v[pi]
    for k, v in fields.items():
        out["field_" + k] = v[gi]
    props = ff.read_dict(os.path.join(CASE, "constant", "dsmcProperties"))
    out["nEquivalentParticles"] = np.float64(props["nEquivalentParticles"])
    for key in ("mass", "diameter", "omega"):
        out[f"Ar_{key}"] = np.float64(props["moleculeProperties"]["Ar"][key])
    ctl = ff.read_dict(os.path.join(CASE, "system", "controlDict"))
    out["deltaT"] = np.float64(ctl["deltaT"])
    n_wall, n_wall_all = walls(out, nT, float(out["nEquivalentParticles"]))
    path = os.path.join(ROOT, "tests", "golden", "hypersonicCorner_resume.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(rows), "cells of", len(pc), "matched uniquely;", n_wall, "wall faces of", n_wall_all)


if __name__ == "__main__":
    main()
