#!/usr/bin/env python
"""Generates tests/golden/couette_N2-O2.npz from the artefacts the reference ships for its
couette_N2-O2 tutorial (run/hyStrath/dsmcFoam+/couette_N2-O2): the OpenFOAM-v1706 polyMesh, the
complete 47 583-parcel cloud state of backup-5/ and the sampled fields written by dsmcFoam+.
These are the only reference-side vectors that pin results on the scoped path.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The GPU box never reads /root/reference; it only sees the committed .npz.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hystrath_b200 import foamfile as ff  # noqa: E402

CASE = "/root/reference/run/hyStrath/dsmcFoam+/couette_N2-O2"


def main():
    pm = os.path.join(CASE, "constant", "polyMesh")
    out = {}
    out["points"] = ff.read_vector_list(os.path.join(pm, "points"))
    out["face_offsets"], out["face_points"] = ff.read_faces(os.path.join(pm, "faces"))
    out["owner"] = ff.read_scalar_list(os.path.join(pm, "owner"), np.int32)
    out["neighbour"] = ff.read_scalar_list(os.path.join(pm, "neighbour"), np.int32)
    bnd = ff.read_boundary(os.path.join(pm, "boundary"))
    out["patch_names"] = np.array([b["name"] for b in bnd])
    out["patch_types"] = np.array([b["type"] for b in bnd])
    out["patch_start"] = np.array([b["startFace"] for b in bnd], np.int32)
    out["patch_size"] = np.array([b["nFaces"] for b in bnd], np.int32)
    out["patch_neighbour"] = np.array([b.get("neighbourPatch", "") for b in bnd])

    cl = os.path.join(CASE, "backup-5", "lagrangian", "dsmc")
    out["positions"], out["cell"] = ff.read_positions(os.path.join(cl, "positions"))
    out["U"] = ff.read_vector_list(os.path.join(cl, "U"))
    out["ERot"] = ff.read_scalar_list(os.path.join(cl, "ERot"))
    out["typeId"] = ff.read_scalar_list(os.path.join(cl, "typeId"), np.int32)
    out["vibLevel"] = ff.read_label_list_list(os.path.join(cl, "vibLevel"))
    out["classification"] = ff.read_scalar_list(os.path.join(cl, "classification"), np.int32)
    out["newParcel"] = ff.read_scalar_list(os.path.join(cl, "newParcel"), np.int32)
    out["origId"] = ff.read_scalar_list(os.path.join(cl, "origId"), np.int32)

    fd = os.path.join(CASE, "backup-5")
    for name in ("rhoN", "rhoM", "p", "Ttra", "Trot", "Tvib", "Tov", "dsmcNMean", "mfp", "mct", "Ma"):
        for inst in ("mixture", "N2", "O2"):
            out[f"{name}_{inst}"] = ff.read_internal_field(os.path.join(fd, f"{name}_{inst}"))
    out["U_mixture"] = ff.read_internal_field(os.path.join(fd, "U_mixture"))
    # wall-face values (boundaryField of the two wall patches, 5 faces each) and the two remaining mean-free-path outputs
    for name in ("wallHeatFlux", "wallShearStress", "p", "rhoN", "rhoM", "Ttra", "Trot", "Tvib", "Tov", "Ma", "U", "fD"):
        for patch in ("upperWall", "lowerWall"):
            out[f"wall_{name}_{patch}"] = ff.read_patch_field(os.path.join(fd, f"{name}_mixture"), patch)
    for name in ("mfpToDx", "SOFP"):
        out[f"{name}_mixture"] = ff.read_internal_field(os.path.join(fd, f"{name}_mixture"))
    out["dsmcSigmaTcRMax"] = ff.read_internal_field(os.path.join(fd, "dsmcSigmaTcRMax"))

    props = ff.read_dict(os.path.join(CASE, "constant", "dsmcProperties"))
    out["nEquivalentParticles"] = np.float64(props["nEquivalentParticles"])
    mp = props["moleculeProperties"]
    for sp in ("N2", "O2"):
        for key in ("mass", "diameter", "omega"):
            out[f"{sp}_{key}"] = np.float64(mp[sp][key])
    path = os.path.join(ROOT, "tests", "golden", "couette_N2-O2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(out["cell"]), "parcels,", len(out["owner"]), "faces")


if __name__ == "__main__":
    main()
