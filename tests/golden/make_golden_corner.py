#!/usr/bin/env python
"""Generates tests/golden/hypersonicCorner.npz from the fields dsmcFoam+ wrote for its hypersonicCorner tutorial
(run/hyStrath/dsmcFoam+/hypersonicCorner/backup-0.003: Bird's supersonic corner flow in argon, averages over t = 1.5..3 ms).
The polyMesh itself is not shipped (the tutorial runs blockMesh); its two-block blockMeshDict fixes the cell numbering
(block 1: 5 x 18 x 18 cells, block 2: 25 x 18 x 18, i fastest inside a block), which tests/test_gpu_reference_fields.py rebuilds.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_corner.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hystrath_b200 import foamfile as ff  # noqa: E402

CASE = "/root/reference/run/hyStrath/dsmcFoam+/hypersonicCorner"


def main():
    d = os.path.join(CASE, "backup-0.003")
    out = {}
    for n in ("rhoN", "rhoM", "Ttra", "p", "Ma", "mfp", "mct", "dsmcNMean", "SOFP"):
        out[n] = ff.read_internal_field(os.path.join(d, f"{n}_Ar")).astype(np.float32)
    out["U"] = ff.read_internal_field(os.path.join(d, "U_Ar")).astype(np.float32)
    for n in ("wallHeatFlux", "wallShearStress", "p", "rhoN", "Ttra", "fD", "U"):
        out[f"wall_{n}"] = ff.read_patch_field(os.path.join(d, f"{n}_Ar"), "walls").astype(np.float32)
    props = ff.read_dict(os.path.join(CASE, "constant", "dsmcProperties"))
    out["nEquivalentParticles"] = np.float64(props["nEquivalentParticles"])
    for key in ("mass", "diameter", "omega", "alpha"):
        out[f"Ar_{key}"] = np.float64(props["moleculeProperties"]["Ar"][key])
    path = os.path.join(ROOT, "tests", "golden", "hypersonicCorner.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def flat_plate():
    """tests/golden/supersonicFlatPlate.npz from run/hyStrath/dsmcFoam+/supersonicFlatPlate/backup-0.04 (averages over t = 8..40 ms)."""
    case = "/root/reference/run/hyStrath/dsmcFoam+/supersonicFlatPlate"
    d = os.path.join(case, "backup-0.04")
    out = {}
    for n in ("rhoN", "rhoM", "Ttra", "Trot", "Tov", "p", "Ma", "mfp", "dsmcNMean"):
        out[n] = ff.read_internal_field(os.path.join(d, f"{n}_N2cold")).astype(np.float32)
    out["U"] = ff.read_internal_field(os.path.join(d, "U_N2cold")).astype(np.float32)
    for n in ("wallHeatFlux", "wallShearStress", "p", "rhoN", "Ttra", "Trot", "fD", "U"):
        out[f"wall_{n}"] = ff.read_patch_field(os.path.join(d, f"{n}_N2cold"), "plate").astype(np.float32)
    props = ff.read_dict(os.path.join(case, "constant", "dsmcProperties"))
    out["nEquivalentParticles"] = np.float64(props["nEquivalentParticles"])
    path = os.path.join(ROOT, "tests", "golden", "supersonicFlatPlate.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
    flat_plate()
