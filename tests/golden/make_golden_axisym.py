#!/usr/bin/env python
"""Generates tests/golden/axisymmetricFlatnosedCylinder.npz from the fields the reference ships for its axisymmetric tutorial
(run/hyStrath/dsmcFoam+/axisymmetricFlatnosedCylinder/backup-0.004: Mach-5.4 argon onto a flat-nosed cylinder, 5-degree wedge mesh of
4000 cells, dsmcAxisymmetric with maxRadialWeightingFactor 1000; fields averaged from t = 8e-4 to 4e-3, 40 000 steps): the per-cell
radial weighting factors, the mean number of parcels per cell, density, velocity, temperature and pressure, and the heat flux on the
cylinder.  These are the only reference-side vectors of the radial-weighting path.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_axisym.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hystrath_b200 import foamfile as ff  # noqa: E402

CASE = "/root/reference/run/hyStrath/dsmcFoam+/axisymmetricFlatnosedCylinder"


def main():
    d = os.path.join(CASE, "backup-0.004")
    out = {}
    for k in ("RWF", "dsmcNMean_Ar", "rhoN_Ar", "Ttra_Ar", "p_Ar", "dsmcSigmaTcRMax"):
        out[k] = ff.read_internal_field(os.path.join(d, k))
    out["U_Ar"] = ff.read_internal_field(os.path.join(d, "U_Ar"))
    out["wallHeatFlux_cylinder"] = ff.read_patch_field(os.path.join(d, "wallHeatFlux_Ar"), "cylinder")
    out.update(nEquivalentParticles=2e7, maxRadialWeightingFactor=1000.0, deltaT=8e-8, numberDensity=1.0e21, temperature=100.0,
               velocity=np.array([1000.0, 0.0, 0.0]), wallTemperature=300.0, mass=66.3e-27, diameter=4.17e-10, omega=0.81, alpha=1.4)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "axisymmetricFlatnosedCylinder.npz"), **out)
    print({k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
