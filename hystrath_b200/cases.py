"""Synthetic case definitions of BASELINE.json's configs, as (mesh, species, models, fill) bundles for the C ABI.

C2  `lofthouse_cylinder`: 2-D Mach-10 argon flow over a 0.3048 m cylinder (Kn = 0.01), VHS, diffuse 500 K wall,
     free-stream inflow + deletion on the outer boundary.  The free-stream values are not in the reference tree;
     they are the literature case (Lofthouse, Boyd & Wright 2007): U = 2634.1 m/s, T = 200 K, n = 4.247e20 m^-3,
     argon d_ref = 3.595e-10 m at T_ref = 1000 K, omega = 0.734.
C5  `periodic_box`: equilibrium gas at rest in a periodic brick (argon VHS or 5-species air LB-VHS).
"""
from __future__ import annotations

import numpy as np

from . import capi, meshgen

KB = 1.38065e-23


def argon_lofthouse():
    return capi.make_species("Ar", 66.3e-27, 3.595e-10, 0.734)


def lofthouse_cylinder(nr=640, ntheta=1250, ppc=25, r_out=0.6096, seed=0xD5C00002, dt=None):
    r_in = 0.1524
    mesh = meshgen.cylinder_ogrid(nr, ntheta, r_in, r_out, grading=3.0)
    n_inf, T_inf, U_inf, T_w = 4.247e20, 200.0, 2634.1, 500.0
    sp = [argon_lofthouse()]
    vol = np.pi * (r_out ** 2 - r_in ** 2) * mesh.thickness
    fnum = n_inf * vol / (mesh.n_cells * ppc)
    dr_min = mesh.r[1] - mesh.r[0]
    if dt is None:
        dt = 0.2 * dr_min / U_inf
    pm = [dict(patch=mesh.patch_index("cylinder"), boundaryModel="dsmcDiffuseWallPatch", temperature=T_w, velocity=(0.0, 0.0, 0.0)),
          dict(patch=mesh.patch_index("outer"), boundaryModel="dsmcDeletionPatch")]
    inflow = [dict(patch=mesh.patch_index("outer"), typeIds=[0], numberDensities=[n_inf], velocity=(U_inf, 0.0, 0.0),
                   translationalTemperature=T_inf)]
    models = capi.build_models("VariableHardSphere", nEquivalentParticles=fnum, deltaT=dt, seed=seed, Tref=1000.0,
                               patch_models=pm, inflows=inflow)
    fill = dict(type_ids=[0], number_densities=[n_inf], Ttra=T_inf, velocity=(U_inf, 0.0, 0.0))
    return mesh, sp, models, fill
