"""Synthetic case definitions of BASELINE.json's configs, as (mesh, species, models, fill) bundles for the C ABI.

C2  `lofthouse_cylinder`: 2-D Mach-10 argon flow over a 0.3048 m cylinder (Kn = 0.01), VHS, diffuse 500 K wall,
     free-stream inflow + deletion on the outer boundary.  The free-stream values are not in the reference tree;
     they are the literature case (Lofthouse, Boyd & Wright 2007): U = 2634.1 m/s, T = 200 K, n = 4.247e20 m^-3,
     argon d_ref = 3.595e-10 m at T_ref = 1000 K, omega = 0.734.
C3  `air_wedge`: 5-species air (N2, O2, NO, N, O; moleculeProperties of run/hyStrath/dsmcFoam+/heatBath-5species/constant/
     dsmcProperties) over a sharp 15-degree wedge at the free stream of run/hyStrath/dsmcFoam+/orion107kmNR/system/boundariesDict
     (6053.4 m/s, 217.63 K, N2 2.318e18 + O2 6.161e17 m^-3, Mach 20), Larsen-Borgnakke with variable Zv, diffuse 1000 K wall;
     NO, N, O enter at 1e-3 mole fraction each so that every species path of the collision kernel executes (no chemistry).
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


def air5_species():
    """typeIdList (N2 O2 NO N O) with the moleculeProperties of the shipped heatBath-5species case."""
    return [
        capi.make_species("N2", 46.5e-27, 4.17e-10, 0.74, 1.36, 2, (3371,), (52560,), (3371,), 113500),
        capi.make_species("O2", 53.12e-27, 4.07e-10, 0.77, 1.4, 2, (2256,), (17900,), (2256,), 59500),
        capi.make_species("NO", 49.81e-27, 4.2e-10, 0.79, 1.0, 2, (2719,), (1400,), (2719,), 75500),
        capi.make_species("N", 23.25e-27, 3.0e-10, 0.8, 1.0),
        capi.make_species("O", 26.56e-27, 3.0e-10, 0.8, 1.0),
    ]


def air_wedge(nx=2000, ny=1000, nz=4, ppc=25, dx=0.15, angle_deg=15.0, seed=0xD5C00003, procs=1, rank=0, density_scale=1.0):
    """BASELINE configs[2].  Cell size dx ~ lambda_inf / 3 (lambda_inf = 0.46 m at 2.93e18 m^-3), dt = 0.3 dx / U_inf.
    density_scale multiplies the free-stream number densities (tests use a denser stream so that a few steps hold many collisions)."""
    U_inf, T_inf, T_w = 6053.4, 217.63, 1000.0
    n_N2, n_O2 = 2.318e18 * density_scale, 6.161e17 * density_scale
    trace = 1e-3 * (n_N2 + n_O2)
    dens = [n_N2, n_O2, trace, trace, trace]
    mesh = meshgen.wedge_mesh((nx, ny, nz), nx * dx, ny * dx, nz * dx, angle_deg, procs, rank)
    # equal statistical weight everywhere: FNUM from the inlet cell volume (the cells shrink towards the outlet)
    fnum = sum(dens) * dx ** 3 / ppc
    dt = 0.3 * dx / U_inf
    pm = [dict(patch=mesh.patch_index("wedge"), boundaryModel="dsmcDiffuseWallPatch", temperature=T_w, velocity=(0.0, 0.0, 0.0)),
          dict(patch=mesh.patch_index("flow"), boundaryModel="dsmcDeletionPatch")]
    inflow = [dict(patch=mesh.patch_index("flow"), typeIds=[0, 1, 2, 3, 4], numberDensities=dens, velocity=(U_inf, 0.0, 0.0),
                   translationalTemperature=T_inf, rotationalTemperature=T_inf, vibrationalTemperature=T_inf)]
    models = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=fnum, deltaT=dt, seed=seed,
                               rotationalRelaxationCollisionNumber=5.0, patch_models=pm, inflows=inflow)
    fill = dict(type_ids=[0, 1, 2, 3, 4], number_densities=dens, Ttra=T_inf, Trot=T_inf, Tvib=T_inf, velocity=(U_inf, 0.0, 0.0))
    return mesh, air5_species(), models, fill


def capsule_forebody(n_local=(200, 200, 200), ppc=31, dx=0.15, seed=0xD5C00004, procs=(1, 1, 1), rank=0, density_scale=1.0):
    """BASELINE configs[3]: 3-D re-entry capsule forebody in 5-species air at the orion107kmNR free stream (6053.4 m/s, 217.63 K;
    run/hyStrath/dsmcFoam+/orion107kmNR/system/boundariesDict), Larsen-Borgnakke with variable Zv, diffuse 1000 K heat shield.
    Cell size dx ~ lambda_inf / 3, dt = 0.3 dx / U_inf; procs = (px, py, pz) bricks of n_local cells each (decomposePar simple):
    the shock layer in front of the shield makes the bricks next to the body heavier than the upstream ones."""
    U_inf, T_inf, T_w = 6053.4, 217.63, 1000.0
    n_N2, n_O2 = 2.318e18 * density_scale, 6.161e17 * density_scale
    trace = 1e-3 * (n_N2 + n_O2)
    dens = [n_N2, n_O2, trace, trace, trace]
    mesh = meshgen.capsule_mesh(n_local, dx, procs, rank)
    fnum = sum(dens) * dx ** 3 / ppc          # equal statistical weight everywhere, from the undistorted inlet cell
    dt = 0.3 * dx / U_inf
    pm = [dict(patch=mesh.patch_index("capsule"), boundaryModel="dsmcDiffuseWallPatch", temperature=T_w, velocity=(0.0, 0.0, 0.0)),
          dict(patch=mesh.patch_index("flow"), boundaryModel="dsmcDeletionPatch"),
          dict(patch=mesh.patch_index("outflow"), boundaryModel="dsmcDeletionPatch")]
    inflow = [dict(patch=mesh.patch_index("flow"), typeIds=[0, 1, 2, 3, 4], numberDensities=dens, velocity=(U_inf, 0.0, 0.0),
                   translationalTemperature=T_inf, rotationalTemperature=T_inf, vibrationalTemperature=T_inf)]
    models = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=fnum, deltaT=dt, seed=seed,
                               rotationalRelaxationCollisionNumber=5.0, patch_models=pm, inflows=inflow)
    fill = dict(type_ids=[0, 1, 2, 3, 4], number_densities=dens, Ttra=T_inf, Trot=T_inf, Tvib=T_inf, velocity=(U_inf, 0.0, 0.0))
    return mesh, air5_species(), models, fill
