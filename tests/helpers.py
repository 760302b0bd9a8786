"""Shared case builders for the parity tests: species tables copied from the shipped case
dictionaries (run/hyStrath/dsmcFoam+/{hypersonicCorner,heatBath-5species}/constant/dsmcProperties)."""
import numpy as np

from hystrath_b200 import capi, meshgen

KB = 1.38065e-23


def argon():
    return capi.make_species("Ar", 66.3e-27, 4.17e-10, 0.81)


def air5():
    return [
        capi.make_species("N2", 46.5e-27, 4.17e-10, 0.74, 1.36, 2, (3371,), (52560,), (3371,), 113500),
        capi.make_species("O2", 53.12e-27, 4.07e-10, 0.77, 1.4, 2, (2256,), (17900,), (2256,), 59500),
        capi.make_species("NO", 49.81e-27, 4.2e-10, 0.79, 1.0, 2, (2719,), (1400,), (2719,), 75500),
        capi.make_species("N", 23.25e-27, 3.0e-10, 0.8, 1.0),
        capi.make_species("O", 26.56e-27, 3.0e-10, 0.8, 1.0),
    ]


def setup_pair(mesh, species, models, engine_cls, oracle_cls):
    eng, ora = engine_cls(0), oracle_cls()
    for x in (eng, ora):
        x.set_mesh(mesh)
        x.set_species(species)
        x.set_models(models)
    return eng, ora


def same_start(eng, ora, type_ids, densities, T, Trot=0.0, Tvib=0.0, velocity=(0, 0, 0)):
    """Fill with the oracle's dsmcMeshFill and give the identical cloud + cell state to the engine."""
    ora.mesh_fill(type_ids, densities, T, Trot, Tvib, 0.0, velocity)
    start = ora.download_parcels()
    sig, rem = ora.download_cellstate()
    eng.upload_parcels(start)
    eng.upload_cellstate(sig, rem)
    return start


def by_id(p):
    """Sort a ParcelData by origId -> dict of arrays (order-independent comparison)."""
    o = np.argsort(p.origId, kind="stable")
    return {k: getattr(p, k)[o] for k in ("position", "U", "ERot", "cell", "tetFace", "tetPt", "typeId", "vibLevel", "ELevel", "origId")}


def vhs_equilibrium_collision_rate(n, T, sp, Tref=273.0):
    """Collisions per unit volume and time in an equilibrium simple gas, Bird (1994) eq. 4.64:
    N_c = 2 sqrt(pi) d_ref^2 n^2 (T/T_ref)^(1-omega) sqrt(k T_ref / m) ... written via the mean rate nu:
    nu = 4 d^2 n sqrt(pi k Tref / m) (T/Tref)^(1-omega)  (SURVEY 8c iii), collisions/volume/time = n nu / 2."""
    nu = 4.0 * sp.diameter ** 2 * n * np.sqrt(np.pi * KB * Tref / sp.mass) * (T / Tref) ** (1.0 - sp.omega)
    return 0.5 * n * nu
