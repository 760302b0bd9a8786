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


# ---- the reacting tutorial (run/hyStrath/dsmcFoam+/heatBath-5species): fixtures of tests/golden/make_golden_heatbath.py ----
def heatbath_case():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "heatBath_5species_reactions.json")) as f:
        return json.load(f)


def heatbath_setup(x, scale=1.0, seed=7, reactions=True):
    """One adiabatic cell (specular walls) of N2/O2 at 30 000 K with the tutorial's 12 QK reactions on engine / oracle `x`;
    scale < 1 runs with fewer, heavier parcels (nEquivalentParticles = 100 / scale).  Returns (case, species dicts, fnum, cell volume)."""
    case = heatbath_case()
    L = case["cellSize"]
    mesh = meshgen.box_mesh((1, 1, 1), (L,) * 3, sides={s: ("wall", "fixedWalls") for s in meshgen.SIDES})
    fnum = case["nEquivalentParticles"] / scale
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=fnum, deltaT=case["deltaT"], seed=seed,
                           rotationalRelaxationCollisionNumber=1.0, vibrationalRelaxationCollisionNumber=1.0,
                           electronicRelaxationCollisionNumber=1.0, patch_models=[dict(patch=0, boundaryModel="dsmcSpecularWallPatch")])
    sp = air5()
    x.set_mesh(mesh); x.set_species(sp); x.set_models(md)
    if reactions:
        x.set_reactions(capi.build_reactions(case["typeIdList"], case["reactions"]))
    spd = [dict(mass=s.mass, diameter=s.diameter, omega=s.omega, rotDof=s.rotationalDegreesOfFreedom,
                thetaV=[s.thetaV[i] for i in range(s.nVibrationalModes)]) for s in sp]
    return case, spd, fnum, L ** 3


def heatbath_series(x, spd, fnum, vol, steps):
    """Evolve to each step of `steps` (ascending) and return the fields of THAT step alone (the tutorial samples and resets every
    step): dict of arrays rhoN_<species>, Ttra, Trot, Tvib, N (parcels)."""
    from oracle import fields_ref
    names = ["N2", "O2", "NO", "N", "O"]
    out = {k: [] for k in ["Ttra", "Trot", "Tvib", "N"] + [f"rhoN_{n}" for n in names]}
    done = 0
    for s in steps:
        if s - 1 > done:
            x.evolve(s - 1 - done)
        a0, _, _ = x.accumulators()
        x.evolve(1)
        done = s
        a1, _, _ = x.accumulators()
        acc = a1 - a0
        f = fields_ref.derive(acc, None, 1.0, spd, [0, 1, 2, 3, 4], fnum, np.array([vol]))
        out["Ttra"].append(f["Ttra"][0]); out["Trot"].append(f["Trot"][0]); out["Tvib"].append(f["Tvib"][0])
        out["N"].append(acc[0, :, 0].sum())
        for k, n in enumerate(names):
            out[f"rhoN_{n}"].append(acc[0, k, 0] * fnum / vol)
    return {k: np.array(v) for k, v in out.items()}


def heatbath_check(series, gold, steps, fnum, vol, case):
    """How far the run is from the shipped time series, per quantity the largest deviation over the sampled steps in units of
    (sigma + 1 %).  Densities: sigma from the two parcel counts (a reaction converts whole parcels, so a count scatters at most like
    a Poisson variable; the shipped run is one realisation as well); temperatures: the sqrt(2 / (3 N)) scatter of both runs.  The 1 %
    covers the history the two realisations do not share."""
    idx = np.searchsorted(gold["step"], steps)
    assert np.array_equal(gold["step"][idx], steps)
    pp_ours, pp_ref = fnum / vol, case["nEquivalentParticles"] / vol   # number density one parcel stands for
    worst = {}
    for n in ["N2", "O2", "NO", "N", "O"]:
        ours, ref = series[f"rhoN_{n}"], gold[f"rhoN_{n}"][idx]
        sigma = np.sqrt(np.maximum(ours, pp_ours) * pp_ours + np.maximum(ref, pp_ref) * pp_ref)
        worst[n] = (np.abs(ours - ref) / (sigma + 0.01 * ref.max())).max()
    n_ref = sum(gold[f"rhoN_{n}"][idx] for n in ["N2", "O2", "NO", "N", "O"]) / pp_ref
    for k in ["Ttra", "Trot"]:
        ours, ref = series[k], gold[f"{k}_mixture"][idx]
        sigma = ref * np.sqrt(2.0 / (3.0 * series["N"]) + 2.0 / (3.0 * n_ref))
        worst[k] = (np.abs(ours - ref) / (sigma + 0.01 * ref)).max()
    return worst


# ---- the axisymmetric tutorial (run/hyStrath/dsmcFoam+/axisymmetricFlatnosedCylinder): fixture of tests/golden/make_golden_axisym.py ----
def axisym_gold():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "axisymmetricFlatnosedCylinder.npz"))


def axisym_setup(x, gold, geometry, seed=11, fnum_scale=1.0, mesh=None):
    """The tutorial on engine / oracle `x`: Mach-5.4 argon (1e21 m^-3, 100 K, 1000 m/s) onto a flat-nosed cylinder with a 300 K diffuse
    wall, free-stream inflow + deletion on `flow`, symmetry planes on the wedge sides, dsmcAxisymmetric about x with radial weighting
    method "cell" and maxRadialWeightingFactor 1000.  geometry: callable returning (cell centres, ..., face centres) after set_mesh.
    Returns (mesh, species dicts, per-cell nParticles = F_N * RWF)."""
    mesh = mesh if mesh is not None else meshgen.axisymmetric_cylinder_mesh()
    sp = [capi.make_species("Ar", float(gold["mass"]), float(gold["diameter"]), float(gold["omega"]), float(gold["alpha"]))]
    rev, pol, ang = capi.axisymmetric_axes()
    flow, cyl = mesh.patch_index("flow"), mesh.patch_index("cylinder")
    fnum = float(gold["nEquivalentParticles"]) * fnum_scale
    md = capi.build_models("VariableHardSphere", nEquivalentParticles=fnum, deltaT=float(gold["deltaT"]), seed=seed,
                           coordinateSystem="dsmcAxisymmetric", angularCoordinate=ang,
                           patch_models=[dict(patch=cyl, boundaryModel="dsmcDiffuseWallPatch", temperature=float(gold["wallTemperature"]), velocity=(0, 0, 0)),
                                         dict(patch=flow, boundaryModel="dsmcDeletionPatch")],
                           inflows=[dict(patch=flow, typeIds=[0], numberDensities=[float(gold["numberDensity"])], velocity=tuple(gold["velocity"]),
                                         translationalTemperature=float(gold["temperature"]))])
    x.set_mesh(mesh); x.set_species(sp); x.set_models(md)
    cc, cv, fc, *_ = geometry()
    rwf, _ = capi.axisymmetric_rwf(cc, fc, pol, float(gold["maxRadialWeightingFactor"]))
    x.set_cell_fields(RWF=rwf)
    spd = [dict(mass=sp[0].mass, diameter=sp[0].diameter, omega=sp[0].omega, rotDof=0.0, thetaV=[])]
    return mesh, spd, fnum * rwf, cv


def axisym_envelope(g):
    """Per cell the smallest and largest value of field g over the cell and its four neighbours inside its block (the tutorial mesh is
    three structured blocks of 40 x 20, 40 x 40 and 40 x 40 cells): what a cell may show when a steep front sits a fraction of a cell
    away from where the shipped run has it."""
    lo, hi = g.copy(), g.copy()
    off = 0
    for nx, nz in ((40, 20), (40, 40), (40, 40)):
        b = g[off:off + nx * nz].reshape(nz, nx)
        l, h = b.copy(), b.copy()
        for r in (np.vstack([b[:1], b[:-1]]), np.vstack([b[1:], b[-1:]]), np.hstack([b[:, :1], b[:, :-1]]), np.hstack([b[:, 1:], b[:, -1:]])):
            l, h = np.minimum(l, r), np.maximum(h, r)
        lo[off:off + nx * nz], hi[off:off + nx * nz] = l.ravel(), h.ravel()
        off += nx * nz
    return lo, hi
