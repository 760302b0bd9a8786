"""numpy restatement of the per-cell field reduction of dsmcVolFields::calculateField
(DSMC/macroscopicProperties/derived/combined/dsmcVolFields/dsmcVolFields.C:1242-1290 every step,
:1401-1508 and :1663-1790 at output time), fed with the per-species moment sums of the sampling
stage (layout of include/dsmcb200.h: N, sum U (3), sum U.U, ERot, Eelec, Evib per mode).

TEST INFRASTRUCTURE (oracle side).  Pinned against the fields the reference ships for
couette_N2-O2/backup-5 (tests/test_oracle_golden.py): rhoN = dsmcNMean*FN/V, p = rhoN k Ttra with
k = 1.38065e-23, and the mean-free-path / mean-collision-time formulas (Bird 4.74/4.76/4.77/1.38).
"""
import numpy as np

SMALL, VSMALL, GREAT = 1e-15, 1e-300, 1e15


def derive(acc, coll, n_time_steps, species, type_ids, fnum, cell_volumes, kB=1.38065e-23, deltaT=1.0,
           mfp_tref=273.0, has_internal=True, n_modes=1):
    """acc: [nCells, nSpecies, nQ]; species: list of dicts(mass, diameter, omega, rotDof, thetaV[list]);
    type_ids: the `typeIds` of the dsmcVolFields instance (subset of species indices); fnum: nEquivalentParticles, or
    cloud_.nParticles(cell) per cell for radial weighting / variable time steps (dsmcVolFields.C:1098,1128)."""
    acc = np.asarray(acc, float)
    V = np.asarray(cell_volumes, float)
    nT = float(n_time_steps)
    ids = list(type_ids)
    m = np.array([species[s]["mass"] for s in ids])
    zeta = np.array([species[s].get("rotDof", 0.0) for s in ids])
    Ns = acc[:, ids, 0]                                  # dsmcNSpeciesCum
    dsmcNCum = Ns.sum(1)
    nCum = fnum * dsmcNCum
    mCum = fnum * (Ns * m).sum(1)
    momentumCum = np.reshape(fnum, (-1, 1)) * (acc[:, ids, 1:4] * m[None, :, None]).sum(1)   # fnum: scalar or per cell (nParticles(cell))
    linearKECum = fnum * (acc[:, ids, 4] * m).sum(1)
    out = {}
    ok = dsmcNCum > 1e-3
    with np.errstate(divide="ignore", invalid="ignore"):
        out["dsmcNMean"] = np.where(ok, dsmcNCum / nT, 0.001)
        rhoN = np.where(ok, nCum / (nT * V), 0.0)
        rhoM = np.where(ok, mCum / (nT * V), 0.0)
        UMean = np.where(ok[:, None], momentumCum / mCum[:, None], 0.0)
        linearKEMean = 0.5 * linearKECum / (V * nT)
        Ttra = np.where(ok, 2.0 / (3.0 * kB * rhoN) * (linearKEMean - 0.5 * rhoM * (UMean * UMean).sum(1)), 0.0)
        out.update(rhoN=rhoN, rhoM=rhoM, UMean=UMean, Ttra=Ttra, p=rhoN * kB * Ttra)
        if has_internal:
            ErotCum = acc[:, ids, 5].sum(1)
            ZetaRotCum = (Ns * zeta).sum(1)
            zetaRotTot = np.where(dsmcNCum > SMALL, ZetaRotCum / dsmcNCum, 0.0)
            Trot = np.where(ZetaRotCum > SMALL, 2.0 * ErotCum / (kB * ZetaRotCum), 0.0)
            Tvib = np.zeros_like(Trot)
            zetaVib = np.zeros_like(Trot)
            moleculesRhoN = np.zeros_like(Trot)
            for k, s in enumerate(ids):
                thv = species[s].get("thetaV", [])
                spZeta = np.zeros_like(Trot)
                zetaByT = np.zeros_like(Trot)
                for mode, th in enumerate(thv):
                    E = acc[:, s, 7 + mode]
                    good = (E > VSMALL) & (Ns[:, k] > SMALL)
                    iMean = np.where(good, E / (kB * th * Ns[:, k]), 0.0)
                    good &= iMean > SMALL
                    logF = np.where(good, np.log(1.0 + 1.0 / np.where(good, iMean, 1.0)), 1.0)
                    TvibMod = np.where(good, th / logF, 0.0)
                    zMod = np.where(good, 2.0 * iMean * logF, 0.0)
                    spZeta = spZeta + zMod
                    zetaByT = np.where(good, zMod * TvibMod, zetaByT)   # assigned, not accumulated (:1469)
                has = spZeta > SMALL
                nS = fnum * Ns[:, k]
                moleculesRhoN += np.where(has, nS, 0.0)
                Tvib += np.where(has, nS * zetaByT / np.where(has, spZeta, 1.0), 0.0)
                zetaVib += np.where(has, nS * spZeta, 0.0)
            Tvib = np.where(moleculesRhoN > SMALL, Tvib / moleculesRhoN, Tvib)
            zetaVib = np.where(moleculesRhoN > SMALL, zetaVib / moleculesRhoN, zetaVib)
            Tov = (3.0 * Ttra + zetaRotTot * Trot + zetaVib * Tvib) / (3.0 + zetaRotTot + zetaVib)
            out.update(Trot=Trot, Tvib=Tvib, Tov=Tov, zetaVib=zetaVib)
        # mean free path, mean collision rate (dsmcVolFields.C:1663-1790)
        mfp = np.zeros_like(rhoN)
        mcr = np.zeros_like(rhoN)
        valid = Ttra > 1.0
        T = np.where(valid, Ttra, 1.0)
        for k, sp in enumerate(ids):
            spMfp = np.zeros_like(rhoN)
            spMcr = np.zeros_like(rhoN)
            for r, sq in enumerate(ids):
                dPQ = 0.5 * (species[sp]["diameter"] + species[sq]["diameter"])
                omegaPQ = 0.5 * (species[sp]["omega"] + species[sq]["omega"])
                massRatio = species[sp]["mass"] / species[sq]["mass"]
                reduced = species[sp]["mass"] * species[sq]["mass"] / (species[sp]["mass"] + species[sq]["mass"])
                has = Ns[:, r] > SMALL
                nDensQ = fnum * Ns[:, r] / (V * nT)
                spMfp += np.where(has, np.pi * dPQ ** 2 * nDensQ * (mfp_tref / T) ** (omegaPQ - 0.5) * np.sqrt(1.0 + massRatio), 0.0)
                spMcr += np.where(has, 2.0 * np.sqrt(np.pi) * dPQ ** 2 * nDensQ * (T / mfp_tref) ** (1.0 - omegaPQ)
                                  * np.sqrt(2.0 * kB * mfp_tref / reduced), 0.0)
            spMfp = np.where(spMfp > SMALL, 1.0 / np.where(spMfp > SMALL, spMfp, 1.0), spMfp)
            w = np.where(nCum > 0, fnum * Ns[:, k] / np.where(nCum > 0, nCum, 1.0), 0.0)
            mfp += spMfp * w
            mcr += spMcr * w
        mfp = np.where(valid & (mfp >= SMALL), mfp, GREAT)
        out["mfp"] = mfp
        out["meanCollisionRate"] = np.where(valid, mcr, 0.0)
        out["mct"] = np.where(valid & (mcr > SMALL), 1.0 / np.where(mcr > SMALL, mcr, 1.0), GREAT)
        if coll is not None:
            coll = np.asarray(coll, float)
            out["measuredCollisionRate"] = np.where(nCum > SMALL, coll[:, 0] * fnum / (np.where(nCum > SMALL, nCum, 1.0) * deltaT), 0.0)
            out["meanCollisionSeparation"] = np.where(coll[:, 0] > SMALL, coll[:, 1] / np.where(coll[:, 0] > SMALL, coll[:, 0], 1.0), GREAT)
    return out


def wall_fields(wall, n_time_steps, species, type_ids, fnum, face_areas, face_centres, first_points, kB=1.38065e-23):
    """Wall-face values of one dsmcVolFields instance from the boundary accumulators (dsmcVolFields.C:1878-2141 and the unit vectors
    of calculateWallUnitVectors :52-80).  wall: [nFaces, nSpecies, nWallQ] in the engine's WallQ order (rhoN 0, rhoNInt 1,
    rhoNElec 2, rhoM 3, linearKE 4, mcc 5, momentum 6-8, Erot 9, zetaRot 10, Evib 11, Eelec 12, q 13, fD 14-16, EvibMod 17+);
    face_areas / face_centres / first_points: [nFaces, 3] of the same faces.  Returns a dict of per-face arrays."""
    w = np.asarray(wall, float)
    ids = list(type_ids)
    nT = float(n_time_steps)
    nF = w.shape[0]
    rhoNBF = w[:, ids, 0].sum(1); rhoMBF = w[:, ids, 3].sum(1); linearKEBF = w[:, ids, 4].sum(1)
    momBF = w[:, ids, 6:9].sum(1); ErotBF = w[:, ids, 9].sum(1); zetaRotBF = w[:, ids, 10].sum(1)
    rhoN, rhoM, lke = rhoNBF * fnum / nT, rhoMBF * fnum / nT, linearKEBF * fnum / nT
    out = {"rhoN": rhoN, "rhoM": rhoM}
    with np.errstate(divide="ignore", invalid="ignore"):
        has = rhoM > VSMALL
        U = np.where(has[:, None], momBF / np.where(has, rhoMBF, 1.0)[:, None], 0.0)
        Ttra = np.where(has, 2.0 / (3.0 * kB * np.where(has, rhoN, 1.0)) * (lke - 0.5 * rhoM * (U * U).sum(1)), 0.0)
        zetaRotTot = np.where(rhoNBF > SMALL, zetaRotBF / np.where(rhoNBF > SMALL, rhoNBF, 1.0), 0.0)
        Trot = np.where(zetaRotBF > SMALL, 2.0 * ErotBF / (kB * np.where(zetaRotBF > SMALL, zetaRotBF, 1.0)), 0.0)
        Tvib = np.zeros(nF); zetaVib = np.zeros(nF); molecules = np.zeros(nF)
        for s in ids:
            spRhoN = w[:, s, 0]
            spZeta = np.zeros(nF); zByT = np.zeros(nF)
            for mod, thetaV in enumerate(species[s].get("thetaV", [])):
                ev = w[:, s, 17 + mod] if 17 + mod < w.shape[2] else np.zeros(nF)
                iMean = np.where(spRhoN > SMALL, ev / (kB * thetaV * np.where(spRhoN > SMALL, spRhoN, 1.0)), 0.0)
                okm = iMean > SMALL
                logF = np.log(1.0 + 1.0 / np.where(okm, iMean, 1.0))
                Tm = np.where(okm, thetaV / logF, 0.0); zm = np.where(okm, 2.0 * iMean * logF, 0.0)
                spZeta += zm; zByT += zm * Tm
            oks = spZeta > SMALL
            molecules += np.where(oks, spRhoN, 0.0)
            Tvib += np.where(oks, spRhoN * zByT / np.where(oks, spZeta, 1.0), 0.0)
            zetaVib += np.where(oks, spRhoN * spZeta, 0.0)
        okM = molecules > SMALL
        Tvib = np.where(okM, Tvib / np.where(okM, molecules, 1.0), Tvib)
        zetaVib = np.where(okM, zetaVib / np.where(okM, molecules, 1.0), zetaVib)
        Tov = (3.0 * Ttra + zetaRotTot * Trot + zetaVib * Tvib) / (3.0 + zetaRotTot + zetaVib)
        X = w[:, ids, 0] / np.where(rhoNBF > SMALL, rhoNBF, 1.0)[:, None]
        mass = np.array([species[s]["mass"] for s in ids]); zr = np.array([species[s].get("rotDof", 0.0) for s in ids])
        mm = (X * mass).sum(1); cv = (X * (3.0 + zr)).sum(1); cp = (X * (5.0 + zr)).sum(1)
        R = np.where(Ttra > SMALL, kB / np.where(mm > 0, mm, 1.0), 0.0)
        a = np.sqrt(cp / np.where(cv > 0, cv, 1.0) * R * Ttra)
        Ma = np.where(rhoNBF > SMALL, np.sqrt((U * U).sum(1)) / a, 0.0)
    Sf = np.asarray(face_areas, float)
    n = Sf / np.linalg.norm(Sf, axis=1)[:, None]
    t1 = np.asarray(face_centres, float) - np.asarray(first_points, float)
    t1 /= np.linalg.norm(t1, axis=1)[:, None]
    t2 = np.cross(n, t1)
    t2 /= np.linalg.norm(t2, axis=1)[:, None]
    fD = w[:, ids, 14:17].sum(1) / nT
    out.update(U=U, Ttra=Ttra, Trot=Trot, Tvib=Tvib, Tov=Tov, Ma=Ma, fD=fD, p=(fD * n).sum(1),
               wallShearStress=np.sqrt((fD * t1).sum(1) ** 2 + (fD * t2).sum(1) ** 2), wallHeatFlux=w[:, ids, 13].sum(1) / nT)
    return out


def flux_fields(acc, n_time_steps, species, type_ids, fnum, cell_volumes, q_flux, n_modes=1, has_internal=True):
    """pressureTensor, shearStressTensor and heatFluxVector of dsmcVolFields.C:1509-1622 from the twelve extra moments the sampling
    stage keeps with measureHeatFluxShearStress (uu uv uw vv vw ww, c^2 u c^2 v c^2 w, E_int u E_int v E_int w at index q_flux..)."""
    acc = np.asarray(acc, float)
    ids = list(type_ids)
    V, nT = np.asarray(cell_volumes, float), float(n_time_steps)
    m = np.array([species[s]["mass"] for s in ids])
    Ns = acc[:, ids, 0]
    dsmcNCum = Ns.sum(1)
    dsmcMCum = (Ns * m).sum(1)
    ok = dsmcNCum > SMALL
    safeN = np.where(ok, dsmcNCum, 1.0)
    rhoN = fnum * dsmcNCum / (nT * V)
    U = (acc[:, ids, 1:4] * m[None, :, None]).sum(1) / np.where(ok, dsmcMCum, 1.0)[:, None]
    M = (acc[:, ids, q_flux:q_flux + 6] * m[None, :, None]).sum(1)
    Mcc = (acc[:, ids, q_flux + 6:q_flux + 9] * m[None, :, None]).sum(1)
    E = acc[:, ids, q_flux + 9:q_flux + 12].sum(1)
    MccAll = (acc[:, ids, 4] * m).sum(1)
    ECum = np.zeros(len(V))
    if has_internal:
        ECum = acc[:, ids, 5].sum(1)
        for k, s in enumerate(ids):
            for mode in range(len(species[s].get("thetaV", []))):
                ECum = ECum + acc[:, s, 7 + mode]
    k0 = rhoN / safeN
    idx = [[0, 1, 2], [1, 3, 4], [2, 4, 5]]
    P = np.zeros((len(V), 3, 3))
    for a in range(3):
        for b in range(3):
            P[:, a, b] = k0 * (M[:, idx[a][b]] - dsmcMCum * U[:, a] * U[:, b])
    sp = (P[:, 0, 0] + P[:, 1, 1] + P[:, 2, 2]) / 3.0
    T = -P.copy()
    for a in range(3):
        T[:, a, a] += sp
    q = np.zeros((len(V), 3))
    for a in range(3):
        q[:, a] = k0 * (0.5 * Mcc[:, a] - 0.5 * MccAll * U[:, a] + E[:, a] - ECum * U[:, a]) - (P[:, a, :] * U).sum(1)
    z = ~ok
    P[z] = 0; T[z] = 0; q[z] = 0
    return dict(pressureTensor=P.reshape(-1, 9), shearStressTensor=T.reshape(-1, 9), heatFluxVector=q)
