"""GPU tests of the quantum-kinetic chemistry inside the NTC loop (noTimeCounter.C:250-303 with DSMC/reactions/derived/{dissociationQK,
exchangeQK, mixed/dissociationExchangeQK}): the CUDA path through the C ABI against the CPU oracle on the same seeded input, and against
the time series the reference ships for its reacting tutorial (tests/golden/heatBath_5species.npz)."""
import os

import numpy as np
import pytest

from hystrath_b200 import capi, meshgen
from oracle.pyoracle import Oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def reacting_box(n=(4, 4, 4), ppc=60, T=25000.0, dens=2e22, dt=2e-8):
    case = H.heatbath_case()
    L = 4e-4
    mesh = meshgen.box_mesh(n, (L,) * 3)
    fnum = dens * L ** 3 / (np.prod(n) * ppc)
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=fnum, deltaT=dt, seed=0xD5C0C4E3,
                           rotationalRelaxationCollisionNumber=1.0, vibrationalRelaxationCollisionNumber=1.0, electronicRelaxationCollisionNumber=1.0)
    sp = H.air5()
    rx = capi.build_reactions(case["typeIdList"], case["reactions"])
    return mesh, sp, md, rx, T, dens


@pytest.mark.parametrize("n,ppc", [((4, 4, 4), 60), ((2, 1, 1), 70000)], ids=["64-cells", "two-cells-of-70000"])
def test_reacting_steps_match_the_oracle_parcel_by_parcel(n, ppc):
    """Five full steps of a periodic box of hot air in which every one of the 12 reactions can fire: reaction counts per reaction and
    channel, species of every parcel, the parcels created by dissociations (identity, order in the cloud, cell) and the occupancy equal
    the oracle's exactly; velocities and internal energies to the accuracy of pow()/exp() on the two sides.  The second case puts
    70 000 parcels in each of two cells: the bitmap sort and the one-block-per-cell candidate loop (512 candidates per batch) of cells beyond 65 536 parcels."""
    mesh, sp, md, rx, T, dens = reacting_box(n=n, ppc=ppc)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    for x in (eng, ora):
        x.set_reactions(rx)
    # molecules and atoms, so that exchange reactions have partners from the first step
    H.same_start(eng, ora, [0, 1, 2, 3, 4], [0.55 * dens, 0.15 * dens, 0.05 * dens, 0.15 * dens, 0.10 * dens], T, T, T)
    n0 = ora.num_parcels()
    for _ in range(5):
        eng.evolve(1)
        ora.evolve(1)
        assert np.array_equal(eng.reaction_counts(), ora.reaction_counts())
    rc = ora.reaction_counts()
    assert rc[:, :2].sum() > 100 and rc[:, 2].sum() > 10, rc        # dissociations and exchanges took place
    g, o = eng.download_parcels(), ora.download_parcels()
    assert g.n == o.n == n0 + rc[:, :2].sum()                        # one new parcel per dissociation
    assert np.array_equal(g.origId, o.origId)                        # the same cloud order, new parcels included
    assert np.array_equal(g.typeId, o.typeId)
    assert np.array_equal(g.cell, o.cell)
    assert np.array_equal(g.vibLevel, o.vibLevel)
    assert np.array_equal(eng.occupancy(), ora.occupancy())
    assert np.allclose(g.position, o.position, rtol=0, atol=1e-15)
    assert np.allclose(g.U, o.U, rtol=1e-9, atol=1e-6)
    assert np.allclose(g.ERot, o.ERot, rtol=1e-9, atol=1e-30)
    # atoms carry no internal energy, whatever they were before
    atoms = g.typeId >= 3
    assert np.all(g.ERot[atoms] == 0) and np.all(g.vibLevel[atoms] == 0)
    eng.close()


def test_reactions_conserve_mass_momentum_and_total_energy():
    """One collide stage with chemistry: mass and momentum of every cell are unchanged, and the energy a cell loses equals the heats
    of the reactions that took place in it (dissociation: k theta_d of the molecule; exchange: the dictionary's heat of reaction)."""
    mesh, sp, md, rx, T, dens = reacting_box(ppc=120)
    eng, ora = H.setup_pair(mesh, sp, md, capi.Engine, Oracle)
    eng.set_reactions(rx)
    H.same_start(eng, ora, [0, 1, 2, 3, 4], [0.55 * dens, 0.15 * dens, 0.05 * dens, 0.15 * dens, 0.10 * dens], T, T, T)
    mass = np.array([s.mass for s in sp]); thv = np.array([s.thetaV[0] for s in sp]); thd = np.array([s.thetaD for s in sp])

    def totals(p):
        m = mass[p.typeId]
        e = 0.5 * m * (p.U ** 2).sum(1) + p.ERot + p.vibLevel[:, 0] * H.KB * thv[p.typeId]
        return m.sum(), (m[:, None] * p.U).sum(0), e.sum()

    eng.stage(capi.STAGE_SORT)
    a = eng.download_parcels()
    eng.stage(capi.STAGE_COLLIDE)
    b = eng.download_parcels()
    rc = eng.reaction_counts()
    assert b.n == a.n + rc[:, :2].sum() and rc.sum() > 200
    m0, p0, e0 = totals(a)
    m1, p1, e1 = totals(b)
    assert abs(m1 / m0 - 1) < 1e-12
    assert np.abs(p1 - p0).max() < 1e-10 * np.abs(mass[a.typeId][:, None] * a.U).sum()
    case = H.heatbath_case()
    ids = {n: i for i, n in enumerate(case["typeIdList"])}
    heat = 0.0
    for k, r in enumerate(case["reactions"]):
        for ch in range(2):
            heat += rc[k, ch] * H.KB * thd[ids[r["reactants"][ch]]]
        if "heatOfReactionExchange" in r:
            heat -= rc[k, 2] * H.KB * r["heatOfReactionExchange"]
    assert abs((e0 - e1) - heat) < 1e-9 * e0, ((e0 - e1) / heat)
    eng.close()


def test_reacting_heat_bath_follows_the_shipped_series():
    """The reference's heatBath-5species tutorial at its own size (one cell, 154 000 parcels, 12 reactions) on the GPU: species
    densities and temperatures within 3 (sigma + 1 %) of the time series the reference ships, over the first 1000 steps.  The single
    cell of 1.5e5 parcels also exercises the large-cell sort and collide paths."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "heatBath_5species.npz"))
    eng = capi.Engine(0)
    case, spd, fnum, vol = H.heatbath_setup(eng, scale=1.0)
    eng.mesh_fill([0, 1], [case["numberDensities"]["N2"], case["numberDensities"]["O2"]], case["temperature"], case["temperature"], case["temperature"])
    n0 = eng.num_parcels()
    assert abs(n0 - 154118) < 2000
    steps = np.arange(100, 1001, 100)
    s = H.heatbath_series(eng, spd, fnum, vol, steps)
    worst = H.heatbath_check(s, gold, steps, fnum, vol, case)
    assert max(worst.values()) < 3.0, worst
    rc = eng.reaction_counts()
    assert eng.num_parcels() == n0 + rc[:, :2].sum()
    eng.close()


def test_single_cell_heat_bath_relaxes_to_equilibrium_and_is_reproducible():
    """The heatBath-5species set-up with `reactions ()`: one adiabatic cell of 154 000 parcels (block-level sort, one-warp candidate loop
    with 1e4 candidates per step), translational energy only at the start.  With all relaxation numbers 1 the three temperatures meet at
    the value energy conservation dictates; the total energy is conserved to rounding; two runs give the same cloud bit for bit."""
    from oracle import fields_ref

    def run(n_relax):
        eng = capi.Engine(0)
        case, spd, fnum, vol = H.heatbath_setup(eng, scale=1.0, reactions=False)
        eng.mesh_fill([0, 1], [case["numberDensities"]["N2"], case["numberDensities"]["O2"]], 9000.0, 0.0, 0.0)
        e0 = eng.counters()
        eng.evolve(n_relax)       # ~0.011 collisions per molecule and step: 1000 steps are eleven collision times
        eng.reset_accumulators()
        eng.evolve(20)
        acc, coll, nt = eng.accumulators()
        f = fields_ref.derive(acc, coll, nt, spd, [0, 1], fnum, np.array([vol]))
        e1 = eng.counters()
        p = eng.download_parcels()
        eng.close()
        return f, e0, e1, p

    f, e0, e1, p = run(1000)
    tot0 = e0.linearKineticEnergy + e0.rotationalEnergy + e0.vibrationalEnergy
    tot1 = e1.linearKineticEnergy + e1.rotationalEnergy + e1.vibrationalEnergy
    assert e0.rotationalEnergy == 0 and e0.vibrationalEnergy == 0 and abs(tot1 / tot0 - 1) < 1e-10
    Ttra, Trot, Tvib = f["Ttra"][0], f["Trot"][0], f["Tvib"][0]
    assert abs(Trot / Ttra - 1) < 0.02 and abs(Tvib / Ttra - 1) < 0.03, (Ttra, Trot, Tvib)
    # energy balance per molecule: 3/2 k T0 = (3/2 + 1) k T + <e_vib>(T) with the harmonic-oscillator mean of the two species
    x = np.array([case_n for case_n in (1.21753030168e22, 3.23647295384e21)]); x = x / x.sum()
    thv = np.array([3371.0, 2256.0])
    evib = (x * thv / np.expm1(thv / Ttra)).sum()
    assert abs((2.5 * Ttra + evib) / (1.5 * 9000.0) - 1) < 0.01
    _, _, _, p = run(100)
    _, _, _, p2 = run(100)
    for k in ("origId", "cell", "typeId", "position", "U", "ERot", "vibLevel"):
        assert np.array_equal(getattr(p, k), getattr(p2, k)), k


def test_reaction_lists_are_checked_like_the_reference():
    """dsmcReactions::initialConfiguration and <model>::setProperties (dsmcReactions.C:137-170, dissociationQK.C:44-195, exchangeQK.C:44-176):
    two models for one typeId pair, molecular dissociation products of a diatomic, an exchange without an atom."""
    names = ["N2", "O2", "NO", "N", "O"]

    def engine_with(reactions):
        eng = capi.Engine(0)
        mesh = meshgen.box_mesh((2, 2, 2), (1e-4,) * 3)
        eng.set_mesh(mesh); eng.set_species(H.air5())
        eng.set_reactions(capi.build_reactions(names, reactions))
        eng.set_models(capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=1e6, deltaT=1e-9, seed=1))
        return eng

    diss = dict(reactionModel="dissociationQK", reactants=["O2", "N2"], dissociationProducts=[["O", "O"], ["N", "N"]])
    eng = engine_with([diss, dict(diss, reactants=["N2", "O2"], dissociationProducts=[["N", "N"], ["O", "O"]])])
    with pytest.raises(capi.Dsmcb200Error, match="more than one reaction model specified for the typeId pair: 0 and 1"):
        eng.mesh_fill([0, 1], [1e22, 1e22], 1000.0)
    eng.close()
    eng = engine_with([dict(diss, dissociationProducts=[["O", "NO"], ["N", "N"]])])
    with pytest.raises(capi.Dsmcb200Error, match="Dissociation product of a diatomic molecule must be an atom"):
        eng.mesh_fill([0, 1], [1e22, 1e22], 1000.0)
    eng.close()
    eng = engine_with([dict(reactionModel="exchangeQK", reactants=["O2", "N2"], exchangeProducts=["NO", "NO"], heatOfReactionExchange=1.0, aCoeff=0.1, bCoeff=0.1)])
    with pytest.raises(capi.Dsmcb200Error, match="None of the reactants is an atom"):
        eng.mesh_fill([0, 1], [1e22, 1e22], 1000.0)
    eng.close()
    with pytest.raises(capi.Dsmcb200Error, match="Valid reaction types are"):
        capi.build_reactions(names, [dict(diss, reactionModel="ionisationQK")])
    with pytest.raises(capi.Dsmcb200Error, match="Cannot find type id: Xe"):
        capi.build_reactions(names, [dict(diss, reactants=["O2", "Xe"])])
