#!/usr/bin/env python
"""Generates tests/golden/heatBath_5species.npz and heatBath_5species_reactions.json from the reacting tutorial the reference ships
(run/hyStrath/dsmcFoam+/heatBath-5species): the probe time series its own dsmcFoam+ run wrote (gnuplot/solution/*: one adiabatic cell of
N2/O2 at 30 000 K relaxing through the 12 quantum-kinetic reactions of system/chemReactDict, sampled and reset every step) and that
reaction list.  These are the only reference-side vectors of the chemistry path.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_heatbath.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from hystrath_b200 import foamfile as ff  # noqa: E402

CASE = "/root/reference/run/hyStrath/dsmcFoam+/heatBath-5species"
SERIES = ["Ttra_mixture", "Trot_mixture", "Tvib_mixture", "Tov_mixture", "rhoN_N2", "rhoN_O2", "rhoN_NO", "rhoN_N", "rhoN_O"]
STRIDE, LAST = 10, 5000   # every 10th of the first 5000 steps (the relaxation is over by then)


def reactions():
    d = ff.read_dict(os.path.join(CASE, "system", "chemReactDict"))
    out = []
    for name, r in d["reactions"]:
        e = dict(name=name, reactionModel=r["reactionModel"], reactants=list(r["reactants"]),
                 allowSplitting=str(r.get("allowSplitting", "yes")) in ("yes", "on", "true"))
        if "dissociationQKProperties" in r:
            e["dissociationProducts"] = [list(x) for x in r["dissociationQKProperties"]["dissociationProducts"]]
        if "exchangeQKProperties" in r:
            x = r["exchangeQKProperties"]
            e.update(exchangeProducts=list(x["exchangeProducts"]), heatOfReactionExchange=float(x["heatOfReactionExchange"]),
                     aCoeff=float(x["aCoeff"]), bCoeff=float(x["bCoeff"]))
        out.append(e)
    return out


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    out = {}
    for k in SERIES:
        a = np.loadtxt(os.path.join(CASE, "gnuplot", "solution", k))
        assert np.allclose(a[:LAST, 0], 1e-9 * np.arange(1, LAST + 1), rtol=1e-6)   # row i = state after step i + 1
        out[k] = a[STRIDE - 1:LAST:STRIDE, 1]
    out["step"] = np.arange(STRIDE, LAST + 1, STRIDE)
    np.savez_compressed(os.path.join(here, "heatBath_5species.npz"), **out)
    with open(os.path.join(here, "heatBath_5species_reactions.json"), "w") as f:
        json.dump(dict(typeIdList=["N2", "O2", "NO", "N", "O"], reactions=reactions(),
                       numberDensities=dict(N2=1.21753030168e22, O2=3.23647295384e21), temperature=30000.0, nEquivalentParticles=100.0,
                       deltaT=1e-9, cellSize=1e-5), f, indent=1)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
