"""GPU: the standalone driver (dsmcb200_run) on the couette_N2-O2 case directory -- dictionaries, polyMesh and
cloud files in the reference's layout -- against the oracle fed with the same state and seed."""
import os
import subprocess

import numpy as np
import pytest

from hystrath_b200 import capi, foamfile as ff
from oracle import fields_ref
from oracle.pyoracle import Oracle
from tests import casegen, helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "hystrath_b200", "dsmcb200_run")


def test_driver_runs_couette_case_and_matches_oracle(tmp_path):
    n_steps = 4
    g, mesh, p = casegen.couette_case(str(tmp_path), n_steps=n_steps, seed=5, nto=2)
    r = subprocess.run([RUN, "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    log = r.stdout
    # log strings the reference's monitor scripts grep for (dsmcCloud.C:960-981, noTimeCounter.C:320-337)
    assert "Number of DSMC particles        = 47583" in log
    assert "Collisions                      = " in log and "Average linear kinetic energy   = " in log
    assert "Time = 5.00002" in log and "End stage 0" in log

    # the same run on the oracle
    sp = H.air5()[:2]
    pm = [dict(patch=mesh.patch_index("upperWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=3000.0, velocity=(300.0, 0, 0)),
          dict(patch=mesh.patch_index("lowerWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=2000.0, velocity=(0, 0, 0))]
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=float(g["nEquivalentParticles"]), deltaT=1e-5,
                           seed=5, patch_models=pm, inverseZvFormulation="pre-2008", rotationalRelaxationCollisionNumber=5.0,
                           measureHeatFluxShearStress=True)       # the mixture field asks for it in fieldPropertiesDict
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
    o.upload_parcels(p)
    o.upload_cellstate(g["dsmcSigmaTcRMax"], None)
    o.set_step(500000)   # the driver keys its Philox streams by the global time index: startTime 5 / deltaT 1e-5
    o.evolve(n_steps)
    ref = o.download_parcels()

    tdir = os.path.join(str(tmp_path), "5.00004")
    assert os.path.isdir(tdir), os.listdir(str(tmp_path))
    cdir = os.path.join(tdir, "lagrangian", "dsmc")
    for name in ("positions", "U", "ERot", "typeId", "vibLevel", "newParcel", "classification", "origId", "origProcId"):
        assert os.path.exists(os.path.join(cdir, name)), name
    xyz, cell = ff.read_positions(os.path.join(cdir, "positions"))
    ids = ff.read_scalar_list(os.path.join(cdir, "origId"), np.int32)
    assert len(cell) == ref.n
    assert np.array_equal(ids, ref.origId)            # same physical order as the oracle
    assert np.array_equal(cell, ref.cell)             # cell indexing identical after 4 steps with walls + LB collisions
    assert np.allclose(xyz, ref.position, rtol=0, atol=5e-10)      # files carry 10 significant digits
    U = ff.read_vector_list(os.path.join(cdir, "U"))
    assert np.allclose(U, ref.U, rtol=1e-9, atol=1e-6)

    # sampled fields written by the driver == reduction of the oracle's accumulators
    acc, coll, nt = o.accumulators()
    _, cv, *_ = o.geometry()
    spd = [dict(mass=s.mass, diameter=s.diameter, omega=s.omega, rotDof=2.0, thetaV=[s.thetaV[0]]) for s in sp]
    for inst, idl in (("mixture", [0, 1]), ("N2", [0]), ("O2", [1])):
        f = fields_ref.derive(acc, coll, nt, spd, idl, float(g["nEquivalentParticles"]), cv, deltaT=1e-5)
        for name, key in (("rhoN", "rhoN"), ("rhoM", "rhoM"), ("Ttra", "Ttra"), ("p", "p"), ("Trot", "Trot"), ("Tvib", "Tvib"), ("Tov", "Tov"),
                          ("mfp", "mfp"), ("mct", "mct")):
            got = ff.read_internal_field(os.path.join(tdir, f"{name}_{inst}"))
            assert np.allclose(got, f[key], rtol=2e-9, atol=1e-300), (name, inst)
    # wall-face fields of the two wall patches == numpy restatement of dsmcVolFields.C:1878-2141 on the oracle's boundary accumulators
    wacc = o.wall_accumulators()
    _, _, fc, fa, _ = o.geometry()
    for inst, idl in (("mixture", [0, 1]), ("N2", [0])):
        for patch, row0 in (("upperWall", 0), ("lowerWall", 5)):     # patch models in boundariesDict order: 5 faces each
            start = mesh.patches[mesh.patch_index(patch)]["start"]
            faces = np.arange(start, start + 5)
            first = mesh.points[mesh.face_points[mesh.face_offsets[faces]]]
            wf = fields_ref.wall_fields(wacc[row0:row0 + 5], nt, spd, idl, float(g["nEquivalentParticles"]), fa[faces], fc[faces], first)
            for name, key in (("rhoN", "rhoN"), ("rhoM", "rhoM"), ("p", "p"), ("Ttra", "Ttra"), ("Trot", "Trot"), ("Tvib", "Tvib"), ("Tov", "Tov"),
                              ("Ma", "Ma"), ("wallHeatFlux", "wallHeatFlux"), ("wallShearStress", "wallShearStress"), ("U", "U"), ("fD", "fD")):
                got = ff.read_patch_field(os.path.join(tdir, f"{name}_{inst}"), patch)
                scale = np.abs(wf[key]).max() + 1e-300
                assert np.abs(got - wf[key]).max() / scale < 5e-9, (name, inst, patch)
        assert np.abs(wf["wallHeatFlux"]).max() > 0 and np.abs(wf["wallShearStress"]).max() > 0
    # measureHeatFluxShearStress / measureErrors of the mixture instance (dsmcVolFields.C:1509-1622, :1857-1873)
    fx = fields_ref.flux_fields(acc, nt, spd, [0, 1], float(g["nEquivalentParticles"]), cv, q_flux=8)
    for name, key in (("heatFluxVector", "heatFluxVector"), ("pressureTensor", "pressureTensor"), ("shearStressTensor", "shearStressTensor")):
        got = ff.read_internal_field(os.path.join(tdir, f"{name}_mixture"))
        assert got.shape == fx[key].shape and np.allclose(got, fx[key], rtol=2e-8, atol=1e-9 * np.abs(fx[key]).max()), name
    # measureErrors: the reference guards the four estimates with particleCv > SMALL where particleCv = molarCv/(0.5 k)/N_A ~ 1e-26
    # (dsmcVolFields.C:1647,1859-1864), so its error fields are written but always zero; the driver keeps that
    for name in ("rhoMError", "UError", "TError", "pError"):
        assert not ff.read_internal_field(os.path.join(tdir, f"{name}_mixture")).any()
    assert not os.path.exists(os.path.join(tdir, "heatFluxVector_N2"))
    # dsmcN_: the instantaneous parcel count per cell of the instance (AUTO_WRITE in the reference)
    nN2 = ff.read_internal_field(os.path.join(tdir, "dsmcN_N2"))
    assert np.array_equal(nN2, np.bincount(ref.cell[ref.typeId == 0], minlength=500))
    assert ff.read_internal_field(os.path.join(tdir, "dsmcN_mixture")).sum() == ref.n
    for name in ("mfpToDx", "SOFP"):
        v = ff.read_internal_field(os.path.join(tdir, f"{name}_mixture"))
        assert v.shape == (500,) and np.all(v > 0)
    text = open(os.path.join(tdir, "wallHeatFlux_mixture")).read()
    assert "upperWall" in text and "nonuniform List<scalar>" in text.split("boundaryField")[1]
    assert os.path.exists(os.path.join(tdir, "dsmcSigmaTcRMax"))
    assert os.path.exists(os.path.join(tdir, "uniform", "lagrangian", "dsmc", "cloudProperties"))


def test_driver_resume_sampling_round_trip(tmp_path):
    """averagingAcrossManyRuns: the run writes uniform/resumeSampling_<fieldName> with the reference's key set
    (dsmcVolFields::writeOut, dsmcVolFields.C:745-835; key list taken from the shipped
    hypersonicCorner/.../uniform/resumeSampling_Ar) and a restarted run continues the averages (readIn, :647-743)."""
    import json

    n_steps = 3
    casegen.couette_case(str(tmp_path), n_steps=n_steps, seed=7, nto=1)
    fp = os.path.join(str(tmp_path), "system", "fieldPropertiesDict")
    text = open(fp).read().replace("resetAtOutput       on;", "resetAtOutput       off;")
    text = text.replace("measureMeanFreePath     true;", "measureMeanFreePath     true;\n            averagingAcrossManyRuns true;")
    with open(fp, "w") as fh:
        fh.write(text)
    r = subprocess.run([RUN, "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    t1 = os.path.join(str(tmp_path), "5.00003")
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "resumeSampling_keys.json")))["keys"]
    for inst in ("mixture", "N2", "O2"):
        d = open(os.path.join(t1, "uniform", f"resumeSampling_{inst}")).read()
        body = d.split("// * * *")[1] if "// * * *" in d else d
        got = [ln.split()[0] for ln in body.splitlines() if ln and ln[0].isalpha()]
        assert got == keys, (inst, got)
        assert "nTimeSteps      3;" in d
    n1 = ff.read_internal_field(os.path.join(t1, "dsmcNMean_mixture"))
    # second run: starts from the latest time, reads the accumulators back and keeps averaging
    cd = os.path.join(str(tmp_path), "system", "controlDict")
    control = open(cd).read()
    assert "5.00003;" in control
    with open(cd, "w") as fh:
        fh.write(control.replace("5.00003;", "5.00006;"))
    r2 = subprocess.run([RUN, "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stderr + r2.stdout
    assert "Resuming sampling" in r2.stdout and "nTimeSteps = 3" in r2.stdout
    t2 = os.path.join(str(tmp_path), "5.00006")
    d = open(os.path.join(t2, "uniform", "resumeSampling_mixture")).read()
    assert "nTimeSteps      6;" in d
    n2 = ff.read_internal_field(os.path.join(t2, "dsmcNMean_mixture"))
    # 47 583 parcels in a closed box: the six-step mean number per cell still sums to the parcel count
    assert abs(n1.sum() - 47583) < 1e-3 and abs(n2.sum() - 47583) < 1e-3   # files carry 10 significant digits
    assert not np.allclose(n1, n2)


def test_driver_purge_write_and_sample_interval(tmp_path):
    """controlDict purgeWrite (Time::writeObject keeps the N most recent time directories of the run) and
    dsmcVolFieldsProperties.sampleInterval (dsmcVolFields.C:1073-1081,1362) through the case directory."""
    casegen.couette_case(str(tmp_path), n_steps=6, seed=7, nto=1)
    cd = os.path.join(str(tmp_path), "system", "controlDict")
    control = open(cd).read()
    assert "writeInterval   6;" in control
    with open(cd, "w") as fh:
        fh.write(control.replace("writeInterval   6;", "writeInterval   2;\npurgeWrite      2;"))
    fp = os.path.join(str(tmp_path), "system", "fieldPropertiesDict")
    text = open(fp).read().replace("resetAtOutput       on;", "resetAtOutput       off;")
    with open(fp, "w") as fh:
        fh.write(text.replace("fieldName", "sampleInterval 3;\n            fieldName").replace(
            "measureMeanFreePath     true;", "measureMeanFreePath     true;\n            averagingAcrossManyRuns true;"))
    r = subprocess.run([RUN, "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    times = sorted(d for d in os.listdir(str(tmp_path)) if d[0].isdigit())
    assert times == ["5", "5.00004", "5.00006"], times            # 5.00002 purged; the start time was not written by this run
    d = open(os.path.join(str(tmp_path), "5.00006", "uniform", "resumeSampling_mixture")).read()
    assert "nTimeSteps      2;" in d                               # steps 3 and 6 of 6
    n = ff.read_internal_field(os.path.join(str(tmp_path), "5.00006", "dsmcNMean_mixture"))
    assert abs(n.sum() - 47583) < 1e-3


def test_driver_initialise_step_then_run(tmp_path):
    """The dsmcInitialise+ step of every shipped Allrun (blockMesh; dsmcInitialise+; dsmcFoam+): `dsmcb200_run -initialise` fills the
    mesh from system/dsmcInitialiseDict (dsmcMeshFill) and writes the start-time cloud with 15 significant digits
    (dsmcInitialise+.C:57-88); the solver then starts from it."""
    import shutil

    g, mesh, _ = casegen.couette_case(str(tmp_path), n_steps=2, seed=3, nto=1, start_time="0")
    shutil.rmtree(os.path.join(str(tmp_path), "0"))                      # a fresh case: no cloud yet
    init = """
configurations
(
    configuration
    {
        type            dsmcMeshFill;

            numberDensities
            {
                  N2         3.2e19;
                  O2         0.8e19;
            };

            translationalTemperature        2500;
            rotationalTemperature           2500;
            vibrationalTemperature          2500;
        electronicTemperature           0;

            velocity        (150 0 0);
      }
);
"""
    from hystrath_b200 import case as casew

    casew.write_dict(os.path.join(str(tmp_path), "system", "dsmcInitialiseDict"), "system", "dsmcInitialiseDict", init)
    r = subprocess.run([RUN, "-initialise", "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "Initialising dsmc for Time = 0" in r.stdout and "End" in r.stdout
    n = int(r.stdout.split("total no. of parcels:")[1].split()[0])
    expect = 4.0e19 * 500 * 1e-6 / float(g["nEquivalentParticles"])       # n V / F_N, 500 cells of 1e-6 m^3: ~46 500 parcels
    assert abs(n / expect - 1) < 0.02
    cdir = os.path.join(str(tmp_path), "0", "lagrangian", "dsmc")
    xyz, cell = ff.read_positions(os.path.join(cdir, "positions"))
    tid = ff.read_scalar_list(os.path.join(cdir, "typeId"), np.int32)
    assert len(cell) == n and abs((tid == 0).mean() - 0.8) < 0.02
    toks = [ln for ln in open(os.path.join(cdir, "positions")).read().splitlines() if ln.startswith("(") and ln[-1].isdigit()][:50]
    digits = max(len(t.strip("()").split()[0].split("e")[0].replace(".", "").replace("-", "").lstrip("0")) for t in toks)
    assert digits >= 14                                                   # 15 significant digits (defaultPrecision(15)), not 10
    ijk = np.floor(xyz / 0.01).astype(int)
    assert np.array_equal(ijk[:, 0] + 5 * ijk[:, 1], cell)
    assert os.path.exists(os.path.join(str(tmp_path), "0", "dsmcSigmaTcRMax"))
    U = ff.read_vector_list(os.path.join(cdir, "U"))
    assert abs(U[:, 0].mean() - 150.0) < 15.0
    # unknown configuration types fail like dsmcConfiguration::New
    bad = init.replace("dsmcMeshFill", "dsmcZoneFill")
    casew.write_dict(os.path.join(str(tmp_path), "system", "dsmcInitialiseDict"), "system", "dsmcInitialiseDict", bad)
    rb = subprocess.run([RUN, "-initialise", "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert rb.returncode == 1 and "unknown dsmcConfiguration type dsmcZoneFill" in rb.stderr
    # and the solver runs from the initialised state
    r2 = subprocess.run([RUN, "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stderr + r2.stdout
    assert f"Number of DSMC particles        = {n}" in r2.stdout and "End stage 0" in r2.stdout


def test_driver_runs_the_axisymmetric_tutorial(tmp_path):
    """`dsmcb200_run -initialise` then the solver on the reference's axisymmetric tutorial: the written RWF field is the shipped one, the
    start cloud holds n V / (F_N RWF) parcels per cell with their radial weights, the run clones / deletes parcels after every move and
    writes radialWeight with the cloud; rhoN of the free stream ahead of the shock comes out as the inflow density."""
    from tests import helpers as H
    gold = H.axisym_gold()
    mesh = casegen.axisym_case(str(tmp_path), n_steps=60)
    r = subprocess.run([RUN, "-initialise", "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "Axisymmetric simulation:" in r.stdout and "radial extent\t0.03" in r.stdout and "maximum radial weighting factor\t1000" in r.stdout
    rwf = ff.read_internal_field(os.path.join(str(tmp_path), "0", "RWF"))
    assert np.allclose(rwf, gold["RWF"], rtol=1e-9)
    cdir = os.path.join(str(tmp_path), "0", "lagrangian", "dsmc")
    xyz, cell = ff.read_positions(os.path.join(cdir, "positions"))
    w = ff.read_scalar_list(os.path.join(cdir, "radialWeight"))
    assert np.allclose(w, rwf[cell], rtol=1e-9)
    n0 = len(cell)
    assert abs(n0 - 130000) < 3000
    r2 = subprocess.run([RUN, "-case", str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert r2.returncode == 0, r2.stderr + r2.stdout
    end = [d for d in os.listdir(str(tmp_path)) if d.startswith("4.8e-06")][0]
    rhoN = ff.read_internal_field(os.path.join(str(tmp_path), end, "rhoN_Ar"))
    # the three columns of cells next to the inlet still see the undisturbed stream after 60 steps: rhoN = 1e21 whatever the cell's weight
    inlet = np.concatenate([np.arange(0, 800, 40), 800 + np.arange(0, 1600, 40)])
    assert abs(rhoN[inlet].mean() / 1e21 - 1) < 0.02 and np.abs(rhoN[inlet] / 1e21 - 1).max() < 0.25
    cdir = os.path.join(str(tmp_path), end, "lagrangian", "dsmc")
    xyz, cell = ff.read_positions(os.path.join(cdir, "positions"))
    w = ff.read_scalar_list(os.path.join(cdir, "radialWeight"))
    assert np.allclose(w, rwf[cell], rtol=1e-9) and len(cell) > n0       # every parcel carries its cell's weight after the weighting stage
    assert os.path.exists(os.path.join(str(tmp_path), end, "RWF"))


def test_driver_fields_with_different_reset_policies(tmp_path):
    """timeProperties per field (dsmcField.C:113-152): O2 keeps averaging (resetAtOutput off) while N2 and mixture reset at every write.
    Two writes of two steps each: the second O2 field averages all four steps, the second N2 field only the last two -- both from the
    one set of accumulators the engine keeps."""
    casegen.couette_case(str(tmp_path), n_steps=4, seed=7, nto=1)
    fp = os.path.join(str(tmp_path), "system", "fieldPropertiesDict")
    text = open(fp).read()
    with open(fp, "w") as fh:
        # the first field{} is O2; the case starts at t = 5, beyond the tutorial's resetAtOutputUntilTime
        fh.write(text.replace("resetAtOutput       on;", "resetAtOutput       off;", 1).replace("resetAtOutputUntilTime       0.5;", "resetAtOutputUntilTime       100;"))
    cd = os.path.join(str(tmp_path), "system", "controlDict")
    control = open(cd).read()
    with open(cd, "w") as fh:
        fh.write(control.replace("writeInterval   4;", "writeInterval   2;"))
    r = subprocess.run([RUN, "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    t1, t2 = os.path.join(str(tmp_path), "5.00002"), os.path.join(str(tmp_path), "5.00004")
    o2_a, o2_b = (ff.read_internal_field(os.path.join(t, "dsmcNMean_O2")) for t in (t1, t2))
    n2_a, n2_b = (ff.read_internal_field(os.path.join(t, "dsmcNMean_N2")) for t in (t1, t2))
    mix_b = ff.read_internal_field(os.path.join(t2, "dsmcNMean_mixture"))
    g = np.load(casegen.GOLD)
    n_o2, n_n2 = int((g["typeId"] == 1).sum()), int((g["typeId"] == 0).sum())
    # closed box: every mean sums to the species' parcel count, whatever the averaging window
    for a, n in ((o2_a, n_o2), (o2_b, n_o2), (n2_a, n_n2), (n2_b, n_n2), (mix_b, n_o2 + n_n2)):
        assert abs(a.sum() - n) < 1e-3 * 50
    # the instantaneous per-cell counts of steps 3-4 differ from those of steps 1-2 ...
    assert not np.allclose(n2_a, n2_b)
    # ... the O2 field of the second write is the mean over all four steps: 2 x (four-step mean) - (mean of steps 1-2) = mean of steps
    # 3-4 is non-negative and a multiple of 1/2; the N2 field (two-step window) is itself a multiple of 1/2
    late = 2.0 * o2_b - o2_a
    assert late.min() > -1e-6 and np.abs(late * 2 - np.round(late * 2)).max() < 1e-5
    assert np.abs(n2_b * 2 - np.round(n2_b * 2)).max() < 1e-5
    assert np.abs(o2_b * 4 - np.round(o2_b * 4)).max() < 1e-5 and np.abs(o2_b * 2 - np.round(o2_b * 2)).max() > 0.2   # quarter steps occur
    assert np.allclose(mix_b - n2_b, late, atol=1e-5)      # mixture (two-step window) minus N2 = O2 over the same two steps


def test_driver_variable_time_step_on_a_uniform_mesh_equals_the_constant_one(tmp_path):
    """`timeStepModel variable` (dsmcVariableTimeStepModel.C:48-100) scales nParticles and deltaT of a cell with V / V_min.  On the couette
    mesh every cell has the same volume (to a few ulp), so the per-cell fields equal the uniform values to 1e-14 and the run -- through the
    kernel instances that read the cell fields -- reproduces the constant-time-step run: the same parcels in the same cells in the same
    order, fields equal to the printed precision (V_min stands in for V in rhoN: 1e-14).  The model also writes nParticles and deltaT."""
    import shutil

    a, b = os.path.join(str(tmp_path), "const"), os.path.join(str(tmp_path), "var")
    os.makedirs(a)
    casegen.couette_case(a, n_steps=3, seed=11, nto=1)
    shutil.copytree(a, b)
    dp = os.path.join(b, "constant", "dsmcProperties")
    txt = open(dp).read()
    with open(dp, "w") as fh:
        fh.write(txt.replace("collisionPartnerSelectionModel", "timeStepModel variable;\ncollisionPartnerSelectionModel", 1))
    for d in (a, b):
        r = subprocess.run([RUN, "-case", d], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr + r.stdout
    assert "Variable time-step model:" in r.stdout
    ta, tb = os.path.join(a, "5.00003"), os.path.join(b, "5.00003")
    for f in ("rhoN_mixture", "Ttra_mixture", "Trot_N2", "dsmcNMean_O2"):
        x, y = ff.read_internal_field(os.path.join(ta, f)), ff.read_internal_field(os.path.join(tb, f))
        assert np.allclose(x, y, rtol=1e-9, atol=0), (f, np.abs(x / np.where(y != 0, y, 1) - 1).max(), int((x != y).sum()))
    (xa, ca), (xb, cb) = (ff.read_positions(os.path.join(t, "lagrangian", "dsmc", "positions")) for t in (ta, tb))
    assert np.array_equal(ca, cb) and np.allclose(xa, xb, rtol=0, atol=2e-10)      # the same parcels in the same cells, in the same order
    n = ff.read_internal_field(os.path.join(tb, "nParticles"))
    dt = ff.read_internal_field(os.path.join(tb, "deltaT"))
    g = np.load(casegen.GOLD)
    assert np.allclose(n, float(g["nEquivalentParticles"]), rtol=1e-9) and np.allclose(dt, 1e-5, rtol=1e-9)
    assert not os.path.exists(os.path.join(ta, "nParticles"))


def test_driver_binary_case_writes_what_the_ascii_case_writes(tmp_path):
    """`writeFormat binary;` (Time::writeFormat_; BASIC/particle/particleIO.C:121-143, BASIC/IOPosition/IOPosition.C:65-83): the couette case
    with binary polyMesh and cloud files runs to the same cloud and volume fields as its ASCII twin written at 17 digits, bit for bit
    (wall fields to the rounding of their atomic sums), and the driver starts again from the binary time directory it wrote."""
    a_dir, b_dir = os.path.join(str(tmp_path), "ascii"), os.path.join(str(tmp_path), "binary")
    for d in (a_dir, b_dir):
        os.makedirs(d)
        casegen.couette_case(d, n_steps=4, seed=5, nto=2)
    for d, old, new in ((a_dir, "writePrecision  10;", "writePrecision  17;"), (b_dir, "writeFormat     ascii;", "writeFormat     binary;")):
        cd = os.path.join(d, "system", "controlDict")
        text = open(cd).read()
        assert old in text
        open(cd, "w").write(text.replace(old, new))
    ff.convert_case_to_binary(b_dir, "5")
    for d in (a_dir, b_dir):
        r = subprocess.run([RUN, "-case", d], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr + r.stdout
    ta, tb = os.path.join(a_dir, "5.00004"), os.path.join(b_dir, "5.00004")
    ca, cb = os.path.join(ta, "lagrangian", "dsmc"), os.path.join(tb, "lagrangian", "dsmc")
    assert b"format      binary;" in open(os.path.join(cb, "positions"), "rb").read(1200)
    assert b"format      binary;" in open(os.path.join(tb, "rhoN_mixture"), "rb").read(1200)
    pa, pb = ff.read_positions(os.path.join(ca, "positions")), ff.read_positions(os.path.join(cb, "positions"))
    assert len(pa[1]) > 40000 and np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1])
    for name in ("U",):
        assert np.array_equal(ff.read_vector_list(os.path.join(ca, name)), ff.read_vector_list(os.path.join(cb, name))), name
    assert np.array_equal(ff.read_scalar_list(os.path.join(ca, "ERot")), ff.read_scalar_list(os.path.join(cb, "ERot")))
    for name in ("typeId", "origId", "classification", "newParcel"):
        assert np.array_equal(ff.read_scalar_list(os.path.join(ca, name), np.int32), ff.read_scalar_list(os.path.join(cb, name), np.int32)), name
    assert np.array_equal(ff.read_label_list_list(os.path.join(ca, "vibLevel")), ff.read_label_list_list(os.path.join(cb, "vibLevel")))
    for name in ("rhoN_mixture", "Ttra_mixture", "U_mixture", "p_N2", "dsmcSigmaTcRMax"):
        fa, fb = ff.read_internal_field(os.path.join(ta, name)), ff.read_internal_field(os.path.join(tb, name))
        assert fa.shape == fb.shape and np.array_equal(fa, fb), name
    for name in ("wallHeatFlux_mixture", "wallShearStress_mixture", "fD_mixture"):
        fa, fb = ff.read_patch_field(os.path.join(ta, name), "upperWall"), ff.read_patch_field(os.path.join(tb, name), "upperWall")
        # wall faces sum their hits with FP64 atomics: the order, and with it the last bits, differ from run to run
        assert fa.shape == fb.shape and np.abs(fa).max() > 0 and np.allclose(fa, fb, rtol=1e-11, atol=0), name
    # latestTime is now the binary directory: the driver reads back what it wrote
    r = subprocess.run([RUN, "-case", b_dir, "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "startTime 5.00004" in r.stdout and f"parcels {len(pb[1])}" in r.stdout, r.stdout + r.stderr
    ra = subprocess.run([RUN, "-case", a_dir, "-dryRun"], capture_output=True, text=True, timeout=120)
    assert [l for l in r.stdout.splitlines() if "checksum" in l] == [l for l in ra.stdout.splitlines() if "checksum" in l]


def test_driver_renumber_cells_keeps_the_case_labels(tmp_path):
    """`dsmcb200_run -renumberCells` (dsmcb200_set_cell_order: renumberMesh in memory): the engine works on relabelled cells, the files keep
    the case's labels -- every written parcel lies in the cell its `positions` entry names, and the sampled fields are those of the
    plain run up to the collisions' different random streams (keyed by the engine's labels)."""
    a_dir, b_dir = os.path.join(str(tmp_path), "plain"), os.path.join(str(tmp_path), "renumbered")
    out = {}
    for d, extra in ((a_dir, []), (b_dir, ["-renumberCells"])):
        os.makedirs(d)
        g, mesh, p = casegen.couette_case(d, n_steps=4, seed=5, nto=2)
        r = subprocess.run([RUN, "-case", d] + extra, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr + r.stdout
        t = os.path.join(d, "5.00004")
        xyz, cell = ff.read_positions(os.path.join(t, "lagrangian", "dsmc", "positions"))
        out[d] = dict(xyz=xyz, cell=cell, ids=ff.read_scalar_list(os.path.join(t, "lagrangian", "dsmc", "origId"), np.int32),
                      rhoN=ff.read_internal_field(os.path.join(t, "rhoN_mixture")), sig=ff.read_internal_field(os.path.join(t, "dsmcSigmaTcRMax")))
    a, b = out[a_dir], out[b_dir]
    assert len(a["cell"]) == len(b["cell"]) == 47583
    o = Oracle()
    o.set_mesh(mesh)
    cc, *_ = o.geometry()
    box = np.abs(cc[1:] - cc[:-1]).max(axis=0)                   # spacing of the structured mesh per direction
    for run in (a, b):
        d = np.abs(run["xyz"] - cc[run["cell"]])
        assert np.all(d <= 0.5 * box * (1 + 1e-9) + 1e-15), "a parcel is not in the cell its label names"
    # the plain run is cell-major in the case's labels, the renumbered one in the engine's: a different order of the same kind of cloud
    assert np.all(np.diff(a["cell"]) >= 0) and not np.all(np.diff(b["cell"]) >= 0)
    assert sorted(a["ids"].tolist()) == sorted(b["ids"].tolist())
    # ~95 parcels per cell: the two runs are independent samples of the same flow after the first collisions
    assert np.abs(a["rhoN"] - b["rhoN"]).mean() < 0.1 * a["rhoN"].mean() and abs(a["rhoN"].sum() / b["rhoN"].sum() - 1) < 1e-3
    assert np.allclose(a["sig"], b["sig"], rtol=0.5)


def test_driver_fields_with_different_sample_intervals(tmp_path):
    """sampleInterval is per field{} (dsmcField.C:113-152, dsmcVolFields.C:1073-1081): N2 samples every second step, the other two fields
    every step.  Each written field equals the reduction of an oracle run made with that field's interval (the cloud does not know
    about sampling, so the runs share their trajectory)."""
    n_steps = 4
    g, mesh, p = casegen.couette_case(str(tmp_path), n_steps=n_steps, seed=5, nto=2)
    path = os.path.join(str(tmp_path), "system", "fieldPropertiesDict")
    text = open(path).read()
    assert "fieldName               N2;" in text
    open(path, "w").write(text.replace("fieldName               N2;", "fieldName               N2;\n            sampleInterval 2;"))
    r = subprocess.run([RUN, "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    tdir = os.path.join(str(tmp_path), "5.00004")
    sp = H.air5()[:2]
    pm = [dict(patch=mesh.patch_index("upperWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=3000.0, velocity=(300.0, 0, 0)),
          dict(patch=mesh.patch_index("lowerWall"), boundaryModel="dsmcDiffuseWallPatch", temperature=2000.0, velocity=(0, 0, 0))]
    spd = [dict(mass=s.mass, diameter=s.diameter, omega=s.omega, rotDof=2.0, thetaV=[s.thetaV[0]]) for s in sp]
    for inst, idl, interval in (("mixture", [0, 1], 1), ("N2", [0], 2), ("O2", [1], 1)):
        md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=float(g["nEquivalentParticles"]), deltaT=1e-5, seed=5,
                               patch_models=pm, inverseZvFormulation="pre-2008", rotationalRelaxationCollisionNumber=5.0,
                               measureHeatFluxShearStress=True, sampleInterval=interval)
        o = Oracle()
        o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
        o.upload_parcels(p)
        o.upload_cellstate(g["dsmcSigmaTcRMax"], None)
        o.set_step(500000)
        o.evolve(n_steps)
        acc, coll, nt = o.accumulators()
        assert nt == n_steps // interval
        _, cv, *_ = o.geometry()
        f = fields_ref.derive(acc, coll, nt, spd, idl, float(g["nEquivalentParticles"]), cv, deltaT=1e-5)
        for name in ("rhoN", "rhoM", "Ttra", "p", "Trot", "Tvib"):
            got = ff.read_internal_field(os.path.join(tdir, f"{name}_{inst}"))
            assert np.allclose(got, f[name], rtol=2e-9, atol=1e-300), (name, inst)
    # the two N2 samples are not the four of the other fields
    assert not np.allclose(ff.read_internal_field(os.path.join(tdir, "rhoN_N2")) + ff.read_internal_field(os.path.join(tdir, "rhoN_O2")),
                           ff.read_internal_field(os.path.join(tdir, "rhoN_mixture")), rtol=1e-6)


CLL_BOUNDARIES = """
dsmcPatchBoundaries
(
    boundary
    {
        patchBoundaryProperties { patchName upperWall; }
        boundaryModel   dsmcCLLWallPatch;
        dsmcCLLWallPatchProperties
        {
            normalAccommodationCoefficient              0.8;
            tangentialAccommodationCoefficient          0.5;
            rotationalEnergyAccommodationCoefficient    0.9;
            vibrationalEnergyAccommodationCoefficient   1.0;
            temperature     3000.0;
            velocity        (300.0 0.0 0.0);
        }
    }
    boundary
    {
        patchBoundaryProperties { patchName lowerWall; }
        boundaryModel   dsmcCLLWallPatch;
        dsmcCLLWallPatchProperties
        {
            normalAccommodationCoefficient              1.0;
            tangentialAccommodationCoefficient          1.0;
            rotationalEnergyAccommodationCoefficient    1.0;
            vibrationalEnergyAccommodationCoefficient   1.0;
            temperature     2000.0;
            velocity        (0.0 0.0 0.0);
        }
    }
);
dsmcCyclicBoundaries ( );
dsmcGeneralBoundaries ( );
"""


def test_driver_reads_cll_wall_patches(tmp_path):
    """system/boundariesDict with boundaryModel dsmcCLLWallPatch (dsmcCLLWallPatch.C:45-75,330-334): the driver's cloud after four steps is the
    oracle's with the same coefficients; a dictionary without one of the mandatory keywords stops with the reference's lookup message."""
    from hystrath_b200 import case as casew

    n_steps = 4
    g, mesh, p = casegen.couette_case(str(tmp_path), n_steps=n_steps, seed=5, nto=2)
    casew.write_dict(os.path.join(str(tmp_path), "system", "boundariesDict"), "system", "boundariesDict", CLL_BOUNDARIES)
    r = subprocess.run([RUN, "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr + r.stdout
    sp = H.air5()[:2]
    pm = [dict(patch=mesh.patch_index("upperWall"), boundaryModel="dsmcCLLWallPatch", temperature=3000.0, velocity=(300.0, 0, 0),
               normalAccommodationCoefficient=0.8, tangentialAccommodationCoefficient=0.5, rotationalEnergyAccommodationCoefficient=0.9),
          dict(patch=mesh.patch_index("lowerWall"), boundaryModel="dsmcCLLWallPatch", temperature=2000.0, velocity=(0, 0, 0),
               normalAccommodationCoefficient=1.0, tangentialAccommodationCoefficient=1.0, rotationalEnergyAccommodationCoefficient=1.0)]
    md = capi.build_models("LarsenBorgnakkeVariableHardSphere", nEquivalentParticles=float(g["nEquivalentParticles"]), deltaT=1e-5,
                           seed=5, patch_models=pm, inverseZvFormulation="pre-2008", rotationalRelaxationCollisionNumber=5.0,
                           measureHeatFluxShearStress=True)
    o = Oracle()
    o.set_mesh(mesh); o.set_species(sp); o.set_models(md)
    o.upload_parcels(p)
    o.upload_cellstate(g["dsmcSigmaTcRMax"], None)
    o.set_step(500000)
    o.evolve(n_steps)
    ref = o.download_parcels()
    cdir = os.path.join(str(tmp_path), "5.00004", "lagrangian", "dsmc")
    xyz, cell = ff.read_positions(os.path.join(cdir, "positions"))
    assert np.array_equal(ff.read_scalar_list(os.path.join(cdir, "origId"), np.int32), ref.origId)
    assert np.array_equal(cell, ref.cell) and np.allclose(xyz, ref.position, rtol=0, atol=5e-10)
    assert np.allclose(ff.read_vector_list(os.path.join(cdir, "U")), ref.U, rtol=1e-9, atol=1e-6)
    assert np.allclose(ff.read_scalar_list(os.path.join(cdir, "ERot"), np.float64), ref.ERot, rtol=1e-8, atol=1e-30)

    casew.write_dict(os.path.join(str(tmp_path), "system", "boundariesDict"), "system", "boundariesDict",
                     CLL_BOUNDARIES.replace("vibrationalEnergyAccommodationCoefficient   1.0;", "", 1))
    r = subprocess.run([RUN, "-case", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "vibrationalEnergyAccommodationCoefficient" in r.stderr + r.stdout
