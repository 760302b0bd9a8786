"""CPU tests of the standalone driver's case reader (C++: foam_io + dsmcCloud) via `dsmcb200_run -dryRun`, and of
the Python foamfile reader/writer round trip, on a case directory in the reference's layout."""
import os
import subprocess

import numpy as np
import pytest

from hystrath_b200 import foamfile as ff
from tests import casegen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "hystrath_b200", "dsmcb200_run")


def test_driver_parses_couette_case(tmp_path):
    g, mesh, p = casegen.couette_case(str(tmp_path))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "1212 points 2105 faces 895 internal 500 cells 5 patches" in out
    assert "patch periodicX_half0 cyclic 100 @905" in out and "patch frontAndBack empty 1000 @1105" in out
    assert "species: N2 O2" in out
    assert "collisionModel 2 invZv 0" in out and "seed 5" in out
    assert "patchModels 2 inflows 0 fields 3" in out
    assert "field mixture typeIds 0 1 mfp 1" in out
    assert "parcels 47583" in out
    assert "startTime 5 " in out


def test_driver_rejects_unknown_models_like_the_reference(tmp_path):
    casegen.couette_case(str(tmp_path))
    path = os.path.join(str(tmp_path), "constant", "dsmcProperties")
    text = open(path).read().replace("BinaryCollisionModel            LarsenBorgnakkeVariableHardSphere;", "BinaryCollisionModel            VariableSofterSphere;")
    open(path, "w").write(text)
    r = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1
    assert "FOAM FATAL ERROR" in r.stderr and "unknown BinaryCollisionModel type VariableSofterSphere" in r.stderr
    assert "Valid BinaryCollisionModel types are" in r.stderr
    # a missing keyword fails with OpenFOAM's message shape
    open(path, "w").write(text.replace("VariableSofterSphere", "VariableHardSphere").replace("nEquivalentParticles", "nEquivParticles"))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "keyword nEquivalentParticles is undefined in dictionary" in r.stderr


def test_driver_accepts_fields_with_different_reset_policies(tmp_path):
    """timeProperties are per field (dsmcField.C:113-152): one field may keep averaging while the others reset at every write (the
    GPU driver test checks the sums), and so is sampleInterval (fields with another interval get their own set of sums)."""
    casegen.couette_case(str(tmp_path))
    path = os.path.join(str(tmp_path), "system", "fieldPropertiesDict")
    text = open(path).read()
    assert text.count("resetAtOutput       on;") == 3
    open(path, "w").write(text.replace("resetAtOutput       on;", "resetAtOutput       off;", 1))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "field O2 typeIds 1 mfp 1 reset 0" in r.stdout and "field N2 typeIds 0 mfp 1 reset 1" in r.stdout
    open(path, "w").write(text.replace("fieldName               N2;", "fieldName               N2;\n            sampleInterval 2;"))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr       # its own sample set (dsmcb200_set_sample_sets)
    assert "field N2 typeIds 0 mfp 1 reset 1 sampleInterval 2" in r.stdout and "field O2 typeIds 1 mfp 1 reset 1 sampleInterval 1" in r.stdout


def test_driver_reads_linear_wall_temperature(tmp_path):
    """groundLevelTemperature / formationLevelTemperature / depthAxis of dsmcDiffuseWallPatchProperties (dsmcDiffuseWallPatch.C:49-64,169-179)."""
    casegen.couette_case(str(tmp_path))
    path = os.path.join(str(tmp_path), "system", "boundariesDict")
    text = open(path).read()
    assert text.count("temperature") >= 2
    open(path, "w").write(text.replace("temperature", "formationLevelTemperature 250; depthAxis x; groundLevelTemperature", 1))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "linearTemperature formationLevel 250 depthAxis 0" in r.stdout


def test_driver_reads_sample_interval(tmp_path):
    """dsmcVolFieldsProperties.sampleInterval (dsmcVolFields.C:1038) reaches the engine; fields that disagree get sample sets of their own."""
    casegen.couette_case(str(tmp_path))
    path = os.path.join(str(tmp_path), "system", "fieldPropertiesDict")
    text = open(path).read()
    assert "sampleInterval" not in text
    open(path, "w").write(text.replace("fieldName", "sampleInterval 4;\n            fieldName"))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count(" sampleInterval 4") == 3
    open(path, "w").write(text.replace("fieldName", "sampleInterval 4;\n            fieldName", 1))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count(" sampleInterval 4") == 1 and r.stdout.count(" sampleInterval 1") == 2


def test_python_reader_round_trip(tmp_path):
    g, mesh, p = casegen.couette_case(str(tmp_path))
    d = os.path.join(str(tmp_path), "5", "lagrangian", "dsmc")
    xyz, cell = ff.read_positions(os.path.join(d, "positions"))
    assert np.array_equal(xyz, g["positions"]) and np.array_equal(cell, g["cell"])       # %.17g round-trips doubles
    assert np.array_equal(ff.read_vector_list(os.path.join(d, "U")), g["U"])
    assert np.array_equal(ff.read_label_list_list(os.path.join(d, "vibLevel")), g["vibLevel"])
    assert np.array_equal(ff.read_scalar_list(os.path.join(d, "classification"), np.int32), np.zeros(47583, np.int32))  # N{0} form
    bnd = ff.read_boundary(os.path.join(str(tmp_path), "constant", "polyMesh", "boundary"))
    assert [b["name"] for b in bnd] == ["upperWall", "lowerWall", "periodicX_half0", "periodicX_half1", "frontAndBack"]
    props = ff.read_dict(os.path.join(str(tmp_path), "constant", "dsmcProperties"))
    assert props["typeIdList"] == ["N2", "O2"]
    assert props["moleculeProperties"]["O2"]["Zref"] == [17900]
    assert props["LarsenBorgnakkeVariableHardSphereCoeffs"]["inverseZvFormulation"] == "pre-2008"
    bd = ff.read_dict(os.path.join(str(tmp_path), "system", "boundariesDict"))
    assert len(bd["dsmcPatchBoundaries"]) == 2 and bd["dsmcPatchBoundaries"][0][1]["boundaryModel"] == "dsmcDiffuseWallPatch"
    assert bd["dsmcGeneralBoundaries"] == []


def test_driver_parses_dsmc_initialise_dict(tmp_path):
    """`-initialise -dryRun`: system/dsmcInitialiseDict in the shipped layout (a stray ';' after the numberDensities block, as in
    run/hyStrath/dsmcFoam+/hypersonicCorner/system/dsmcInitialiseDict) and the dsmcConfiguration::New failure for other types."""
    from hystrath_b200 import case as casew

    casegen.couette_case(str(tmp_path))
    init = """
configurations
(
    configuration
    {
        type			dsmcMeshFill;

		    numberDensities
		    {
			      N2         1.0e20;
			      O2         2.5e19;
		    };

		    translationalTemperature     	300;
		    rotationalTemperature			      290;
		    vibrationalTemperature			    280;
        electronicTemperature           0;

		    velocity        (1936 0 0);
	  }
);
"""
    path = os.path.join(str(tmp_path), "system", "dsmcInitialiseDict")
    casew.write_dict(path, "system", "dsmcInitialiseDict", init)
    r = subprocess.run([RUN, "-initialise", "-dryRun", "-case", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "configuration dsmcMeshFill: Ttra 300 Trot 290 Tvib 280 Telec 0 velocity (1936 0 0)" in r.stdout
    assert "numberDensity N2 1e+20" in r.stdout and "numberDensity O2 2.5e+19" in r.stdout
    casew.write_dict(path, "system", "dsmcInitialiseDict", init.replace("dsmcMeshFill", "dsmcZoneFill"))
    r = subprocess.run([RUN, "-initialise", "-dryRun", "-case", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "unknown dsmcConfiguration type dsmcZoneFill" in r.stderr and "(dsmcMeshFill)" in r.stderr
    casew.write_dict(path, "system", "dsmcInitialiseDict", init.replace("O2 ", "Xe "))
    r = subprocess.run([RUN, "-initialise", "-dryRun", "-case", str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Cannot find typeId: Xe" in r.stderr


ORION = "/root/reference/run/hyStrath/dsmcFoam+/orion107kmNR"


@pytest.mark.skipif(not os.path.isdir(ORION), reason="the shipped case directory only exists next to the reference checkout")
def test_driver_parses_the_shipped_orion_dictionaries(tmp_path):
    """The unchanged system/ and constant/ of the shipped 5-species-air capsule case (its snappyHexMesh polyMesh is not shipped, so a
    box with the same patch names stands in): LB N2/O2, two dsmcFreeStreamInflowPatch + dsmcDeletionPatch boundaries, a diffuse wall,
    three dsmcVolFields, dsmcMeshFill, empty chemReactDict, runTime write control with purgeWrite."""
    import shutil

    from hystrath_b200 import case as casew
    from hystrath_b200 import meshgen

    for sub in ("system", "constant"):
        shutil.copytree(os.path.join(ORION, sub), os.path.join(str(tmp_path), sub))
    os.chmod(os.path.join(str(tmp_path), "constant"), 0o755)
    sides = {"xmin": ("patch", "inlet"), "xmax": ("patch", "spline"), "ymin": ("wall", "orion"), "ymax": ("patch", "spline"),
             "zmin": ("patch", "spline"), "zmax": ("patch", "spline")}
    casew.write_poly_mesh(str(tmp_path), meshgen.box_mesh((6, 5, 5), (1.2, 1.0, 1.0), sides=sides))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-initialise", "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "deltaT 7e-06 endTime 4 writeControl runTime writeInterval 0.007 nTerminalOutputs 10 purgeWrite 2" in out
    assert "species: N2 O2" in out and "collisionModel 2 invZv 2 nEquivalentParticles 2.6708e+15" in out
    assert "patchModels 3 inflows 2 fields 3" in out
    assert "numberDensity N2 2.318e+18" in out and "velocity (6053.4 0 0)" in out


HEATBATH = "/root/reference/run/hyStrath/dsmcFoam+/heatBath-5species"


@pytest.mark.skipif(not os.path.isdir(HEATBATH), reason="the shipped case directory only exists next to the reference checkout")
def test_driver_parses_the_shipped_reacting_heat_bath_dictionaries(tmp_path):
    """The unchanged system/ and constant/ of the shipped reacting tutorial: 12 quantum-kinetic reactions in system/chemReactDict
    (8 dissociationQK, 4 dissociationExchangeQK), five species, specular walls, per-species dsmcVolFields that reset at every write."""
    import shutil

    from hystrath_b200 import case as casew
    from hystrath_b200 import meshgen

    for sub in ("system", "constant"):
        shutil.copytree(os.path.join(HEATBATH, sub), os.path.join(str(tmp_path), sub))
    os.chmod(os.path.join(str(tmp_path), "constant"), 0o755)
    os.chmod(os.path.join(str(tmp_path), "system"), 0o755)
    casew.write_poly_mesh(str(tmp_path), meshgen.box_mesh((1, 1, 1), (1e-5,) * 3, sides={s: ("wall", "fixedWalls") for s in meshgen.SIDES}))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-initialise", "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "Creating dsmcReactions" in out and out.count("Selecting the reaction model dissociationQK") == 8
    assert out.count("Selecting the reaction model dissociationExchangeQK") == 4 and "Number of reactions created: 12" in out
    assert "species: N2 O2 NO N O" in out
    # a reaction model the engine does not have stops like dsmcReaction::New
    p = os.path.join(str(tmp_path), "system", "chemReactDict")
    txt = open(p).read()
    open(p, "w").write(txt.replace("reactionModel   dissociationQK;", "reactionModel   ionisationQK;", 1))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-initialise", "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "unknown dsmc reaction model type ionisationQK" in r.stderr and "Valid reaction types are" in r.stderr
    open(p, "w").write(txt.replace("(O2 N2)", "(O2 Xe)", 1))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-initialise", "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Cannot find type id: Xe" in r.stderr


AXISYM = "/root/reference/run/hyStrath/dsmcFoam+/axisymmetricFlatnosedCylinder"


@pytest.mark.skipif(not os.path.isdir(AXISYM), reason="the shipped case directory only exists next to the reference checkout")
def test_driver_parses_the_shipped_axisymmetric_dictionaries(tmp_path):
    """The unchanged system/ and constant/ of the shipped axisymmetric tutorial (coordinateSystem dsmcAxisymmetric with
    maxRadialWeightingFactor 1000) on the mesh its blockMeshDict describes; other coordinate systems / time-step models stop with the
    selector's message."""
    import shutil

    from hystrath_b200 import case as casew
    from hystrath_b200 import meshgen

    for sub in ("system", "constant"):
        shutil.copytree(os.path.join(AXISYM, sub), os.path.join(str(tmp_path), sub))
    os.chmod(os.path.join(str(tmp_path), "constant"), 0o755)
    casew.write_poly_mesh(str(tmp_path), meshgen.axisymmetric_cylinder_mesh())
    r = subprocess.run([RUN, "-case", str(tmp_path), "-initialise", "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "4000 cells 4 patches" in out and "patch wedgeFront symmetry 4000" in out and "patch cylinder wall 60" in out
    assert "coordinateSystem dsmcAxisymmetric polarAxis 1 angularCoordinate 2 maxRadialWeightingFactor 1000 timeStepModel constant" in out
    assert "nEquivalentParticles 2e+07" in out and "patchModels 2 inflows 1 fields 1" in out
    p = os.path.join(str(tmp_path), "constant", "dsmcProperties")
    txt = open(p).read()
    os.chmod(p, 0o644)
    open(p, "w").write(txt.replace("dsmcAxisymmetric;", "dsmcCylindrical;"))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-initialise", "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "dsmcCylindrical" in r.stderr and "dsmcAxisymmetric" in r.stderr and "dsmcCartesian" in r.stderr and "dsmcSpherical" in r.stderr
    # dsmcSpherical reads sphericalProperties (dsmcSpherical.C:325-386): missing sub-dictionary -> OpenFOAM's message; with it, accepted
    open(p, "w").write(txt.replace("dsmcAxisymmetric;", "dsmcSpherical;"))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-initialise", "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "sphericalProperties" in r.stderr
    open(p, "w").write(txt.replace("dsmcAxisymmetric;", "dsmcSpherical;\nsphericalProperties { maxRadialWeightingFactor 50; origin (0 0 0); }"))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-initialise", "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "coordinateSystem dsmcSpherical" in r.stdout and "maxRadialWeightingFactor 50" in r.stdout, r.stderr
    open(p, "w").write(txt.replace("dsmcAxisymmetric;", "dsmcSpherical;\nsphericalProperties { maxRadialWeightingFactor 50; radialWeightingMethod particleAverage; }"))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-initialise", "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "particleAverage is not supported" in r.stderr
    open(p, "w").write(txt.replace("coordinateSystem   dsmcAxisymmetric;", "coordinateSystem   dsmcAxisymmetric;\ntimeStepModel adaptive;"))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-initialise", "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "dsmcAdaptiveTimeStepModel" in r.stderr and "dsmcVariableTimeStepModel" in r.stderr
    open(p, "w").write(txt.replace("coordinateSystem   dsmcAxisymmetric;", "coordinateSystem   dsmcAxisymmetric;\ntimeStepModel variable;"))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-initialise", "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "timeStepModel variable" in r.stdout


def test_driver_reads_a_binary_case_like_the_ascii_one(tmp_path):
    """`writeFormat binary;` cases (BASIC/particle/particleIO.C:121-143, BASIC/IOPosition/IOPosition.C:65-83: a contiguous list is its size and
    the raw bytes in round brackets, a particle one block of position / cellI / faceI / stepFraction, faces a faceCompactList): the mesh and
    the cloud the driver reads from the binary files are, byte for byte, those it reads from the ASCII files of the same case; the Python
    reader agrees."""
    casegen.couette_case(str(tmp_path))
    r = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    ascii_sum = [l for l in r.stdout.splitlines() if "checksum" in l]
    assert len(ascii_sum) == 1 and "parcels 47583" in r.stdout
    d = os.path.join(str(tmp_path), "5", "lagrangian", "dsmc")
    before = {"positions": ff.read_positions(os.path.join(d, "positions")), "U": ff.read_vector_list(os.path.join(d, "U")),
              "vibLevel": ff.read_label_list_list(os.path.join(d, "vibLevel")), "typeId": ff.read_scalar_list(os.path.join(d, "typeId"), np.int32),
              "faces": ff.read_faces(os.path.join(str(tmp_path), "constant", "polyMesh", "faces"))}
    ff.convert_case_to_binary(str(tmp_path), "5")
    raw = open(os.path.join(d, "positions"), "rb").read()
    assert b"format      binary;" in raw and len(raw) > 47583 * 43      # '(' 40 bytes ')' newline per particle
    r2 = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r2.returncode == 0, r2.stderr
    assert [l for l in r2.stdout.splitlines() if "checksum" in l] == ascii_sum
    assert "1212 points 2105 faces 895 internal 500 cells 5 patches" in r2.stdout and "parcels 47583" in r2.stdout
    after = {"positions": ff.read_positions(os.path.join(d, "positions")), "U": ff.read_vector_list(os.path.join(d, "U")),
             "vibLevel": ff.read_label_list_list(os.path.join(d, "vibLevel")), "typeId": ff.read_scalar_list(os.path.join(d, "typeId"), np.int32),
             "faces": ff.read_faces(os.path.join(str(tmp_path), "constant", "polyMesh", "faces"))}
    for k in before:
        a, b = before[k], after[k]
        if isinstance(a, tuple):
            assert all(np.array_equal(x, y) for x, y in zip(a, b)), k
        else:
            assert np.array_equal(a, b), k
    # the same case as a 64-bit-label build writes it (arch "LSB;label=64;scalar=64")
    ff.convert_case_to_binary(str(tmp_path), "5", label64=True)
    assert b"label=64" in open(os.path.join(d, "typeId"), "rb").read(1200)
    r64 = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r64.returncode == 0, r64.stderr
    assert [l for l in r64.stdout.splitlines() if "checksum" in l] == ascii_sum
    # a truncated block is an error, not a short cloud
    whole = open(os.path.join(d, "U"), "rb").read()
    open(os.path.join(d, "U"), "wb").write(whole[:-5000])
    r3 = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r3.returncode == 1 and "binary block is truncated" in r3.stderr


def test_python_binary_round_trip(tmp_path):
    """The Python reader / writer of `format binary;` files: every list kind of a cloud, an empty list, rows of different length, and a
    file written by a 64-bit-label build (arch "...label=64...")."""
    d = str(tmp_path)
    rng = np.random.default_rng(3)
    n = 257
    xyz, cell = rng.random((n, 3)), rng.integers(0, 1000, n).astype(np.int32)
    ff.write_positions(os.path.join(d, "positions"), "0/lagrangian/dsmc", xyz, cell, binary=True)
    x2, c2 = ff.read_positions(os.path.join(d, "positions"))
    assert np.array_equal(x2, xyz) and np.array_equal(c2, cell)
    U = rng.standard_normal((n, 3))
    ff.write_vector_list(os.path.join(d, "U"), "vectorField", "0/lagrangian/dsmc", "U", U, binary=True)
    assert np.array_equal(ff.read_vector_list(os.path.join(d, "U")), U)
    e = rng.random(n)
    ff.write_scalar_list(os.path.join(d, "ERot"), "scalarField", "0/lagrangian/dsmc", "ERot", e, binary=True)
    assert np.array_equal(ff.read_scalar_list(os.path.join(d, "ERot")), e)
    ff.write_scalar_list(os.path.join(d, "typeId"), "labelField", "0/lagrangian/dsmc", "typeId", cell, binary=True)
    assert np.array_equal(ff.read_scalar_list(os.path.join(d, "typeId"), np.int32), cell)
    ff.write_scalar_list(os.path.join(d, "none"), "labelField", "0/lagrangian/dsmc", "none", np.zeros(0, np.int32), binary=True)
    assert len(ff.read_scalar_list(os.path.join(d, "none"), np.int32)) == 0
    rows = [np.array([1, 2, 3]), np.array([], int), np.array([7])]
    ff.write_label_list_list(os.path.join(d, "vibLevel"), "labelFieldField", "0/lagrangian/dsmc", "vibLevel", np.array(rows, dtype=object), binary=True)
    v = ff.read_label_list_list(os.path.join(d, "vibLevel"))
    assert v.shape == (3, 3) and v.tolist() == [[1, 2, 3], [0, 0, 0], [7, 0, 0]]
    offs, labels = np.array([0, 4, 7, 12], np.int32), rng.integers(0, 50, 12).astype(np.int32)
    ff.write_faces(os.path.join(d, "faces"), "constant/polyMesh", offs, labels, binary=True)
    o2, l2 = ff.read_faces(os.path.join(d, "faces"))
    assert np.array_equal(o2, offs) and np.array_equal(l2, labels)
    # 64-bit labels: same layout with 8-byte integers, announced in the header's arch entry
    raw = ff.header("labelList", "constant/polyMesh", "owner", True).replace("label=32", "label=64").encode() + b"\n5\n(" + np.arange(5, dtype=np.int64).tobytes() + b")\n"
    open(os.path.join(d, "owner"), "wb").write(raw)
    assert ff.read_scalar_list(os.path.join(d, "owner"), np.int32).tolist() == [0, 1, 2, 3, 4]


def test_driver_reads_cll_wall_patch(tmp_path):
    """boundaryModel dsmcCLLWallPatch: the four accommodation coefficients, temperature and velocity are all mandatory lookups
    (dsmcCLLWallPatch.C:57-62,330-334), the vibrational one included although nothing uses it."""
    from hystrath_b200 import case as casew
    from tests.test_gpu_driver import CLL_BOUNDARIES

    casegen.couette_case(str(tmp_path))
    path = os.path.join(str(tmp_path), "system", "boundariesDict")
    casew.write_dict(path, "system", "boundariesDict", CLL_BOUNDARIES)
    r = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "patchModel upperWall dsmcCLLWallPatch temperature 3000 accommodation normal 0.8 tangential 0.5 rotational 0.9" in r.stdout
    assert "patchModel lowerWall dsmcCLLWallPatch temperature 2000 accommodation normal 1 tangential 1 rotational 1" in r.stdout
    for key in ("normalAccommodationCoefficient", "vibrationalEnergyAccommodationCoefficient", "velocity"):
        casew.write_dict(path, "system", "boundariesDict", CLL_BOUNDARIES.replace(key, key + "X", 1))
        r = subprocess.run([RUN, "-case", str(tmp_path), "-dryRun"], capture_output=True, text=True, timeout=120)
        assert r.returncode != 0 and f"keyword {key} is undefined in dictionary" in r.stderr + r.stdout
