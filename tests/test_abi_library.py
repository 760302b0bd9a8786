"""The C-ABI library loads on a machine without a GPU, exports every symbol include/dsmcb200.h declares, and the
product path refuses to run without CUDA (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from hystrath_b200 import capi, meshgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "dsmcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsmcb200_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = header_functions()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dsmcb200.h but not exported"
    assert set(capi.EXPORTED_SYMBOLS) == set(names)
    assert lib.dsmcb200_abi_version() == 6


def test_struct_layouts_match_the_header():
    # sizes the C compiler gives the PODs (gcc x86-64); a mismatch would corrupt every call
    assert C.sizeof(capi.Patch) == 64 + 8 * 4 + 24
    assert C.sizeof(capi.Species) == 64 + 5 * 8 + 8 + 9 * 8 + 8 + 8 + 16 * 8 + 16 * 4
    assert C.sizeof(capi.PatchModel) == 8 + 8 + 24 + 8 + 8 + 8 + 24   # + diffuseFraction, linearTemperature / depthAxis, formationLevelTemperature, the three CLL coefficients
    assert C.sizeof(capi.ParcelsSoA) == 12 * 8 + 8 + 8 + 8   # + radialWeight
    assert C.sizeof(capi.Counters) == 9 * 8 + 5 * 8 + 8 * 8 + 8 + 16 * 4 + 2 * 16 * 8 + 16
    assert C.sizeof(capi.AccumInfo) == 24


def test_struct_sizes_agree_with_the_c_compiler(tmp_path):
    """sizeof() of every POD of include/dsmcb200.h as gcc lays it out against the ctypes mirror in capi.py."""
    import subprocess

    names = {"dsmcb200_patch": capi.Patch, "dsmcb200_species": capi.Species, "dsmcb200_patch_model": capi.PatchModel,
             "dsmcb200_inflow": capi.Inflow, "dsmcb200_models": capi.Models, "dsmcb200_parcels_soa": capi.ParcelsSoA,
             "dsmcb200_counters": capi.Counters, "dsmcb200_accum_info": capi.AccumInfo, "dsmcb200_mesh": capi.Mesh,
             "dsmcb200_reaction": capi.Reaction}
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "dsmcb200.h"\nint main(void) {\n' +
                   "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in names) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n, cls in names.items():
        assert int(out[n]) == C.sizeof(cls), (n, out[n], C.sizeof(cls))


def test_no_cpu_fallback_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.Dsmcb200Error, match="no CPU fallback"):
        capi.Engine(0)


def test_unknown_model_names_are_rejected_like_the_reference():
    with pytest.raises(capi.Dsmcb200Error, match="Valid BinaryCollisionModel types are"):
        capi.build_models("VariableSoftSphereTypo")
    with pytest.raises(capi.Dsmcb200Error, match="Valid patch boundary types are"):
        capi.build_models("VariableHardSphere", patch_models=[dict(patch=0, boundaryModel="dsmcStickingWallPatch")])
    with pytest.raises(ValueError):
        capi.make_species("X", 1e-26, 1e-10, 0.7, thetaV=(1000.0,), Zref=(), TrefZv=())


def test_box_mesh_conventions():
    m = meshgen.box_mesh((3, 4, 5), (0.3, 0.4, 0.5))
    assert m.n_cells == 60 and m.n_internal == 2 * 4 * 5 + 3 * 3 * 5 + 3 * 4 * 4
    # upper-triangular order, owner < neighbour
    assert np.all(m.owner[:m.n_internal] < m.neighbour)
    assert np.all(np.diff(m.owner[:m.n_internal]) >= 0)
    # coupled halves: matched first vertex (same transverse position) and opposite circulation
    for a, b, d in (("cyclicX_half0", "cyclicX_half1", 0), ("cyclicY_half0", "cyclicY_half1", 1), ("cyclicZ_half0", "cyclicZ_half1", 2)):
        pa, pb = m.patches[m.patch_index(a)], m.patches[m.patch_index(b)]
        assert pa["neighbPatch"] == m.patch_index(b) and pa["size"] == pb["size"]
        fa = m.face_points.reshape(-1, 4)[pa["start"]:pa["start"] + pa["size"]]
        fb = m.face_points.reshape(-1, 4)[pb["start"]:pb["start"] + pb["size"]]
        xa, xb = m.points[fa], m.points[fb]
        other = [k for k in range(3) if k != d]
        assert np.allclose(xa[:, 0][:, other], xb[:, 0][:, other])
        assert np.allclose(xa[:, 1][:, other], xb[:, 3][:, other]) and np.allclose(xa[:, 3][:, other], xb[:, 1][:, other])


def test_decomposed_box_patch_ordering():
    m = meshgen.decomposed_box((4, 4, 4), (0.04, 0.04, 0.04), (2, 2, 1), rank=1)
    types = [p["type"] for p in m.patches]
    assert types == ["cyclic", "cyclic", "processor", "processor", "processorCyclic", "processorCyclic"]
    nb = [p.get("neighbProcNo") for p in m.patches]
    assert nb[2:] == [0, 3, 0, 3]
    sep = [p.get("separation") for p in m.patches if p["type"] == "processorCyclic"]
    assert sep == [(-0.08, 0.0, 0.0), (0.0, 0.08, 0.0)]


def test_the_openfoam_shim_only_uses_declared_entry_points_and_fields():
    """shim/dsmcCloudB200.C cannot be compiled here (no OpenFOAM): at least every ABI function it calls is declared in the header and
    exported by the library, and every struct member it touches exists in the ctypes mirror."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "shim", "dsmcCloudB200.C")).read() + open(os.path.join(root, "shim", "dsmcCloudB200.H")).read()
    header = open(os.path.join(root, "include", "dsmcb200.h")).read()
    lib = capi.load_library()
    called = set(re.findall(r"\b(dsmcb200_[a-z_0-9]+)\s*\(", src))
    assert len(called) >= 15
    for f in called:
        assert re.search(r"\b%s\s*\(" % f, header), f
        assert hasattr(lib, f), f
    for var, cls in (("models_", capi.Models), ("soa", capi.ParcelsSoA), ("pm", capi.PatchModel), ("in", capi.Inflow), ("m", capi.Mesh),
                     ("out", capi.Patch), ("ai", capi.AccumInfo)):
        members = {n for n, _ in cls._fields_}
        for mem in set(re.findall(r"\b%s\.([A-Za-z_0-9]+)\b" % var, src)):
            assert mem in members, (var, mem)
