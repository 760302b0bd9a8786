"""Writes dsmcFoam+ case directories in the reference's on-disk layout (constant/polyMesh, constant/dsmcProperties,
system/{controlDict,boundariesDict,fieldPropertiesDict}, <time>/lagrangian/dsmc/*, <time>/dsmcSigmaTcRMax) so that
synthetic loads and fixtures can be fed to the standalone driver exactly like a user's case."""
from __future__ import annotations

import os

import numpy as np

from . import foamfile as ff

_TYPE = {0: "wall", 1: "patch", 2: "cyclic", 3: "processor", 4: "empty", 5: "symmetryPlane", 6: "symmetry", 7: "wedge", 8: "processorCyclic"}


def write_poly_mesh(case_dir, mesh):
    pm = os.path.join(case_dir, "constant", "polyMesh")
    os.makedirs(pm, exist_ok=True)
    loc = "constant/polyMesh"
    ff.write_vector_list(os.path.join(pm, "points"), "vectorField", loc, "points", mesh.points, fmt="%.17g")
    ff.write_faces(os.path.join(pm, "faces"), loc, mesh.face_offsets, mesh.face_points)
    with open(os.path.join(pm, "neighbour"), "w") as f:
        f.write(ff.header("labelList", loc, "neighbour"))
        f.write(f"{len(mesh.neighbour)}\n(\n" + "\n".join(str(int(v)) for v in mesh.neighbour) + "\n)\n")
    with open(os.path.join(pm, "owner"), "w") as f:
        f.write(ff.header("labelList", loc, "owner"))
        f.write(f"{len(mesh.owner)}\n(\n" + "\n".join(str(int(v)) for v in mesh.owner) + "\n)\n")
    with open(os.path.join(pm, "boundary"), "w") as f:
        f.write(ff.header("polyBoundaryMesh", loc, "boundary"))
        f.write(f"{len(mesh.patches)}\n(\n")
        for p in mesh.patches:
            t = p["type"] if isinstance(p["type"], str) else _TYPE[p["type"]]
            f.write(f"    {p['name']}\n    {{\n        type            {t};\n        nFaces          {p['size']};\n        startFace       {p['start']};\n")
            if t == "cyclic":
                f.write(f"        neighbourPatch  {mesh.patches[p['neighbPatch']]['name']};\n")
            if t in ("processor", "processorCyclic"):
                f.write(f"        myProcNo        {p['myProcNo']};\n        neighbProcNo    {p['neighbProcNo']};\n")
            if t == "processorCyclic":
                s = p["separation"]
                f.write(f"        separationVector ({s[0]:.17g} {s[1]:.17g} {s[2]:.17g});\n")
            f.write("    }\n")
        f.write(")\n")


def write_cloud(case_dir, time_name, parcels, sigma_tcr_max, mesh, cloud="dsmc"):
    d = os.path.join(case_dir, time_name, "lagrangian", cloud)
    os.makedirs(d, exist_ok=True)
    loc = f"{time_name}/lagrangian/{cloud}"
    n = parcels.n
    ff.write_positions(os.path.join(d, "positions"), loc, parcels.position[:n], parcels.cell[:n], fmt="%.17g")
    ff.write_vector_list(os.path.join(d, "U"), "vectorField", loc, "U", parcels.U[:n], fmt="%.17g")
    ff.write_scalar_list(os.path.join(d, "typeId"), "labelField", loc, "typeId", parcels.typeId[:n], fmt="%d")
    if parcels.ERot is not None:
        ff.write_scalar_list(os.path.join(d, "ERot"), "scalarField", loc, "ERot", parcels.ERot[:n], fmt="%.17g")
    if parcels.vibLevel is not None:
        ff.write_label_list_list(os.path.join(d, "vibLevel"), "labelFieldField", loc, "vibLevel", np.asarray(parcels.vibLevel[:n]).reshape(n, -1))
    if parcels.origId is not None:
        ff.write_scalar_list(os.path.join(d, "origId"), "labelField", loc, "origId", parcels.origId[:n], fmt="%d")
    ff.write_scalar_list(os.path.join(d, "newParcel"), "labelField", loc, "newParcel", np.full(n, -1), fmt="%d")
    ff.write_scalar_list(os.path.join(d, "classification"), "labelField", loc, "classification", np.zeros(n, int), fmt="%d")
    sig = np.broadcast_to(np.asarray(sigma_tcr_max, float), (mesh.n_cells,))
    with open(os.path.join(case_dir, time_name, "dsmcSigmaTcRMax"), "w") as f:
        f.write(ff.header("volScalarField", time_name, "dsmcSigmaTcRMax"))
        f.write("dimensions      [0 3 -1 0 0 0 0];\n\ninternalField   nonuniform List<scalar> \n")
        f.write(f"{mesh.n_cells}\n(\n" + "\n".join("%.17g" % v for v in sig) + "\n)\n;\n\nboundaryField\n{\n")
        for p in mesh.patches:
            t = p["type"] if isinstance(p["type"], str) else _TYPE[p["type"]]
            f.write(f"    {p['name']}\n    {{\n        type            {t if t in ('cyclic', 'empty', 'processor', 'processorCyclic', 'symmetryPlane') else 'zeroGradient'};\n    }}\n")
        f.write("}\n")


def write_dict(path, location, obj, body):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(ff.header("dictionary", location, obj))
        f.write(body)
