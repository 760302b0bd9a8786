"""Structured hex polyMesh generator (a blockMesh equivalent for box-shaped blocks) and the
brick decomposition used by the weak-scaling cases.

The reference meshes its cases with OpenFOAM's blockMesh and partitions them with decomposePar
(run/hyStrath/dsmcFoam+/*/Allrun); neither tool exists outside an OpenFOAM install, so synthetic
loads are meshed here with the same conventions: points x-fastest, cells x-fastest, internal faces
in upper-triangular order (owner ascending, then neighbour ascending), face normals pointing from
owner to neighbour / out of the domain, boundary faces grouped per patch, processor patches after
all other patches with plain `processor` before `processorCyclic`.  Coupled faces (cyclic halves,
processor pairs) have matched first vertices and opposite circulation, which is what
particle::hitCyclicPatch / correctAfterParallelTransfer rely on (tetPtI -> nPts - 1 - tetPtI,
BASIC/particle/particleTemplates.C:90-111,1539).
"""
from __future__ import annotations

import numpy as np

from .capi import MeshData

SIDES = ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")


def _pid(i, j, k, nx, ny):
    return i + (nx + 1) * (j + (ny + 1) * k)


def _side_faces(side, nx, ny, nz):
    """(faces[n,4] point labels, owner[n]) of one side, outward normals, coupled-compatible ordering."""
    if side in ("xmin", "xmax"):
        k, j = np.meshgrid(np.arange(nz), np.arange(ny), indexing="ij")
        j, k = j.ravel(), k.ravel()
        i = np.zeros_like(j) if side == "xmin" else np.full_like(j, nx)
        a, b, c, d = _pid(i, j, k, nx, ny), _pid(i, j + 1, k, nx, ny), _pid(i, j + 1, k + 1, nx, ny), _pid(i, j, k + 1, nx, ny)
        faces = np.stack([a, d, c, b], 1) if side == "xmin" else np.stack([a, b, c, d], 1)
        own = (0 if side == "xmin" else nx - 1) + nx * (j + ny * k)
    elif side in ("ymin", "ymax"):
        k, i = np.meshgrid(np.arange(nz), np.arange(nx), indexing="ij")
        i, k = i.ravel(), k.ravel()
        j = np.zeros_like(i) if side == "ymin" else np.full_like(i, ny)
        a, b, c, d = _pid(i, j, k, nx, ny), _pid(i, j, k + 1, nx, ny), _pid(i + 1, j, k + 1, nx, ny), _pid(i + 1, j, k, nx, ny)
        faces = np.stack([a, d, c, b], 1) if side == "ymin" else np.stack([a, b, c, d], 1)
        own = i + nx * ((0 if side == "ymin" else ny - 1) + ny * k)
    else:
        j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
        i, j = i.ravel(), j.ravel()
        k = np.zeros_like(i) if side == "zmin" else np.full_like(i, nz)
        a, b, c, d = _pid(i, j, k, nx, ny), _pid(i + 1, j, k, nx, ny), _pid(i + 1, j + 1, k, nx, ny), _pid(i, j + 1, k, nx, ny)
        faces = np.stack([a, d, c, b], 1) if side == "zmin" else np.stack([a, b, c, d], 1)
        own = i + nx * (j + ny * (0 if side == "zmin" else nz - 1))
    return faces.astype(np.int32), own.astype(np.int32)


def box_mesh(n, lengths, origin=(0.0, 0.0, 0.0), sides=None, my_proc=-1):
    """Hex mesh of a box.

    sides: dict side -> spec, spec one of
      ("cyclic",)                        paired with the opposite side
      ("wall", name) / ("patch", name) / ("empty", name) / ("symmetryPlane", name) / ("symmetry", name)
      ("processor", neighbProcNo)
      ("processorCyclic", neighbProcNo, separation_xyz)
    Sides sharing (type, name) are merged into one patch (e.g. frontAndBack).
    """
    nx, ny, nz = (int(v) for v in n)
    lx, ly, lz = (float(v) for v in lengths)
    sides = dict(sides or {s: ("cyclic",) for s in SIDES})
    xs = origin[0] + lx * np.arange(nx + 1) / nx
    ys = origin[1] + ly * np.arange(ny + 1) / ny
    zs = origin[2] + lz * np.arange(nz + 1) / nz
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    points = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)

    # internal faces: per cell (+x, +y, +z), upper-triangular order
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    cell = (i + nx * (j + ny * k)).astype(np.int32)
    fx = np.stack([_pid(i + 1, j, k, nx, ny), _pid(i + 1, j + 1, k, nx, ny), _pid(i + 1, j + 1, k + 1, nx, ny), _pid(i + 1, j, k + 1, nx, ny)], 1)
    fy = np.stack([_pid(i, j + 1, k, nx, ny), _pid(i, j + 1, k + 1, nx, ny), _pid(i + 1, j + 1, k + 1, nx, ny), _pid(i + 1, j + 1, k, nx, ny)], 1)
    fz = np.stack([_pid(i, j, k + 1, nx, ny), _pid(i + 1, j, k + 1, nx, ny), _pid(i + 1, j + 1, k + 1, nx, ny), _pid(i, j + 1, k + 1, nx, ny)], 1)
    allf = np.stack([fx, fy, fz], 1)  # [nCells, 3, 4]
    mask = np.stack([i < nx - 1, j < ny - 1, k < nz - 1], 1)
    nbr = np.stack([cell + 1, cell + nx, cell + nx * ny], 1)
    own_int = np.repeat(cell[:, None], 3, 1)[mask]
    nbr_int = nbr[mask].astype(np.int32)
    faces_int = allf[mask].astype(np.int32)

    # group sides into patches
    groups = []  # (type, name, [sides], extra)
    for s in SIDES:
        spec = sides[s]
        t = spec[0]
        if t == "cyclic":
            groups.append((t, {"xmin": "cyclicX_half0", "xmax": "cyclicX_half1", "ymin": "cyclicY_half0", "ymax": "cyclicY_half1",
                               "zmin": "cyclicZ_half0", "zmax": "cyclicZ_half1"}[s], [s], None))
        elif t in ("processor", "processorCyclic"):
            groups.append((t, f"proc{t[9:]}{my_proc}to{spec[1]}_{s}", [s], spec))
        else:
            name = spec[1] if len(spec) > 1 else s
            for g in groups:
                if g[0] == t and g[1] == name:
                    g[2].append(s)
                    break
            else:
                groups.append((t, name, [s], None))
    order = [g for g in groups if g[0] not in ("processor", "processorCyclic")]
    order += [g for g in groups if g[0] == "processor"] + [g for g in groups if g[0] == "processorCyclic"]

    bfaces, bown, patches = [], [], []
    start = len(own_int)
    for t, name, ss, spec in order:
        size = 0
        for s in ss:
            f, o = _side_faces(s, nx, ny, nz)
            bfaces.append(f)
            bown.append(o)
            size += len(o)
        p = {"name": name, "type": t, "start": start, "size": size, "sides": ss}
        if t in ("processor", "processorCyclic"):
            p["myProcNo"] = my_proc
            p["neighbProcNo"] = spec[1]
            p["separation"] = tuple(spec[2]) if t == "processorCyclic" else (0.0, 0.0, 0.0)
        patches.append(p)
        start += size
    opposite = {"xmin": "xmax", "xmax": "xmin", "ymin": "ymax", "ymax": "ymin", "zmin": "zmax", "zmax": "zmin"}
    for idx, p in enumerate(patches):
        if p["type"] == "cyclic":
            opp = opposite[p["sides"][0]]
            if sides[opp][0] != "cyclic":
                raise ValueError(f"cyclic side {p['sides'][0]} needs a cyclic opposite side")
            p["neighbPatch"] = next(q for q, pp in enumerate(patches) if pp["type"] == "cyclic" and pp["sides"][0] == opp)

    faces = np.concatenate([faces_int] + bfaces) if bfaces else faces_int
    owner = np.concatenate([own_int] + bown).astype(np.int32) if bown else own_int
    face_offsets = (4 * np.arange(len(owner) + 1)).astype(np.int32)
    mesh = MeshData(points, face_offsets, faces.ravel(), owner, nbr_int, patches)
    mesh.shape = (nx, ny, nz)
    mesh.lengths = (lx, ly, lz)
    mesh.origin = tuple(float(v) for v in origin)
    return mesh


def decomposed_box(n_local, lengths_local, procs, rank, outer=("cyclic", "cyclic", "cyclic")):
    """Brick `rank` of a procs=(px,py,pz) tiling of identical bricks (weak-scaling layout).

    outer[d] is the treatment of the global domain boundary in direction d: "cyclic" (periodic:
    a same-rank cyclic pair when p==1, processorCyclic otherwise), a (type, name) spec for both ends, or a pair of
    (type, name) specs ((lo), (hi)).
    """
    px, py, pz = procs
    rx, ry, rz = rank % px, (rank // px) % py, rank // (px * py)
    coords, P = (rx, ry, rz), (px, py, pz)
    origin = tuple(coords[d] * lengths_local[d] for d in range(3))
    total = tuple(P[d] * lengths_local[d] for d in range(3))

    def rank_of(c):
        return c[0] + px * (c[1] + py * c[2])

    sides = {}
    for d, (lo, hi) in enumerate((("xmin", "xmax"), ("ymin", "ymax"), ("zmin", "zmax"))):
        for side, step in ((lo, -1), (hi, +1)):
            c = list(coords)
            c[d] += step
            if 0 <= c[d] < P[d]:
                sides[side] = ("processor", rank_of(c))
            elif outer[d] == "cyclic":
                if P[d] == 1:
                    sides[side] = ("cyclic",)
                else:
                    c[d] %= P[d]
                    sep = [0.0, 0.0, 0.0]
                    # separation of the receiving patch = C_send - C_recv
                    sep[d] = -total[d] if step > 0 else total[d]
                    sides[side] = ("processorCyclic", rank_of(c), tuple(sep))
            elif isinstance(outer[d][0], (tuple, list)):
                sides[side] = tuple(outer[d][0 if step < 0 else 1])
            else:
                sides[side] = tuple(outer[d])
    return box_mesh(n_local, lengths_local, origin, sides, my_proc=rank)


def wedge_mesh(n, length, height, depth, angle_deg, procs=1, rank=0):
    """Hypersonic wedge (BASELINE configs[2]): a sharp wedge of half-angle `angle_deg` whose surface is the lower boundary of a
    structured n = (nx, ny, nz) hexahedral mesh, a few cells deep in z.  The grid lines x = const stay vertical; the rows are
    stretched between the ramp y = x tan(angle) and the top y = height, so every face is planar and the cells are general
    (non axis-aligned) hexahedra.  Patches: `wedge` (wall, lower boundary), `flow` (patch: inlet x = 0, top, outlet x = length --
    free-stream inflow + deletion), `sides` (symmetryPlane, z = 0 and z = depth).
    procs > 1: brick `rank` of a decomposition into `procs` slabs along x (the `simple` method of decomposePar, equal cells per
    slab -- the shock layer makes the parcel load per slab unequal), with processor patches between the slabs."""
    nx, ny, nz = (int(v) for v in n)
    if nx % procs:
        raise ValueError("nx must be a multiple of the number of slabs")
    outer = ((("patch", "flow"), ("patch", "flow")), (("wall", "wedge"), ("patch", "flow")), ("symmetryPlane", "sides"))
    mesh = decomposed_box((nx // procs, ny, nz), (length / procs, height, depth), (procs, 1, 1), rank, outer=outer)
    t = np.tan(np.radians(angle_deg))
    x, y = mesh.points[:, 0], mesh.points[:, 1]
    yb = x * t
    mesh.points[:, 1] = yb + (height - yb) * (y / height)
    mesh.wedge = dict(length=length, height=height, depth=depth, tan=t)
    return mesh


def cylinder_ogrid(nr, ntheta, r_in, r_out, thickness=None, grading=1.0):
    """2-D body-fitted O-grid around a circular cylinder (one cell thick in z, `empty` front and back): the mesh
    topology of BASELINE configs[1] (Mach-10 argon flow over a cylinder, Lofthouse).  Cells are hexahedra with
    non-axis-aligned faces; the grid closes on itself in theta (no cyclic patch needed).

    Patches: `cylinder` (wall), `outer` (patch: free-stream inflow + deletion), `frontAndBack` (empty).
    grading = (last radial cell size) / (first radial cell size), geometric.
    """
    nr, nt = int(nr), int(ntheta)
    if grading == 1.0:
        r = r_in + (r_out - r_in) * np.arange(nr + 1) / nr
    else:
        q = grading ** (1.0 / max(nr - 1, 1))
        w = np.concatenate([[0.0], np.cumsum(q ** np.arange(nr))])
        r = r_in + (r_out - r_in) * w / w[-1]
    th = 2.0 * np.pi * np.arange(nt) / nt
    t = thickness if thickness is not None else (r[1] - r[0])
    # points p(i, j, k) = i + (nr+1) * (j + nt * k)
    R, TH = np.meshgrid(r, th, indexing="xy")           # [nt, nr+1]
    xy = np.stack([(R * np.cos(TH)).ravel(), (R * np.sin(TH)).ravel()], 1)
    points = np.concatenate([np.column_stack([xy, np.zeros(len(xy))]), np.column_stack([xy, np.full(len(xy), t)])])

    def pid(i, j, k):
        return i + (nr + 1) * ((j % nt) + nt * k)

    def cid(i, j):
        return i + nr * (j % nt)

    J, I = np.meshgrid(np.arange(nt), np.arange(nr), indexing="ij")
    I, J = I.ravel(), J.ravel()
    faces, own, nei = [], [], []
    # radial internal faces at radius index i = 1..nr-1 between (i-1, j) and (i, j): normal +r
    m = I >= 1
    i, j = I[m], J[m]
    faces.append(np.stack([pid(i, j, 0), pid(i, j + 1, 0), pid(i, j + 1, 1), pid(i, j, 1)], 1))
    own.append(cid(i - 1, j)); nei.append(cid(i, j))
    # angular internal faces between (i, j) and (i, j+1), j+1 < nt: normal +theta
    m = J < nt - 1
    i, j = I[m], J[m]
    faces.append(np.stack([pid(i, j + 1, 0), pid(i, j + 1, 1), pid(i + 1, j + 1, 1), pid(i + 1, j + 1, 0)], 1))
    own.append(cid(i, j)); nei.append(cid(i, j + 1))
    # the closing faces at theta index 0: owner (i, 0) (lower label), neighbour (i, nt-1): normal -theta
    i = np.arange(nr)
    z = np.zeros_like(i)
    faces.append(np.stack([pid(i, z, 0), pid(i + 1, z, 0), pid(i + 1, z, 1), pid(i, z, 1)], 1))
    own.append(cid(i, z)); nei.append(cid(i, z + nt - 1))
    n_int = sum(len(o) for o in own)
    j = np.arange(nt)
    z = np.zeros_like(j)
    # cylinder wall (i = 0), outward = -r
    f_wall = np.stack([pid(z, j, 0), pid(z, j, 1), pid(z, j + 1, 1), pid(z, j + 1, 0)], 1)
    o_wall = cid(z, j)
    # outer boundary (i = nr), outward = +r
    f_out = np.stack([pid(z + nr, j, 0), pid(z + nr, j + 1, 0), pid(z + nr, j + 1, 1), pid(z + nr, j, 1)], 1)
    o_out = cid(z + nr - 1, j)
    # back (k = 0, -z) and front (k = 1, +z)
    f_back = np.stack([pid(I, J, 0), pid(I, J + 1, 0), pid(I + 1, J + 1, 0), pid(I + 1, J, 0)], 1)
    f_front = np.stack([pid(I, J, 1), pid(I + 1, J, 1), pid(I + 1, J + 1, 1), pid(I, J + 1, 1)], 1)
    o_fb = cid(I, J)
    patches = [dict(name="cylinder", type="wall", start=n_int, size=nt),
               dict(name="outer", type="patch", start=n_int + nt, size=nt),
               dict(name="frontAndBack", type="empty", start=n_int + 2 * nt, size=2 * nr * nt)]
    allf = np.concatenate(faces + [f_wall, f_out, f_back, f_front]).astype(np.int32)
    owner = np.concatenate(own + [o_wall, o_out, o_fb, o_fb]).astype(np.int32)
    neighbour = np.concatenate(nei).astype(np.int32)
    mesh = MeshData(points, (4 * np.arange(len(owner) + 1)).astype(np.int32), allf.ravel(), owner, neighbour, patches)
    mesh.shape = (nr, nt, 1)
    mesh.r = r
    mesh.thickness = t
    return mesh


def _morton3(i, j, k, bits=10):
    """Interleave the low `bits` bits of three non-negative integer arrays (z-order curve key)."""
    key = np.zeros(len(i), dtype=np.int64)
    i, j, k = i.astype(np.int64), j.astype(np.int64), k.astype(np.int64)
    for b in range(bits):
        key |= ((i >> b) & 1) << (3 * b)
        key |= ((j >> b) & 1) << (3 * b + 1)
        key |= ((k >> b) & 1) << (3 * b + 2)
    return key


def morton_order(mesh, bits=10):
    """new_of_old cell labels following a z-order curve through the cell centres (approximated by the mean of the
    cell's face-point coordinates), quantised to 2^bits per direction.  Works for any polyMesh."""
    nF = mesh.n_faces
    nPts = np.diff(mesh.face_offsets)
    fc = np.add.reduceat(mesh.points[mesh.face_points], mesh.face_offsets[:-1], axis=0) / nPts[:, None]
    cc = np.zeros((mesh.n_cells, 3))
    cnt = np.zeros(mesh.n_cells)
    np.add.at(cc, mesh.owner, fc)
    np.add.at(cnt, mesh.owner, 1.0)
    np.add.at(cc, mesh.neighbour, fc[: mesh.n_internal])
    np.add.at(cnt, mesh.neighbour, 1.0)
    cc /= cnt[:, None]
    lo, hi = cc.min(0), cc.max(0)
    span = np.where(hi > lo, hi - lo, 1.0)
    # one common cell-size scale so that the curve's bricks are cubes, not slabs
    scale = ((1 << bits) - 1) / span.max()
    q = np.minimum(((cc - lo) * scale).astype(np.int64), (1 << bits) - 1)
    key = _morton3(q[:, 0], q[:, 1], q[:, 2], bits)
    order = np.argsort(key, kind="stable")          # order[new] = old
    new_of_old = np.empty(mesh.n_cells, dtype=np.int32)
    new_of_old[order] = np.arange(mesh.n_cells, dtype=np.int32)
    return new_of_old


def renumber_cells(mesh, new_of_old):
    """The polyMesh with cell `c` relabelled `new_of_old[c]` -- what OpenFOAM's renumberMesh does to the mesh files: owner / neighbour
    relabelled, an internal face whose owner would exceed its neighbour is flipped (same first vertex, opposite circulation), internal
    faces re-sorted into upper-triangular order; boundary faces keep their order (patch starts and coupled-face matching unchanged).
    Returns (mesh, new_of_old)."""
    new_of_old = np.asarray(new_of_old, dtype=np.int32)
    nI = mesh.n_internal
    own = new_of_old[mesh.owner]
    nei = new_of_old[mesh.neighbour]
    if not np.all(np.diff(mesh.face_offsets) == np.diff(mesh.face_offsets)[0]):
        raise NotImplementedError("renumber_cells: faces of mixed size")
    npf = int(mesh.face_offsets[1] - mesh.face_offsets[0])
    faces = mesh.face_points.reshape(-1, npf).copy()
    flip = own[:nI] > nei
    o_int = np.where(flip, nei, own[:nI])
    n_int = np.where(flip, own[:nI], nei)
    fi = faces[:nI]
    fi[flip] = np.concatenate([fi[flip][:, :1], fi[flip][:, :0:-1]], axis=1)
    perm = np.lexsort((n_int, o_int))
    faces[:nI] = fi[perm]
    owner = np.concatenate([o_int[perm], own[nI:]]).astype(np.int32)
    out = MeshData(mesh.points, mesh.face_offsets, faces.ravel(), owner, n_int[perm].astype(np.int32), [dict(p) for p in mesh.patches])
    for a in ("shape", "lengths", "origin", "r", "thickness"):
        if hasattr(mesh, a):
            setattr(out, a, getattr(mesh, a))
    out.cell_numbering = "renumbered"
    return out, new_of_old


def relabel_cells(mesh, new_of_old):
    """The polyMesh with cell `c` called `new_of_old[c]` and nothing else changed: owner / neighbour entries only, no face is flipped or
    moved (so an owner label may exceed its neighbour's).  This is the mesh the engine works on after dsmcb200_set_cell_order."""
    new_of_old = np.asarray(new_of_old, dtype=np.int32)
    out = MeshData(mesh.points, mesh.face_offsets, mesh.face_points, new_of_old[mesh.owner].astype(np.int32),
                   new_of_old[mesh.neighbour].astype(np.int32), [dict(p) for p in mesh.patches])
    for a in ("shape", "lengths", "origin", "r", "thickness"):
        if hasattr(mesh, a):
            setattr(out, a, getattr(mesh, a))
    out.cell_numbering = "relabelled"
    return out


def poly_mesh_from_cells(points, cells, patch_of_face, patch_specs):
    """A polyMesh from an explicit cell list -- the general (non-hex) path of the tracker's tests.

    cells: list of cells, each a list of faces, each face a tuple of point labels with the normal pointing OUT of that cell.
    A face shared by two cells becomes an internal face (stored as seen from the lower-labelled cell, starting at its lowest point
    label); internal faces are put in upper-triangular order.  A face seen once is a boundary face of patch
    patch_of_face(face_points) -> name; patch_specs: ordered list of (name, type).  Faces may mix sizes (triangles, quads, ...)."""
    seen = {}
    for c, faces in enumerate(cells):
        for f in faces:
            key = tuple(sorted(f))
            seen.setdefault(key, []).append((c, tuple(f)))

    def canon(f):
        k = f.index(min(f))
        return f[k:] + f[:k]

    internal, boundary = [], {name: [] for name, _ in patch_specs}
    for key, lst in seen.items():
        if len(lst) == 2:
            (c0, f0), (c1, f1) = sorted(lst)
            internal.append((c0, c1, canon(f0)))
        elif len(lst) == 1:
            c0, f0 = lst[0]
            boundary[patch_of_face(f0)].append((c0, canon(f0)))
        else:
            raise ValueError("face shared by more than two cells")
    internal.sort(key=lambda t: (t[0], t[1]))
    faces = [f for _, _, f in internal]
    owner = [o for o, _, _ in internal]
    neighbour = [n for _, n, _ in internal]
    patches = []
    for name, typ in patch_specs:
        lst = sorted(boundary[name], key=lambda t: t[0])
        patches.append({"name": name, "type": typ, "start": len(faces), "size": len(lst)})
        faces += [f for _, f in lst]
        owner += [o for o, _ in lst]
    offsets = np.concatenate([[0], np.cumsum([len(f) for f in faces])]).astype(np.int32)
    flat = np.array([p for f in faces for p in f], dtype=np.int32)
    return MeshData(points, offsets, flat, np.array(owner, np.int32), np.array(neighbour, np.int32), patches)


def split_box_mesh(n, lengths, kind="prism", wall_type="wall"):
    """A box of n hexahedra cut into triangular prisms (2 per hex, cut along the x-y diagonal) or tetrahedra (6 per hex, Kuhn
    subdivision, conforming), one `wall`-type patch `walls` all round: meshes with triangular faces and non-hex cells.
    Returns (mesh, locate) where locate(xyz[n,3]) -> cell labels by direct point location."""
    nx, ny, nz = (int(v) for v in n)
    lx, ly, lz = (float(v) for v in lengths)
    xs, ys, zs = lx * np.arange(nx + 1) / nx, ly * np.arange(ny + 1) / ny, lz * np.arange(nz + 1) / nz
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    points = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)

    def pid(i, j, k):
        return i + (nx + 1) * (j + (ny + 1) * k)

    def outward(face, centre):
        p = points[list(face)]
        nrm = np.cross(p[1] - p[0], p[2] - p[0])
        return tuple(face) if np.dot(nrm, p.mean(0) - centre) > 0 else tuple(reversed(face))

    cells = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                v = {(a, b, c): pid(i + a, j + b, k + c) for a in (0, 1) for b in (0, 1) for c in (0, 1)}
                if kind == "prism":
                    # lower-right (y - y0 <= x - x0) and upper-left prisms, triangles in the z planes
                    tris = [((0, 0), (1, 0), (1, 1)), ((0, 0), (1, 1), (0, 1))]
                    for t in tris:
                        bot = tuple(v[(a, b, 0)] for a, b in t)
                        top = tuple(v[(a, b, 1)] for a, b in t)
                        fs = [bot, top] + [(bot[e], bot[(e + 1) % 3], top[(e + 1) % 3], top[e]) for e in range(3)]
                        ctr = points[list(bot + top)].mean(0)
                        cells.append([outward(f, ctr) for f in fs])
                elif kind == "tet":
                    # Kuhn: one tet per permutation of the axes, all sharing the diagonal (0,0,0)-(1,1,1)
                    import itertools

                    for perm in itertools.permutations(range(3)):
                        cur = [0, 0, 0]
                        verts = [tuple(cur)]
                        for ax in perm:
                            cur[ax] = 1
                            verts.append(tuple(cur))
                        t = [v[x] for x in verts]
                        fs = [(t[1], t[2], t[3]), (t[0], t[2], t[3]), (t[0], t[1], t[3]), (t[0], t[1], t[2])]
                        ctr = points[t].mean(0)
                        cells.append([outward(f, ctr) for f in fs])
                else:
                    raise ValueError(kind)
    mesh = poly_mesh_from_cells(points, cells, lambda f: "walls", [("walls", wall_type)])
    mesh.shape = (nx, ny, nz)
    mesh.lengths = (lx, ly, lz)

    def locate(xyz):
        q = np.asarray(xyz) / np.array([lx / nx, ly / ny, lz / nz])
        ijk = np.minimum(np.floor(q).astype(int), [nx - 1, ny - 1, nz - 1])
        r = q - ijk
        hexid = ijk[:, 0] + nx * (ijk[:, 1] + ny * ijk[:, 2])
        if kind == "prism":
            return 2 * hexid + (r[:, 1] > r[:, 0]).astype(int)
        import itertools

        out = np.zeros(len(q), dtype=int)
        for t, perm in enumerate(itertools.permutations(range(3))):
            # the tet of `perm` holds the points with r[perm[0]] >= r[perm[1]] >= r[perm[2]]
            m = (r[:, perm[0]] >= r[:, perm[1]]) & (r[:, perm[1]] >= r[:, perm[2]])
            out[m] = t
        return 6 * hexid + out

    return mesh, locate


def refined_interface_mesh(nx_coarse=2, h=0.004, wall_type="wall"):
    """nx_coarse unit hexahedra in a row followed by one hexahedron refined 2 x 2 in y, z (an OpenFOAM 2:1 refinement interface):
    the coarse cell next to the interface has nine faces -- four small quads towards the fine cells, one quad, and four FIVE-point
    faces whose extra point is the hanging node in the middle of a straight edge, i.e. polygons with a degenerate fan triangle, the
    case polyMeshTetDecomposition's base-point search exists for.  One patch `walls`.  Returns (mesh, locate)."""
    pts, index = [], {}

    def P(x, y, z):
        key = (round(2 * x), round(2 * y), round(2 * z))
        if key not in index:
            index[key] = len(pts)
            pts.append((x * h, y * h, z * h))
        return index[key]

    cells = []
    X = nx_coarse
    for i in range(X):
        x0, x1 = float(i), float(i + 1)
        last = i == X - 1
        if not last:
            fs = [
                (P(x0, 0, 0), P(x0, 0, 1), P(x0, 1, 1), P(x0, 1, 0)),      # x-
                (P(x1, 0, 0), P(x1, 1, 0), P(x1, 1, 1), P(x1, 0, 1)),      # x+
                (P(x0, 0, 0), P(x1, 0, 0), P(x1, 0, 1), P(x0, 0, 1)),      # y-
                (P(x0, 1, 0), P(x0, 1, 1), P(x1, 1, 1), P(x1, 1, 0)),      # y+
                (P(x0, 0, 0), P(x0, 1, 0), P(x1, 1, 0), P(x1, 0, 0)),      # z-
                (P(x0, 0, 1), P(x1, 0, 1), P(x1, 1, 1), P(x0, 1, 1)),      # z+
            ]
        else:
            fs = [(P(x0, 0, 0), P(x0, 0, 1), P(x0, 1, 1), P(x0, 1, 0))]
            for j in (0, 0.5):
                for k in (0, 0.5):
                    fs.append((P(x1, j, k), P(x1, j + .5, k), P(x1, j + .5, k + .5), P(x1, j, k + .5)))   # four small x+ faces
            fs += [
                (P(x0, 0, 0), P(x1, 0, 0), P(x1, 0, .5), P(x1, 0, 1), P(x0, 0, 1)),     # y-  (hanging node on the x1 edge)
                (P(x0, 1, 0), P(x0, 1, 1), P(x1, 1, 1), P(x1, 1, .5), P(x1, 1, 0)),     # y+
                (P(x0, 0, 0), P(x0, 1, 0), P(x1, 1, 0), P(x1, .5, 0), P(x1, 0, 0)),     # z-
                (P(x0, 0, 1), P(x1, 0, 1), P(x1, .5, 1), P(x1, 1, 1), P(x0, 1, 1)),     # z+
            ]
        cells.append(fs)
    x0, x1 = float(X), float(X + 1)
    for j in (0, 0.5):
        for k in (0, 0.5):
            j1, k1 = j + .5, k + .5
            cells.append([
                (P(x0, j, k), P(x0, j, k1), P(x0, j1, k1), P(x0, j1, k)),
                (P(x1, j, k), P(x1, j1, k), P(x1, j1, k1), P(x1, j, k1)),
                (P(x0, j, k), P(x1, j, k), P(x1, j, k1), P(x0, j, k1)),
                (P(x0, j1, k), P(x0, j1, k1), P(x1, j1, k1), P(x1, j1, k)),
                (P(x0, j, k), P(x0, j1, k), P(x1, j1, k), P(x1, j, k)),
                (P(x0, j, k1), P(x1, j, k1), P(x1, j1, k1), P(x0, j1, k1)),
            ])
    points = np.array(pts, dtype=np.float64)
    mesh = poly_mesh_from_cells(points, cells, lambda f: "walls", [("walls", wall_type)])

    def locate(xyz):
        q = np.asarray(xyz) / h
        i = np.minimum(np.floor(q[:, 0]).astype(int), X)
        fine = i >= X
        sub = 2 * (q[:, 1] >= 0.5).astype(int) + (q[:, 2] >= 0.5).astype(int)
        return np.where(fine, X + sub, i)

    return mesh, locate


def corner_mesh(h=0.01):
    """The mesh of the reference's hypersonicCorner tutorial (Bird 1994, 16.2: supersonic corner flow), restated from its
    blockMeshDict (run/hyStrath/dsmcFoam+/hypersonicCorner/system/blockMeshDict): block 1 = 5 x 18 x 18 cells over x in [0, 0.05],
    block 2 = 25 x 18 x 18 over x in [0.05, 0.30], y and z in [0, 0.18], uniform 1 cm cells; cells numbered block by block, i fastest.
    Patches in the dictionary's order: `flow` (patch: inlet, outlet, top y and top z), `entrance` (symmetry: the y = 0 and z = 0 faces
    of block 1), `walls` (wall: the y = 0 and z = 0 faces of block 2 -- the two plates forming the corner)."""
    nxs, ny, nz = (5, 25), 18, 18
    pts, index = [], {}

    def P(i, j, k):
        key = (i, j, k)
        if key not in index:
            index[key] = len(pts)
            pts.append((i * h, j * h, k * h))
        return index[key]

    cells = []
    x0 = 0
    for nx in nxs:
        for k in range(nz):
            for j in range(ny):
                for ii in range(nx):
                    i = x0 + ii
                    cells.append([
                        (P(i, j, k), P(i, j, k + 1), P(i, j + 1, k + 1), P(i, j + 1, k)),              # x-
                        (P(i + 1, j, k), P(i + 1, j + 1, k), P(i + 1, j + 1, k + 1), P(i + 1, j, k + 1)),  # x+
                        (P(i, j, k), P(i + 1, j, k), P(i + 1, j, k + 1), P(i, j, k + 1)),              # y-
                        (P(i, j + 1, k), P(i, j + 1, k + 1), P(i + 1, j + 1, k + 1), P(i + 1, j + 1, k)),  # y+
                        (P(i, j, k), P(i, j + 1, k), P(i + 1, j + 1, k), P(i + 1, j, k)),              # z-
                        (P(i, j, k + 1), P(i + 1, j, k + 1), P(i + 1, j + 1, k + 1), P(i, j + 1, k + 1)),  # z+
                    ])
        x0 += nx
    points = np.array(pts, dtype=np.float64)

    def patch_of(face):
        c = points[list(face)].mean(0)
        on_floor = c[1] < 1e-9 or c[2] < 1e-9
        if on_floor:
            return "entrance" if c[0] < nxs[0] * h else "walls"
        return "flow"

    mesh = poly_mesh_from_cells(points, cells, patch_of, [("flow", "patch"), ("entrance", "symmetry"), ("walls", "wall")])
    mesh.shape = (sum(nxs), ny, nz)
    return mesh


def _graded(x0, x1, n, ratio):
    """blockMesh simpleGrading: n cells between x0 and x1 whose sizes form a geometric progression with last/first = ratio."""
    if n == 1 or abs(ratio - 1.0) < 1e-12:
        return x0 + (x1 - x0) * np.arange(n + 1) / n
    r = ratio ** (1.0 / (n - 1))
    w = np.concatenate([[0.0], np.cumsum(r ** np.arange(n))])
    return x0 + (x1 - x0) * w / w[-1]


def flat_plate_mesh():
    """The mesh of the reference's supersonicFlatPlate tutorial, restated from its blockMeshDict
    (run/hyStrath/dsmcFoam+/supersonicFlatPlate/system/blockMeshDict): block 1 = 5 x 60 x 1 cells over x in [0, 50 mm] graded (0.5 2 1),
    block 2 = 95 x 60 x 1 over x in [50 mm, 1 m] graded (2 2 1), y in [0, 0.6 m], one 1-mm cell in z; cells numbered block by block,
    i fastest.  Patches: `plate` (wall: y = 0 of block 2), `inlet` (patch: everything else in the x-y outline), `defaultFaces`
    (front and back; the tutorial's Allrun turns it from empty into wall and gives it a specular model)."""
    xs = np.concatenate([_graded(0.0, 0.05, 5, 0.5), _graded(0.05, 1.0, 95, 2.0)[1:]])
    ys = _graded(0.0, 0.6, 60, 2.0)
    zs = np.array([0.0, 0.001])
    nx, ny = len(xs) - 1, len(ys) - 1
    pts, index = [], {}

    def P(i, j, k):
        key = (i, j, k)
        if key not in index:
            index[key] = len(pts)
            pts.append((xs[i], ys[j], zs[k]))
        return index[key]

    cells = []
    for i0, i1 in ((0, 5), (5, nx)):
        for j in range(ny):
            for i in range(i0, i1):
                k = 0
                cells.append([
                    (P(i, j, k), P(i, j, k + 1), P(i, j + 1, k + 1), P(i, j + 1, k)),
                    (P(i + 1, j, k), P(i + 1, j + 1, k), P(i + 1, j + 1, k + 1), P(i + 1, j, k + 1)),
                    (P(i, j, k), P(i + 1, j, k), P(i + 1, j, k + 1), P(i, j, k + 1)),
                    (P(i, j + 1, k), P(i, j + 1, k + 1), P(i + 1, j + 1, k + 1), P(i + 1, j + 1, k)),
                    (P(i, j, k), P(i, j + 1, k), P(i + 1, j + 1, k), P(i + 1, j, k)),
                    (P(i, j, k + 1), P(i + 1, j, k + 1), P(i + 1, j + 1, k + 1), P(i, j + 1, k + 1)),
                ])
    points = np.array(pts, dtype=np.float64)

    def patch_of(face):
        p = points[list(face)]
        if np.ptp(p[:, 2]) < 1e-12:
            return "defaultFaces"
        c = p.mean(0)
        if c[1] < 1e-12 and c[0] > 0.05:
            return "plate"
        return "inlet"

    mesh = poly_mesh_from_cells(points, cells, patch_of, [("plate", "wall"), ("inlet", "patch"), ("defaultFaces", "wall")])
    mesh.shape = (nx, ny, 1)
    mesh.xs, mesh.ys = xs, ys
    return mesh


def axisymmetric_cylinder_mesh(n_axial=40, n_inner=20, n_outer=40, half_depth=(0.000437, 0.00131)):
    """The blockMesh of the reference's axisymmetric tutorial (run/hyStrath/dsmcFoam+/axisymmetricFlatnosedCylinder/system/blockMeshDict):
    a 5-degree wedge about the x axis around a flat-nosed cylinder of radius 0.01 whose face sits at x = 0.  Three blocks -- in front of
    the face (x in [-0.02, 0], r in [0, 0.01]; its innermost row of cells are prisms on the axis), above it (r in [0.01, 0.03]) and above
    the cylinder (x in [0, 0.02], r in [0.01, 0.03]) -- one cell thick, cells numbered like blockMesh does (block by block, x fastest, then
    the radial index), so cell fields written by the reference line up.  Patches: flow (patch), cylinder (wall), wedgeFront / wedgeBack
    (symmetry), as in the dictionary."""
    zi, zo = half_depth
    V = {0: (-0.02, 0.01, zi), 1: (0.0, 0.01, zi), 2: (0.0, 0.01, -zi), 3: (-0.02, 0.01, -zi), 4: (-0.02, 0.0, 0.0), 5: (0.0, 0.0, 0.0),
         6: (-0.02, 0.03, zo), 7: (0.0, 0.03, zo), 8: (-0.02, 0.03, -zo), 9: (0.0, 0.03, -zo), 10: (0.02, 0.01, zi), 11: (0.02, 0.03, zo),
         12: (0.02, 0.01, -zi), 13: (0.02, 0.03, -zo)}
    blocks = [((4, 5, 5, 4, 0, 1, 2, 3), (n_axial, 1, n_inner)), ((0, 1, 2, 3, 6, 7, 9, 8), (n_axial, 1, n_outer)),
              ((1, 10, 12, 2, 7, 11, 13, 9), (n_axial, 1, n_outer))]
    pts, index = [], {}
    front, back = set(), set()   # points of the j = 0 / j = 1 planes of the blocks (the axis belongs to both)

    def pid(x):
        key = tuple(np.round(np.asarray(x) / 1e-12).astype(np.int64))
        if key not in index:
            index[key] = len(pts)
            pts.append(tuple(float(v) for v in x))
        return index[key]

    cells = []
    for verts, (nx, ny, nz) in blocks:
        c = np.array([V[v] for v in verts])

        def point(i, j, k):
            a, b, g = i / nx, j / ny, k / nz
            w = [(1 - a) * (1 - b) * (1 - g), a * (1 - b) * (1 - g), a * b * (1 - g), (1 - a) * b * (1 - g),
                 (1 - a) * (1 - b) * g, a * (1 - b) * g, a * b * g, (1 - a) * b * g]
            q = pid((np.array(w)[:, None] * c).sum(0))
            (front if j == 0 else back).add(q)
            return q

        for k in range(nz):
            for j in range(ny):
                for i in range(nx):
                    v = [point(i, j, k), point(i + 1, j, k), point(i + 1, j + 1, k), point(i, j + 1, k),
                         point(i, j, k + 1), point(i + 1, j, k + 1), point(i + 1, j + 1, k + 1), point(i, j + 1, k + 1)]
                    faces = []
                    for f in ((0, 4, 7, 3), (1, 2, 6, 5), (0, 1, 5, 4), (3, 7, 6, 2), (0, 3, 2, 1), (4, 5, 6, 7)):
                        lab = []
                        for q in f:   # a collapsed edge leaves a triangle, a collapsed face nothing
                            if v[q] not in lab:
                                lab.append(v[q])
                        if len(lab) >= 3:
                            faces.append(tuple(lab))
                    cells.append(faces)
    points = np.array(pts)

    def patch_of_face(f):
        p = points[list(f)]
        if all(q in front for q in f):
            return "wedgeFront"
        if all(q in back for q in f):
            return "wedgeBack"
        if np.all(np.abs(p[:, 0]) < 1e-12) and np.all(p[:, 1] <= 0.01 + 1e-12):
            return "cylinder"   # the flat face
        if np.all(np.abs(p[:, 1] - 0.01) < 1e-12) and np.all(p[:, 0] >= -1e-12):
            return "cylinder"   # the side
        return "flow"

    mesh = poly_mesh_from_cells(points, cells, patch_of_face, [("flow", "patch"), ("cylinder", "wall"), ("wedgeFront", "symmetry"), ("wedgeBack", "symmetry")])
    return mesh


def split_patch(mesh, name, keep, new_name, new_type):
    """Split boundary patch `name`: the faces for which keep(face centres [n, 3]) holds stay (in their order), the others form the new
    patch `new_name` right behind it.  Faces of one patch are contiguous, so this is a stable partition of the patch's range."""
    k = mesh.patch_index(name)
    p = mesh.patches[k]
    s, n = p["start"], p["size"]
    if not np.all(np.diff(mesh.face_offsets[s:s + n + 1]) == 4):
        raise ValueError("split_patch expects quadrilateral faces")
    o0 = mesh.face_offsets[s]
    quads = mesh.face_points[o0:o0 + 4 * n].reshape(n, 4)
    centres = mesh.points[quads].mean(1)
    m = np.asarray(keep(centres), bool)
    order = np.concatenate([np.nonzero(m)[0], np.nonzero(~m)[0]])
    mesh.face_points[o0:o0 + 4 * n] = quads[order].ravel()
    mesh.owner[s:s + n] = mesh.owner[s:s + n][order]
    n_keep = int(m.sum())
    p["size"] = n_keep
    mesh.patches.insert(k + 1, {"name": new_name, "type": new_type, "start": s + n_keep, "size": n - n_keep})
    return mesh


def order_patches(mesh, names):
    """Put the physical (non-processor) boundary patches into the order `names` = [(name, type), ...]; a patch the mesh does not have
    is listed empty.  The face ranges of the patches are moved accordingly (owner, face lists)."""
    first_proc = next((i for i, p in enumerate(mesh.patches) if p["type"].startswith("processor")), len(mesh.patches))
    phys = {p["name"]: p for p in mesh.patches[:first_proc]}
    if set(phys) - {n for n, _ in names}:
        raise ValueError("order_patches: the mesh has physical patches that are not in the list")
    b0 = mesh.n_internal
    end_phys = mesh.patches[first_proc]["start"] if first_proc < len(mesh.patches) else mesh.n_faces
    face_ids, ordered, cursor = [], [], b0
    for nm, ty in names:
        if nm in phys:
            q = dict(phys[nm])
            face_ids.append(np.arange(q["start"], q["start"] + q["size"]))
            q["start"] = cursor
            cursor += q["size"]
            ordered.append(q)
        else:
            ordered.append({"name": nm, "type": ty, "start": cursor, "size": 0})
    if cursor != end_phys:
        raise ValueError("order_patches: physical patches do not cover the boundary range")
    ids = np.concatenate(face_ids) if face_ids else np.zeros(0, np.int64)
    if len(ids):
        sizes = np.diff(mesh.face_offsets)
        lab = [mesh.face_points[mesh.face_offsets[f]:mesh.face_offsets[f + 1]] for f in ids] if not np.all(sizes[ids] == 4) else None
        if lab is None:
            o0 = mesh.face_offsets[b0]
            quads = mesh.face_points[o0:o0 + 4 * len(ids)].reshape(-1, 4)
            mesh.face_points[o0:o0 + 4 * len(ids)] = quads[ids - b0].ravel()
        else:
            o0 = mesh.face_offsets[b0]
            flat = np.concatenate(lab)
            mesh.face_points[o0:o0 + len(flat)] = flat
            mesh.face_offsets[b0:end_phys + 1] = o0 + np.concatenate([[0], np.cumsum(sizes[ids])])
        mesh.owner[b0:end_phys] = mesh.owner[ids]
    mesh.patches = ordered + mesh.patches[first_proc:]
    return mesh


def capsule_mesh(n_local, cell_size, procs=(1, 1, 1), rank=0, cap_radius_frac=0.3, sphere_to_cap=1.5):
    """Re-entry capsule forebody (BASELINE configs[3]): the heat shield of a capsule -- a spherical segment of base radius R_b and sphere
    radius sphere_to_cap * R_b -- facing a stream along +x.  The domain is a box of procs * n_local cells of size `cell_size`; its x = max
    boundary is the body: the spherical segment (wall patch `capsule`, inside R_b of the axis through the middle of the y-z section) and,
    around it, an open plane through which the gas that has gone round the shoulder leaves (patch `outflow`).  The grid lines are
    compressed along x between the inlet plane and the body surface, so the cells are general hexahedra with warped faces (tracked
    through their tet decomposition).  Inlet and the four lateral boundaries: patch `flow` (free stream in, deletion out).
    procs = (px, py, pz): brick `rank` of the decomposition into identical bricks of n_local cells (decomposePar simple), with
    processor patches between them; the patch list is the same on every rank (empty where a rank does not touch the boundary)."""
    nx, ny, nz = (int(v) for v in n_local)
    px, py, pz = procs
    lengths = (nx * cell_size, ny * cell_size, nz * cell_size)
    outer = ((("patch", "flow"), ("wall", "capsule")), ("patch", "flow"), ("patch", "flow"))
    mesh = decomposed_box((nx, ny, nz), lengths, procs, rank, outer=outer)
    Lx, Ly, Lz = px * lengths[0], py * lengths[1], pz * lengths[2]
    Rb = cap_radius_frac * min(Ly, Lz)
    Rs = sphere_to_cap * Rb
    yc, zc = 0.5 * Ly, 0.5 * Lz

    def bulge(y, z):
        r2 = (y - yc) ** 2 + (z - zc) ** 2
        return np.where(r2 < Rb * Rb, np.sqrt(np.maximum(Rs * Rs - r2, 0.0)) - np.sqrt(Rs * Rs - Rb * Rb), 0.0)

    try:
        split_patch(mesh, "capsule", lambda c: (c[:, 1] - yc) ** 2 + (c[:, 2] - zc) ** 2 < Rb * Rb, "outflow", "patch")
    except KeyError:
        pass   # this brick does not touch the body plane
    # every rank lists the same physical patches, in the same order, empty where it has no faces of them (as decomposePar writes them)
    order_patches(mesh, [("flow", "patch"), ("capsule", "wall"), ("outflow", "patch")])
    x, y, z = mesh.points[:, 0], mesh.points[:, 1], mesh.points[:, 2]
    mesh.points[:, 0] = x * (Lx - bulge(y, z)) / Lx
    mesh.capsule = dict(Rb=Rb, Rs=Rs, height=Rs - np.sqrt(Rs * Rs - Rb * Rb), L=(Lx, Ly, Lz))
    return mesh
