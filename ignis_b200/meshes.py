"""Triangle-mesh producers for the scene tables handed to the device.

Host-side data producers only (not on the hot path): they restate what the
reference's mesh generators and file readers emit, so that the `shapes`
dyn-table the device receives has the reference's content.

Reference: src/runtime/mesh/TriMesh.cpp (generators :772-1129, normals :96-115,
texcoords :123-142, face normals :152-197, subdivide :352-474, plane detection
:521-635), src/runtime/mesh/PlyFile.cpp:96-377, src/runtime/mesh/ObjFile.cpp:24-200,
src/runtime/shape/TriMeshProvider.cpp:17-104.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

F = np.float32
PI = F(3.14159265358979323846)
FLT_EPS = F(1.1920928955e-07)


@dataclass
class TriMesh:
    vertices: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), F))
    normals: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), F))
    texcoords: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), F))
    indices: np.ndarray = field(default_factory=lambda: np.zeros((0, 4), np.uint32))  # i0,i1,i2,0

    @property
    def face_count(self) -> int:
        return int(self.indices.shape[0])

    # -- TriMesh.cpp:34-43
    def flip_normals(self) -> None:
        self.indices[:, [1, 2]] = self.indices[:, [2, 1]]
        self.normals = (-self.normals).astype(F)

    # -- TriMesh.cpp:96-115
    def compute_vertex_normals(self) -> None:
        n = np.zeros_like(self.vertices, dtype=F)
        v = self.vertices
        for f in range(self.face_count):
            i0, i1, i2 = (int(x) for x in self.indices[f, :3])
            N = _normalized(np.cross((v[i1] - v[i0]).astype(F), (v[i2] - v[i0]).astype(F)).astype(F))
            n[i0] += N
            n[i1] += N
            n[i2] += N
        ln = np.sqrt((n * n).sum(axis=1, dtype=F)).astype(F)
        ln[ln == 0] = 1
        self.normals = (n / ln[:, None]).astype(F)

    # -- TriMesh.cpp:123-142
    def make_texcoords_normalized(self) -> None:
        lo = self.vertices.min(axis=0)
        d = (self.vertices.max(axis=0) - lo).astype(F)
        t = (self.vertices - lo).astype(F)
        uv = np.zeros((len(self.vertices), 2), F)
        for k in range(2):
            if d[k] > FLT_EPS:
                uv[:, k] = t[:, k] / d[k]
        self.texcoords = uv

    # -- TriMesh.cpp:152-197
    def setup_face_normals_as_vertex_normals(self) -> None:
        idx = self.indices[:, :3].astype(np.int64)
        nv = self.vertices[idx.reshape(-1)].astype(F)
        tri = nv.reshape(-1, 3, 3)
        N = np.cross((tri[:, 1] - tri[:, 0]).astype(F), (tri[:, 2] - tri[:, 0]).astype(F)).astype(F)
        ln = np.sqrt((N * N).sum(axis=1, dtype=F)).astype(F)
        ln[ln == 0] = 1
        N = (N / ln[:, None]).astype(F)
        if len(self.texcoords):
            self.texcoords = self.texcoords[idx.reshape(-1)].astype(F)
        self.vertices = nv
        self.normals = np.repeat(N, 3, axis=0).astype(F)
        nf = self.face_count
        self.indices = np.stack([np.arange(nf) * 3, np.arange(nf) * 3 + 1, np.arange(nf) * 3 + 2,
                                 np.zeros(nf)], axis=1).astype(np.uint32)

    # -- TriMesh.cpp:263-273 (TransformCache: points by M, normals by inverse transpose, re-normalised)
    def transform(self, m44: np.ndarray) -> None:
        m = m44.astype(np.float64)
        if np.array_equal(m, np.eye(4)):
            return
        self.vertices = (self.vertices.astype(np.float64) @ m[:3, :3].T + m[:3, 3]).astype(F)
        nm = np.linalg.inv(m[:3, :3]).T
        n = self.normals.astype(np.float64) @ nm.T
        ln = np.linalg.norm(n, axis=1)
        ln[ln == 0] = 1
        self.normals = (n / ln[:, None]).astype(F)

    # -- TriMesh.cpp:352-474 (mask-less variant: every triangle splits in four)
    def subdivide(self) -> None:
        edges: dict[tuple[int, int], int] = {}
        tris = self.indices[:, :3].astype(np.int64)

        def key(a: int, b: int) -> tuple[int, int]:
            return (a, b) if a < b else (b, a)

        for t in tris:
            for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                edges.setdefault(key(int(a), int(b)), len(edges))
        pv = len(self.vertices)
        nv = np.zeros((len(edges), 3), F)
        for (a, b), e in edges.items():
            nv[e] = (self.vertices[a] + self.vertices[b]) / F(2)
        had_normals = len(self.normals) == pv
        had_tex = len(self.texcoords) == pv
        if had_normals:
            nn = np.zeros((len(edges), 3), F)
            for (a, b), e in edges.items():
                nn[e] = _normalized((self.normals[a] + self.normals[b]).astype(F))
            self.normals = np.concatenate([self.normals, nn]).astype(F)
        if had_tex:
            nt = np.zeros((len(edges), 2), F)
            for (a, b), e in edges.items():
                nt[e] = (self.texcoords[a] + self.texcoords[b]) / F(2)
            self.texcoords = np.concatenate([self.texcoords, nt]).astype(F)
        self.vertices = np.concatenate([self.vertices, nv]).astype(F)
        out = []
        for t in tris:
            v0, v1, v2 = int(t[0]), int(t[1]), int(t[2])
            e01 = pv + edges[key(v0, v1)]
            e12 = pv + edges[key(v1, v2)]
            e20 = pv + edges[key(v2, v0)]
            out += [(v0, e01, e20, 0), (v1, e12, e01, 0), (v2, e20, e12, 0), (e01, e12, e20, 0)]
        self.indices = np.asarray(out, np.uint32)

    def compute_area(self) -> float:
        v = self.vertices
        i = self.indices[:, :3].astype(np.int64)
        c = np.cross(v[i[:, 1]] - v[i[:, 0]], v[i[:, 2]] - v[i[:, 0]])
        return float(np.sqrt((c * c).sum(axis=1)).sum() / 2)

    # -- TriMesh.cpp:521-635: a mesh of exactly two coplanar triangles over four points is a plane
    def get_as_plane(self):
        eps = 1e-5
        if self.face_count != 2:
            return None
        v = self.vertices.astype(np.float64)
        if len(v) == 4:
            uv = [v[i] for i in range(4)]
            uid = [0, 1, 2, 3]
        elif 4 < len(v) <= 6:
            uv, uid = [], []
            for k, p in enumerate(v):
                if not any(np.linalg.norm(p - q) <= eps * min(np.linalg.norm(p), np.linalg.norm(q)) or np.allclose(p, q, atol=eps, rtol=0) for q in uv):
                    if len(uv) >= 4:
                        return None
                    uv.append(p)
                    uid.append(k)
            if len(uv) != 4:
                return None
        else:
            return None
        idx = self.indices[:, :3].astype(np.int64)

        def fnorm(t):
            c = np.cross(v[t[1]] - v[t[0]], v[t[2]] - v[t[0]])
            return c / np.linalg.norm(c)

        fn0, fn1 = fnorm(idx[0]), fnorm(idx[1])
        if not np.allclose(fn0, fn1, atol=eps, rtol=0):
            return None
        origin = uv[0]

        def angle(start):
            x = uv[(start + 0) % 3 + 1] - origin
            y = uv[(start + 1) % 3 + 1] - origin
            x /= np.linalg.norm(x)
            y /= np.linalg.norm(y)
            return abs(math.acos(max(-1.0, min(1.0, float(x @ y)))))

        a12, a23, a31 = angle(0), angle(1), angle(2)
        if a12 >= a23 and a12 >= a31:
            sel = 0
        elif a23 >= a31 and a23 >= a12:
            sel = 1
        else:
            sel = 2
        xa = uv[(sel + 0) % 3 + 1] - origin
        ya = uv[(sel + 1) % 3 + 1] - origin
        if fn0 @ np.cross(xa, ya) < 0:
            xa, ya = ya, xa
            uid[1], uid[2] = uid[2], uid[1]
        if len(self.texcoords):
            tc = [None] * 4
            tc[0] = self.texcoords[uid[0]]
            tc[(0 + sel) % 3 + 1] = self.texcoords[uid[1]]
            tc[(1 + sel) % 3 + 1] = self.texcoords[uid[2]]
            tc[(2 + sel) % 3 + 1] = self.texcoords[uid[3]]
        else:
            tc = [np.array(t, F) for t in ((0, 0), (1, 0), (0, 1), (1, 1))]
        return dict(origin=origin.astype(F), x_axis=xa.astype(F), y_axis=ya.astype(F),
                    texcoords=np.asarray(tc, F))


def _normalized(v: np.ndarray) -> np.ndarray:
    ln = F(np.sqrt(F((v * v).sum(dtype=F))))
    return (v / ln).astype(F) if ln > 0 else v.astype(F)


def _frame(N: np.ndarray):
    """Tangent::frame (math/Tangent.h:53-61,70-75): Duff et al. basis, normalised."""
    n0, n1, n2 = (F(x) for x in N)
    sign = F(math.copysign(1.0, float(n2)))
    a = F(-1.0) / (sign + n2)
    b = n0 * n1 * a
    nx = np.array([F(1) + sign * n0 * n0 * a, sign * b, -sign * n0], F)
    ny = np.array([b, sign + n1 * n1 * a, -n1], F)
    return _normalized(nx), _normalized(ny)


def _v(x) -> np.ndarray:
    return np.asarray(x, F)


def _add_triangle(m: TriMesh, origin, xa, ya) -> None:
    origin, xa, ya = _v(origin), _v(xa), _v(ya)
    N = _normalized(np.cross(xa, ya).astype(F))
    off = len(m.vertices)
    m.vertices = np.concatenate([m.vertices, np.stack([origin, origin + xa, origin + ya]).astype(F)])
    m.normals = np.concatenate([m.normals, np.stack([N, N, N])])
    m.texcoords = np.concatenate([m.texcoords, _v([[0, 0], [1, 0], [0, 1]])])
    m.indices = np.concatenate([m.indices, np.asarray([[off, off + 1, off + 2, 0]], np.uint32)])


def _add_grid(m: TriMesh, origin, xa, ya, cx: int, cy: int) -> None:
    origin, xa, ya = _v(origin), _v(xa), _v(ya)
    N = _normalized(np.cross(xa, ya).astype(F))
    off = len(m.vertices)
    vs, ns, ts = [], [], []
    for j in range(cy + 1):
        for i in range(cx + 1):
            u, v = F(i) / F(cx), F(j) / F(cy)
            vs.append(origin + xa * u + ya * v)
            ns.append(N)
            ts.append((u, v))
    ids = []
    for j in range(cy):
        for i in range(cx):
            i1 = j * (cx + 1) + i + off
            i2 = (j + 1) * (cx + 1) + i + off
            ids += [(i1, i1 + 1, i2 + 1, 0), (i1, i2 + 1, i2, 0)]
    m.vertices = np.concatenate([m.vertices, _v(vs)])
    m.normals = np.concatenate([m.normals, _v(ns)])
    m.texcoords = np.concatenate([m.texcoords, _v(ts)])
    m.indices = np.concatenate([m.indices, np.asarray(ids, np.uint32)])


def _add_disk(m: TriMesh, origin, N, Nx, Ny, radius, sections: int, fill_cap: bool, flip: bool = False) -> None:
    origin, N, Nx, Ny = _v(origin), _v(N), _v(Nx), _v(Ny)
    step = F(1.0) / F(sections)
    off = len(m.vertices)
    vs, ns, ts = [], [], []
    if fill_cap:
        vs.append(origin)
        ns.append(N)
        ts.append((0, 0))
    for i in range(sections):
        x = F(math.cos(float(F(2) * PI * step * F(i))))
        y = F(math.sin(float(F(2) * PI * step * F(i))))
        vs.append(F(radius) * Nx * x + F(radius) * Ny * y + origin)
        ns.append(N)
        ts.append((F(0.5) * (x + 1), F(0.5) * (y + 1)))
    m.vertices = np.concatenate([m.vertices, _v(vs)])
    m.normals = np.concatenate([m.normals, _v(ns)])
    m.texcoords = np.concatenate([m.texcoords, _v(ts)])
    if not fill_cap:
        return
    ids = []
    for i in range(sections):
        C = i + 1
        NC = (i + 1 if i + 1 < sections else 0) + 1
        ids.append((off, NC + off, C + off, 0) if flip else (off, C + off, NC + off, 0))
    m.indices = np.concatenate([m.indices, np.asarray(ids, np.uint32)])


def make_triangle(p0, p1, p2) -> TriMesh:
    m = TriMesh()
    _add_triangle(m, p0, _v(p1) - _v(p0), _v(p2) - _v(p0))
    return m


def make_plane(origin, xa, ya) -> TriMesh:
    m = TriMesh()
    _add_grid(m, origin, xa, ya, 1, 1)
    return m


def make_rectangle(p0, p1, p2, p3) -> TriMesh:
    p0, p1, p2, p3 = _v(p0), _v(p1), _v(p2), _v(p3)
    m = TriMesh()
    _add_triangle(m, p0, p1 - p0, p3 - p0)
    _add_triangle(m, p1, p2 - p1, p3 - p1)
    return m


def make_box(origin, xa, ya, za) -> TriMesh:
    origin, xa, ya, za = _v(origin), _v(xa), _v(ya), _v(za)
    lll, hhh = origin, origin + xa + ya + za
    m = TriMesh()
    _add_grid(m, lll, ya, xa, 1, 1)
    _add_grid(m, lll, xa, za, 1, 1)
    _add_grid(m, lll, za, ya, 1, 1)
    _add_grid(m, hhh, -xa, -ya, 1, 1)
    _add_grid(m, hhh, -za, -xa, 1, 1)
    _add_grid(m, hhh, -ya, -za, 1, 1)
    return m


def make_uv_sphere(center, radius, stacks: int, slices: int) -> TriMesh:
    stacks, slices = max(2, stacks), max(2, slices)
    center = _v(center)
    drho, dtheta = PI / F(stacks), F(2) * PI / F(slices)
    vs, ns, ts = [], [], []
    for i in range(stacks + 1):
        rho = F(i) * drho
        srho, crho = F(math.sin(float(rho))), F(math.cos(float(rho)))
        for j in range(slices):
            theta = F(j) * dtheta
            st, ct = F(-math.sin(float(theta))), F(math.cos(float(theta)))
            N = np.array([st * srho, ct * srho, crho], F)
            vs.append(N * F(radius) + center)
            ns.append(N)
            ts.append((F(0.5) * theta / PI, rho / PI))
    ids = []
    for i in range(stacks):
        c, n = i * slices, (i + 1) * slices
        for j in range(slices):
            nj = (j + 1) % slices
            ids += [(n + j, n + nj, c + nj, 0), (n + j, c + nj, c + j, 0)]
    return TriMesh(_v(vs), _v(ns), _v(ts), np.asarray(ids, np.uint32))


def make_ico_sphere(center, radius, subdivisions: int) -> TriMesh:
    G = F(1.618033989)
    verts = []
    for d in range(3):
        for s1 in (-1, 1):
            for s2 in (-1, 1):
                vec = np.zeros(3, F)
                vec[(d + 1) % 3] = G * F(s1)
                vec[(d + 2) % 3] = F(s2)
                verts.append(_normalized(vec))

    def gi(d, s1, s2):
        return d * 4 + (s1 + 1) + ((s2 + 1) >> 1)

    ids = []
    for s1 in (-1, 1):
        for s2 in (-1, 1):
            for s3 in (-1, 1):
                rev = s1 * s2 * s3 == -1
                i1, i2, i3 = gi(0, s1, s2), gi(1, s2, s3), gi(2, s3, s1)
                ids.append((i1, i3 if rev else i2, i2 if rev else i3))
    for d in range(3):
        for s1 in (-1, 1):
            for s2 in (-1, 1):
                rev = s1 * s2 == 1
                i2, i1, i3 = gi(d, s1, -1), gi(d, s1, 1), gi((d + 2) % 3, s2, s1)
                ids.append((i1, i3 if rev else i2, i2 if rev else i3))
    for _ in range(subdivisions):
        emap: dict[tuple[int, int], int] = {}
        for tri in ids:
            for j in range(3):
                a, b = tri[j], tri[(j + 1) % 3]
                if a >= b:
                    continue
                emap[(a, b)] = len(verts)
                verts.append(_normalized((verts[a] + verts[b]).astype(F)))
        new = []
        for tri in ids:
            ec = [emap[(min(tri[j], tri[(j + 1) % 3]), max(tri[j], tri[(j + 1) % 3]))] for j in range(3)]
            new.append((ec[0], ec[1], ec[2]))
            for j in range(3):
                new.append((tri[j], ec[j % 3], ec[(j + 2) % 3]))
        ids = new
    v = _v(verts)
    n = v.copy()
    theta = np.arccos(np.clip(n[:, 2], -1, 1)).astype(F)
    phi = np.arctan2(-n[:, 0], n[:, 1]).astype(F)
    phi = np.where(phi < 0, phi + F(2) * PI, phi).astype(F)
    tc = np.stack([phi / (F(2) * PI), theta / PI], axis=1).astype(F)
    m = TriMesh(v, n, tc, np.asarray([(a, b, c, 0) for a, b, c in ids], np.uint32))
    t = np.eye(4)
    t[:3, 3] = np.asarray(center, np.float64)
    t[:3, :3] *= float(radius)
    m.transform(t)
    return m


def make_disk(center, normal, radius, sections: int) -> TriMesh:
    sections = max(3, sections)
    Nx, Ny = _frame(_v(normal))
    m = TriMesh()
    _add_disk(m, center, normal, Nx, Ny, radius, sections, True)
    return m


def make_cone(base, radius, tip, sections: int, fill_cap: bool) -> TriMesh:
    sections = max(3, sections)
    H = _normalized((_v(base) - _v(tip)).astype(F))
    Nx, Ny = _frame(H)
    m = TriMesh()
    _add_disk(m, base, H, Nx, Ny, radius, sections, fill_cap)
    m.vertices = np.concatenate([m.vertices, _v([tip])])
    m.normals = np.concatenate([m.normals, H[None]])
    m.texcoords = np.concatenate([m.texcoords, _v([[0, 0]])])
    start = 1 if fill_cap else 0
    tP = len(m.vertices) - 1
    ids = []
    for i in range(sections):
        C = i + start
        NC = (i + 1 if i + 1 < sections else 0) + start
        ids.append((C, tP, NC, 0))
    m.indices = np.concatenate([m.indices, np.asarray(ids, np.uint32)])
    m.compute_vertex_normals()
    return m


def make_cylinder(base, base_radius, top, top_radius, sections: int, fill_cap: bool) -> TriMesh:
    sections = max(3, sections)
    H = _normalized((_v(base) - _v(top)).astype(F))
    Nx, Ny = _frame(H)
    m = TriMesh()
    _add_disk(m, base, H, Nx, Ny, base_radius, sections, fill_cap)
    off = len(m.vertices)
    _add_disk(m, top, H, Nx, Ny, top_radius, sections, fill_cap, True)
    start = 1 if fill_cap else 0
    ids = []
    for i in range(sections):
        C = i + start
        NC = (i + 1 if i + 1 < sections else 0) + start
        ids += [(C, C + off, NC, 0), (C + off, NC + off, NC, 0)]
    m.indices = np.concatenate([m.indices, np.asarray(ids, np.uint32)])
    m.compute_vertex_normals()
    return m


# ---------------------------------------------------------------- file readers
def _fan(ids):
    return [(ids[0], ids[k - 1], ids[k]) for k in range(2, len(ids))]


def load_ply(path: str) -> TriMesh:
    """PlyFile.cpp:263-377. Polygons with more than four corners are fan-triangulated
    (the reference ear-clips them, Triangulation.cpp; none of the shipped scenes needs it)."""
    with open(path, "rb") as fh:
        data = fh.read()
    end = data.index(b"end_header")
    end = data.index(b"\n", end) + 1
    header = data[:end].decode("ascii", "replace").splitlines()
    if header[0].strip() != "ply":
        raise ValueError(f"{path}: not a ply file")
    fmt, nvert, nface, props, in_vertex, idx_types = "", 0, 0, [], False, ("uchar", "int")
    for line in header[1:]:
        t = line.split()
        if not t or t[0] == "comment":
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            in_vertex = t[1] == "vertex"
            if t[1] == "vertex":
                nvert = int(t[2])
            elif t[1] == "face":
                nface = int(t[2])
        elif t[0] == "property":
            if t[1] == "list":
                idx_types = (t[2], t[3])
            elif in_vertex:
                props.append((t[1], t[2]))
    names = [p[1] for p in props]
    col = {k: names.index(k) for k in names}
    body = data[end:]
    faces = []
    if fmt == "ascii":
        lines = body.decode("ascii").splitlines()
        vt = np.asarray([[float(x) for x in ln.split()] for ln in lines[:nvert]], F)
        for ln in lines[nvert:nvert + nface]:
            t = [int(x) for x in ln.split()]
            faces.append(t[1:1 + t[0]])
    else:
        end_c = ">" if fmt == "binary_big_endian" else "<"
        if any(p[0] != "float" for p in props):
            raise ValueError(f"{path}: only float vertex properties are supported")
        vt = np.frombuffer(body, dtype=end_c + "f4", count=nvert * len(props)).reshape(nvert, len(props)).astype(F)
        off = nvert * len(props) * 4
        isz = {"int": 4, "uint": 4, "uchar": 1, "uint8_t": 1}[idx_types[1]]
        for _ in range(nface):
            n = body[off]
            off += 1
            faces.append([int(x) for x in np.frombuffer(body, dtype=end_c + ("u4" if isz == 4 else "u1"), count=n, offset=off)])
            off += n * isz
    m = TriMesh()
    m.vertices = vt[:, [col["x"], col["y"], col["z"]]].astype(F)
    if all(k in col for k in ("nx", "ny", "nz")):
        n = vt[:, [col["nx"], col["ny"], col["nz"]]].astype(F)
        ln = np.sqrt((n * n).sum(axis=1, dtype=F)).astype(F)
        ln[ln == 0] = 1
        m.normals = (n / ln[:, None]).astype(F)
    ukey = "u" if "u" in col else ("s" if "s" in col else None)
    vkey = "v" if "v" in col else ("t" if "t" in col else None)
    if ukey and vkey:
        m.texcoords = vt[:, [col[ukey], col[vkey]]].astype(F)
    ids = []
    for f in faces:
        if len(f) == 3:
            ids.append((f[0], f[1], f[2], 0))
        elif len(f) == 4:
            ids += [(f[0], f[1], f[2], 0), (f[0], f[2], f[3], 0)]
        elif len(f) > 4:
            ids += [(a, b, c, 0) for a, b, c in _fan(f)]
    m.indices = np.asarray(ids, np.uint32)
    if not len(m.normals):
        m.compute_vertex_normals()
    else:  # fixNormals, TriMesh.cpp:17-32
        l2 = (m.normals * m.normals).sum(axis=1)
        bad = ~(l2 > FLT_EPS)
        m.normals[bad] = (0, 1, 0)
    if not len(m.texcoords):
        m.make_texcoords_normalized()
    return m


def load_obj(path: str, shape_index: int | None = None) -> TriMesh:
    """ObjFile.cpp:24-200 (via tinyobjloader's triangulating reader): unique (v,vn,vt) triples become
    vertices in first-use order; polygons are fan-triangulated."""
    V, N, T = [], [], []
    shapes: list[list[tuple[tuple[int, int, int], ...]]] = [[]]
    with open(path, "r", errors="replace") as fh:
        for line in fh:
            t = line.split()
            if not t or t[0].startswith("#"):
                continue
            if t[0] == "v":
                V.append([float(x) for x in t[1:4]])
            elif t[0] == "vn":
                N.append([float(x) for x in t[1:4]])
            elif t[0] == "vt":
                T.append([float(x) for x in t[1:3]])
            elif t[0] in ("o", "g"):
                if shapes[-1]:
                    shapes.append([])
            elif t[0] == "f":
                corners = []
                for c in t[1:]:
                    p = c.split("/")

                    def fix(s, n):
                        if s == "":
                            return -1
                        i = int(s)
                        return i - 1 if i > 0 else n + i

                    vi = fix(p[0], len(V))
                    ti = fix(p[1], len(T)) if len(p) > 1 else -1
                    ni = fix(p[2], len(N)) if len(p) > 2 else -1
                    corners.append((vi, ni, ti))
                for a, b, c in _fan(corners):
                    shapes[-1].append((a, b, c))
    sel = [shapes[shape_index]] if shape_index is not None else shapes
    has_n = bool(N) and any(c[1] >= 0 for s in sel for tri in s for c in tri)
    has_t = bool(T) and any(c[2] >= 0 for s in sel for tri in s for c in tri)
    imap: dict[tuple[int, int, int], int] = {}
    vs, ns, ts, ids = [], [], [], []
    for s in sel:
        for tri in s:
            row = []
            for c in tri:
                if c not in imap:
                    imap[c] = len(imap)
                    vs.append(V[c[0]])
                    if has_n:
                        ns.append(N[c[1]] if c[1] >= 0 else (0, 0, 1))
                    if has_t:
                        ts.append(T[c[2]] if c[2] >= 0 else (0, 0))
                row.append(imap[c])
            ids.append((*row, 0))
    m = TriMesh(_v(vs), _v(ns) if has_n else np.zeros((0, 3), F), _v(ts) if has_t else np.zeros((0, 2), F),
                np.asarray(ids, np.uint32))
    if not has_n:
        m.compute_vertex_normals()
    if not has_t:
        m.make_texcoords_normalized()
    return m


def load_external(path: str, shape_index: int | None = None) -> TriMesh:
    ext = os.path.splitext(path)[1].lower()
    if ext == ".ply":
        return load_ply(path)
    if ext == ".obj":
        return load_obj(path, shape_index)
    raise ValueError(f"unsupported external mesh type '{ext}' ({path})")
