"""Asset loaders for the JSON scene config: Wavefront OBJ / MTL meshes and PNG / JPEG / HDR
textures, with the conventions of the reference's import path:

* meshes: what the reference gets from assimp (src/config.cpp:196-228: Triangulate | GenNormals or GenSmoothNormals |
  JoinIdenticalVertices | GenUVCoords | FindDegenerates, then CalcTangentSpace) and pushes through
  Scene::LoadAiNode / LoadAiMesh (src/scene.cpp:96-207): one mesh per (object or group, material) run in file order,
  positions transformed by the object's 4x4, normals / tangents by its 3x3 (not renormalised), texture coordinates as
  stored (no V flip), faces with fewer than 3 distinct corners dropped.  assimp itself is not available here, so this is a
  restatement of its documented OBJ importer and post-processing steps, not a bit-pinned copy: vertex order inside a
  mesh, the diagonal chosen for concave quads and the smoothing of tangents across split vertices may differ.  Triangle
  ORDER (which the kd-tree's tie-breaking sees) follows the file.
* materials: Material::LoadFromAiMaterial (src/bxdf/bxdf.cpp:88-184): every imported material becomes
  BxDFLTCDiffuse<GGX>(diffuse = Kd or map_Kd, color = Ks or map_Ks, roughness = sqrt(2 / (2 + Ns))), emission = Ke, bump
  map = map_bump / bump; registered without overriding an existing material unless "override-materials" is set.
* textures: FileTexture::CreateNewFromPNG / JPEG / HDR (src/texture.cpp:189-321): 8-bit values / 255 then gamma-decoded
  (pow 2.2); JPEG rows flipped (y -> h-1-y), PNG and HDR not; single-channel JPEGs replicated; HDR linear floats.

    pack, cfg = scene.load_json_config(path, mesh_loader=assets.load_obj_into, texture_loader=assets.load_image)
"""
import os

import numpy as np

from . import abi

F = np.float32


# ---------------------------------------------------------------------------------------------- textures
def load_image(path):
    """Returns float32 [h, w, 3] in the reference's texture convention (see module docstring)."""
    ext = os.path.splitext(path)[1].lower().lstrip(".")
    if not os.path.exists(path):
        raise FileNotFoundError(f"Failed to load texture '{path}', file does not exist.")
    if ext == "hdr":
        os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
        import cv2
        img = cv2.imread(path, cv2.IMREAD_UNCHANGED | cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)
        if img is None or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError(f"Failed to load texture '{path}', it does not contain exactly 3 color components.")
        return np.ascontiguousarray(img[..., ::-1].astype(F))
    if ext not in ("png", "jpg", "jpeg"):
        raise ValueError(f"ERROR: Texture format '{ext}' is not supported!")
    from PIL import Image
    with Image.open(path) as im:
        if ext == "png":
            a = np.asarray(im.convert("RGB"), dtype=np.uint8)
        else:
            a = np.asarray(im if im.mode in ("L", "RGB") else im.convert("RGB"), dtype=np.uint8)
            if a.ndim == 2:
                a = np.repeat(a[..., None], 3, axis=2)
            a = a[::-1]                                      # SetPixel(x, h-y-1, ...)
    c = a.astype(F) / F(255.0)
    return np.ascontiguousarray(np.power(c, F(2.2), dtype=F))    # Color::gammaDecode


# ---------------------------------------------------------------------------------------------- MTL
def parse_mtl(path):
    """name -> dict(Kd, Ks, Ke, Ns, map_Kd, map_Ks, map_bump), assimp's defaults where a key is absent."""
    mats, cur = {}, None
    if not os.path.exists(path):
        return mats
    with open(path, errors="replace") as f:
        for line in f:
            t = line.split("#", 1)[0].split()
            if not t:
                continue
            k = t[0]
            if k == "newmtl":
                cur = dict(Kd=(0.6, 0.6, 0.6), Ks=(0.0, 0.0, 0.0), Ke=(0.0, 0.0, 0.0), Ns=0.0, map_Kd="", map_Ks="", map_bump="")
                mats[" ".join(t[1:])] = cur
            elif cur is None:
                continue
            elif k in ("Kd", "Ks", "Ke") and len(t) >= 4:
                cur[k] = tuple(float(x) for x in t[1:4])
            elif k == "Ns" and len(t) >= 2:
                cur["Ns"] = float(t[1])
            elif k in ("map_Kd", "map_Ks"):
                cur[k] = t[-1].replace("\\", "/")
            elif k in ("map_bump", "map_Bump", "bump"):
                cur["map_bump"] = t[-1].replace("\\", "/")
    return mats


def import_materials(pack, mtl, texdir, load_texture, override):
    """Scene::LoadAiSceneMaterials + Material::LoadFromAiMaterial."""
    for name, m in mtl.items():
        if name in pack.material_names and not override:
            continue
        def slot(key, fallback):
            # a texture that cannot be read leaves the slot on its solid colour (the reference prints "Failed to load
            # texture ..., ignoring it" and is then left with a null texture)
            if m[key]:
                try:
                    return load_texture(os.path.join(texdir, m[key]))
                except (OSError, ValueError) as e:
                    import warnings
                    warnings.warn(f"{e}; ignoring it")
            return fallback()
        diffuse = slot("map_Kd", lambda: pack.add_solid_texture(m["Kd"]))
        specular = slot("map_Ks", lambda: pack.add_solid_texture(m["Ks"]))
        bump = slot("map_bump", lambda: -1)
        phong_exp = F(F(4.0) * F(m["Ns"])) / F(4.0)         # assimp stores 4 Ns, the reference divides by 4
        rough = float(np.power(F(2.0) / (F(2.0) + phong_exp), F(0.5), dtype=F))
        pack.add_material(name, abi.BXDF_LTC_GGX_DIFFUSE, emission=m["Ke"], roughness=rough, tex_diffuse=diffuse,
                          tex_color=specular, tex_bump=bump)


# ---------------------------------------------------------------------------------------------- OBJ
def _triangulate(poly, P):
    """Corner lists -> triangles.  Triangles pass; quads are cut at the corner that keeps both halves inside a concave
    quad (assimp's TriangulateProcess), larger polygons by ear clipping in the plane of their Newell normal."""
    n = len(poly)
    if n == 3:
        return [tuple(poly)]
    pts = np.array([P[c[0]] for c in poly], dtype=np.float64)
    if n == 4:
        start = 0
        for i in range(4):
            v0, v1, v2 = pts[(i + 3) % 4], pts[(i + 2) % 4], pts[(i + 1) % 4]
            v = pts[i]
            a, b, c = v0 - v, v1 - v, v2 - v
            la, lb, lc = np.linalg.norm(a), np.linalg.norm(b), np.linalg.norm(c)
            if la == 0 or lb == 0 or lc == 0:
                continue
            ang = np.arccos(np.clip(np.dot(a, b) / (la * lb), -1, 1)) + np.arccos(np.clip(np.dot(b, c) / (lb * lc), -1, 1))
            if ang > np.pi:                      # reflex corner: both diagonals must start here
                start = i
                break
        q = [poly[(start + k) % 4] for k in range(4)]
        return [(q[0], q[1], q[2]), (q[0], q[2], q[3])]
    nrm = np.zeros(3)
    for i in range(n):
        a, b = pts[i], pts[(i + 1) % n]
        nrm += np.array([(a[1] - b[1]) * (a[2] + b[2]), (a[2] - b[2]) * (a[0] + b[0]), (a[0] - b[0]) * (a[1] + b[1])])
    idx, out = list(range(n)), []
    guard = 0
    while len(idx) > 3 and guard < 10 * n:
        guard += 1
        m = len(idx)
        for k in range(m):
            i0, i1, i2 = idx[(k - 1) % m], idx[k], idx[(k + 1) % m]
            cr = np.cross(pts[i1] - pts[i0], pts[i2] - pts[i1])
            if np.dot(cr, nrm) <= 0:
                continue
            inside = False
            for j in idx:
                if j in (i0, i1, i2):
                    continue
                p = pts[j]
                s = [np.dot(np.cross(pts[b] - pts[a], p - pts[a]), nrm) for a, b in ((i0, i1), (i1, i2), (i2, i0))]
                if all(x >= 0 for x in s):
                    inside = True
                    break
            if not inside:
                out.append((poly[i0], poly[i1], poly[i2]))
                idx.pop(k)
                break
        else:
            break
    if len(idx) >= 3:                            # what is left (a triangle, or a fan for a degenerate outline)
        for k in range(1, len(idx) - 1):
            out.append((poly[idx[0]], poly[idx[k]], poly[idx[k + 1]]))
    return out


def parse_obj(path):
    """Returns (P, T, N, groups, mtllibs): groups = list of (material name, [polygon = [(v, vt, vn), ...]]) in file order,
    a new group at every o / g / usemtl that is followed by faces (assimp: one mesh per object-material pair)."""
    P, T, N, groups, mtllibs = [], [], [], [], []
    cur_mat, cur = "DefaultMaterial", None
    with open(path, errors="replace") as f:
        for line in f:
            t = line.split("#", 1)[0].split()
            if not t:
                continue
            k = t[0]
            if k == "v":
                P.append((float(t[1]), float(t[2]), float(t[3])))
            elif k == "vt":
                T.append((float(t[1]), float(t[2]) if len(t) > 2 else 0.0))
            elif k == "vn":
                N.append((float(t[1]), float(t[2]), float(t[3])))
            elif k == "mtllib":
                mtllibs.append(" ".join(t[1:]))
            elif k in ("o", "g"):
                cur = None
            elif k == "usemtl":
                cur_mat, cur = " ".join(t[1:]), None
            elif k == "f":
                poly = []
                for c in t[1:]:
                    q = (c.split("/") + ["", ""])[:3]
                    v = int(q[0]); v = v - 1 if v > 0 else len(P) + v
                    vt = -1 if q[1] == "" else (int(q[1]) - 1 if int(q[1]) > 0 else len(T) + int(q[1]))
                    vn = -1 if q[2] == "" else (int(q[2]) - 1 if int(q[2]) > 0 else len(N) + int(q[2]))
                    poly.append((v, vt, vn))
                if cur is None:
                    cur = (cur_mat, [])
                    groups.append(cur)
                cur[1].append(poly)
    return (np.array(P, dtype=F).reshape(-1, 3), np.array(T, dtype=F).reshape(-1, 2), np.array(N, dtype=F).reshape(-1, 3),
            groups, mtllibs)


def build_mesh(P, T, N, polys, smooth_normals=False):
    """One assimp mesh: FindDegenerates, Triangulate, GenNormals / GenSmoothNormals where the file has none,
    CalcTangentSpace where it has texture coordinates, JoinIdenticalVertices.  Returns (pos, nrm, uv, tan, idx)."""
    tris = []
    for poly in polys:
        clean = [c for i, c in enumerate(poly) if not np.array_equal(P[c[0]], P[poly[(i + 1) % len(poly)][0]])]
        if len(clean) < 3:
            continue                                        # degenerate: becomes a line / point, which LoadAiMesh skips
        tris.extend(_triangulate(clean, P))
    if not tris:
        return None
    c = np.array(tris, dtype=np.int64).reshape(-1, 3, 3)    # [tri, corner, (v, vt, vn)]
    pos = P[c[..., 0]]                                      # [tri, 3, 3]
    has_uv, has_n = bool(np.all(c[..., 1] >= 0)) and len(T) > 0, bool(np.all(c[..., 2] >= 0)) and len(N) > 0
    uv = T[c[..., 1]] if has_uv else np.zeros(pos.shape[:2] + (2,), F)
    e1, e2 = pos[:, 1] - pos[:, 0], pos[:, 2] - pos[:, 0]
    fn = np.cross(e1, e2).astype(F)
    if has_n:
        nrm = N[c[..., 2]]
    else:
        ln = np.linalg.norm(fn, axis=1, keepdims=True)
        fnn = np.where(ln > 0, fn / np.where(ln > 0, ln, 1), 0).astype(F)
        if smooth_normals:                                  # sum of the face normals around each position
            acc = np.zeros((len(P), 3), np.float64)
            for k in range(3):
                np.add.at(acc, c[:, k, 0], fnn)
            la = np.linalg.norm(acc, axis=1, keepdims=True)
            acc = np.where(la > 0, acc / np.where(la > 0, la, 1), 0)
            nrm = acc[c[..., 0]].astype(F)
        else:
            nrm = np.repeat(fnn[:, None, :], 3, axis=1)
    tan = np.zeros_like(pos)
    if has_uv:                                              # per-face tangent from the uv deltas, projected per corner
        du1, dv1 = (uv[:, 1] - uv[:, 0]).T
        du2, dv2 = (uv[:, 2] - uv[:, 0]).T
        det = du1 * dv2 - du2 * dv1
        r = np.where(det != 0, 1.0 / np.where(det != 0, det, 1), 1.0)[:, None]
        ft = (e1 * dv2[:, None] - e2 * dv1[:, None]) * r
        for k in range(3):
            t = ft - nrm[:, k] * np.sum(ft * nrm[:, k], axis=1, keepdims=True)
            lt = np.linalg.norm(t, axis=1, keepdims=True)
            tan[:, k] = np.where(lt > 0, t / np.where(lt > 0, lt, 1), 0)
    # JoinIdenticalVertices: unique (position, normal, uv, tangent) rows, first occurrence first
    rows = np.concatenate([pos.reshape(-1, 3), nrm.reshape(-1, 3), uv.reshape(-1, 2), tan.reshape(-1, 3)], axis=1).astype(F)
    rows = rows + F(0.0)
    _, first, inv = np.unique(rows.view(np.uint32), axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first)
    rank = np.empty_like(order); rank[order] = np.arange(len(order))
    uniq = rows[first[order]]
    idx = rank[inv.reshape(-1)].reshape(-1, 3).astype(np.uint32)
    if has_uv:                                              # CalcTangentSpace averages over the faces sharing a vertex
        acc = np.zeros((len(uniq), 3), np.float64)
        np.add.at(acc, idx.reshape(-1), tan.reshape(-1, 3))
        la = np.linalg.norm(acc, axis=1, keepdims=True)
        uniq[:, 8:11] = np.where(la > 0, acc / np.where(la > 0, la, 1), 0).astype(F)
    return uniq[:, 0:3], uniq[:, 3:6], uniq[:, 6:8], uniq[:, 8:11], idx


def load_obj_into(pack, path, transform, obj=None, root=None, texture_loader=load_image):
    """mesh_loader for scene.load_config: ConfigJSON::InstallScene's "file" branch (src/config.cpp:436-486)."""
    from .scene import ConfigFileException
    obj = obj or {}
    if not os.path.exists(path):
        raise ConfigFileException(f'Unable to find model file "{path}"')
    P, T, N, groups, mtllibs = parse_obj(path)
    moddir = os.path.dirname(path)
    if obj.get("import-materials", False):
        cache = pack.__dict__.setdefault("_asset_textures", {})

        def load_texture(p):
            if p not in cache:
                cache[p] = pack.add_image_texture(texture_loader(p))
            return cache[p]
        mtl = {}
        for lib in mtllibs:
            mtl.update(parse_mtl(os.path.join(moddir, lib)))
        used = [g[0] for g in groups]
        if "DefaultMaterial" in used and "DefaultMaterial" not in mtl:
            mtl["DefaultMaterial"] = dict(Kd=(0.6, 0.6, 0.6), Ks=(0.0, 0.0, 0.0), Ke=(0.0, 0.0, 0.0), Ns=0.0, map_Kd="", map_Ks="", map_bump="")
        import_materials(pack, mtl, moddir, load_texture, bool(obj.get("override-materials", False)))
    forced = obj.get("material", "")
    M = np.array(transform, dtype=F)                         # column-major 4x4 as scene.object_transform builds it
    for mat, polys in groups:
        mesh = build_mesh(P, T, N, polys, bool(obj.get("smooth-normals", False)))
        if mesh is None:
            continue
        pos, nrm, uv, tan, idx = mesh
        # vertex = (current_transform * vec4(v, 1)).xyz, normal = mat3(current_transform) * n  (src/scene.cpp:149-160,189-193)
        # (GLM's mat4 * vec4 adds (m0 x + m1 y) + (m2 z + m3 w); mat3 * vec3 adds left to right)
        tp = np.stack([(M[0][k] * pos[:, 0] + M[1][k] * pos[:, 1]) + (M[2][k] * pos[:, 2] + M[3][k] * F(1.0)) for k in range(3)], 1).astype(F)
        tn = np.stack([M[0][k] * nrm[:, 0] + M[1][k] * nrm[:, 1] + M[2][k] * nrm[:, 2] for k in range(3)], 1).astype(F)
        tt = np.stack([M[0][k] * tan[:, 0] + M[1][k] * tan[:, 1] + M[2][k] * tan[:, 2] for k in range(3)], 1).astype(F)
        name = forced or mat
        if name not in pack.material_names:
            raise ValueError(f'Error: Material named "{name}" was not defined')      # Scene::GetMaterialByName, src/scene.cpp:289
        pack.add_mesh(tp, tn, uv, tt, idx, name)
