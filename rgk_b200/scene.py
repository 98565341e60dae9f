"""Scene pack: the host-side input of the hot path (include/rgk_b200.h: rgk_scene_desc).

`ScenePack` collects meshes / materials / textures / lights as numpy arrays and
exposes them as a ctypes SceneDesc.  `load_json_config` reads the reference's
JSON scene config (schema: SURVEY Appendix D; src/config.cpp:260-558,
src/bxdf/bxdf.cpp:52-86,207-330) for scenes made of built-in primitives and OBJ
files.  This is input-format plumbing on the host; nothing here is on the timed path.
"""
import ctypes as C
import json
import math
import os
import re

import numpy as np

from . import abi

F = np.float32
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_ltc_tables():
    """LTC fit tables (GGX, Beckmann): M[4096,9] float32 in mat33::m order, amplitude[4096].

    Extracted once from the reference's data tables (src/LTC/ltc_ggx.cpp, ltc_beckmann.cpp) by
    tools/extract_ltc.py with the double->float cast of src/LTC/ltc.hpp:6-9."""
    z = np.load(os.path.join(_DATA, "ltc_tables.npz"))
    return {k: np.ascontiguousarray(z[k], dtype=F) for k in ("ggx_M", "ggx_amp", "beckmann_M", "beckmann_amp")}


# ---------------------------------------------------------------- GLM-formula transforms (float32, op order as GLM)
def _mat4_identity():
    return [np.array(c, dtype=F) for c in ((1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1))]


def _mat4_mul(a, b):
    return [a[0] * b[j][0] + a[1] * b[j][1] + a[2] * b[j][2] + a[3] * b[j][3] for j in range(4)]


def _mat4_vec4(m, v):
    return (m[0] * v[0] + m[1] * v[1]) + (m[2] * v[2] + m[3] * v[3])


def _scale(v):
    m = _mat4_identity()
    return [m[0] * F(v[0]), m[1] * F(v[1]), m[2] * F(v[2]), m[3]]


def _translate(v):
    m = _mat4_identity()
    return [m[0], m[1], m[2], m[0] * F(v[0]) + m[1] * F(v[1]) + m[2] * F(v[2]) + m[3]]


def _normalize3(v):
    v = np.asarray(v, dtype=F)
    d = F(F(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
    return v * F(F(1.0) / np.sqrt(d, dtype=F))


def _rotate(angle, axis):
    a = F(angle)
    c, s = F(np.cos(a, dtype=F)), F(np.sin(a, dtype=F))
    ax = _normalize3(axis)
    t = (F(1.0) - c) * ax
    R = [[c + t[0] * ax[0], t[0] * ax[1] + s * ax[2], t[0] * ax[2] - s * ax[1]],
         [t[1] * ax[0] - s * ax[2], c + t[1] * ax[1], t[1] * ax[2] + s * ax[0]],
         [t[2] * ax[0] + s * ax[1], t[2] * ax[1] - s * ax[0], c + t[2] * ax[2]]]
    m = _mat4_identity()
    return [m[0] * R[0][0] + m[1] * R[0][1] + m[2] * R[0][2],
            m[0] * R[1][0] + m[1] * R[1][1] + m[2] * R[1][2],
            m[0] * R[2][0] + m[1] * R[2][1] + m[2] * R[2][2], m[3]]


def object_transform(scale=(1, 1, 1), rotate=(0, 0, 0), translate=(0, 0, 0), base=None):
    """scale -> rotate (Z, Y, X about the negated axes, degrees) -> translate, src/config.cpp:513-523."""
    t = base if base is not None else _mat4_identity()
    t = _mat4_mul(_scale(scale), t)
    t = _mat4_mul(_rotate(F(0.0174533) * F(rotate[2]), (0.0, 0.0, -1.0)), t)
    t = _mat4_mul(_rotate(F(0.0174533) * F(rotate[1]), (0.0, -1.0, 0.0)), t)
    t = _mat4_mul(_rotate(F(0.0174533) * F(rotate[0]), (-1.0, 0.0, 0.0)), t)
    t = _mat4_mul(_translate(translate), t)
    return t


# ---------------------------------------------------------------- built-in primitives (src/primitives.cpp:168-228)
_PATTERN = ((1, 1), (1, -1), (-1, 1), (-1, -1), (-1, 1), (1, -1))


def _face(place, normal, tangent):
    v = [place(a, b) for a, b in _PATTERN]
    uv = [((a + 1) / 2, (b + 1) / 2) for a, b in _PATTERN]
    return v, [normal] * 6, uv, [tangent] * 6


def primitive_data(kind):
    """(positions, normals, uvs, tangents) of 'plane' (planeY), 'tri' (trigY), 'cube'."""
    if kind in ("plane", "tri"):
        v, n, uv, t = _face(lambda a, b: (a, 0, b), (0, 1, 0), (0, 0, 1))
        k = 6 if kind == "plane" else 3
        return tuple(np.array(x[:k], dtype=F) for x in (v, n, uv, t))
    if kind == "cube":
        faces = [
            _face(lambda a, b: (1, a, b), (1, 0, 0), (0, 0, 1)), _face(lambda a, b: (-1, a, b), (-1, 0, 0), (0, 0, 1)),
            _face(lambda a, b: (a, 1, b), (0, 1, 0), (1, 0, 0)), _face(lambda a, b: (-a, -1, b), (0, -1, 0), (1, 0, 0)),
            _face(lambda a, b: (b, a, 1), (0, 0, 1), (0, 1, 0)), _face(lambda a, b: (b, a, -1), (0, 0, -1), (0, 1, 0)),
        ]
        return tuple(np.array(sum((f[i] for f in faces), []), dtype=F) for i in range(4))
    raise ValueError("primitive must be 'plane', 'tri' or 'cube'")


def transform_primitive(data, transform, texscale=(1, 1, 1)):
    """Scene::AddPrimitive, src/scene.cpp:218-229."""
    pos, nrm, uv, tan = data
    P, N, T, U = [], [], [], []
    for i in range(len(pos)):
        P.append(_mat4_vec4(transform, np.array([*pos[i], 1.0], dtype=F))[:3])
        N.append(_normalize3(_mat4_vec4(transform, np.array([*nrm[i], 0.0], dtype=F))[:3]))
        T.append(_normalize3(_mat4_vec4(transform, np.array([*tan[i], 0.0], dtype=F))[:3]))
        # texture_transform = mat3(scale(texscale)): diag; (M * vec3(uv,1)).xy
        U.append(np.array([F(texscale[0]) * uv[i][0] + F(0) * uv[i][1] + F(0) * F(1),
                           F(0) * uv[i][0] + F(texscale[1]) * uv[i][1] + F(0) * F(1)], dtype=F))
    return np.array(P, dtype=F), np.array(N, dtype=F), np.array(U, dtype=F), np.array(T, dtype=F)


# ---------------------------------------------------------------- the pack
class ScenePack:
    def __init__(self):
        self.meshes = []        # (positions[n,3], normals[n,3], uvs[n,2], tangents[n,3], indices[m,3], material)
        self.materials = []     # dicts
        self.material_names = {}
        self.textures = []      # ("solid", (r,g,b)) | ("image", ndarray[h,w,3] float32)
        self.point_lights = []  # (pos, color, intensity, size)
        self.sky = dict(mode=0, color=(0.0, 0.0, 0.0), intensity=1.0, rotate=0.0, envmap=-1)
        self.thinglass = 0
        self._keep = None

    # -- textures / materials
    def add_solid_texture(self, rgb):
        self.textures.append(("solid", tuple(float(F(c)) for c in rgb)))
        return len(self.textures) - 1

    def add_image_texture(self, img):
        img = np.ascontiguousarray(img, dtype=F)
        assert img.ndim == 3 and img.shape[2] == 3
        self.textures.append(("image", img))
        return len(self.textures) - 1

    def add_material(self, name, bxdf, emission=(0, 0, 0), no_russian=False, roughness=0.0, ior=1.0, amount=0.0,
                     mix_a=-1, mix_b=-1, tex_diffuse=-1, tex_color=-1, tex_bump=-1):
        m = dict(name=name, bxdf=bxdf, emission=tuple(emission), no_russian=bool(no_russian), roughness=roughness,
                 ior=ior, amount=amount, mix_a=mix_a, mix_b=mix_b, tex_diffuse=tex_diffuse, tex_color=tex_color,
                 tex_bump=tex_bump)
        if name in self.material_names:   # RegisterMaterial(override=true), src/scene.cpp:73-94
            self.materials[self.material_names[name]] = m
        else:
            self.material_names[name] = len(self.materials)
            self.materials.append(m)
        return self.material_names[name]

    def add_mesh(self, positions, normals, uvs, tangents, indices, material):
        if isinstance(material, str):
            if material not in self.material_names:
                raise ValueError(f'Error: Material named "{material}" was not defined')  # src/scene.cpp:289
            material = self.material_names[material]
        # "+ 0.0" canonicalises -0.0 to +0.0, as the reference's transform multiply does (x*1 + y*0 + ...)
        a = [np.ascontiguousarray(np.asarray(x, dtype=F) + F(0.0)) for x in (positions, normals, uvs, tangents)]
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        if len(idx) and (int(idx.max()) >= len(a[0])):
            raise ValueError("mesh index out of range")
        # drop vertices the mesh does not reference? No: assimp meshes are dense; require it.
        self.meshes.append((a[0], a[1], a[2], a[3], idx, int(material)))

    def add_point_light(self, position, color=(1, 1, 1), intensity=1.0, size=0.0):
        self.point_lights.append((tuple(position), tuple(color), float(intensity), float(size)))

    def set_sky_color(self, color, intensity=1.0):
        self.sky = dict(mode=0, color=tuple(color), intensity=float(intensity), rotate=0.0, envmap=-1)

    def set_sky_envmap(self, texture, intensity=1.0, rotate=0.0):
        self.sky = dict(mode=1, color=(0.0, 0.0, 0.0), intensity=float(intensity), rotate=float(rotate),
                        envmap=int(texture))

    # -- flattening
    @property
    def n_triangles(self):
        return sum(len(m[4]) for m in self.meshes)

    def arrays(self):
        pos = np.concatenate([m[0] for m in self.meshes]) if self.meshes else np.zeros((0, 3), F)
        nrm = np.concatenate([m[1] for m in self.meshes]) if self.meshes else np.zeros((0, 3), F)
        uv = np.concatenate([m[2] for m in self.meshes]) if self.meshes else np.zeros((0, 2), F)
        tan = np.concatenate([m[3] for m in self.meshes]) if self.meshes else np.zeros((0, 3), F)
        idx, ranges, voff, toff = [], [], 0, 0
        for m in self.meshes:
            idx.append(m[4] + np.uint32(voff))
            ranges.append((toff, len(m[4]), m[5]))
            voff += len(m[0])
            toff += len(m[4])
        idx = np.concatenate(idx).astype(np.uint32) if idx else np.zeros((0, 3), np.uint32)
        return dict(positions=np.ascontiguousarray(pos, F), normals=np.ascontiguousarray(nrm, F),
                    texcoords=np.ascontiguousarray(uv, F), tangents=np.ascontiguousarray(tan, F),
                    indices=np.ascontiguousarray(idx), mesh_ranges=ranges)

    def desc(self):
        """Build the ctypes SceneDesc; every buffer it points into is kept alive on self."""
        a = self.arrays()
        ltc = load_ltc_tables()
        keep = [a, ltc]
        d = abi.SceneDesc()
        fp = lambda x: x.ctypes.data_as(abi.f32p)
        d.n_vertices = len(a["positions"])
        d.positions, d.normals, d.tangents, d.texcoords = fp(a["positions"]), fp(a["normals"]), fp(a["tangents"]), fp(a["texcoords"])
        d.n_triangles = len(a["indices"])
        d.indices = a["indices"].ctypes.data_as(abi.u32p)
        meshes = (abi.Mesh * max(1, len(a["mesh_ranges"])))()
        for i, (f, n, m) in enumerate(a["mesh_ranges"]):
            meshes[i].first_triangle, meshes[i].n_triangles, meshes[i].material = f, n, m
        d.n_meshes, d.meshes = len(a["mesh_ranges"]), meshes
        mats = (abi.Material * max(1, len(self.materials)))()
        for i, m in enumerate(self.materials):
            mm = mats[i]
            mm.bxdf, mm.no_russian = m["bxdf"], int(m["no_russian"])
            mm.emission = (C.c_float * 3)(*m["emission"])
            mm.roughness, mm.ior, mm.amount = m["roughness"], m["ior"], m["amount"]
            mm.mix_a, mm.mix_b = m["mix_a"], m["mix_b"]
            mm.tex_diffuse, mm.tex_color, mm.tex_bump = m["tex_diffuse"], m["tex_color"], m["tex_bump"]
        d.n_materials, d.materials = len(self.materials), mats
        texs = (abi.Texture * max(1, len(self.textures)))()
        for i, (kind, val) in enumerate(self.textures):
            if kind == "solid":
                texs[i].kind, texs[i].width, texs[i].height = 0, 0, 0
                texs[i].color = (C.c_float * 3)(*val)
            else:
                texs[i].kind, texs[i].height, texs[i].width = 1, val.shape[0], val.shape[1]
                texs[i].texels = fp(val)
        d.n_textures, d.textures = len(self.textures), texs
        pls = (abi.PointLight * max(1, len(self.point_lights)))()
        for i, (p, c, inten, size) in enumerate(self.point_lights):
            pls[i].position = (C.c_float * 3)(*p)
            pls[i].color = (C.c_float * 3)(*c)
            pls[i].intensity, pls[i].size = inten, size
        d.n_point_lights, d.point_lights = len(self.point_lights), pls
        d.sky.mode = self.sky["mode"]
        d.sky.color = (C.c_float * 3)(*self.sky["color"])
        d.sky.intensity, d.sky.rotate, d.sky.envmap = self.sky["intensity"], self.sky["rotate"], self.sky["envmap"]
        d.ltc_ggx.M, d.ltc_ggx.amplitude = fp(ltc["ggx_M"]), fp(ltc["ggx_amp"])
        d.ltc_beckmann.M, d.ltc_beckmann.amplitude = fp(ltc["beckmann_M"]), fp(ltc["beckmann_amp"])
        d.thinglass = int(self.thinglass)
        keep += [meshes, mats, texs, pls]
        self._keep = keep
        return d


    # -- on-disk form (RGKPACK1, little endian; read by rgkb::PackFile in include/rgk_b200_host.hpp and load_pack below)
    def save(self, path, cfg):
        """Writes the flattened scene, the LTC tables, the render configuration and the Camera constructor arguments."""
        a = self.arrays()
        ltc = load_ltc_tables()
        d = self.desc()
        ca = cfg.camera_args()
        with open(path, "wb") as f:
            f.write(b"RGKPACK1")
            f.write(np.array([len(a["positions"]), len(a["indices"]), len(a["mesh_ranges"]), len(self.materials), len(self.textures),
                              len(self.point_lights), int(self.thinglass), 0], np.uint32).tobytes())
            for k in ("positions", "normals", "tangents", "texcoords", "indices"):
                f.write(a[k].tobytes())
            f.write(bytes(C.string_at(d.meshes, C.sizeof(abi.Mesh) * d.n_meshes)))
            f.write(bytes(C.string_at(d.materials, C.sizeof(abi.Material) * d.n_materials)))
            for kind, val in self.textures:
                if kind == "solid":
                    f.write(np.array([0, 0, 0], np.uint32).tobytes()); f.write(np.array(val, F).tobytes())
                else:
                    f.write(np.array([1, val.shape[1], val.shape[0]], np.uint32).tobytes()); f.write(np.zeros(3, F).tobytes())
                    f.write(np.ascontiguousarray(val, F).tobytes())
            f.write(bytes(C.string_at(d.point_lights, C.sizeof(abi.PointLight) * d.n_point_lights)))
            f.write(bytes(d.sky))
            for k in ("ggx_M", "ggx_amp", "beckmann_M", "beckmann_amp"):
                f.write(np.ascontiguousarray(ltc[k], F).tobytes())
            f.write(np.array([cfg.xres, cfg.yres, cfg.multisample, cfg.recursion_level, cfg.rounds, int(cfg.force_fresnell), cfg.reverse, 0],
                             np.uint32).tobytes())
            f.write(np.array([cfg.clamp, cfg.russian, cfg.bumpmap_scale, cfg.output_scale], F).tobytes())
            f.write(np.array(list(ca["pos"]) + list(ca["lookat"]) + list(ca["up"]) + [ca["yview"], ca["xview"], ca["focus_plane"], ca["lens_size"]], F).tobytes())


def load_pack(path):
    """Reads an RGKPACK1 file back into (ScenePack, RenderConfig); the config's camera is returned as ready-made
    Camera constructor arguments (cfg.camera_args() gives them back unchanged)."""
    raw = open(path, "rb").read()
    if raw[:8] != b"RGKPACK1":
        raise ValueError("not an RGKPACK1 file: " + path)
    off = 8

    def take(dtype, n):
        nonlocal off
        arr = np.frombuffer(raw, dtype=dtype, count=n, offset=off).copy()
        off += arr.nbytes
        return arr
    nv, nt, nm, nmat, ntex, npl, thinglass, _ = (int(x) for x in take(np.uint32, 8))
    pos, nrm, tan = take(F, 3 * nv).reshape(-1, 3), take(F, 3 * nv).reshape(-1, 3), take(F, 3 * nv).reshape(-1, 3)
    uv, idx = take(F, 2 * nv).reshape(-1, 2), take(np.uint32, 3 * nt).reshape(-1, 3)
    meshes = take(np.uint32, 4 * nm).reshape(-1, 4)
    mats = (abi.Material * max(1, nmat)).from_buffer_copy(raw[off:off + 64 * nmat].ljust(64, b"\0")); off += 64 * nmat
    pack = ScenePack()
    pack.thinglass = thinglass
    for i in range(nmat):
        m = mats[i]
        pack.add_material("m%d" % i, int(m.bxdf), tuple(m.emission), bool(m.no_russian), float(m.roughness), float(m.ior), float(m.amount),
                          int(m.mix_a), int(m.mix_b), int(m.tex_diffuse), int(m.tex_color), int(m.tex_bump))
    for _ in range(ntex):
        kind, w, h = (int(x) for x in take(np.uint32, 3))
        col = take(F, 3)
        if kind == 0:
            pack.textures.append(("solid", tuple(float(c) for c in col)))
        else:
            pack.textures.append(("image", take(F, 3 * w * h).reshape(h, w, 3)))
    pls = take(F, 8 * npl).reshape(-1, 8)
    for q in pls:
        pack.point_lights.append((tuple(float(x) for x in q[0:3]), tuple(float(x) for x in q[3:6]), float(q[6]), float(q[7])))
    sky = abi.Sky.from_buffer_copy(raw[off:off + C.sizeof(abi.Sky)]); off += C.sizeof(abi.Sky)
    pack.sky = dict(mode=int(sky.mode), color=tuple(sky.color), intensity=float(sky.intensity), rotate=float(sky.rotate), envmap=int(sky.envmap))
    off += 4 * (2 * 4096 * 9 + 2 * 4096)          # LTC tables: load_ltc_tables() supplies the same data
    for first, n, mat, _ in meshes:
        first, n = int(first), int(n)
        tri = idx[first:first + n]
        lo, hi = (int(tri.min()), int(tri.max()) + 1) if n else (0, 0)
        pack.meshes.append((pos[lo:hi], nrm[lo:hi], uv[lo:hi], tan[lo:hi], (tri - np.uint32(lo)).astype(np.uint32), int(mat)))
    c = take(np.uint32, 8)
    cf = take(F, 4)
    ca = take(F, 13)
    cfg = RenderConfig()
    cfg.xres, cfg.yres, cfg.multisample, cfg.recursion_level, cfg.rounds = (int(x) for x in c[:5])
    cfg.force_fresnell, cfg.reverse = bool(c[5]), int(c[6])
    cfg.clamp, cfg.russian, cfg.bumpmap_scale, cfg.output_scale = (float(x) for x in cf)
    args = dict(pos=ca[0:3].copy(), lookat=ca[3:6].copy(), up=ca[6:9].copy(), yview=float(ca[9]), xview=float(ca[10]),
                xres=cfg.xres, yres=cfg.yres, focus_plane=float(ca[11]), lens_size=float(ca[12]))
    cfg.camera_args = lambda: dict(args)
    return pack, cfg


# ---------------------------------------------------------------- JSON config (src/config.cpp:260-558)
class ConfigFileException(RuntimeError):
    pass


def _strip_comments(text):
    out, i, n, in_str = [], 0, len(text), False
    while i < n:
        ch = text[i]
        if in_str:
            out.append(ch)
            if ch == "\\" and i + 1 < n:
                out.append(text[i + 1]); i += 1
            elif ch == '"':
                in_str = False
        elif ch == '"':
            in_str = True; out.append(ch)
        elif text.startswith("//", i):
            while i < n and text[i] != "\n":
                i += 1
            continue
        elif text.startswith("/*", i):
            i = text.find("*/", i + 2)
            i = n if i < 0 else i + 2
            continue
        else:
            out.append(ch)
        i += 1
    text = re.sub(r",(\s*[\]}])", r"\1", "".join(out))
    # jsoncpp also reads numbers with leading zeros ("000.0" in scenes/conference.json:16); strict JSON does not
    return re.sub(r'("(?:\\.|[^"\\])*")|(?<![\w.])(-?)0+(?=\d)', lambda m: m.group(1) if m.group(1) is not None else m.group(2), text)


def _vec3(node, key, default=None, required=False):
    for k, div in ((key, 1.0), (key + "255", 255.0)):
        if k in node:
            v = node[k]
            if isinstance(v, (int, float)):
                v = [v, v, v]
            if not (isinstance(v, list) and len(v) == 3):
                raise ConfigFileException(f'value "{k}" must be an array of 3 numbers or a single number')
            return tuple(float(F(x) / F(div)) if div != 1.0 else float(F(x)) for x in v)
    if required:
        raise ConfigFileException(f'Required value "{key}" is missing')
    return default


class RenderConfig:
    """Config fields (src/config.hpp:25-56) with ConfigJSON's defaults (src/config.cpp:277-323)."""

    def __init__(self):
        self.output_file = ""
        self.xres = self.yres = 0
        self.rounds = 1
        self.render_minutes = None
        self.recursion_level = 40
        self.multisample = 1
        self.clamp = 10000000.0
        self.bumpmap_scale = 1.0
        self.russian = 0.74
        self.reverse = 0
        self.force_fresnell = False
        self.output_scale = -1.0
        self.camera = {}

    def params(self, sampler_mode=abi.SAMPLER_MT19937):
        p = abi.RenderParams()
        p.xres, p.yres, p.multisample, p.depth = self.xres, self.yres, self.multisample, self.recursion_level
        p.clamp, p.russian, p.bumpmap_scale = self.clamp, self.russian, self.bumpmap_scale
        p.force_fresnell, p.reverse, p.sampler_mode = int(self.force_fresnell), self.reverse, sampler_mode
        return p

    def camera_args(self):
        """Arguments of Camera::Camera as ConfigJSON::GetCamera(0) computes them, src/config.cpp:332-370."""
        c = self.camera
        if "focal" in c:
            yview = F(c["focal"]); xview = yview * F(self.xres) / F(self.yres)
        elif "fov" in c:
            xview = F(2.0) * F(np.tan(F(F(c["fov"]) * F(0.0174533)) / F(2.0), dtype=F))
            yview = xview * F(self.yres) / F(self.xres)
        else:
            raise ConfigFileException('Camera must either have a "fov" or "focal" key defined')
        pos = np.array(c["position"], dtype=F)
        la = np.array(c["lookat"], dtype=F)
        up = np.array(c.get("upvector", (0.0, 1.0, 0.0)), dtype=F)
        # GetCamera(rotation=0): p = rotate(lookat - pos, 0, up); pos = lookat - p  (rotation by 0 is the identity
        # up to rounding of the rotation matrix; reproduced with the same formula)
        p = la - pos
        R = _rotate(F(0.0) * F(2.0) * F(np.pi), up)
        m3 = [R[0][:3], R[1][:3], R[2][:3]]
        p = np.array([m3[0][k] * p[0] + m3[1][k] * p[1] + m3[2][k] * p[2] for k in range(3)], dtype=F)
        pos = la - p
        return dict(pos=pos, lookat=la, up=up, yview=float(yview), xview=float(xview), xres=self.xres, yres=self.yres,
                    focus_plane=float(c.get("focus-plane", 1.0)), lens_size=float(c.get("lens-size", 0.0)))


def _load_material(pack, node, texdir, load_texture):
    """Material::LoadFromJson + BxDF*::LoadFromJson, src/bxdf/bxdf.cpp:52-86,207-330."""
    name = node["name"]
    emission = _vec3(node, "emission", (0.0, 0.0, 0.0))
    bump = -1
    if node.get("bump-map", ""):
        bump = load_texture(os.path.join(texdir, node["bump-map"]))
    brdf = node.get("brdf")
    if brdf is None:
        raise ConfigFileException('Required value "brdf" is missing')
    kw = dict(emission=emission, no_russian=node.get("no-russian", False), tex_bump=bump)

    def colour_slot(keys_tex, keys_col, default):
        for k in keys_tex:
            if node.get(k, ""):
                return load_texture(os.path.join(texdir, node[k]))
        for k in keys_col:
            if k in node or k + "255" in node:
                return pack.add_solid_texture(_vec3(node, k, required=True))
        return pack.add_solid_texture(default)

    if brdf in ("diffuse", "diffusecosine"):
        kw["tex_diffuse"] = colour_slot(["diffuse-texture"], ["diffuse"], (0.5, 0.5, 0.5))
        return pack.add_material(name, abi.BXDF_DIFFUSE, **kw)
    if brdf == "mix":
        for k in ("material1", "material2"):
            if node[k] not in pack.material_names:
                raise ConfigFileException(f'Material "{node[k]}", used for mixing, was not (yet) defined')
        return pack.add_material(name, abi.BXDF_MIX, mix_a=pack.material_names[node["material1"]],
                                 mix_b=pack.material_names[node["material2"]], amount=float(node["amount"]), **kw)
    if brdf == "mirror":
        kw["tex_color"] = colour_slot(["color-texture"], ["color"], (1.0, 1.0, 1.0))
        return pack.add_material(name, abi.BXDF_MIRROR, **kw)
    if brdf == "dielectric":
        kw["tex_color"] = colour_slot(["color-texture", "specular-texture"], ["color"], (1.0, 1.0, 1.0))
        return pack.add_material(name, abi.BXDF_DIELECTRIC, ior=float(node["ior"]), **kw)
    if brdf == "transparent":
        return pack.add_material(name, abi.BXDF_TRANSPARENT, **kw)
    ltc = {"ltc_beckmann": abi.BXDF_LTC_BECKMANN, "ltc_ggx": abi.BXDF_LTC_GGX,
           "ltc_beckmann_diffuse": abi.BXDF_LTC_BECKMANN_DIFFUSE, "ltc_ggx_diffuse": abi.BXDF_LTC_GGX_DIFFUSE}
    if brdf in ltc:
        if "roughness" in node:
            rough = float(node["roughness"])
        elif "exponent" in node:
            rough = float(np.power(F(2.0) / (F(2.0) + F(node["exponent"])), F(0.5), dtype=F))
        else:
            raise ConfigFileException('Either "roughness" or "exponent" must be present for LTC BxDF')
        kw["tex_color"] = colour_slot(["color-texture", "specular-texture"], ["color", "specular"], (0.0, 0.0, 0.0))
        if brdf.endswith("_diffuse"):
            kw["tex_diffuse"] = colour_slot(["diffuse-texture"], ["diffuse"], (0.0, 0.0, 0.0))
        return pack.add_material(name, ltc[brdf], roughness=rough, **kw)
    raise ConfigFileException("Unsupported BRDF id in config!")


def load_json_config(path, overrides=None, mesh_loader=None, texture_loader=None):
    """Reads a reference-style JSON scene file. Returns (ScenePack, RenderConfig)."""
    with open(path) as f:
        root = json.loads(_strip_comments(f.read()))
    return load_config(root, os.path.dirname(os.path.abspath(path)), overrides, mesh_loader, texture_loader)


def load_config(root, cfgdir=".", overrides=None, mesh_loader=None, texture_loader=None):
    """Same from an already-parsed config dict. `overrides` patches root keys (e.g. output-width).  Mesh files and
    image textures go through rgk_b200.assets (OBJ / MTL, PNG / JPEG / HDR) unless other loaders are given."""
    if mesh_loader is None or texture_loader is None:
        from . import assets
        mesh_loader = mesh_loader or assets.load_obj_into
        texture_loader = texture_loader or assets.load_image
    root = dict(root)
    root.update(overrides or {})
    cfg = RenderConfig()
    for key in ("output-file", "output-width", "output-height"):
        if key not in root:
            raise ConfigFileException(f'Required value "{key}" is missing from the config file.')
    cfg.output_file, cfg.xres, cfg.yres = root["output-file"], int(root["output-width"]), int(root["output-height"])
    if "rounds" in root and "render-time" in root:
        raise ConfigFileException('The config file may not contain both "rounds" and "render-time" keys simultaneously.')
    cfg.rounds = int(root.get("rounds", 1))
    cfg.render_minutes = root.get("render-time")
    cfg.recursion_level = int(root.get("recursion-max", 40))
    cfg.multisample = int(root.get("multisample", 1))
    cfg.clamp = float(root.get("clamp", 10000000.0))
    cfg.bumpmap_scale = float(root.get("bumpscale", 1.0))
    cfg.russian = float(root.get("russian", 0.74))
    cfg.reverse = int(root.get("reverse", 0))
    cfg.force_fresnell = bool(root.get("force-fresnell", False))
    if "output-scale" in root:                       # src/config.cpp:303-312: a number, or the string "auto" (-1: normalise by 1 / max)
        v = root["output-scale"]
        if isinstance(v, str):
            if v != "auto":
                raise ConfigFileException('The value of "output-scale" must either be a number, or "auto".')
            cfg.output_scale = -1.0
        elif isinstance(v, (int, float)) and not isinstance(v, bool):
            cfg.output_scale = float(np.float32(v))
        else:
            raise ConfigFileException('The value of "output-scale" must either be a number, or "auto".')
    if "camera" not in root:
        raise ConfigFileException('Value "camera" is missing.')
    cfg.camera = root["camera"]

    pack = ScenePack()
    tex_cache = {}

    def load_texture(p):
        if p not in tex_cache:
            if texture_loader is None:
                raise ConfigFileException(f"texture '{p}' requested but no texture_loader given")
            tex_cache[p] = pack.add_image_texture(texture_loader(p))
        return tex_cache[p]

    for m in root.get("materials", []):
        _load_material(pack, m, cfgdir, load_texture)
    if "model-file" in root and "scene" in root:
        raise ConfigFileException('The input file may not contain both "model-file" key and "scene" key.')
    objects = root.get("scene")
    if objects is None:
        if "model-file" in root:
            objects = [{"file": root["model-file"], "import-materials": True}]
        else:
            raise ConfigFileException('The input file contains neither "scene" nor "model-file" key.')
    for obj in objects:
        if "file" in obj and "primitive" in obj:
            raise ConfigFileException('Both "file" and "primitive" keys found')
        scale = _vec3(obj, "scale", (1.0, 1.0, 1.0))
        translate = _vec3(obj, "translate", (0.0, 0.0, 0.0))
        rotate = _vec3(obj, "rotate", (0.0, 0.0, 0.0))
        if "primitive" in obj:
            kind = obj["primitive"]
            base = _mat4_identity()
            if kind == "cube":
                base = _mat4_mul(_scale((0.5, 0.5, 0.5)), base)
            axis = obj.get("axis", "Y")
            if axis == "X":
                base = _mat4_mul(_rotate(F(np.pi) / F(2.0), (0.0, 0.0, 1.0)), base)
            elif axis == "Z":
                base = _mat4_mul(_rotate(F(np.pi) / F(2.0), (1.0, 0.0, 0.0)), base)
            elif axis != "Y":
                raise ConfigFileException('Optional value "axis" must be either X, Y or Z.')
            t = object_transform(scale, rotate, translate, base)
            P, N, U, T = transform_primitive(primitive_data(kind), t, _vec3(obj, "texture-scale", (1.0, 1.0, 1.0)))
            if "material" not in obj:
                raise ConfigFileException('Required value "material" is missing')
            idx = np.arange(len(P), dtype=np.uint32).reshape(-1, 3)
            pack.add_mesh(P, N, U, T, idx, obj["material"])
        elif "file" in obj:
            if mesh_loader is None:
                raise ConfigFileException(f"mesh file '{obj['file']}' requested but no mesh_loader given")
            t = object_transform(scale, rotate, translate)
            mesh_loader(pack, os.path.join(cfgdir, obj["file"]), t, obj, root)
        else:
            raise ConfigFileException('Missing mesh data: needs a "file" key or a "primitive" key.')
    for l in root.get("lights", []):
        pack.add_point_light(_vec3(l, "position", required=True), _vec3(l, "color", (1.0, 1.0, 1.0)),
                             float(l["intensity"]), float(l.get("size", 0.0)))
    sky = root.get("sky")
    if sky is not None:
        if "envmap" in sky:
            pack.set_sky_envmap(load_texture(os.path.join(cfgdir, sky["envmap"])), float(sky.get("intensity", 1.0)),
                                float(sky.get("rotate", 0.0)))
        elif "color" in sky or "color255" in sky:
            pack.set_sky_color(_vec3(sky, "color", required=True), float(sky.get("intensity", 1.0)))
        else:
            raise ConfigFileException('Sky configuration must either contain an "envmap" key or a "color" key')
    pack.thinglass = 1 if root.get("thinglass") else 0
    return pack, cfg
