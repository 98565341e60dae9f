"""Seeded stand-in scenes for the BASELINE configs whose assets are not distributed
(sponza.obj, sibenik.obj, cloudy1.hdr are missing from the reference checkout; conference
and dragon were never in it -- SURVEY D5).  Every number measured on these is tagged
`standin`.  The scenes are deterministic (fixed seeds, pure numpy) so that the CPU
oracle, the reference build and the GPU all see byte-identical input.

An "atrium" generator builds a Sponza-like colonnaded hall from parametric pieces
(tessellated floor/walls, fluted columns, arches, draped curtains, urns) whose
tessellation is scaled to hit a target triangle count; materials are what every
assimp-imported material becomes in the reference, BxDFLTCDiffuse<GGX> with
diffuse/specular/bump textures (src/bxdf/bxdf.cpp:176-180), with procedural textures.
Cameras, lights, sky and render parameters are taken from the corresponding reference
scene files (scenes/sponza.json, sibenik.json, conference.json, dragon-sponza.json).
"""
import numpy as np

from . import abi
from .scene import ScenePack, RenderConfig, F


# ---------------------------------------------------------------- procedural textures
def _value_noise(rng, size, cells):
    g = rng.random((cells + 1, cells + 1)).astype(np.float64)
    g[-1, :] = g[0, :]
    g[:, -1] = g[:, 0]
    t = np.linspace(0, cells, size, endpoint=False)
    i = t.astype(int)
    f = t - i
    f = f * f * (3 - 2 * f)
    a = g[np.ix_(i, i)]
    b = g[np.ix_(i, i + 1)]
    c = g[np.ix_(i + 1, i)]
    d = g[np.ix_(i + 1, i + 1)]
    fx, fy = f[None, :], f[:, None]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def _fbm(rng, size, octaves=4, base=4):
    out = np.zeros((size, size))
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        out += amp * _value_noise(rng, size, base << o)
        tot += amp
        amp *= 0.5
    return out / tot


def texture_set(seed, size=512):
    """Returns dict name -> (albedo[h,w,3], bump[h,w,3]) float32 in linear space."""
    rng = np.random.default_rng(seed)
    sets = {}
    n = _fbm(rng, size, 5, 4)
    stone = np.stack([0.55 + 0.25 * n, 0.50 + 0.24 * n, 0.42 + 0.22 * n], -1)
    sets["stone"] = (stone, np.repeat(n[..., None], 3, -1))
    yy, xx = np.mgrid[0:size, 0:size]
    row = (yy // (size // 16))
    bx = ((xx + (row % 2) * (size // 16)) % (size // 8)) < 3
    by = (yy % (size // 16)) < 3
    mortar = (bx | by)
    n2 = _fbm(rng, size, 4, 8)
    brick = np.where(mortar[..., None], np.array([0.6, 0.58, 0.55]), np.stack([0.45 + 0.2 * n2, 0.2 + 0.1 * n2, 0.13 + 0.08 * n2], -1))
    sets["brick"] = (brick, np.repeat(np.where(mortar, 0.2, 0.7 + 0.3 * n2)[..., None], 3, -1))
    n3 = _fbm(rng, size, 3, 16)
    stripes = 0.5 + 0.5 * np.sin(xx / size * 2 * np.pi * 24)
    fabric = np.stack([0.55 + 0.3 * stripes, 0.08 + 0.05 * n3, 0.08 + 0.05 * n3], -1)
    sets["fabric_red"] = (fabric, np.repeat((0.5 + 0.5 * np.sin(xx / size * 2 * np.pi * 96))[..., None], 3, -1))
    fabric_g = np.stack([0.08 + 0.05 * n3, 0.35 + 0.25 * stripes, 0.12 + 0.05 * n3], -1)
    sets["fabric_green"] = (fabric_g, sets["fabric_red"][1])
    n4 = _fbm(rng, size, 5, 2)
    tiles = (((xx // (size // 8)) + (yy // (size // 8))) % 2).astype(float)
    floor = np.stack([0.35 + 0.3 * tiles + 0.1 * n4, 0.33 + 0.28 * tiles + 0.1 * n4, 0.3 + 0.25 * tiles + 0.1 * n4], -1)
    groove = ((xx % (size // 8)) < 2) | ((yy % (size // 8)) < 2)
    sets["floor"] = (floor, np.repeat(np.where(groove, 0.0, 0.8 + 0.2 * n4)[..., None], 3, -1))
    n5 = _fbm(rng, size, 4, 6)
    sets["bronze"] = (np.stack([0.45 + 0.2 * n5, 0.3 + 0.15 * n5, 0.1 + 0.08 * n5], -1), np.repeat(n5[..., None], 3, -1))
    return {k: (np.ascontiguousarray(a, F), np.ascontiguousarray(b, F)) for k, (a, b) in sets.items()}


def envmap(seed=7, w=2048, h=1024):
    """Synthetic lat-long HDR sky (gradient + sun + soft clouds), stands in for envmap/cloudy1.hdr."""
    rng = np.random.default_rng(seed)
    v = np.linspace(0, 1, h)[:, None]
    u = np.linspace(0, 1, w)[None, :]
    clouds = np.resize(_fbm(rng, 1024, 5, 4), (h, w)) if False else _fbm(rng, 1024, 5, 4)[:h, :][:, np.arange(w) % 1024]
    up = np.clip((v - 0.5) * 2, 0, 1)
    sky = np.stack([0.25 + 0.35 * (1 - up), 0.45 + 0.3 * (1 - up), 0.85 - 0.1 * up], -1) * (0.6 + 0.8 * up[..., None] * 0 + 0.4)
    sky = sky * (0.7 + 0.6 * clouds[..., None])
    ground = np.array([0.18, 0.16, 0.14])
    img = np.where(v[..., None] > 0.5, sky, ground * (0.6 + 0.4 * clouds[..., None]))
    su, sv = 0.3, 0.78
    d2 = ((u - su) * 2) ** 2 + (v - sv) ** 2
    img = img + np.exp(-d2 / 0.0004)[..., None] * np.array([60.0, 52.0, 40.0]) + np.exp(-d2 / 0.01)[..., None] * 1.5
    return np.ascontiguousarray(img, F)


# ---------------------------------------------------------------- mesh pieces: (P, N, UV, T, I)
def _finish(P, N, UV, T, I):
    N = N / np.maximum(np.linalg.norm(N, axis=1, keepdims=True), 1e-20)
    T = T / np.maximum(np.linalg.norm(T, axis=1, keepdims=True), 1e-20)
    return (np.ascontiguousarray(P, F), np.ascontiguousarray(N, F), np.ascontiguousarray(UV, F), np.ascontiguousarray(T, F),
            np.ascontiguousarray(I, np.uint32))


def grid(origin, du, dv, nu, nv, uvscale=(1.0, 1.0), height=None):
    """Tessellated parallelogram origin + s*du + t*dv, s,t in [0,1]; optional displacement along the normal."""
    origin, du, dv = (np.asarray(x, np.float64) for x in (origin, du, dv))
    s, t = np.meshgrid(np.linspace(0, 1, nu + 1), np.linspace(0, 1, nv + 1), indexing="xy")
    s, t = s.ravel(), t.ravel()
    n = np.cross(du, dv)
    n = n / np.linalg.norm(n)
    P = origin + s[:, None] * du + t[:, None] * dv
    N = np.broadcast_to(n, P.shape).copy()
    if height is not None:
        hgt, dhs, dht = height(s, t)
        P = P + hgt[:, None] * n
        N = N - dhs[:, None] * du / np.dot(du, du) - dht[:, None] * dv / np.dot(dv, dv)
    UV = np.stack([s * uvscale[0], t * uvscale[1]], 1)
    T = np.broadcast_to(du / np.linalg.norm(du), P.shape).copy()
    idx = np.arange((nu + 1) * (nv + 1)).reshape(nv + 1, nu + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    I = np.concatenate([np.stack([a, b, c], 1), np.stack([b, d, c], 1)])
    return _finish(P, N, UV, T, I)


def revolve(profile, center, segments, uvscale=(1.0, 1.0), flute=0.0, flutes=0):
    """Surface of revolution about +Y through `center`; profile = [(radius, y)], open ends."""
    profile = np.asarray(profile, np.float64)
    k = len(profile)
    ang = np.linspace(0, 2 * np.pi, segments + 1)
    r = profile[:, 0][:, None] * (1.0 + (flute * np.cos(ang * flutes))[None, :] if flutes else 1.0)
    x = r * np.cos(ang)[None, :]
    z = r * np.sin(ang)[None, :]
    y = np.broadcast_to(profile[:, 1][:, None], x.shape)
    P = np.stack([x, y, z], -1).reshape(-1, 3) + np.asarray(center, np.float64)
    dr = np.gradient(profile[:, 0])
    dy = np.gradient(profile[:, 1])
    nr, ny = dy, -dr                                       # outward normal of the profile curve
    N = np.stack([nr[:, None] * np.cos(ang)[None, :], np.broadcast_to(ny[:, None], x.shape), nr[:, None] * np.sin(ang)[None, :]], -1).reshape(-1, 3)
    UV = np.stack([np.broadcast_to(ang[None, :] / (2 * np.pi) * uvscale[0], x.shape),
                   np.broadcast_to((np.arange(k) / max(1, k - 1))[:, None] * uvscale[1], x.shape)], -1).reshape(-1, 2)
    T = np.stack([-np.sin(ang), np.zeros_like(ang), np.cos(ang)], -1)
    T = np.broadcast_to(T[None, :, :], (k, segments + 1, 3)).reshape(-1, 3).copy()
    idx = np.arange(k * (segments + 1)).reshape(k, segments + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    I = np.concatenate([np.stack([a, c, b], 1), np.stack([b, c, d], 1)])
    return _finish(P, N, UV, T, I)


def box(lo, hi, nseg=1, uvscale=1.0):
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    e = hi - lo
    faces = [
        (lo + [0, 0, e[2]], [e[0], 0, 0], [0, e[1], 0]), (lo + [e[0], 0, 0], [-e[0], 0, 0], [0, e[1], 0]),
        (lo + [e[0], 0, e[2]], [0, 0, -e[2]], [0, e[1], 0]), (lo + [0, 0, 0], [0, 0, e[2]], [0, e[1], 0]),
        (lo + [0, e[1], e[2]], [e[0], 0, 0], [0, 0, -e[2]]), (lo + [0, 0, 0], [e[0], 0, 0], [0, 0, e[2]]),
    ]
    return merge([grid(o, du, dv, nseg, nseg, (uvscale * np.linalg.norm(du), uvscale * np.linalg.norm(dv))) for o, du, dv in faces])


def arch(center, radius, thickness, depth, segments, axis="x"):
    """Half-ring (semicircular arch) standing in the plane spanned by `axis` and +Y, extruded by `depth`."""
    ang = np.linspace(0, np.pi, segments + 1)
    ax = np.array([1.0, 0, 0]) if axis == "x" else np.array([0, 0, 1.0])
    dp = np.array([0, 0, 1.0]) if axis == "x" else np.array([1.0, 0, 0])
    c = np.asarray(center, np.float64)
    pieces = []
    for (r, sign) in ((radius, -1.0), (radius + thickness, 1.0)):     # intrados / extrados
        ring = c + r * (np.cos(ang)[:, None] * ax + np.sin(ang)[:, None] * np.array([0, 1.0, 0]))
        P = np.concatenate([ring - dp * depth / 2, ring + dp * depth / 2])
        nrm = sign * (np.cos(ang)[:, None] * ax + np.sin(ang)[:, None] * np.array([0, 1.0, 0]))
        N = np.concatenate([nrm, nrm])
        UV = np.concatenate([np.stack([ang / np.pi * 4, np.zeros_like(ang)], 1), np.stack([ang / np.pi * 4, np.ones_like(ang)], 1)])
        T = np.broadcast_to(dp, P.shape).copy()
        i0 = np.arange(segments); k = segments + 1
        tri = np.concatenate([np.stack([i0, i0 + 1, i0 + k], 1), np.stack([i0 + 1, i0 + k + 1, i0 + k], 1)])
        if sign < 0:
            tri = tri[:, ::-1]
        pieces.append(_finish(P, N, UV, T, tri))
    for side in (-1.0, 1.0):                                          # front / back faces
        inner = c + radius * (np.cos(ang)[:, None] * ax + np.sin(ang)[:, None] * np.array([0, 1.0, 0])) + side * dp * depth / 2
        outer = c + (radius + thickness) * (np.cos(ang)[:, None] * ax + np.sin(ang)[:, None] * np.array([0, 1.0, 0])) + side * dp * depth / 2
        P = np.concatenate([inner, outer])
        N = np.broadcast_to(side * dp, P.shape).copy()
        UV = np.concatenate([np.stack([ang / np.pi * 4, np.zeros_like(ang)], 1), np.stack([ang / np.pi * 4, np.full_like(ang, 0.3)], 1)])
        T = np.broadcast_to(ax, P.shape).copy()
        i0 = np.arange(segments); k = segments + 1
        tri = np.concatenate([np.stack([i0, i0 + k, i0 + 1], 1), np.stack([i0 + 1, i0 + k, i0 + k + 1], 1)])
        if side < 0:
            tri = tri[:, ::-1]
        pieces.append(_finish(P, N, UV, T, tri))
    return merge(pieces)


def merge(pieces):
    P, N, UV, T, I, off = [], [], [], [], [], 0
    for p in pieces:
        P.append(p[0]); N.append(p[1]); UV.append(p[2]); T.append(p[3]); I.append(p[4] + np.uint32(off)); off += len(p[0])
    return (np.concatenate(P), np.concatenate(N), np.concatenate(UV), np.concatenate(T), np.concatenate(I))


def _drop_degenerate(piece):
    P, N, UV, T, I = piece
    a, b, c = P[I[:, 0]].astype(np.float64), P[I[:, 1]].astype(np.float64), P[I[:, 2]].astype(np.float64)
    area = np.linalg.norm(np.cross(b - a, c - a), axis=1)
    return P, N, UV, T, np.ascontiguousarray(I[area > 1e-10])


# ---------------------------------------------------------------- the atrium
def atrium(pack, detail, seed, hall=(28.0, 12.0, 13.0), with_ceiling=False, extra_objects=0):
    """Adds a colonnaded hall to `pack`: interior x in [-L/2,L/2], z in [-W/2,W/2], y in [0,H].
    `detail` >= 1 scales tessellation.  Returns the triangle count added."""
    L, W, H = hall
    rng = np.random.default_rng(seed)
    tex = texture_set(seed)
    tid = {}
    for name, (alb, bump) in tex.items():
        tid[name] = (pack.add_image_texture(alb), pack.add_image_texture(bump))
    spec = pack.add_solid_texture((0.04, 0.04, 0.04))
    spec_hi = pack.add_solid_texture((0.35, 0.3, 0.2))

    def mat(name, t, exponent, specular=spec, bump=True):
        rough = float(np.power(F(2.0) / (F(2.0) + F(exponent)), F(0.5), dtype=F))
        return pack.add_material(name, abi.BXDF_LTC_GGX_DIFFUSE, roughness=rough, tex_diffuse=tid[t][0], tex_color=specular,
                                 tex_bump=tid[t][1] if bump else -1)

    m_floor = mat("floor", "floor", 40.0)
    m_stone = mat("stone", "stone", 10.0)
    m_brick = mat("brick", "brick", 5.0)
    m_red = mat("fabric_red", "fabric_red", 2.0)
    m_green = mat("fabric_green", "fabric_green", 2.0)
    m_bronze = mat("bronze", "bronze", 80.0, specular=spec_hi)
    d = max(1, int(detail))
    count0 = pack.n_triangles
    add = lambda piece, m: pack.add_mesh(*_drop_degenerate(piece), m)

    # shell
    add(grid((-L / 2, 0, W / 2), (L, 0, 0), (0, 0, -W), 14 * d, 6 * d, (L / 2, W / 2)), m_floor)
    add(grid((-L / 2, 0, -W / 2), (L, 0, 0), (0, H, 0), 14 * d, 6 * d, (L / 3, H / 3)), m_brick)
    add(grid((L / 2, 0, W / 2), (-L, 0, 0), (0, H, 0), 14 * d, 6 * d, (L / 3, H / 3)), m_brick)
    add(grid((-L / 2, 0, W / 2), (0, 0, -W), (0, H, 0), 6 * d, 6 * d, (W / 3, H / 3)), m_brick)
    add(grid((L / 2, 0, -W / 2), (0, 0, W), (0, H, 0), 6 * d, 6 * d, (W / 3, H / 3)), m_brick)
    if with_ceiling:
        add(grid((-L / 2, H, -W / 2), (L, 0, 0), (0, 0, W), 14 * d, 6 * d, (L / 3, W / 3)), m_stone)
    else:   # cornice ring: the roof stays open to the sky / sun like Sponza's courtyard
        for z0, z1 in ((-W / 2, -W / 2 + 2.5), (W / 2 - 2.5, W / 2)):
            add(grid((-L / 2, H, z0), (L, 0, 0), (0, 0, z1 - z0), 14 * d, 2 * d, (L / 3, 1)), m_stone)

    # two storeys of colonnades along both long sides, galleries behind them
    ncol = 10
    zc = W / 2 - 2.5
    storey = [(0.0, 5.0), (5.6, 4.2)]
    seg = 12 * d
    for side in (-1.0, 1.0):
        for (y0, hgt) in storey:
            xs = np.linspace(-L / 2 + 1.6, L / 2 - 1.6, ncol)
            for x in xs:
                prof = [(0.46, y0), (0.46, y0 + 0.25), (0.34, y0 + 0.35)]
                prof += [(0.34 - 0.04 * t, y0 + 0.35 + t * (hgt - 1.0)) for t in np.linspace(0, 1, 2 + 2 * d)[1:]]
                prof += [(0.42, y0 + hgt - 0.5), (0.5, y0 + hgt - 0.3), (0.5, y0 + hgt)]
                add(revolve(prof, (x, 0, side * zc), seg, (2, 4), flute=0.03, flutes=12), m_stone)
            span = xs[1] - xs[0]
            for x in (xs[:-1] + xs[1:]) / 2:
                add(arch((x, y0 + hgt, side * zc), span / 2 - 0.5, 0.5, 0.8, 8 * d, "x"), m_stone)
            # gallery floor slab above this storey
            add(box((-L / 2, y0 + hgt + span / 2 + 0.0, side * zc - 0.4 if side > 0 else -W / 2),
                    (L / 2, y0 + hgt + span / 2 + 0.6, W / 2 if side > 0 else side * zc + 0.4), 2 * d, 0.3), m_stone)

    # draped curtains hanging between upper columns
    ncur = 8
    for i in range(ncur):
        x0 = -L / 2 + 2.5 + i * (L - 5.0) / ncur
        side = -1.0 if i % 2 else 1.0
        ph, amp = rng.uniform(0, 6.28), rng.uniform(0.12, 0.22)
        waves = rng.integers(5, 9)

        def h(s, t, ph=ph, amp=amp, waves=waves):
            a = amp * (0.3 + 0.7 * t)
            arg = s * waves * 2 * np.pi + ph
            return a * np.sin(arg), a * np.cos(arg) * waves * 2 * np.pi, 0.7 * amp * np.sin(arg)
        n = 22 * d
        add(grid((x0, 9.6, side * (zc - 0.6)), (2.6, 0, 0), (0, -3.6, 0), n, n, (2, 2), h), m_red if i % 3 else m_green)

    # urns on the floor and the galleries
    nurn = 10 + extra_objects
    for i in range(nurn):
        x = rng.uniform(-L / 2 + 2, L / 2 - 2)
        z = rng.uniform(-zc + 1.2, zc - 1.2)
        s = rng.uniform(0.5, 0.9)
        ys = np.linspace(0, 1, 10 * d + 2)
        prof = [(s * (0.25 + 0.45 * np.sin(np.pi * (0.15 + 0.8 * t)) ** 2 * (1 - 0.5 * t)), s * 1.6 * t) for t in ys]
        add(revolve([(0.01, 0.0)] + prof + [(0.01, s * 1.6)], (x, 0, z), 14 * d, (2, 2)), m_bronze)
    return pack.n_triangles - count0


def _camera_cfg(cfg, position, lookat, focal=None, fov=None, lens=0.0, focus=1.0):
    cfg.camera = {"position": list(position), "lookat": list(lookat)}
    if focal is not None:
        cfg.camera["focal"] = focal
    if fov is not None:
        cfg.camera["fov"] = fov
    if lens:
        cfg.camera["lens-size"] = lens
        cfg.camera["focus-plane"] = focus


def _detail_for(target, seed, **kw):
    """Smallest integer tessellation level whose triangle count reaches `target`."""
    for d in range(1, 64):
        p = ScenePack()
        n = atrium(p, d, seed, **kw)
        if n >= target:
            return d, n
    raise RuntimeError("target too large")


def sponza(width=1920, height=1080, multisample=64, target_tris=66000):
    """BASELINE configs[1]: scenes/sponza.json (sun point light, constant sky, recursion-max 2, bumpscale 10,
    focal 1.6 camera at (-9.5,1.5,-1.5) -> (3,3,-0.5)) on the ~66 k-triangle atrium stand-in."""
    d, _ = _detail_for(target_tris, 1)
    pack = ScenePack()
    atrium(pack, d, 1)
    pack.add_point_light((-16.0, 100.0, -10.0), (255 / 255.0, 240 / 255.0, 200 / 255.0), 20000.0, 0.0)
    pack.set_sky_color((145 / 255.0, 200 / 255.0, 235 / 255.0), 0.3)
    cfg = RenderConfig()
    cfg.output_file, cfg.xres, cfg.yres = "sponza.exr", width, height
    cfg.recursion_level, cfg.multisample, cfg.rounds, cfg.bumpmap_scale = 2, multisample, 1, 10.0
    _camera_cfg(cfg, (-9.5, 1.5, -1.5), (3.0, 3.0, -0.5), focal=1.6)
    return pack, cfg


def sibenik(width=1920, height=1080, multisample=256, target_tris=75000):
    """BASELINE configs[2]: scenes/sibenik.json (thin-lens camera, sphere light size 0.8, russian 0.8, clamp 0.4,
    bumpscale 10) + envmap sky, on a closed ~75 k-triangle hall with windows (open cornice) stand-in."""
    d, _ = _detail_for(target_tris, 2, hall=(30.0, 12.0, 15.0), extra_objects=8)
    pack = ScenePack()
    atrium(pack, d, 2, hall=(30.0, 12.0, 15.0), extra_objects=8)
    pack.add_point_light((5.0, 6.0, 0.0), (255 / 255.0, 250 / 255.0, 210 / 255.0), 0.8 * 40.0, 0.8)
    pack.set_sky_envmap(pack.add_image_texture(envmap(7)), 1.0, 0.0)
    cfg = RenderConfig()
    cfg.output_file, cfg.xres, cfg.yres = "sibenik.exr", width, height
    cfg.multisample, cfg.rounds, cfg.bumpmap_scale, cfg.russian, cfg.clamp = multisample, 1, 10.0, 0.8, 0.4
    _camera_cfg(cfg, (-12.0, 2.0, 2.0), (0.0, 4.0, 0.0), focal=1.2, lens=0.035, focus=14.0)
    return pack, cfg


def conference(width=3840, height=2160, multisample=1024, target_tris=331000):
    """BASELINE configs[3]: scenes/conference.json parameters (8 sphere lights, recursion-max 4, russian 0.7,
    clamp 5) on a closed ~331 k-triangle hall stand-in."""
    d, _ = _detail_for(target_tris, 3, with_ceiling=True, extra_objects=30)
    pack = ScenePack()
    atrium(pack, d, 3, with_ceiling=True, extra_objects=30)
    for i in range(8):
        pack.add_point_light((-10.5 + 3.0 * i, 11.5, 1.5 if i % 2 else -1.5), (1.0, 0.95, 0.85), 14.0, 0.35)
    pack.set_sky_color((0.0, 0.0, 0.0), 1.0)
    cfg = RenderConfig()
    cfg.output_file, cfg.xres, cfg.yres = "conference.exr", width, height
    cfg.recursion_level, cfg.multisample, cfg.rounds, cfg.russian, cfg.clamp = 4, multisample, 1, 0.7, 5.0
    _camera_cfg(cfg, (-11.0, 3.0, 2.0), (4.0, 2.5, -1.0), fov=70.0)
    return pack, cfg


def dragon_sponza(width=3840, height=2160, multisample=512, target_tris=2000000):
    """BASELINE configs[4]: multi-million-triangle traversal stress (sponza.json lighting, recursion-max 40)."""
    d, _ = _detail_for(target_tris, 4, extra_objects=40)
    pack = ScenePack()
    atrium(pack, d, 4, extra_objects=40)
    pack.add_point_light((-16.0, 100.0, -10.0), (1.0, 240 / 255.0, 200 / 255.0), 20000.0, 0.0)
    pack.set_sky_color((145 / 255.0, 200 / 255.0, 235 / 255.0), 0.3)
    cfg = RenderConfig()
    cfg.output_file, cfg.xres, cfg.yres = "dragon-sponza.exr", width, height
    cfg.multisample, cfg.rounds, cfg.bumpmap_scale = multisample, 1, 10.0
    _camera_cfg(cfg, (-9.5, 1.5, -1.5), (3.0, 3.0, -0.5), focal=1.6)
    return pack, cfg


BUILDERS = {"sponza": sponza, "sibenik": sibenik, "conference": conference, "dragon-sponza": dragon_sponza}
