"""Built-in scene configs (the reference's JSON schema, as Python dicts) and the seeded
stand-in scenes used where the reference's assets are not distributed.

`cornell_box()` is BASELINE configs[0]: the Cornell box of scenes/cornell-box.json (five
walls, two boxes, a two-triangle areal light, four diffuse materials) with the BASELINE
overrides 256x256, multisample 16, recursion-max 40.
"""
import numpy as np

from . import abi
from .scene import ScenePack, load_config, F


def cornell_box(width=256, height=256, multisample=16, recursion_max=40, rounds=1):
    wall = lambda axis, translate, material, rotate=(0, 0, 0): {
        "primitive": "plane", "axis": axis, "translate": list(translate), "rotate": list(rotate), "material": material}
    light = lambda rotate: {"primitive": "tri", "translate": [-0.005, 1.98, -0.03], "scale": [0.235, -1.0, 0.19],
                            "rotate": list(rotate), "material": "Light"}
    return {
        "output-file": "cornell-box.exr", "output-width": width, "output-height": height,
        "multisample": multisample, "rounds": rounds, "russian": 0.74, "recursion-max": recursion_max, "clamp": 20.0,
        "camera": {"position": [0, 1, 6.8], "lookat": [0, 1, 0], "fov": 19.5},
        "materials": [
            {"name": "LeftWall", "diffuse": [0.63, 0.065, 0.05], "brdf": "diffuse"},
            {"name": "RightWall", "diffuse": [0.14, 0.45, 0.091], "brdf": "diffuse"},
            {"name": "WallsAndBoxes", "diffuse": [0.725, 0.71, 0.68], "brdf": "diffuse"},
            {"name": "Light", "emission": [17, 12, 4], "brdf": "diffuse"},
        ],
        "scene": [
            wall("Z", (0, 1, -1), "WallsAndBoxes"),                       # back
            wall("Y", (0, 0, 0), "WallsAndBoxes"),                        # floor
            wall("Y", (0, 2, 0), "WallsAndBoxes", rotate=(180, 0, 0)),    # ceiling
            wall("X", (1, 1, 0), "RightWall"),
            wall("X", (-1, 1, 0), "LeftWall", rotate=(0, 180, 0)),
            {"primitive": "cube", "translate": [-0.335439, 0.6, -0.291415], "scale": [0.607289, 0.597739, 1.2],
             "rotate": [90, 180, -160.812], "material": "WallsAndBoxes"},  # tall box
            {"primitive": "cube", "translate": [0.328631, 0.3, 0.374592], "scale": [0.594811, 0.604394, 0.6],
             "rotate": [0, -163.36, 0], "material": "WallsAndBoxes"},      # short box
            light((0, 180, 0)), light((0, 0, 0)),
        ],
    }


def load_builtin(name, **kw):
    if name == "cornell-box":
        return load_config(cornell_box(**kw))
    raise KeyError(name)


def material_zoo(width=96, height=64, multisample=4, recursion_max=6, lens=0.0, envmap_sky=True):
    """A small test scene that exercises every live BxDF (src/bxdf/bxdf.cpp:63-84), image textures with bump
    maps, a thin-lens camera, a sphere point light, an areal light and a lat-long envmap sky: the coverage the
    Cornell box (diffuse only) lacks.  Deterministic; built directly on a ScenePack."""
    from .scene import RenderConfig, primitive_data, transform_primitive, object_transform
    from . import standin
    rng = np.random.default_rng(11)
    pack = ScenePack()
    tex = standin.texture_set(5, size=64)
    t_stone, t_stone_b = pack.add_image_texture(tex["stone"][0]), pack.add_image_texture(tex["stone"][1])
    t_brick, t_brick_b = pack.add_image_texture(tex["brick"][0]), pack.add_image_texture(tex["brick"][1])
    t_floor = pack.add_image_texture(tex["floor"][0])
    grey = pack.add_solid_texture((0.6, 0.6, 0.6))
    white = pack.add_solid_texture((1.0, 1.0, 1.0))
    spec = pack.add_solid_texture((0.25, 0.22, 0.2))
    black = pack.add_solid_texture((0.0, 0.0, 0.0))
    tint = pack.add_solid_texture((0.9, 0.95, 1.0))
    pack.add_material("floor", abi.BXDF_LTC_GGX_DIFFUSE, roughness=0.35, tex_diffuse=t_floor, tex_color=spec, tex_bump=t_stone_b)
    pack.add_material("wall", abi.BXDF_DIFFUSE, tex_diffuse=t_brick, tex_bump=t_brick_b)
    pack.add_material("diffuse", abi.BXDF_DIFFUSE, tex_diffuse=grey)
    pack.add_material("mirror", abi.BXDF_MIRROR, tex_color=tint)
    pack.add_material("glass", abi.BXDF_DIELECTRIC, ior=1.5, tex_color=white)
    pack.add_material("clear", abi.BXDF_TRANSPARENT)
    pack.add_material("ggx", abi.BXDF_LTC_GGX, roughness=0.2, tex_color=spec)
    pack.add_material("beckmann", abi.BXDF_LTC_BECKMANN, roughness=0.5, tex_color=t_stone)
    pack.add_material("beck_diff", abi.BXDF_LTC_BECKMANN_DIFFUSE, roughness=0.6, tex_diffuse=t_stone, tex_color=spec, tex_bump=t_stone_b)
    pack.add_material("ggx_diff", abi.BXDF_LTC_GGX_DIFFUSE, roughness=0.15, tex_diffuse=grey, tex_color=spec, no_russian=True)
    pack.add_material("mix", abi.BXDF_MIX, mix_a=pack.material_names["diffuse"], mix_b=pack.material_names["mirror"], amount=0.6)
    pack.add_material("mix2", abi.BXDF_MIX, mix_a=pack.material_names["mix"], mix_b=pack.material_names["ggx"], amount=0.3)
    pack.add_material("lamp", abi.BXDF_DIFFUSE, tex_diffuse=black, emission=(6.0, 5.0, 4.0))

    def prim(kind, material, scale=(1, 1, 1), rotate=(0, 0, 0), translate=(0, 0, 0), axis="Y", texscale=(1, 1, 1)):
        from .scene import _mat4_identity, _mat4_mul, _scale, _rotate
        base = _mat4_identity()
        if kind == "cube":
            base = _mat4_mul(_scale((0.5, 0.5, 0.5)), base)
        if axis == "X":
            base = _mat4_mul(_rotate(F(np.pi) / F(2.0), (0.0, 0.0, 1.0)), base)
        elif axis == "Z":
            base = _mat4_mul(_rotate(F(np.pi) / F(2.0), (1.0, 0.0, 0.0)), base)
        P, N, U, T = transform_primitive(primitive_data(kind), object_transform(scale, rotate, translate, base), texscale)
        pack.add_mesh(P, N, U, T, np.arange(len(P), dtype=np.uint32).reshape(-1, 3), material)

    prim("plane", "floor", scale=(6, 1, 4), texscale=(3, 2, 1))
    prim("plane", "wall", scale=(6, 1, 3), axis="Z", translate=(0, 3, -4), texscale=(3, 1.5, 1))
    prim("plane", "wall", scale=(4, 1, 3), axis="X", translate=(6, 3, 0), texscale=(2, 1.5, 1))
    names = ["diffuse", "mirror", "glass", "clear", "ggx", "beckmann", "beck_diff", "ggx_diff", "mix", "mix2"]
    for i, nme in enumerate(names):
        x = -4.5 + i * 1.0
        prim("cube", nme, scale=(0.7, 0.6 + 0.15 * (i % 3), 0.7), rotate=(0, 17.0 * i, 0), translate=(x, 0.35 + 0.075 * (i % 3), -1.0 + 0.8 * (i % 2)))
    prim("plane", "lamp", scale=(1.2, 1, 0.8), rotate=(180, 0, 0), translate=(-1.0, 4.5, 0.5))
    pack.add_point_light((3.0, 3.5, 2.5), (1.0, 0.9, 0.8), 14.0, 0.4)
    if envmap_sky:
        pack.set_sky_envmap(pack.add_image_texture(standin.envmap(3, 128, 64)), 0.6, 30.0)
    else:
        pack.set_sky_color((0.3, 0.4, 0.6), 0.8)
    cfg = RenderConfig()
    cfg.output_file, cfg.xres, cfg.yres = "zoo.exr", width, height
    cfg.multisample, cfg.recursion_level, cfg.rounds, cfg.clamp, cfg.russian, cfg.bumpmap_scale = multisample, recursion_max, 1, 30.0, 0.8, 4.0
    cfg.camera = {"position": [0.5, 2.6, 7.5], "lookat": [0.0, 0.6, -0.5], "fov": 55.0}
    if lens:
        cfg.camera["lens-size"] = lens
        cfg.camera["focus-plane"] = 8.0
    return pack, cfg
