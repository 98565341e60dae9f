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
