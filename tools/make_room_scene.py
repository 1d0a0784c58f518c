"""Writes scenes/synthetic_room.json: the stand-in for BASELINE.json configs[3] (the Bitterli bedroom cannot be obtained
offline). A closed diffuse room lit by an area light and a constant environment seen through an opening, filled
with a grid of finely tessellated icospheres (diffuse and dielectric) -- about a million instanced triangles from four unique
meshes, so that BVH nodes and triangles live in HBM/L2 instead of shared memory."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6          # n x n spheres
sub = int(sys.argv[2]) if len(sys.argv) > 2 else 6         # icosphere subdivisions: 20 * 4^sub triangles
scene = {
    "technique": {"type": "path", "max_depth": 8},
    "camera": {"type": "perspective", "fov": 60, "near_clip": 0.01, "far_clip": 100, "transform": {"lookat": {"origin": [0, -9.5, 4.0], "target": [0, 0, 0.6], "up": [0, 0, 1]}}},
    "film": {"size": [1920, 1080]},
    "bsdfs": [{"type": "diffuse", "name": "white", "reflectance": [0.75, 0.75, 0.75]}, {"type": "diffuse", "name": "red", "reflectance": [0.75, 0.2, 0.2]},
              {"type": "diffuse", "name": "blue", "reflectance": [0.2, 0.3, 0.75]}, {"type": "dielectric", "name": "glass", "int_ior": 1.5, "ext_ior": 1.0},
              {"type": "diffuse", "name": "black", "reflectance": [0, 0, 0]}],
    "shapes": [{"type": "rectangle", "name": "Floor", "width": 12, "height": 12}, {"type": "rectangle", "name": "Lamp", "width": 3, "height": 3},
               {"type": "cube", "name": "Shell", "width": 12, "height": 12, "depth": 6}]
              + [{"type": "icosphere", "name": f"Ico{k}", "radius": 0.35 + 0.05 * k, "subdivisions": sub - (k % 2)} for k in range(4)],
    "entities": [{"name": "Floor", "shape": "Floor", "bsdf": "white", "transform": {"translate": [0, 0, 0]}},
                 {"name": "Lamp", "shape": "Lamp", "bsdf": "black", "transform": [{"translate": [0, 0, 5.0]}, {"rotate": [180, 0, 0]}]}],
    "lights": [{"type": "area", "name": "LampLight", "entity": "Lamp", "radiance": [18, 17, 15]}, {"type": "env", "name": "Sky", "radiance": [0.4, 0.5, 0.7]}],
}
k = 0
for i in range(n):
    for j in range(n):
        x, y = (i - (n - 1) / 2) * 1.6, (j - (n - 1) / 2) * 1.6
        mat = ["white", "red", "blue", "glass"][(i * 7 + j * 3) % 4]
        scene["entities"].append({"name": f"S{k}", "shape": f"Ico{k % 4}", "bsdf": mat, "transform": [{"translate": [x, y, 0.5 + 0.15 * ((i + j) % 3)]}, {"rotate": [10 * i, 15 * j, 0]}]})
        k += 1
out = os.path.join(ROOT, "scenes", "synthetic_room.json")
json.dump(scene, open(out, "w"), indent=1)
print(out, "entities", len(scene["entities"]))
