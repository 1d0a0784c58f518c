import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENES = os.path.join(ROOT, "scenes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def scenes_dir():
    return SCENES


def flat_scene():
    """create_flat_scene() of the reference's integrator tests (src/tests/integrator/common/__init__.py:38-66)."""
    return {
        "technique": {"type": "path", "max_depth": 2},
        "camera": {"type": "perspective", "fov": 90, "near_clip": 0.01, "far_clip": 100,
                   "transform": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, -1]},
        "film": {"size": [1000, 1000]},
        "bsdfs": [{"type": "diffuse", "name": "ground", "reflectance": [1, 1, 1]}],
        "shapes": [{"type": "rectangle", "name": "Bottom", "width": 2, "height": 2, "flip_normals": True}],
        "entities": [{"name": "Bottom", "shape": "Bottom", "bsdf": "ground"}],
        "lights": [],
    }


def furnace_scene():
    """White furnace: a non-absorbing glass box under a constant white environment; every pixel converges to 1."""
    return {
        "technique": {"type": "path", "max_depth": 64},
        "camera": {"type": "perspective", "fov": 40, "near_clip": 0.01, "far_clip": 100, "transform": {"lookat": {"origin": [2.2, -3.1, 1.7], "target": [0, 0, 0], "up": [0, 0, 1]}}},
        "film": {"size": [96, 96]},
        "bsdfs": [{"type": "dielectric", "name": "glass", "int_ior": 1.5, "ext_ior": 1.0}],
        "shapes": [{"type": "cube", "name": "Box", "width": 1.6, "height": 1.2, "depth": 1.0, "origin": [-0.8, -0.6, -0.5]}],
        "entities": [{"name": "Box", "shape": "Box", "bsdf": "glass", "transform": {"rotate": [15, 25, 35]}}],
        "lights": [{"type": "env", "name": "env", "radiance": [1, 1, 1]}],
    }
