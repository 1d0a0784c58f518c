"""Copies the reference's evaluation fixtures this repo's tests need into tests/golden/ and scenes/.

Run once in the build container (where /root/reference is mounted); the outputs are committed so that tests and
bench.py never read /root/reference at run time (it does not exist on the GPU box).

 * tests/golden/ref_images.npz : the converged reference images of scenes/evaluation/references/*.exr that lie
   inside the supported path, stored as float16 RGB (256x256), keyed by scene name.
 * scenes/ : the scene description files + meshes of the benchmark / parity configurations (data, not source).
"""
import os
import shutil
import sys

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402
import numpy as np  # noqa: E402

REF = "/root/reference/scenes"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

IMAGES = {
    "plane-d1": "ref-plane-d1-4096.exr", "plane-d6": "ref-plane-d6-4096.exr", "point": "ref-point-4096.exr",
    "emissive-plane": "ref-emissive-plane-4096.exr", "cbox-d1": "ref-cbox-d1-4096.exr", "cbox-d6": "ref-cbox-d6-4096.exr",
    "multilight-uniform": "ref-multilight-4096.exr",
    # the same scene with the other two light selectors (flux CDF, light hierarchy): same expectation, same reference image
    "multilight-simple": "ref-multilight-4096.exr", "multilight-hierarchy": "ref-multilight-4096.exr",
    # analytic sphere area light. (The only reference image with a pure dielectric, ref-three-planes-dielectric-rad.exr, is a
    # Radiance rendering of a single glass interface; Ignis' own dielectric (bsdf/dielectric.art:15-37) applies no 1/eta^2
    # radiance scaling in the non-adjoint direction, so the reference's algorithm itself does not reproduce that image -- it is
    # not a pin for a restatement of that algorithm. The dielectric is pinned analytically instead: tests/test_oracle_kat.py.)
    "sphere-light-pure": "ref-sphere-light-4096.exr",
    # the same references rendered with other shape providers / without the plane optimisation: pins the shape (trimesh) emitter
    "sphere-light-ico": "ref-sphere-light-4096.exr", "sphere-light-uv": "ref-sphere-light-4096.exr", "sphere-light-ico-nopt": "ref-sphere-light-4096.exr",
    "emissive-plane-nopt": "ref-emissive-plane-4096.exr", "emissive-plane-scale": "ref-emissive-plane-scale-4096.exr",
    "emissive-plane-scale-nopt": "ref-emissive-plane-scale-4096.exr",
    # Radiance rendering of two diffuse planes lit by a tiny analytic sphere light (radius 0.01, radiance 10^4)
    "two-planes-base": "ref-two-planes-rad.exr",
    # OBJ room under a constant light. (Not used: ref-flipped-prim-diffuse-4096.exr -- a convex Lambertian cylinder of albedo 0.8
    # under a uniform environment of radiance 0.8 has radiance 0.64 wherever only the environment lights it; the oracle gives 0.635,
    # that image 0.496 = 0.8 * 0.8^2.2, i.e. it was rendered with the albedo taken as an sRGB value.)
    "room": "ref-room-4096.exr",
    # Radiance rendering of a diffuse plane under the sun (cone of 0.533 degrees, radiance 1e5): pins make_sun_light. (Not used:
    # ref-sun-on-plane-and-stick-rad.exr -- the Ignis scene file puts the sun on the horizon, direction (0.707, -0.707, 0) in a z-up
    # scene, and renders a plane at grazing incidence (mean 0.03); the Radiance image shows a fully lit plane (mean 0.24).)
    "sun-on-plane": "ref-sun-on-plane-rad.exr",
    # textured environment lights (light/env.art:112-167): a diffuse floor under an environment map, sampled through the 2-D cdf
    # ("conditional", the default) and uniformly ("none"); `env` is a 100 x 50 map with one bright pixel, nearest filter, scale 100.
    # The 4k map (93 MB) is stored box-filtered to 1024 x 512 -- the resolution the reference bakes it to for its cdf anyway.
    "env4k-conditional": "ref-env4k-4096.exr", "env4k-none": "ref-env4k-4096.exr", "env": "ref-env-4096.exr",
}
SCENES = ["single_triangle.json", "diamond_scene.json", "primitives.json", "primitives_data.json", "flipped_prim.json",
          "many_point_lights.json", "meshes/Pillar.ply", "textures/bumpmap.png", "textures/environment/single_bright_pixel.png",
          "meshes/Room.obj", "meshes/Bottom.ply", "meshes/Top.ply", "meshes/Left.ply", "meshes/Right.ply", "meshes/Back.ply", "meshes/Diamond.ply"]
EVAL = ["plane-base.json", "plane-d1.json", "plane-d6.json", "point.json", "emissive-plane.json", "cbox-base.json", "cbox-d1.json",
        "cbox-d6.json", "multilight.json", "multilight-uniform.json", "multilight-simple.json", "multilight-hierarchy.json", "flipped-prim-base.json", "flipped-prim-diffuse.json",
        "sphere-light-base.json", "sphere-light-pure.json", "sphere-light-ico.json", "sphere-light-uv.json", "sphere-light-ico-nopt.json",
        "emissive-plane-nopt.json", "emissive-plane-scale.json", "emissive-plane-scale-nopt.json", "two-planes-base.json", "two-planes-mirror.json", "room.json", "sun-on-plane.json",
        "env4k-base.json", "env4k-conditional.json", "env4k-none.json", "env.json"]


def main():
    out = {}
    for fn in sorted(set(IMAGES.values())):   # one array per reference file, keyed by its name; tests map scene -> file with IMAGES
        img = cv2.imread(os.path.join(REF, "evaluation", "references", fn), cv2.IMREAD_UNCHANGED)
        if img is None:
            print("skip", fn)
            continue
        out[fn[:-4]] = img[..., 2::-1][..., :3].astype(np.float16) if img.shape[-1] >= 3 else img.astype(np.float16)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_images.npz"), **out)
    for f in SCENES:
        dst = os.path.join(ROOT, "scenes", f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(os.path.join(REF, f)):
            shutil.copyfile(os.path.join(REF, f), dst)
    for f in EVAL:
        dst = os.path.join(ROOT, "scenes", "evaluation", f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(os.path.join(REF, "evaluation", f)):
            shutil.copyfile(os.path.join(REF, "evaluation", f), dst)
    mdir = os.path.join(REF, "evaluation", "meshes")
    for f in os.listdir(mdir):
        if True:
            os.makedirs(os.path.join(ROOT, "scenes", "evaluation", "meshes"), exist_ok=True)
            shutil.copyfile(os.path.join(mdir, f), os.path.join(ROOT, "scenes", "evaluation", "meshes", f))
    # the 4k environment map, box-filtered 4 x 4 -> 1024 x 512 float16 (rows top-down as in the file); the scene file is pointed at it
    env = cv2.imread(os.path.join(REF, "textures", "environment", "phalzer_forest_01_4k.exr"), cv2.IMREAD_UNCHANGED)
    if env is not None:
        rgb = env[..., 2::-1].astype(np.float64)
        k = rgb.shape[1] // 1024
        small = rgb.reshape(rgb.shape[0] // k, k, rgb.shape[1] // k, k, 3).mean(axis=(1, 3))
        assert small.max() < 60000
        os.makedirs(os.path.join(ROOT, "scenes", "textures", "environment"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "scenes", "textures", "environment", "phalzer_forest_01_1k.npz"), rgb=small.astype(np.float16))
        base = os.path.join(ROOT, "scenes", "evaluation", "env4k-base.json")
        txt = open(base).read().replace("phalzer_forest_01_4k.exr", "phalzer_forest_01_1k.npz")
        open(base, "w").write(txt)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
