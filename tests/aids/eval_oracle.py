"""Pins the CPU oracle against the reference's converged evaluation images (RelMSE rule of
scripts/RunEvaluations.py:83-92 in the reference). Usage: python tools/eval_oracle.py [spp] [scene ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tools"))
from make_golden import IMAGES  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ignis_b200.scene import load_scene  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

EPS = {"cbox-d1": 5e-3, "cbox-d6": 5e-3, "multilight-uniform": 3e-4}


def relmse(img, ref):
    mask = ref != 0
    err = np.zeros_like(ref)
    err[mask] = np.square((img[mask] - ref[mask]) / ref[mask])
    err[~mask] = np.square(img[~mask])
    mx = np.percentile(err, 99)
    return float(np.average(np.clip(err, 0, mx)))


def main():
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    refs = np.load(os.path.join(ROOT, "tests", "golden", "ref_images.npz"))
    names = sys.argv[2:] or [k for k in refs.files if k != "flipped-prim-glass"]
    for name in names:
        t = load_scene(os.path.join(ROOT, "scenes", "evaluation", name + ".json"))
        w, h = t.film_size
        o = Oracle(t)
        fb = np.zeros((h, w, 3), np.float32)
        spi = 8
        t0 = time.time()
        for it in range(spp // spi):
            o.render(w, h, spi=spi, iteration=it, fb=fb)
        img = fb / (spp // spi)
        ref = refs[IMAGES[name][:-4]].astype(np.float32)
        e = relmse(img, ref)
        eps = EPS.get(name, 1e-3)
        print(f"{name:24s} spp={spp} relmse={e:.3e} eps={eps:.0e} {'OK' if e < eps else 'FAIL'} mean={img.mean():.5f} ref_mean={ref.mean():.5f} "
              f"({time.time() - t0:.1f}s, {o.counters.sum() / (time.time() - t0) / 1e6:.1f} Mrays/s)", flush=True)


if __name__ == "__main__":
    main()
