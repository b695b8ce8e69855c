#!/usr/bin/env python
"""Regenerate tests/golden/weekend_mean_300x200_4096spp.npz (run here, on CPU; ~2-4 minutes of oracle time:
python tests/golden/make_golden_converged.py).

SURVEY.md §8c check 3 asks for converged renders at >= 4096 vs >= 4096 spp with the Monte-Carlo error estimated
from two independent oracle runs.  This is that fixture for the BASELINE config-2 scene at 300x200 (the 1200x800
frame's aspect and camera; 2350 warp tiles, ragged bottom/top rows of tiles included):

  mean_a, mean_b   the f64 oracle's linear (pre-gamma) mean images of two independent 4096-spp runs (float32)
  floor            RMSE(mean_a, mean_b) — the oracle-vs-oracle floor = sqrt(2) x the per-estimate standard error
  segments         path segments per primary sample of run a

tests/test_gpu_scale.py renders the same frame on the GPU at 4096 spp and requires
RMSE(gpu, mean_a) and RMSE(gpu, mean_b) <= 1.25 x floor (an unbiased 4096-spp estimate sits AT the floor)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import shirley_raytracing_rs_b200 as rt   # noqa: E402  (host side only: scene loader + camera)
from oracle import pyoracle as po          # noqa: E402

W, SPP = 300, 4096


def main():
    scene = rt.Scene.from_json(open(os.path.join(HERE, "weekend_scene.json")).read(), 0x5EED)
    cam = rt.default_camera(W)
    o = po.OracleScene(scene.desc)
    a, sa = o.render(cam, SPP, max_depth=50, seed=9001, threads=0)
    b, sb = o.render(cam, SPP, max_depth=50, seed=9002, threads=0)
    ma, mb = a / SPP, b / SPP
    floor = float(np.sqrt(np.mean((ma - mb) ** 2)))
    np.savez_compressed(os.path.join(HERE, "weekend_mean_300x200_4096spp.npz"), mean_a=ma.astype(np.float32), mean_b=mb.astype(np.float32),
                        floor=np.float64(floor), segments=np.float64(sa.rays / sa.paths), spp=np.int64(SPP))
    print(f"{cam.image_width}x{cam.image_height} x {SPP} spp twice: {sa.rays + sb.rays} rays in {sa.seconds + sb.seconds:.1f} s; "
          f"oracle-vs-oracle RMSE {floor:.5f}; segments/sample {sa.rays / sa.paths:.4f}")


if __name__ == "__main__":
    main()
