"""A/B runs of one launch-time environment variable on a named workload, one process per library build:
    [B200RT_LIB=build/variants/libX.so] python scripts/ab.py weekend|c4:G|earth VAR v1 v2 ... [-- spp [width]]
Prints Mrays/s (scene resident, CUDA-event kernel time, best of 3) and a hash of the accumulation buffer for each value:
images are bit-deterministic, so every variant of the work distribution must print the same hash."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import shirley_raytracing_rs_b200 as rt

args = sys.argv[1:]
tail = []
if "--" in args:
    k = args.index("--"); args, tail = args[:k], args[k + 1:]
what, var, values = args[0], args[1], args[2:]
if what.startswith("c4"):
    G = int(what.split(":")[1]) if ":" in what else 500
    scene = rt.Scene.named("scaled", seed=3, param=G)
    spp = int(tail[0]) if tail else 64
    cam = rt.camera((0.9 * G, 0.18 * G + 2, 0.35 * G), (0, 0, 0), vfov=30, aperture=0.001, width=int(tail[1]) if len(tail) > 1 else 3840,
                    aspect_ratio=(16, 9), focus_length=10.0)
elif what == "earth":
    scene = rt.Scene.named("earth"); spp = int(tail[0]) if tail else 256
    cam = rt.default_camera(int(tail[1]) if len(tail) > 1 else 1920, aspect_ratio=(16, 9))
else:
    scene = rt.Scene.named("random", seed=0xDEADBEEF); spp = int(tail[0]) if tail else 500
    cam = rt.default_camera(int(tail[1]) if len(tail) > 1 else 1200)
rt.render(scene, cam, samples=2, seed=1)
for v in values:
    os.environ[var] = "" if v == "-" else v
    best, h = 0.0, None
    for rep in range(3):
        acc, st = rt.render(scene, cam, samples=spp, seed=5)
        best = max(best, st.rays / st.kernel_ms / 1e3)
        if h is None:
            h = hashlib.sha1(np.ascontiguousarray(acc).tobytes()).hexdigest()[:12]
    print(f"{what} lib={os.path.basename(os.environ.get('B200RT_LIB', 'default'))} {var}={v}: {cam.image_width}x{cam.image_height} {spp} spp  "
          f"best of 3: {best:.0f} Mrays/s ({st.kernel_ms:.2f} ms)  image {h}", flush=True)
