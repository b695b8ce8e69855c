"""A/B sweeps of the render kernel's tunables on the bench scene (scene resident, CUDA-event kernel time):
    python scripts/sweep_env.py VAR v1 v2 ... [-- width spp]
prints Mrays/s for each value of the environment variable VAR (read by launch_render per launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shirley_raytracing_rs_b200 as rt

args = sys.argv[1:]
tail = []
if "--" in args:
    k = args.index("--"); args, tail = args[:k], args[k + 1:]
var, values = args[0], args[1:]
width = int(tail[0]) if tail else 1200
spp = int(tail[1]) if len(tail) > 1 else 500
scene = rt.Scene.named("random", seed=0xDEADBEEF)
cam = rt.default_camera(width)
rt.render(scene, cam, samples=8, seed=1)
for v in values:
    os.environ[var] = v
    best = 0.0
    for rep in range(3):
        _, st = rt.render(scene, cam, samples=spp, seed=5 + rep)
        best = max(best, st.rays / st.kernel_ms / 1e3)
    print(f"{var}={v}: {cam.image_width}x{cam.image_height} {spp} spp  best of 3: {best:.0f} Mrays/s  ({st.kernel_ms:.2f} ms, launches {st.launches})", flush=True)
