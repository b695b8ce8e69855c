import sys, time, ctypes as C, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
if mode in ("torch", "torchflush"):
    import torch
    torch.cuda.set_device(0); x = torch.zeros(10, device="cuda")
import shirley_raytracing_rs_b200 as rt
F, lib = rt._ffi, rt._ffi.lib
scene = rt.Scene.named("random", seed=0xDEADBEEF)
cam = rt.default_camera(1200)
H, W = cam.image_height, cam.image_width
rgb = np.empty((H, W, 3), dtype=np.uint8)
if mode == "smi":
    pr = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.DEVNULL)
    time.sleep(1.0); pr.terminate(); pr.wait()
if mode == "torchflush":
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.fill_(1); torch.cuda.synchronize()
for it in range(5):
    t0 = time.perf_counter()
    h = C.c_void_p(); F.check(lib.b200rt_scene_create(scene.desc, 0, C.byref(h)))
    t1 = time.perf_counter()
    st = F.Stats(); p = F.RenderParams(samples=500, max_depth=50, seed=it, device=-1)
    F.check(lib.b200rt_render_rgb8(h, C.byref(cam), C.byref(p), rgb.ctypes.data, None, C.byref(st)))
    t2 = time.perf_counter()
    lib.b200rt_scene_destroy(h)
    t3 = time.perf_counter()
    print(f"{mode}: create {1e3*(t1-t0):.2f} ms  render_rgb8 {1e3*(t2-t1):.2f} ms (kernel {st.kernel_ms:.2f}, total_ms {st.total_ms:.2f})  destroy {1e3*(t3-t2):.2f} ms")
