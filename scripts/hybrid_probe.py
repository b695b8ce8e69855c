"""What one of N GPUs renders of the 1200x800x500 bench frame under hybrid partitions (T interleaved tile shards x N/T sample
ranges), timed on ONE GPU: kernel ms (CUDA events, best of 5) for every (T, shard) so the slowest rank of each plan is visible.
    python scripts/hybrid_probe.py [N ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shirley_raytracing_rs_b200 as rt

scene = rt.Scene.named("random", seed=0xDEADBEEF)
cam = rt.default_camera(1200)
rt.render(scene, cam, samples=8, seed=1)
total = 500
for N in [int(x) for x in sys.argv[1:]] or [8, 4, 2]:
    T = 1
    while T <= N:
        R = N // T                      # sample ranges
        base, rem = divmod(total, R)
        worst = 0.0; cells = []
        for spp in sorted({base + (1 if rem else 0), base}, reverse=True):
            for shard in range(min(T, 2)):
                best = 1e9
                for rep in range(5):
                    _, st = rt.render(scene, cam, samples=spp, seed=5, shard=(T, shard), sample_offset=0)
                    best = min(best, st.kernel_ms)
                cells.append(f"{spp} spp shard {shard}/{T}: {best:.2f} ms")
                worst = max(worst, best)
        print(f"N={N}: {T} tile shards x {R} sample ranges -> slowest rank {worst:.2f} ms   [" + "; ".join(cells) + "]", flush=True)
        T *= 2
