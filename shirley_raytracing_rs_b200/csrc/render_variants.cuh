// render_variants.cuh — K2 design experiments kept selectable (B200RT_KERNEL=1|3): v1 straight per-lane loops, v3 warp-level wavefront over a path pool; numbers in profiles/README.md
// (included by b200rt.cu; everything lives in namespace b200rt)
#pragma once
#include "render_kernel.cuh"

namespace b200rt {

// ------------------------------------------------------------------------------------------
// K2: persistent path-tracing megakernel.
// Grid = (#SMs x resident CTAs); each warp pulls 8x4-pixel tiles from a global counter
// (bottom rows first: the geometry-heavy tiles are scheduled before the cheap sky tiles).
// One lane owns one pixel and walks its samples in order with path regeneration: a lane
// whose path ends starts its next sample at the top of the loop instead of idling until
// the warp's longest path finishes, so every traversal round has as many live lanes as the
// tile still has work for.  Per-pixel sums stay in registers and are written once
// (render.rs:59,66,68), as float4 {r, g, b, n}.
// ------------------------------------------------------------------------------------------
template <class Acc, bool COUNT>
__global__ void __launch_bounds__(BLOCK) path_trace_kernel(const __grid_constant__ RenderArgs a) {
    extern __shared__ float4 smem[];
    int* stack_base;
    Acc acc = Stager<Acc>::stage(a.scene, a.plan, smem, &stack_base);
    int* stack = stack_base + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const TopPrims top = top_of(a.scene);

    unsigned long long w_rays = 0, w_paths = 0, w_exh = 0, w_nodes = 0, w_prims = 0;

    for (;;) {
        unsigned int j = 0;
        if (lane == 0) j = atomicAdd(&a.counters->tile_counter, 1u);
        j = __shfl_sync(FULL, j, 0);
        unsigned long long t64 = (unsigned long long)j * a.shard_count + a.shard_index;
        if (t64 >= a.n_tiles) break;
        uint32_t t = (uint32_t)t64;
        uint32_t ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
        uint32_t px = tx * TILE_W + (lane & (TILE_W - 1));
        uint32_t py = (a.tile_row0 + ty) * TILE_H + (lane >> 3);
        bool valid = px < a.cam.width && py >= a.row_begin && py < a.row_end;
        uint32_t pix = py * a.cam.width + px;

        float3 sum = f3(0.f, 0.f, 0.f);
        uint32_t s = 0, nrays = 0, nexh = 0;
        TravCounters tc; tc.nodes = 0; tc.prims = 0;
        bool alive = false;
        Rng rng; rng.state = 0; rng.inc = 1;
        RayF ray = make_ray(f3(0, 0, 0), f3(0, 0, 1));
        float3 atten = f3(1, 1, 1), emit = f3(0, 0, 0);
        uint32_t depth = 0;

        for (;;) {
            if (!alive && valid && s < a.samples) {
                // render_scanline body, render.rs:60-66
                rng.init(a.keys, pix, a.sample_offset + s);
                float jx = (float)px + rng.gen();
                float jy = (float)py + rng.gen();
                float3 o, d;
                pixel_ray(a.cam, rng, jx, jy, &o, &d);
                ray = make_ray(o, d);
                atten = f3(1, 1, 1); emit = f3(0, 0, 0);
                depth = a.max_depth;
                alive = depth > 0;
                ++s;
            }
            if (!__any_sync(FULL, alive)) break;
            if (alive) {
                // one iteration of ray_color's loop, render.rs:30-46
                Closest c; c.t = INFINITY; c.code = -1; c.face = 0;
                closest_hit<COUNT>(ray, acc, top, stack, BLOCK, 0.001f, c, tc);
                ++nrays;
                bool done;
                if (c.code < 0) {
                    emit = emit + atten * background(a.scene, ray.d);
                    done = true;
                } else {
                    HitRec h = make_hit(ray, acc, c);
                    ShadeOut so = shade(a.scene, acc, ray, h, rng, atten, emit);
                    done = !so.scattered;
                    if (!done) {
                        ray = make_ray(so.o, so.d);
                        if (--depth == 0) { done = true; ++nexh; }
                    }
                }
                if (done) { sum = sum + emit; alive = false; }
            }
        }
        if (valid) {
            float4* dst = a.accum + pix;
            float4 v = make_float4(sum.x, sum.y, sum.z, (float)a.samples);
            if (a.accumulate) { float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            *dst = v;
        }
        w_rays += nrays; w_exh += nexh; w_paths += valid ? a.samples : 0;
        if (COUNT) { w_nodes += tc.nodes; w_prims += tc.prims; }
    }
    // one set of atomics per warp per kernel
    for (int o = 16; o > 0; o >>= 1) {
        w_rays += __shfl_down_sync(FULL, w_rays, o);
        w_paths += __shfl_down_sync(FULL, w_paths, o);
        w_exh += __shfl_down_sync(FULL, w_exh, o);
        if (COUNT) { w_nodes += __shfl_down_sync(FULL, w_nodes, o); w_prims += __shfl_down_sync(FULL, w_prims, o); }
    }
    if (lane == 0) {
        atomicAdd(&a.counters->rays, w_rays);
        atomicAdd(&a.counters->paths, w_paths);
        atomicAdd(&a.counters->exhausted, w_exh);
        if (COUNT) { atomicAdd(&a.counters->nodes, w_nodes); atomicAdd(&a.counters->prims, w_prims); }
    }
}

// ------------------------------------------------------------------------------------------
// K2 (v3): warp-level wavefront over a shared-memory path pool.
//
// Measured on v2 (scripts/diag.py): 26 lanes traverse per round but only 12.6 run each
// inner-node step (lanes differ in how many nodes they visit), 17 % of the lanes sit out of
// samples at the end of a tile, and shading batches hold ~21 lanes of mixed materials.  A
// lane-owns-a-pixel design can only trade traversal efficiency against shading efficiency, so
// here lanes are decoupled from paths:
//   * every warp owns a pool of P path slots in shared memory (SoA: ray, attenuation, RNG,
//     depth, hit) and the (pixel, sample) work list of its 8x4 tile;
//   * GENERATE, TRAVERSE and SHADE run as warp-wide batches over slots picked with
//     __ballot_sync/__popc compaction, so each phase runs on (up to) 32 live lanes;
//   * in TRAVERSE a lane that finishes stores its hit and immediately claims the next waiting
//     slot; when the queue is dry and only stragglers run, they are parked (cursor stays in
//     registers) while the warp shades / generates, and resume afterwards;
//   * per-pixel sums are 64-bit fixed point (2^-32) in shared memory: integer adds commute, so
//     the image stays bit-deterministic and independent of scheduling and tile sharding.
// ------------------------------------------------------------------------------------------
// Slot states.  A finished traversal parks its slot under the shading KIND it needs, so
// SHADE batches can be claimed one material at a time (no divergence inside a batch).
enum { SLOT_EMPTY = 0, SLOT_TRAV = 1, SLOT_RUN = 2, SLOT_SHADE0 = 3 };
enum { SK_MISS = 0, SK_METAL = 1, SK_DIELECTRIC = 2, SK_LAMBERT_SOLID = 3, SK_OTHER = 4, SK_COUNT = 5 };
constexpr int POOL_FIELDS = 15;
__host__ __device__ constexpr int pool_words(int P) { return 192 + 32 + POOL_FIELDS * P; }

template <int P> struct Pool {
    long long* acc;          // [32 pixels][3] fixed-point sums
    int* list;               // [32] compaction scratch
    float *ox, *oy, *oz, *dx, *dy, *dz, *ax, *ay, *az, *t;
    int* code;
    uint32_t *rs, *ri, *meta, *state;   // meta: pixel (5) | box face (3) << 5 | depth << 8
    __device__ __forceinline__ explicit Pool(uint32_t* base) {
        acc = reinterpret_cast<long long*>(base);
        list = reinterpret_cast<int*>(base + 192);
        float* f = reinterpret_cast<float*>(base + 224);
        ox = f; oy = f + P; oz = f + 2 * P; dx = f + 3 * P; dy = f + 4 * P; dz = f + 5 * P;
        ax = f + 6 * P; ay = f + 7 * P; az = f + 8 * P; t = f + 9 * P;
        code = reinterpret_cast<int*>(f + 10 * P);
        rs = reinterpret_cast<uint32_t*>(f + 11 * P); ri = rs + P; meta = rs + 2 * P; state = rs + 3 * P;
    }
};

// Compaction: the lanes flagged `want` receive distinct slots currently in `state_wanted`,
// lowest slot first; -1 when the pool has no more.  Returns how many such slots exist in *total.
// Out of line (it is called from every phase) and scalar in/out only, so the call costs no
// local-memory traffic: returns (slot & 0xffff) | (total << 16), slot 0xffff = none.
template <int P>
__device__ __noinline__ uint32_t claim_slots_packed(const uint32_t* state, int* list, uint32_t state_wanted, bool want) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int n = 0;
#pragma unroll
    for (int k = 0; k < P / 32; ++k) {
        bool is = state[k * 32 + lane] == state_wanted;
        unsigned m = __ballot_sync(FULL, is);
        int rank = n + __popc(m & lt);
        if (is && rank < 32) list[rank] = k * 32 + lane;
        n += __popc(m);
    }
    __syncwarp();
    unsigned wm = __ballot_sync(FULL, want);
    int j = __popc(wm & lt);
    uint32_t slot = (want && j < n) ? (uint32_t)list[j] : 0xffffu;
    __syncwarp();
    return slot | ((uint32_t)n << 16);
}
template <int P>
__device__ __forceinline__ int claim_slots(const Pool<P>& pool, uint32_t state_wanted, bool want, int lane, int* total) {
    uint32_t r = claim_slots_packed<P>(pool.state, pool.list, state_wanted, want);
    *total = (int)(r >> 16);
    uint32_t slot = r & 0xffffu;
    return slot == 0xffffu ? -1 : (int)slot;
}

// Store a freshly produced ray (camera ray or scattered ray) into its slot, after testing it
// against the scene-spanning primitives: that test runs here, in a full uniform batch.
template <bool COUNT, int P, class Acc>
__device__ __forceinline__ void produce_ray(const Pool<P>& pool, int slot, const Acc& acc, const TopPrims& top, float3 o, float3 d, float3 atten,
                                            const Rng& rng, uint32_t pixel, uint32_t depth, TravCounters& tc) {
    RayF r = make_ray_shade(o, d);
    Closest c; c.t = INFINITY; c.code = -1; c.face = 0;
    hit_top_prims<COUNT>(r, acc, top, 0.001f, c, tc);
    pool.ox[slot] = o.x; pool.oy[slot] = o.y; pool.oz[slot] = o.z;
    pool.dx[slot] = d.x; pool.dy[slot] = d.y; pool.dz[slot] = d.z;
    pool.ax[slot] = atten.x; pool.ay[slot] = atten.y; pool.az[slot] = atten.z;
    pool.rs[slot] = rng.state; pool.ri[slot] = rng.inc;
    pool.t[slot] = c.t; pool.code[slot] = c.code;
    pool.meta[slot] = pixel | ((uint32_t)c.face << 5) | (depth << 8);
    pool.state[slot] = SLOT_TRAV;
}

template <class Acc, bool COUNT, bool FAST, int BLK, int P>
__global__ void __launch_bounds__(BLK, 1) path_trace_kernel_v3(const __grid_constant__ RenderArgs a) {
    extern __shared__ float4 smem[];
    int* stack_base;
    Acc acc = Stager<Acc>::stage(a.scene, a.plan, smem, &stack_base);
    int* stack = stack_base + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = (1u << lane) - 1u;
    Pool<P> pool(reinterpret_cast<uint32_t*>(stack_base + a.plan.stack_depth * BLK) + warp * pool_words(P));
    const TopPrims top = top_of(a.scene);
    const float T_MIN = 0.001f;                      // render.rs:31
    const bool has_perlin = a.scene.perlin != nullptr;
    const int T_INNER = (int)a.wf_inner, F_FETCH = (int)a.wf_fetch, T_PARK = (int)a.wf_park;

    unsigned long long w_rays = 0, w_paths = 0, w_exh = 0, w_nodes = 0, w_prims = 0;
    // diagnostics (COUNT): d0 policy iterations, d1 shade batches, d2 kinds per shade batch,
    // d3 inner warp-steps, d4 lanes in inner steps, d5 lanes shaded, d6 lanes generated, d7 fetch batches
    unsigned long long d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0, d5 = 0, d6 = 0, d7 = 0;
    TravCounters tc; tc.nodes = 0; tc.prims = 0;

    for (;;) {
        unsigned int j = 0;
        if (lane == 0) j = atomicAdd(&a.counters->tile_counter, 1u);
        j = __shfl_sync(FULL, j, 0);
        unsigned long long t64 = (unsigned long long)j * a.shard_count + a.shard_index;
        if (t64 >= a.n_tiles) break;
        uint32_t tile = (uint32_t)t64;
        uint32_t ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
        const uint32_t px0 = tx * TILE_W, py0 = (a.tile_row0 + ty) * TILE_H;
        // this lane's own pixel (validity and the final write); slots carry arbitrary pixels
        const uint32_t my_px = px0 + (lane & (TILE_W - 1)), my_py = py0 + (lane >> 3);
        const bool my_valid = my_px < a.cam.width && my_py >= a.row_begin && my_py < a.row_end;
        const unsigned valid_mask = __ballot_sync(FULL, my_valid);

        // reset the pool
#pragma unroll
        for (int k = 0; k < P / 32; ++k) pool.state[k * 32 + lane] = SLOT_EMPTY;
        for (int k = lane; k < 96; k += 32) pool.acc[k] = 0;
        __syncwarp();

        // work list: item i = sample * nv + k, k-th valid pixel of the tile (nv = 32 for full tiles)
        const uint32_t nv = (uint32_t)__popc(valid_mask);
        uint32_t next_item = 0;
        const uint32_t n_items = (a.max_depth == 0) ? 0u : a.samples * nv;

        // warp-uniform slot census, maintained incrementally
        int nE = P, nT = 0, nS = 0;
        // per-lane traversal context (may stay parked across shade / generate phases)
        bool running = false;
        int cur = -1, node = B200RT_TRAV_DONE, sp = 0;
        RayF ray = make_ray_shade(f3(0, 0, 0), f3(0, 0, 1));
        Closest c; c.t = INFINITY; c.code = -1; c.face = 0;
        uint32_t nrays = 0, nexh = 0, npaths = 0;

        for (;;) {
            __syncwarp();
            const int nRun = __popc(__ballot_sync(FULL, running));
            const bool can_gen = next_item < n_items && nE > 0;
            const bool full_gen = next_item < n_items && nE >= 32;
            if (COUNT) d0 += 1;

            int phase;   // 0 shade, 1 generate, 2 traverse
            if (nS >= 32) phase = 0;
            else if (full_gen) phase = 1;
            else if (nT > 0 || nRun >= T_PARK || (nRun > 0 && nS == 0 && !can_gen)) phase = 2;
            else if (nS > 0) phase = 0;
            else if (can_gen) phase = 1;
            else if (nRun > 0) phase = 2;
            else break;

            // a ray produced by GENERATE or by SHADE's scatter; stored by the one produce_ray below
            bool produced = false;
            int p_slot = -1; uint32_t p_pixel = 0, p_depth = 0;
            float3 p_o = f3(0, 0, 0), p_d = f3(0, 0, 1), p_atten = f3(1, 1, 1);
            Rng p_rng; p_rng.state = 0; p_rng.inc = 1;

            if (phase == 1) {
                // ---- GENERATE: render_scanline's sample loop body, render.rs:60-66 ----
                int total;
                int slot = claim_slots<P>(pool, SLOT_EMPTY, true, lane, &total);
                unsigned gm = __ballot_sync(FULL, slot >= 0);
                uint32_t take = min((uint32_t)__popc(gm), n_items - next_item);
                uint32_t my_rank = (uint32_t)__popc(gm & lt);
                bool have = slot >= 0 && my_rank < take;
                uint32_t item = next_item + my_rank;
                next_item += take;
                if (have) {
                    uint32_t sidx = (nv == 32u) ? (item >> 5) : item / nv;
                    uint32_t kth = item - sidx * nv;
                    uint32_t pl = (nv == 32u) ? kth : (uint32_t)__fns(valid_mask, 0, (int)kth + 1);
                    uint32_t px = px0 + (pl & (TILE_W - 1)), py = py0 + (pl >> 3);
                    Rng rng; rng.init(a.keys, py * a.cam.width + px, a.sample_offset + sidx);
                    float jx = (float)px + rng.gen();
                    float jy = (float)py + rng.gen();
                    float3 o, d;
                    pixel_ray(a.cam, rng, jx, jy, &o, &d);
                    produced = true; p_slot = slot; p_o = o; p_d = d; p_rng = rng; p_pixel = pl; p_depth = a.max_depth;
                    ++npaths;
                }
                nE -= (int)take; nT += (int)take;
                if (COUNT) d6 += take;
            } else if (phase == 2) {
                // ---- TRAVERSE: BboxTree::hit_workspace, bvh/bbox_tree.rs:56-91 ----
                bool do_fetch = true;
                for (;;) {
                    // fetch: idle lanes claim waiting slots; the scene-spanning primitives were
                    // already tested by the producer, so a fetch is 9 loads and 3 reciprocals
                    if (do_fetch && nT > 0 && __any_sync(FULL, !running)) {
                        int total;
                        int slot = claim_slots<P>(pool, SLOT_TRAV, !running, lane, &total);
                        if (slot >= 0) {
                            pool.state[slot] = SLOT_RUN;
                            float3 o = f3(pool.ox[slot], pool.oy[slot], pool.oz[slot]), d = f3(pool.dx[slot], pool.dy[slot], pool.dz[slot]);
                            ray = FAST ? make_ray_fast(o, d) : make_ray(o, d);
                            c.t = pool.t[slot]; c.code = pool.code[slot]; c.face = (int)((pool.meta[slot] >> 5) & 7u);
                            node = 0; sp = 0; cur = slot; running = true;
                        }
                        nT -= __popc(__ballot_sync(FULL, slot >= 0));
                        if (COUNT) d7 += 1;
                    }
                    // inner-node steps while enough lanes want one (always at least one step)
                    bool in = running && node >= 0 && node != B200RT_TRAV_DONE;
                    if (__any_sync(FULL, in)) {
                        do {
                            if (COUNT) { d3 += 1; d4 += __popc(__ballot_sync(FULL, in)); }
                            if (in) trav_inner<COUNT, FAST>(ray, acc, stack, BLK, T_MIN, c, node, sp, tc);
                            in = running && node >= 0 && node != B200RT_TRAV_DONE;
                        } while (__popc(__ballot_sync(FULL, in)) >= T_INNER);
                    }
                    // postponed leaves
                    if (running && node < 0) trav_leaf<COUNT>(ray, acc, stack, BLK, T_MIN, c, node, sp, tc);
                    // finished traversals: park the slot under its shading kind, release the lane
                    bool fin = running && node == B200RT_TRAV_DONE;
                    if (fin) {
                        uint32_t kind = SK_MISS;
                        if (c.code >= 0) {
                            MatRec m = acc.mat((int)((uint32_t)c.code & B200RT_LEAF_ID_MASK));
                            kind = m.kind == B200RT_MAT_METAL ? SK_METAL : (m.kind == B200RT_MAT_DIELECTRIC ? SK_DIELECTRIC
                                   : ((m.kind == B200RT_MAT_LAMBERTIAN && m.tex < 0) ? SK_LAMBERT_SOLID : SK_OTHER));
                        }
                        pool.t[cur] = c.t; pool.code[cur] = c.code;
                        pool.meta[cur] = (pool.meta[cur] & ~(7u << 5)) | ((uint32_t)c.face << 5);
                        pool.state[cur] = SLOT_SHADE0 + kind;
                        running = false; cur = -1;
                    }
                    nS += __popc(__ballot_sync(FULL, fin));
                    int n_run = __popc(__ballot_sync(FULL, running));
                    if (nT > 0) do_fetch = (32 - n_run >= F_FETCH) || n_run == 0;   // refill when enough lanes idle
                    else {
                        do_fetch = false;
                        if (n_run == 0) break;
                        // park the stragglers when a full batch of other work is ready
                        if (n_run < T_PARK && (nS >= 32 || (next_item < n_items && nE >= 32))) break;
                    }
                }
                continue;
            } else {
                // ---- SHADE: ray_color's loop body, render.rs:31-46, one or two kinds per batch ----
                int cnt[SK_COUNT];
#pragma unroll
                for (int q = 0; q < SK_COUNT; ++q) cnt[q] = 0;
#pragma unroll
                for (int k = 0; k < P / 32; ++k) {
                    uint32_t st = pool.state[k * 32 + lane];
#pragma unroll
                    for (int q = 0; q < SK_COUNT; ++q) cnt[q] += __popc(__ballot_sync(FULL, st == (uint32_t)(SLOT_SHADE0 + q)));
                }
                int k1 = 0;
#pragma unroll
                for (int q = 1; q < SK_COUNT; ++q) if (cnt[q] > cnt[k1]) k1 = q;
                int total;
                int slot = claim_slots<P>(pool, SLOT_SHADE0 + k1, true, lane, &total);
                if (COUNT) { d1 += 1; d2 += 1; }
                if (total < 32) {
                    int k2 = -1;
#pragma unroll
                    for (int q = 0; q < SK_COUNT; ++q) if (q != k1 && cnt[q] > 0 && (k2 < 0 || cnt[q] > cnt[k2])) k2 = q;
                    if (k2 >= 0) {
                        int total2;
                        int slot2 = claim_slots<P>(pool, SLOT_SHADE0 + k2, slot < 0, lane, &total2);
                        if (slot < 0) slot = slot2;
                        if (COUNT) d2 += 1;
                    }
                }
                bool have = slot >= 0;
                int n_have = __popc(__ballot_sync(FULL, have));
                if (COUNT) d5 += n_have;
                bool hit = false;
                HitRec h; h.p = f3(0.f, 0.f, 0.f);
                ShadePrep sp_; sp_.tex.need_perlin = false; sp_.tex.perlin_idx = 0;
                RayF r = make_ray_shade(f3(0, 0, 0), f3(0, 0, 1));
                float3 atten = f3(1, 1, 1);
                uint32_t meta = 0;
                bool cont = false;
                if (have) {
                    r = make_ray_shade(f3(pool.ox[slot], pool.oy[slot], pool.oz[slot]), f3(pool.dx[slot], pool.dy[slot], pool.dz[slot]));
                    atten = f3(pool.ax[slot], pool.ay[slot], pool.az[slot]);
                    meta = pool.meta[slot];
                    Closest hc; hc.t = pool.t[slot]; hc.code = pool.code[slot]; hc.face = (int)((meta >> 5) & 7u);
                    hit = hc.code >= 0;
                    if (!hit) {
                        acc_add(pool.acc, meta & 31u, atten * background(a.scene, r.d));
                    } else {
                        h = make_hit(r, acc, hc);
                        sp_ = shade_prepare(a.scene, acc, h);
                    }
                }
                float turb = 0.0f;
                if (has_perlin && __any_sync(FULL, hit && sp_.tex.need_perlin))
                    turb = coop_turbulence(a.scene.perlin, hit && sp_.tex.need_perlin, h.p, sp_.tex.perlin_idx);
                if (hit) {
                    Rng rng; rng.state = pool.rs[slot]; rng.inc = pool.ri[slot];
                    float3 albedo = sp_.tex.need_perlin ? marble(sp_.tex.perlin_scale, h.p, turb) : sp_.tex.rgb;
                    float3 emit = f3(0.f, 0.f, 0.f);
                    ShadeOut so = shade_finish(r, h, sp_.m, albedo, rng, atten, emit);
                    if (emit.x != 0.f || emit.y != 0.f || emit.z != 0.f) acc_add(pool.acc, meta & 31u, emit);
                    uint32_t depth = meta >> 8;
                    cont = so.scattered;
                    if (cont) { --depth; if (depth == 0) { cont = false; ++nexh; } }
                    if (cont) { produced = true; p_slot = slot; p_o = so.o; p_d = so.d; p_atten = atten; p_rng = rng; p_pixel = meta & 31u; p_depth = depth; }
                }
                if (have && !cont) pool.state[slot] = SLOT_EMPTY;
                int n_cont = __popc(__ballot_sync(FULL, cont));
                nS -= n_have; nT += n_cont; nE += n_have - n_cont;
            }
            // the new rays meet the scene-spanning primitives here, in one full uniform batch
            if (produced) { produce_ray<COUNT, P>(pool, p_slot, acc, top, p_o, p_d, p_atten, p_rng, p_pixel, p_depth, tc); ++nrays; }
        }
        __syncwarp();
        if (my_valid) {
            const float inv = 1.0f / 4294967296.0f;
            float4 v = make_float4((float)pool.acc[lane * 3 + 0] * inv, (float)pool.acc[lane * 3 + 1] * inv, (float)pool.acc[lane * 3 + 2] * inv, (float)a.samples);
            float4* dst = a.accum + (my_py * a.cam.width + my_px);
            if (a.accumulate) { float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            *dst = v;
        }
        w_rays += nrays; w_exh += nexh; w_paths += npaths;
    }
    if (COUNT) { w_nodes += tc.nodes; w_prims += tc.prims; }
    for (int o = 16; o > 0; o >>= 1) {
        w_rays += __shfl_down_sync(FULL, w_rays, o);
        w_paths += __shfl_down_sync(FULL, w_paths, o);
        w_exh += __shfl_down_sync(FULL, w_exh, o);
        if (COUNT) { w_nodes += __shfl_down_sync(FULL, w_nodes, o); w_prims += __shfl_down_sync(FULL, w_prims, o); }
    }
    if (lane == 0) {
        atomicAdd(&a.counters->rays, w_rays);
        atomicAdd(&a.counters->paths, w_paths);
        atomicAdd(&a.counters->exhausted, w_exh);
        if (COUNT) {
            atomicAdd(&a.counters->nodes, w_nodes); atomicAdd(&a.counters->prims, w_prims);
            atomicAdd(&a.counters->diag[0], d0); atomicAdd(&a.counters->diag[1], d1); atomicAdd(&a.counters->diag[2], d2); atomicAdd(&a.counters->diag[3], d3);
            atomicAdd(&a.counters->diag[4], d4); atomicAdd(&a.counters->diag[5], d5); atomicAdd(&a.counters->diag[6], d6); atomicAdd(&a.counters->diag[7], d7);
        }
    }
}


}  // namespace b200rt
