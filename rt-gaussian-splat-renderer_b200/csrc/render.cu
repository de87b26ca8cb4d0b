// render.cu — the fused per-ray render kernel (sm_100a) and its small companions.
//
// One persistent kernel renders a whole sample of RayTracer.sample (ray_tracer.py:39-104):
//   camera ray generation        camera.py:31-71              (registers; rays never stored)
//   BVH traversal                scene.py:406-450             (warp-coherent frustum traversal)
//   ray-Gaussian intersection    gaussian.py:203-230          (local-frame quadratic)
//   response + SH colour         gaussian.py:140-201
//   front-to-back compositing    ray_tracer.py:79-104         (register-resident k-buffer)
//
// Execution model.  A warp owns one 4x8-pixel tile at a time (lane = pixel) and pulls tiles from
// a global atomic counter (persistent threads).  All 32 primary rays share the camera origin, so
// the tile is a thin pyramid bounded by 4 planes through the origin.  The warp traverses the LBVH
// ONCE for the tile: up to 32 nodes are popped from a shared-memory stack per step, each lane
// tests the two child boxes of its node against the 4 planes, and survivors are compacted back with
// ballot/popc (internal children -> stack, the nearer one on top; leaves -> candidate queue).  Once rays hold
// K hits the traversal prunes by distance as well - the K-nearest form of the reference's far pruning
// (scene.py:417-419): a box beyond the farthest kept hit of every full ray that also misses the pyramid of the
// rays still lacking hits is dropped.  That makes this kernel the right one for tiles whose frustum holds
// thousands of Gaussians (k_tile_lists sends it the groups whose list overflows).  Candidates are staged 32 at
// a time (one lane each, float64: origin shifted to the closest point of the tile's centre ray)
// and then every lane tests its own ray against every staged candidate with broadcast
// shared-memory reads: first a conservative 5-FMA quadratic (q - 3 as a polynomial of the pixel
// offset), then - in warp-wide rounds, one pending candidate per lane - the precise test, the entry
// distance and alpha.  Hits are appended to a per-lane K-entry buffer in shared memory (replace-max
// when full); after the traversal each lane loads its entries into registers, sorts them with a
// bitonic network and composites front to back (register-resident k-buffer compositor).
//
// Numerics.  The reference's f32 formulation (B^2 - 4AC with a cofactor inverse) is ill-conditioned
// (SURVEY.md §7 hard part 1), and parity is defined against a float64 evaluation of the
// reference's maths.  The fast path is float32 but expressed relative to (Gaussian centre, tile
// centre ray), which keeps all magnitudes O(tile size / sigma); a hit/miss decision within a
// small band of the sqrt(3)-sigma surface, an entry distance within a band of 0, and adjacent
// k-buffer entries closer than a few ulp are re-evaluated in float64 from the raw parameters.
#include <stdlib.h>

#include "render_common.cuh"

using namespace rtgs_dev;

namespace {

#ifndef RTGS_STACK_CAP
#define RTGS_STACK_CAP 512
#endif
constexpr int STACK_CAP = RTGS_STACK_CAP;
constexpr int STACK_SINGLE = STACK_CAP - 100;   // above this pop one node at a time: growth/step <= 32, then DFS depth <= 62
constexpr int CQ_CAP = 96;
constexpr int BATCH = 32;
constexpr int REC_Q = 5;                 // quads per staged record (80-byte stride: conflict-free gathers)

struct __align__(16) TraversalScratch {
    float4 rec[BATCH][REC_Q];   // precise records (render_common.cuh: stage_candidate)
    float4 polyA[BATCH];        // coarse quadratics {c0 c1 c2 c3}
    float2 polyB[BATCH];        //                   {c4 c5}
    int stack[STACK_CAP];
    int cq[CQ_CAP];
    Frustum open;               // pyramid of the rays that still lack hits (distance pruning, below)
};

template <int K>
struct __align__(16) WarpShared {
    union {
        TraversalScratch t;        // traversal phase
        struct {                   // compositing phase: hits in ascending entry distance
            int so_i[K][32];
            float so_a[K][32];
        } c;
    };
    float kb_t[K][32];     // per-lane hit buffer (unsorted): entry distance, sorted position, alpha
    int kb_i[K][32];
    float kb_a[K][32];
};

// STATS = true compiles the per-render counters in (rtgs_render with a stats pointer); the timed path
// uses STATS = false so that the 64-bit counters do not occupy registers.
template <int K, bool STATS>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, (K <= 16 ? 2 : 1)) k_render(const __grid_constant__ RenderParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpShared<K>& ws = reinterpret_cast<WarpShared<K>*>(smem_raw)[threadIdx.x >> 5];
    TraversalScratch& tr = ws.t;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const CamD& cam = P.cam;
    const int xe = P.x0 + P.w, ye = P.y0 + P.h;
    // work items: every tile id, or (fallback mode) the tiles k_tile_lists could not store a list for
    const int nwork = P.use_fallback_list ? (int)min(P.counters[CTR_FALLBACK], (unsigned)P.ntiles) : P.ntiles;

    if (P.use_fallback_list && blockIdx.x == 0 && threadIdx.x == 0) {
        // pool demand of this frame (for the host's sizing) and whether any frame so far needed the fallback
        *reinterpret_cast<volatile int*>(P.mirror) = (int)min(P.counters[CTR_POOL], 0x7fffffffu);
        if (P.counters[CTR_FALLBACK] != 0) *reinterpret_cast<volatile int*>(P.mirror + 1) = 1;
    }

    unsigned long long st_nodes = 0, st_cands = 0, st_pairs = 0, st_f64 = 0, st_layers = 0, st_hit = 0,
                       st_rays = 0, st_tiles = 0, st_steps = 0, st_ins = 0;
#define ST(expr) do { if (STATS) { expr; } } while (0)

#pragma unroll 1
    for (;;) {
        int tile = 0;
        if (lane == 0) {
            tile = (int)atomicAdd(P.counters + CTR_WORK3, 1u);
            if (P.use_fallback_list && tile < nwork) tile = P.fallback_tiles[tile];
            else if (P.use_fallback_list) tile = P.ntiles;
            else tile = work_to_id(P, tile, P.macro_cols * TILES_PER_MACRO, P.ntiles);
        }
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= P.ntiles) break;
        int i0, j0;
        if (!tile_origin(P, tile, i0, j0) || i0 >= xe || j0 >= ye) {
            tile_done(P, tile, lane);
            continue;
        }
        const int pi = i0 + lane / TILE_J, pj = j0 + lane % TILE_J;
        const bool active = pi < xe && pj < ye;

        TileRays ry;
        make_tile_rays(cam, i0, j0, pi, pj, active, ry);
        Frustum fr;
        make_frustum(cam, i0, i0 + TILE_I, j0, j0 + TILE_J, fr);

        // ---- per-lane hit buffer (shared memory, unsorted; replace-max once K entries are held) --
        int cnt = 0;
        float kmax_t = INFINITY;
        int kmax_slot = 0;

        // one pending candidate per lane; the precise test and the hit-only work (entry distance,
        // alpha, float64 refinement, buffer append) run in warp-wide rounds
        bool pend = false;
        int pend_c = 0;

        // Distance pruning - the K-nearest form of the reference's "skip a node whose entry distance exceeds the best
        // hit so far" (scene.py:417-419).  A ray that holds K hits needs nothing farther than its farthest one, so a
        // box is dropped when it lies beyond `cut` = the largest such distance among the rays that are full (with a
        // margin far above float32 rounding, so that near-ties at the K-th place still see both contenders) AND
        // misses the pyramid of the rays that are not full yet (`tr.open`, the bounding pixel rectangle of those
        // rays; initially the whole tile).  Children are pushed far one first.  This is what keeps a tile that looks
        // along a surface - thousands of splats in its frustum, the first few dozen of them opaque - affordable.
        float cut2 = -1.0f;            // (cut * (1 + 1e-4))^2, < 0: no ray is full yet
        unsigned open_mask = __ballot_sync(FULL, active);
        bool open_all = true;          // tr.open == fr
        const float ox = (float)cam.o[0], oy = (float)cam.o[1], oz = (float)cam.o[2];
        auto box_dist2 = [&](float cx, float cy, float cz, float hx, float hy, float hz) {
            const float dx = fmaxf(fabsf(ox - cx) - hx, 0.0f), dy = fmaxf(fabsf(oy - cy) - hy, 0.0f),
                        dz = fmaxf(fabsf(oz - cz) - hz, 0.0f);
            return dx * dx + dy * dy + dz * dz;
        };

        int top = 1, ncq = 0;
        if (lane == 0) tr.stack[0] = 0;
        __syncwarp();

        // ================================ traversal ==========================================
#pragma unroll 1
        while (top > 0 || ncq > 0) {
            if (top > 0) {
                const int take = top > STACK_SINGLE ? 1 : min(32, top);
                int node = -1;
                if (lane < take) node = tr.stack[top - 1 - lane];
                top -= take;
                __syncwarp();
                bool h0 = false, h1 = false;
                int c0 = 0, c1 = 0;
                if (node >= 0) {
                    float4 a, b, c, d;
                    ldg256(P.nodes + (int64_t)node * 4 + 0, a, b);
                    ldg256(P.nodes + (int64_t)node * 4 + 2, c, d);
                    c0 = __float_as_int(d.x);
                    c1 = __float_as_int(d.y);
                    h0 = box_in_frustum(fr, a.x, a.y, a.z, a.w, b.x, b.y);
                    h1 = box_in_frustum(fr, b.z, b.w, c.x, c.y, c.z, c.w);
                    const float d0 = box_dist2(a.x, a.y, a.z, a.w, b.x, b.y);
                    const float d1 = box_dist2(b.z, b.w, c.x, c.y, c.z, c.w);
                    if (cut2 >= 0.0f) {
                        if (h0 && d0 > cut2)
                            h0 = open_mask != 0 && (open_all || box_in_frustum(tr.open, a.x, a.y, a.z, a.w, b.x, b.y));
                        if (h1 && d1 > cut2)
                            h1 = open_mask != 0 && (open_all || box_in_frustum(tr.open, b.z, b.w, c.x, c.y, c.z, c.w));
                    }
                    if (d1 > d0) {   // child 1 is pushed last, i.e. popped first: make it the nearer one
                        const int ci = c0; c0 = c1; c1 = ci;
                        const bool hi = h0; h0 = h1; h1 = hi;
                    }
                }
                ST(st_nodes += 2ull * (unsigned)take);
                ST(st_steps += 1);
                const unsigned mI0 = __ballot_sync(FULL, h0 && c0 >= 0), mI1 = __ballot_sync(FULL, h1 && c1 >= 0);
                const unsigned mL0 = __ballot_sync(FULL, h0 && c0 < 0), mL1 = __ballot_sync(FULL, h1 && c1 < 0);
                if (h0 && c0 >= 0) tr.stack[top + __popc(mI0 & lt_mask)] = c0;
                const int topa = top + __popc(mI0);
                if (h1 && c1 >= 0) tr.stack[topa + __popc(mI1 & lt_mask)] = c1;
                top = topa + __popc(mI1);
                if (h0 && c0 < 0) tr.cq[ncq + __popc(mL0 & lt_mask)] = ~c0;
                const int ncqa = ncq + __popc(mL0);
                if (h1 && c1 < 0) tr.cq[ncqa + __popc(mL1 & lt_mask)] = ~c1;
                ncq = ncqa + __popc(mL1);
                __syncwarp();
            }
            // -------- candidate batch: stage (one lane each, float64) then test (all lanes) --
#pragma unroll 1
            while (ncq >= BATCH || (top == 0 && ncq > 0)) {
                const int m = min(BATCH, ncq);
                ncq -= m;
                if (lane < m) {
                    float4 rec[5];
                    float poly[6];
                    stage_candidate(P, ry, tr.cq[ncq + lane], rec, poly);
#pragma unroll
                    for (int k = 0; k < 5; ++k) tr.rec[lane][k] = rec[k];
                    tr.polyA[lane] = make_float4(poly[0], poly[1], poly[2], poly[3]);
                    tr.polyB[lane] = make_float2(poly[4], poly[5]);
                }
                __syncwarp();
                ST(st_cands += (unsigned)m);
                ST(st_pairs += 32ull * (unsigned)m);
#pragma unroll 1
                for (int c = 0; c <= m; ++c) {
                    bool cand = false;
                    if (c < m) {
                        const float4 pA = tr.polyA[c];
                        const float2 pB = tr.polyB[c];
                        const float ta = fmaf(ry.pa, pA.w, fmaf(ry.pb, pB.x, pA.y));   // c1 + a c3 + b c4
                        const float tb = fmaf(ry.pb, pB.y, pA.z);                      // c2 + b c5
                        const float S = fmaf(ry.pa, ta, fmaf(ry.pb, tb, pA.x));
                        cand = active && (S < 0.0f);
                    }
                    // flush when a lane gets a second candidate, and once at the end of the batch
                    // (the staged records are about to be overwritten)
                    if (__any_sync(FULL, pend && (cand || c == m))) {
                        if (pend) {
                            const PreciseHit h = precise_test(P, tr.rec[pend_c], ry.dlx, ry.dly, ry.dlz, pi, pj);
                            ST(st_f64 += h.refined);
                            if (h.hit) {
                                auto find_farthest = [&]() {
                                    float mt = -INFINITY;
                                    int ms = 0;
#pragma unroll 4
                                    for (int k = 0; k < K; ++k) {
                                        const float t = ws.kb_t[k][lane];
                                        if (t > mt) { mt = t; ms = k; }
                                    }
                                    kmax_t = mt;
                                    kmax_slot = ms;
                                };
                                if (cnt < K) {
                                    const int slot = cnt++;
                                    ws.kb_t[slot][lane] = h.t1;
                                    ws.kb_i[slot][lane] = h.s;
                                    ws.kb_a[slot][lane] = h.alpha;
                                    if (cnt == K) find_farthest();
                                } else {
                                    // full: the candidate replaces the farthest entry if it is nearer.  Whenever
                                    // two contenders for the last place are within float32 rounding of each other
                                    // - the candidate and the farthest entry, or the evicted entry and the new
                                    // farthest one - their float64 entry distances decide (exact_less; rare).
                                    float ct = h.t1, ca = h.alpha;
                                    int cs = h.s;
#pragma unroll 1
                                    for (;;) {
                                        bool nearer = ct < kmax_t;
                                        if (fabsf(ct - kmax_t) <= 2e-6f * kmax_t) {
                                            nearer = exact_less(P.raw, cam, cs, ws.kb_i[kmax_slot][lane], pi, pj);
                                            ST(st_f64 += 2);
                                        }
                                        if (!nearer) break;
                                        const float et = kmax_t, ea = ws.kb_a[kmax_slot][lane];
                                        const int es = ws.kb_i[kmax_slot][lane];
                                        ws.kb_t[kmax_slot][lane] = ct;
                                        ws.kb_i[kmax_slot][lane] = cs;
                                        ws.kb_a[kmax_slot][lane] = ca;
                                        find_farthest();
                                        if (!(et - kmax_t <= 2e-6f * et)) break;
                                        ct = et; cs = es; ca = ea;   // the evicted entry ties with the new farthest
                                    }
                                }
                            }
                            pend = false;
                        }
                        ST(st_ins += 1);
                    }
                    if (cand) {
                        pend = true;
                        pend_c = c;
                    }
                }
                __syncwarp();
                // ---- pruning state after the batch: who is full, how far their farthest hit is ----
                {
                    const bool full = active && cnt == K;
                    float cm = full ? kmax_t : -1.0f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(FULL, cm, o));
                    if (cm >= 0.0f) {
                        cm *= 1.0001f;
                        cut2 = cm * cm;
                    }
                    const unsigned om = __ballot_sync(FULL, active && cnt < K);
                    if (om != open_mask) {
                        open_mask = om;
                        if (om != 0) {
                            int il = cnt < K && active ? pi : 0x7fffffff, ih = cnt < K && active ? pi : -1;
                            int jl = cnt < K && active ? pj : 0x7fffffff, jh = cnt < K && active ? pj : -1;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                il = min(il, __shfl_xor_sync(FULL, il, o));
                                ih = max(ih, __shfl_xor_sync(FULL, ih, o));
                                jl = min(jl, __shfl_xor_sync(FULL, jl, o));
                                jh = max(jh, __shfl_xor_sync(FULL, jh, o));
                            }
                            Frustum fo;
                            make_frustum(cam, il, ih + 1, jl, jh + 1, fo);
                            if (lane == 0) tr.open = fo;
                            open_all = false;
                        }
                        __syncwarp();
                    }
                }
            }
        }

        // ---- order the hits by ascending entry distance: rank counting into the compositing list --
        // rank_i = #{j : t_j < t_i}; pairs within float32 rounding of each other are ordered by their
        // float64 entry distances (exact_less).  The traversal scratch is dead from here on.
        int maxcnt = cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxcnt = max(maxcnt, __shfl_xor_sync(FULL, maxcnt, o));
        __syncwarp();
        if (maxcnt > 0) {
            float tk[K];
#pragma unroll
            for (int k = 0; k < K; ++k) tk[k] = k < cnt ? ws.kb_t[k][lane] : INFINITY;
#pragma unroll 1
            for (int i = 0; i < maxcnt; ++i) {
                if (i < cnt) {
                    const float ti_ = ws.kb_t[i][lane];
                    const float band = 2e-6f * fabsf(ti_);
                    int rank = 0, nnear = 0;
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        rank += tk[j] < ti_;
                        nnear += fabsf(tk[j] - ti_) <= band;
                    }
                    const int id = ws.kb_i[i][lane];
                    if (nnear > 1) {   // rare: resolve near ties exactly
                        rank = 0;
#pragma unroll 1
                        for (int j = 0; j < cnt; ++j) {
                            if (j == i) continue;
                            const float tj_ = ws.kb_t[j][lane];
                            if (fabsf(tj_ - ti_) <= band) {
                                rank += exact_less(P.raw, cam, ws.kb_i[j][lane], id, pi, pj);
                                ST(st_f64 += 2);
                            } else {
                                rank += tj_ < ti_;
                            }
                        }
                    }
                    ws.c.so_i[rank][lane] = id;
                    ws.c.so_a[rank][lane] = ws.kb_a[i][lane];
                }
            }
        }
        __syncwarp();

        // ================================ compositing ========================================
        // accum += T * alpha * rgb ; T *= 1 - alpha   (ray_tracer.py:96-98), rgb = color +
        // eval_sh(normalize(dir)) (gaussian.py:199-200).
        float T = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;
        int nl = 0;
        {
            float Y[15];
            sh_basis(ry.dnx, ry.dny, ry.dnz, Y);
            const int nmine = min(cnt, P.depth);
            const int nloop = min(maxcnt, P.depth);
#pragma unroll 1
            for (int k = 0; k < nloop; ++k) {
                if (k < nmine && T >= P.t_cut) {
                    const int s = ws.c.so_i[k][lane];
                    const float alpha = ws.c.so_a[k][lane];
                    float r, g, b;
                    eval_colour(P, s, Y, r, g, b);
                    const float wgt = T * alpha;
                    cr = fmaf(wgt, r, cr);
                    cg = fmaf(wgt, g, cg);
                    cb = fmaf(wgt, b, cb);
                    T *= 1.0f - alpha;
                    ++nl;
                }
            }
        }
        store_tile(P, reinterpret_cast<float*>(&ws.kb_t[0][0]), lane, i0, j0, pi, pj, active, cr, cg, cb, T);
        tile_done(P, tile, lane);
        if (active) {
            ST(st_rays += 1);
            ST(st_hit += nl > 0);
            ST(st_layers += (unsigned)nl);
        }
        ST(if (lane == 0) st_tiles += 1);
    }

    if (STATS && P.stats) {
        unsigned long long v[ST_COUNT] = {0};
        v[ST_RAYS] = st_rays; v[ST_RAYS_HIT] = st_hit; v[ST_LAYERS] = st_layers; v[ST_F64] = st_f64;
        // warp-uniform counters are taken from lane 0 only
        if (lane == 0) {
            v[ST_NODES] = st_nodes;
            v[ST_CANDS] = st_cands;
            v[ST_PAIRS] = st_pairs;
            v[ST_TILES] = st_tiles;
            v[ST_STEPS] = st_steps;
            v[ST_INSERTS] = st_ins;
            v[ST_FALLBACK] = P.use_fallback_list ? st_tiles : 0;
        }
#pragma unroll
        for (int k = 0; k < ST_COUNT; ++k) {
            unsigned long long x = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
            if (lane == 0 && x) atomicAdd(P.stats + k, x);
        }
    }
}
#undef ST

// ---- Camera.generate_ray_field (camera.py:57-71): (W,H,8) = origin, direction, start, end -----
__global__ void k_generate_rays(const CamD cam, float* __restrict__ rays) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)cam.W * cam.H) return;
    const int i = (int)(idx / cam.H), j = (int)(idx % cam.H);
    const d3 d = cam_dir(cam, (double)i + 0.5, (double)j + 0.5);
    float4* o = reinterpret_cast<float4*>(rays + idx * 8);
    o[0] = make_float4((float)cam.o[0], (float)cam.o[1], (float)cam.o[2], (float)d.x);
    o[1] = make_float4((float)d.y, (float)d.z, 0.0f, INFINITY);   // ray.py:20-41 start 0, end inf
}

// ---- Scene.hit for a batch of arbitrary rays (scene.py:406-450) -------------------------------
// Per-ray stack traversal, near child first, far-node pruning against the best entry distance.
// Box tests are float32 slabs on slightly inflated boxes; the ray-Gaussian test is the float64
// exact evaluation (this entry point exists for parity checks of closest-hit semantics, not for
// the fused renderer, which never calls it).
__global__ void k_trace_closest(const float4* __restrict__ nodes, const float4* __restrict__ raw, int64_t nrays,
                                const float* __restrict__ rays, int32_t* __restrict__ out_idx,
                                float* __restrict__ out_t12) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nrays) return;
    const float* rp = rays + r * 8;
    const double o[3] = {rp[0], rp[1], rp[2]};
    const d3 d = d3make(rp[3], rp[4], rp[5]);
    const double t_start = rp[6], t_end = rp[7];
    const float of[3] = {rp[0], rp[1], rp[2]};
    const float idx_[3] = {1.0f / rp[3], 1.0f / rp[4], 1.0f / rp[5]};
    double best = INFINITY, best_t2 = INFINITY;
    int best_s = -1;
    int stack[64];
    int sp = 0;
    stack[sp++] = 0;
    auto slab = [&](float mnx, float mny, float mnz, float mxx, float mxy, float mxz) -> float {
        // returns entry distance or +inf on a miss (bounding_box.py:50-89; hit iff t_min < t_max)
        float t0 = -INFINITY, t1 = INFINITY;
        const float mn[3] = {mnx, mny, mnz}, mx[3] = {mxx, mxy, mxz};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float ta = (mn[a] - of[a]) * idx_[a], tb = (mx[a] - of[a]) * idx_[a];
            float lo = fminf(ta, tb), hi = fmaxf(ta, tb);
            // inflate by a few ulp so float32 rounding can never cull a true hit
            lo -= 4e-6f * fabsf(lo) + 1e-30f;
            hi += 4e-6f * fabsf(hi) + 1e-30f;
            if (ta != ta || tb != tb) { lo = -INFINITY; hi = INFINITY; }   // 0 * inf: ray inside slab plane
            t0 = fmaxf(t0, lo);
            t1 = fminf(t1, hi);
        }
        return (t0 <= t1) ? t0 : INFINITY;
    };
    while (sp > 0) {
        const int node = stack[--sp];
        const float4 a = __ldg(nodes + (int64_t)node * 4 + 0), b = __ldg(nodes + (int64_t)node * 4 + 1),
                     c = __ldg(nodes + (int64_t)node * 4 + 2), dd = __ldg(nodes + (int64_t)node * 4 + 3);
        const int ch[2] = {__float_as_int(dd.x), __float_as_int(dd.y)};
        // child boxes are (centre, half extent): min = c - h, max = c + h
        float te[2] = {slab(a.x - a.w, a.y - b.x, a.z - b.y, a.x + a.w, a.y + b.x, a.z + b.y),
                       slab(b.z - c.y, b.w - c.z, c.x - c.w, b.z + c.y, b.w + c.z, c.x + c.w)};
        // visit near child first: push far first
        const int first = te[0] <= te[1] ? 0 : 1;
        for (int k = 1; k >= 0; --k) {
            const int w = k == 0 ? first : 1 - first;
            if (!(te[w] < INFINITY) || (double)te[w] > best) continue;   // miss or farther than best
            if (ch[w] >= 0) {
                stack[sp++] = ch[w];
            } else {
                const int s = ~ch[w];
                const float4 ra = __ldg(raw + (int64_t)s * 3 + 0), rb = __ldg(raw + (int64_t)s * 3 + 1),
                             rc = __ldg(raw + (int64_t)s * 3 + 2);
                const double p[3] = {ra.x, ra.y, ra.z}, q[4] = {ra.w, rb.x, rb.y, rb.z}, sc[3] = {rb.w, rc.x, rc.y};
                const ExactHit e = exact_intersect(p, q, sc, o, d);
                if (e.hit && e.t1 > t_start && e.t1 < t_end && e.t1 < best) {   // scene.py:433-437
                    best = e.t1;
                    best_t2 = e.t2;
                    best_s = s;
                }
            }
        }
    }
    if (best_s >= 0) {
        out_idx[r] = __float_as_int(__ldg(raw + (int64_t)best_s * 3 + 2).z);
        out_t12[r * 2 + 0] = (float)best;
        out_t12[r * 2 + 1] = (float)best_t2;
    } else {
        out_idx[r] = -1;
        out_t12[r * 2 + 0] = INFINITY;
        out_t12[r * 2 + 1] = INFINITY;
    }
}

// ---- PLY rows -> stored parameters (scene.py:101-114 on the device) ---------------------------
__global__ void k_activate_ply(int64_t n, const float* __restrict__ rows, int stride, const int32_t* __restrict__ col,
                               float scale_arg, int sh_layout, float* __restrict__ pos, float* __restrict__ rot,
                               float* __restrict__ sca, float* __restrict__ color, float* __restrict__ opacity,
                               float* __restrict__ sh) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* row = rows + i * stride;
    auto get = [&](int k) -> float { int c = col[k]; return c >= 0 ? row[c] : 0.0f; };
    // column order: x,y,z (0-2), f_dc (3-5), f_rest (6-50), opacity (51), scale (52-54), rot_0..3 (55-58)
    pos[i * 3 + 0] = get(0); pos[i * 3 + 1] = get(1); pos[i * 3 + 2] = get(2);
    // scalar-first (rot_0 = w) -> scalar-last, normalised in float32 (scene.py:103,110-111)
    const float qx = get(56), qy = get(57), qz = get(58), qw = get(55);
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz)),
                                           __fmul_rn(qw, qw)));
    rot[i * 4 + 0] = __fdiv_rn(qx, nrm); rot[i * 4 + 1] = __fdiv_rn(qy, nrm);
    rot[i * 4 + 2] = __fdiv_rn(qz, nrm); rot[i * 4 + 3] = __fdiv_rn(qw, nrm);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        sca[i * 3 + a] = __fmul_rn(expf(get(52 + a)), scale_arg);                  // scene.py:112
        color[i * 3 + a] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-get(3 + a))));     // scene.py:113
    }
    opacity[i] = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-get(51))));                 // scene.py:114
    if (sh != nullptr) {
        for (int k = 0; k < 15; ++k)
            for (int c = 0; c < 3; ++c)
                sh[i * 45 + k * 3 + c] = sh_layout == 0 ? get(6 + 15 * c + k) : get(6 + 3 * k + c);
    }
}

CamD make_camd(const rtgs_camera* cam) {
    CamD c;
    for (int a = 0; a < 3; ++a) c.o[a] = cam->position[a];
    for (int a = 0; a < 4; ++a) c.q[a] = cam->rotation[a];
    double Rm[3][3];
    quat_to_mat(c.q, Rm);   // utils/quaternion.py:99-121 in float64 (rotation used as given, not normalised)
    for (int r = 0; r < 3; ++r)
        for (int k = 0; k < 3; ++k) c.R[r * 3 + k] = Rm[r][k];
    c.fx = cam->focal[0];
    c.fy = cam->focal[1];
    c.ifx = 1.0 / c.fx;
    c.ify = 1.0 / c.fy;
    c.W = cam->width;
    c.H = cam->height;
    return c;
}

template <int K, bool STATS>
int launch_render_k(rtgs_scene* s, const RenderParams& P, cudaStream_t stream) {
    static int blocks_per_sm[16] = {0};
    const size_t smem = sizeof(WarpShared<K>) * WARPS_PER_CTA;
    int dev = s->device;
    if (dev < 0 || dev >= 16) dev = 0;
    if (blocks_per_sm[dev] == 0) {
        CUDA_TRY(cudaFuncSetAttribute(k_render<K, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_render<K, STATS>, WARPS_PER_CTA * 32, smem));
        if (nb < 1) {
            rtgs_set_error("render kernel does not fit on an SM (smem %zu)", smem);
            return RTGS_ERR_CUDA;
        }
        blocks_per_sm[dev] = nb;
    }
    int grid = s->sm_count * blocks_per_sm[dev];
    const int need = (P.ntiles + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (grid > need) grid = need;
    if (P.use_fallback_list) {
        // The fallback launch is empty unless the list pool overflowed.  While no finished frame has reported
        // fallback tiles it runs on a quarter of the SMs (a cheaper launch); the first frame that does overflow
        // is still rendered correctly, just with fewer warps on its fallback tiles.
        const int seen = *reinterpret_cast<volatile int*>(s->band_flags + RTGS_MAX_BANDS + 1);
        static const char* e = getenv("RTGS_FB_GRID");
        const int small = e ? atoi(e) : (s->sm_count + 3) / 4;
        if (!seen && grid > small) grid = small;
    }
    if (grid < 1) grid = 1;
    k_render<K, STATS><<<grid, WARPS_PER_CTA * 32, smem, stream>>>(P);
    CUDA_TRY(cudaGetLastError());
    return RTGS_OK;
}

// Which kernels render a frame: 0 = k_tile_lists + k_shade_tiles (+ k_render on the fallback list),
// 1 = the fused k_render alone.  depth > 16 always takes the fused kernel (its k-buffer has 32 entries).
int render_mode(const rtgs_scene* s) {
    if (s->opt_render_mode >= 0) return s->opt_render_mode;
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("RTGS_RENDER_MODE");
        mode = e ? atoi(e) : 0;
    }
    return mode;
}

// Candidate-list scratch of a scene, sized for the tile count of the largest region rendered so far.
// The pool starts at 16 chunks (496 candidates) per tile on average - tiles may take any share of it - and
// doubles whenever the demand reported by the last finished frame (k_tile_lists keeps counting past the end of
// the pool; the closing k_render launch mirrors the count into mapped host memory) exceeded 70 % of it.  A frame
// that overflows is still rendered correctly (fallback list), the next one has the larger pool.
int ensure_lists(rtgs_scene* s, int ntiles) {
    int64_t want = s->pool_chunks;
    if (s->opt_pool_chunks < 0 && s->list_tiles > 0) {
        const int64_t used = *reinterpret_cast<volatile int*>(s->band_flags + RTGS_MAX_BANDS);
        while (used * 10 > want * 7 && want < (1ll << 26)) want *= 2;
    }
    if (s->list_tiles >= ntiles && want == s->pool_chunks) return RTGS_OK;
    const bool grow_only = s->list_tiles >= ntiles;
    const int tiles = grow_only ? s->list_tiles : ntiles;
    cudaFree(s->tile_desc); cudaFree(s->list_pool); cudaFree(s->fallback_tiles);
    s->tile_desc = nullptr; s->list_pool = nullptr; s->fallback_tiles = nullptr;
    s->list_tiles = 0;
    int64_t chunks = s->opt_pool_chunks >= 0 ? s->opt_pool_chunks : (int64_t)tiles * 16;
    if (grow_only && want > chunks) chunks = want;
    if (chunks > (1ll << 26)) chunks = 1ll << 26;
    CUDA_TRY(cudaMalloc((void**)&s->tile_desc, (size_t)tiles * sizeof(TileDesc)));
    CUDA_TRY(cudaMalloc((void**)&s->list_pool, (size_t)(chunks > 0 ? chunks : 1) * CHUNK_INTS * sizeof(int)));
    CUDA_TRY(cudaMalloc((void**)&s->fallback_tiles, (size_t)tiles * sizeof(int)));
    s->list_tiles = tiles;
    s->pool_chunks = (int)chunks;
    s->band_flags[RTGS_MAX_BANDS] = 0;
    return RTGS_OK;
}

}  // namespace

static int launch_render_on(rtgs_scene* s, const rtgs_camera* cam, int x0, int y0, int w, int h, int depth,
                            float t_cut, int accumulate, int full_pitch, float* out_rgb, float* out_T,
                            cudaStream_t stream, bool want_stats);

// A scene's render scratch (work counters, candidate lists, band counters) is shared by all its frames, which is
// safe in stream order.  A frame launched on ANOTHER stream than the previous one (rtgs_render on the caller's
// stream, then rtgs_render_host on the library's) first waits for that frame's kernels.
int rtgs_launch_render(rtgs_scene* s, const rtgs_camera* cam, int x0, int y0, int w, int h, int depth,
                       float t_cut, int accumulate, int full_pitch, float* out_rgb, float* out_T,
                       cudaStream_t stream, bool want_stats) {
    if (!s->scratch_free) CUDA_TRY(cudaEventCreateWithFlags(&s->scratch_free, cudaEventDisableTiming));
    if (s->scratch_used && s->scratch_stream != stream) CUDA_TRY(cudaStreamWaitEvent(stream, s->scratch_free, 0));
    const int r = launch_render_on(s, cam, x0, y0, w, h, depth, t_cut, accumulate, full_pitch, out_rgb, out_T, stream,
                                   want_stats);
    // (also after a failed launch sequence: whatever was queued still uses the scratch)
    if (cudaEventRecord(s->scratch_free, stream) == cudaSuccess) {
        s->scratch_stream = stream;
        s->scratch_used = true;
    } else {
        cudaGetLastError();
    }
    return r;
}

static int launch_render_on(rtgs_scene* s, const rtgs_camera* cam, int x0, int y0, int w, int h, int depth,
                            float t_cut, int accumulate, int full_pitch, float* out_rgb, float* out_T,
                            cudaStream_t stream, bool want_stats) {
    RenderParams P;
    P.nodes = s->nodes;
    P.nodes4 = s->nodes4;
    P.geo = s->geo;
    P.shp = s->shp;
    P.raw = s->raw;
    P.leafbox = s->leafbox;
    P.cam = make_camd(cam);
    P.x0 = x0; P.y0 = y0; P.w = w; P.h = h;
    const int mrows = (w + MACRO_PX_I - 1) / MACRO_PX_I;
    const int mcols = (h + MACRO_PX_J - 1) / MACRO_PX_J;
    P.macro_cols = mcols;
    P.stripe_mod = s->opt_stripe_mod;
    P.stripe_rem = s->opt_stripe_rem;
    P.ntiles = mrows * mcols * TILES_PER_MACRO;
    P.depth = depth;
    P.t_cut = t_cut;
    P.accumulate = accumulate;
    P.full_pitch = full_pitch;
    P.has_sh = s->has_sh ? 1 : 0;
    P.out_rgb = out_rgb;
    P.out_T = out_T;
    P.counters = s->counters;
    P.stats = want_stats ? s->stats_dev : nullptr;
    P.desc = nullptr;
    P.pool = nullptr;
    P.pool_chunks = 0;
    P.fallback_tiles = nullptr;
    P.use_fallback_list = 0;
    static const int heavy_fused = getenv("RTGS_HEAVY_FUSED") ? atoi(getenv("RTGS_HEAVY_FUSED")) : 1;
    P.heavy_fused = heavy_fused;
    static const int heavy_limit = getenv("RTGS_HEAVY_LIMIT") ? atoi(getenv("RTGS_HEAVY_LIMIT")) : 1 << 30;
    P.heavy_limit = heavy_limit;
    P.nbands = 0;
    P.band_macro_cols = 1;
    P.macro_rows = mrows;
    P.schedule = s->bands_active > 0 ? s->band_schedule : 0;
    P.band_done = s->band_done;
    P.band_flags = s->band_flags_cur_dev ? s->band_flags_cur_dev : s->band_flags_dev;
    P.mirror = s->band_flags_dev + RTGS_MAX_BANDS;
    if (s->bands_active > 0) {
        P.nbands = s->bands_active;
        P.band_macro_cols = s->band_macro_cols;
        CUDA_TRY(cudaMemsetAsync(s->band_done, 0, RTGS_MAX_BANDS * sizeof(unsigned int), stream));
    }
    CUDA_TRY(cudaMemsetAsync(s->counters, 0, CTR_COUNT * sizeof(unsigned int), stream));
    if (want_stats) CUDA_TRY(cudaMemsetAsync(s->stats_dev, 0, ST_COUNT * sizeof(unsigned long long), stream));

    // optional per-kernel timing: events e[0..3] of this frame's ring slot bracket the three kernels
    cudaEvent_t* ev = nullptr;
    unsigned char* ran = nullptr;
    if (!s->timing_ran.empty()) {
        const int64_t slot = s->timing_frames % (int64_t)s->timing_ran.size();
        ev = &s->timing_events[slot * 4];
        ran = &s->timing_ran[slot];
        *ran = 0;
        ++s->timing_frames;
    }
    auto mark = [&](int k) -> int {
        if (ev) CUDA_TRY(cudaEventRecord(ev[k], stream));
        return RTGS_OK;
    };
    auto fused = [&](int K) -> int {
        if (K > 16) return want_stats ? launch_render_k<32, true>(s, P, stream) : launch_render_k<32, false>(s, P, stream);
        return want_stats ? launch_render_k<16, true>(s, P, stream) : launch_render_k<16, false>(s, P, stream);
    };
    int r;
    if (depth > 16 || render_mode(s) == 1) {
        if ((r = mark(0)) != RTGS_OK || (r = mark(1)) != RTGS_OK || (r = mark(2)) != RTGS_OK) return r;
        if ((r = fused(depth > 16 ? 32 : 16)) != RTGS_OK) return r;
        if (ran) *ran = 4;
        return mark(3);
    }
    // traversal -> per-tile candidate lists -> shading; tiles whose list did not fit the pool are
    // rendered by the fused kernel afterwards (it returns at once when there are none)
    if ((r = ensure_lists(s, P.ntiles)) != RTGS_OK) return r;
    P.desc = reinterpret_cast<TileDesc*>(s->tile_desc);
    P.pool = s->list_pool;
    P.pool_chunks = s->pool_chunks;
    P.fallback_tiles = s->fallback_tiles;
    if ((r = mark(0)) != RTGS_OK) return r;
    if ((r = rtgs_launch_tile_lists(s, P, stream, want_stats)) != RTGS_OK) return r;
    if ((r = mark(1)) != RTGS_OK) return r;
    if ((r = rtgs_launch_shade_tiles(s, P, stream, want_stats)) != RTGS_OK) return r;
    if ((r = mark(2)) != RTGS_OK) return r;
    P.use_fallback_list = 1;
    if ((r = fused(16)) != RTGS_OK) return r;
    if (ran) *ran = 7;
    return mark(3);
}

int rtgs_launch_generate_rays(const rtgs_camera* cam, float* rays, cudaStream_t stream) {
    const CamD c = make_camd(cam);
    const int64_t n = (int64_t)c.W * c.H;
    k_generate_rays<<<(int)((n + 255) / 256), 256, 0, stream>>>(c, rays);
    CUDA_TRY(cudaGetLastError());
    return RTGS_OK;
}

int rtgs_launch_trace_closest(rtgs_scene* s, int64_t nrays, const float* rays, int32_t* idx, float* t12,
                              cudaStream_t stream) {
    if (nrays == 0) return RTGS_OK;
    k_trace_closest<<<(int)((nrays + 127) / 128), 128, 0, stream>>>(s->nodes, s->raw, nrays, rays, idx, t12);
    CUDA_TRY(cudaGetLastError());
    return RTGS_OK;
}

int rtgs_launch_activate_ply(int64_t n, const float* rows_dev, int stride, const int32_t* col, float scale,
                             int sh_layout, float* pos, float* rot, float* sca, float* color, float* opacity,
                             float* sh, cudaStream_t stream) {
    k_activate_ply<<<(int)((n + 127) / 128), 128, 0, stream>>>(n, rows_dev, stride, col, scale, sh_layout, pos, rot,
                                                               sca, color, opacity, sh);
    CUDA_TRY(cudaGetLastError());
    return RTGS_OK;
}
