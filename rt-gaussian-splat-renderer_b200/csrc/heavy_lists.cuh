// heavy_lists.cuh — depth-capped candidate lists for HEAVY groups: 8x16-pixel groups whose frustum holds more Gaussians
// than the shared-memory list of tile_lists.cuh (typically because they look along a surface: thousands of splats, of
// which every ray composites the nearest 16, ray_tracer.py:96-104).  Listing them all is the wrong plan, and so is the
// fused kernel's per-tile near-first traversal (fused.cuh: ~77 k warp instructions per tile).  Instead the group's
// traversal is run in DEPTH SLABS:
//
//   depth z(box) = min over the box of (x - o).c / L,  c = unit direction of the group's centre ray, L >= |d| of every
//   ray of the group: a lower bound of the ray parameter t of any point of the box on any ray of the group.
//
//   1. traverse with a cap: children with z < cap are expanded / listed as in lists_group, children with z >= cap are
//      DEFERRED (appended, with their depth, to a per-warp list in global memory);
//   2. the new candidates are intersected (float64) with 32 SAMPLE rays of the group (a 4x8 lattice, one per lane, eight
//      per tile), each keeping the entry distances of its 16 nearest hits;
//   3. while some sample ray holds fewer than 16 hits the cap is raised (geometrically from the nearest deferred depth,
//      with a step that adapts to the yield), the deferred list is re-scanned and the nodes now below the cap are
//      expanded; once every sample ray is full the cap goes to  margin x (farthest 16th hit)  and the walk ends when
//      nothing nearer than that is deferred any more.
//
// The lists are then filtered per tile as in lists_group, additionally by depth < the tile's own limit (margin x the
// farthest 16th hit of ITS sample rays, at most the depth of the nearest deferred node), and published with that limit
// (TileDesc::count | TILE_CAPPED, tile_cap[tile]).  The limit is a guess for the rays between the samples, so the
// shading checks it: a tile is complete iff every ray holds K hits nearer than the limit - everything that could enter
// a ray nearer than the limit is in the list by construction; otherwise the tile goes to the fused kernel like before
// (shade.cuh).  Nothing depends on the guess being right, only the speed does.
//
// Measured on the surface-like scene (1 M Gaussians, 1080p): 12 % of the tiles are in heavy groups; their frusta hold a
// median of 2300 boxes, the slab that decides them 300-450.
#pragma once
#include "tile_lists.cuh"

namespace rtgs_dev {

constexpr int HEAVY_K = 16;              // hits a sample ray keeps (the k-buffer depth of the shading)
constexpr int HEAVY_MAX_ITERS = 96;      // cap raises per group before it is handed to the fused kernel
constexpr int HEAVY_LIST_LIMIT = GLIST_CAP - 4 * LISTS_TAKE;

struct HeavyStats {
    unsigned long long groups = 0, failed = 0, iters = 0, tested = 0, retries = 0, fail_list = 0, fail_defer = 0, fail_passes = 0;
    unsigned max_deferred = 0;
};

// One heavy group.  `def` = this warp's deferred lists: 2 x P.heavy_defer_cap entries {child id (< 0: ~leaf), depth bits}.
template <bool STATS>
__device__ __noinline__ void heavy_group(const RenderParams& P, ListsShared& ws, ListsState& S, HeavyStats& HS,
                                         int2* __restrict__ def, int group, int gi0, int gj0, int lane) {
    const unsigned lt_mask = (1u << lane) - 1u;
    const CamD& cam = P.cam;
    const int xe = P.x0 + P.w, ye = P.y0 + P.h;
    const int dcap = P.heavy_defer_cap;
#define ST(expr) do { if (STATS) { expr; } } while (0)

    Frustum fg;
    make_frustum(cam, gi0, min(gi0 + GPX_I, xe), gj0, min(gj0 + GPX_J, ye), fg);

    // ---- this lane's sample ray (lattice: i offset 2a + (b & 1), j offset 2b + (a & 1)) and the depth functional ----
    const int la = lane >> 3, lb = lane & 7;
    const int si = gi0 + 2 * la + (lb & 1), sj = gj0 + 2 * lb + (la & 1);
    const bool s_active = si < xe && sj < ye;
    const int s_tile = (la >> 1) * GROUP_TJ + (lb >> 2);
    const d3 sd = cam_dir(cam, (double)si + 0.5, (double)sj + 0.5);
    float zc[3], za[3], zo;
    {
        const d3 dc = cam_dir(cam, (double)gi0 + 0.5 * GPX_I, (double)gj0 + 0.5 * GPX_J);
        // L: the longest direction among the group's corner and centre rays (1 for a unit camera quaternion), plus a
        // margin for the rays in between
        double l2 = d3dot(dc, dc);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const d3 e = cam_dir(cam, (double)(gi0 + (k & 1) * GPX_I), (double)(gj0 + (k >> 1) * GPX_J));
            l2 = fmax(l2, d3dot(e, e));
        }
        const double sc = 1.0 / (sqrt(d3dot(dc, dc)) * sqrt(l2) * 1.0001);
        zc[0] = (float)(dc.x * sc); zc[1] = (float)(dc.y * sc); zc[2] = (float)(dc.z * sc);
#pragma unroll
        for (int k = 0; k < 3; ++k) za[k] = fabsf(zc[k]);
        zo = zc[0] * (float)cam.o[0] + zc[1] * (float)cam.o[1] + zc[2] * (float)cam.o[2];
    }
    auto box_depth = [&](float cx, float cy, float cz, float hx, float hy, float hz) {
        return zc[0] * cx + zc[1] * cy + zc[2] * cz - za[0] * hx - za[1] * hy - za[2] * hz - zo;
    };

    // ---- sample ray state: entry distances of the nearest HEAVY_K hits (unordered), the farthest of them -----------
    float kb[HEAVY_K];
    int kcnt = 0, kslot = 0;
    float kmax = INFINITY;   // the HEAVY_K-th nearest entry distance once kcnt == HEAVY_K

    int ng = 0, ntested = 0;
    int top = 0;
    bool failed = false;
    int why = 0;   // 1 = the slab overflowed the shared list, 2 = the deferred list overflowed, 3 = too many passes
    int n_in = 0, n_out = 0;
    int2* din = def;
    int2* dout = def + dcap;
    float zmin = INFINITY;   // per lane: nearest depth among the entries this lane deferred in this pass

    // expand what is on the stack, depth first, against `cap` (traverse_step of tile_lists.cuh plus the depth test)
    auto drain = [&](float cap) {
#pragma unroll 1
        while (top > 0 && !failed) {
            const int take = top > P.lists_single ? 1 : min(16, top);
            int node = -1;
            if ((lane >> 1) < take) node = ws.stack[top - 1 - (lane >> 1)];
            top -= take;
            __syncwarp();
            bool h0 = false, h1 = false;
            int c0 = 0, c1 = 0;
            float z0 = 0.0f, z1 = 0.0f;
            if (node >= 0) {
                const float4* rec = P.nodes4 + (int64_t)node * 8 + (lane & 1) * 4;
                float4 a, b, c, d;
                ldg256(rec + 0, a, b);
                ldg256(rec + 2, c, d);
                c0 = __float_as_int(d.x);
                c1 = __float_as_int(d.y);
                h0 = box_in_frustum(fg, a.x, a.y, a.z, a.w, b.x, b.y);
                h1 = box_in_frustum(fg, b.z, b.w, c.x, c.y, c.z, c.w);
                z0 = box_depth(a.x, a.y, a.z, a.w, b.x, b.y);
                z1 = box_depth(b.z, b.w, c.x, c.y, c.z, c.w);
            }
            ST(S.st_nodes += 4ull * (unsigned)take);
            ST(S.st_steps += 1);
            const bool n0 = h0 && z0 < cap, n1 = h1 && z1 < cap;   // near: expand / list now
            const bool f0 = h0 && !n0, f1 = h1 && !n1;             // far: deferred
            const unsigned mI0 = __ballot_sync(FULL, n0 && c0 >= 0), mI1 = __ballot_sync(FULL, n1 && c1 >= 0);
            const unsigned mL0 = __ballot_sync(FULL, n0 && c0 < 0), mL1 = __ballot_sync(FULL, n1 && c1 < 0);
            const unsigned mF0 = __ballot_sync(FULL, f0), mF1 = __ballot_sync(FULL, f1);
            if (n0 && c0 >= 0) ws.stack[top + __popc(mI0 & lt_mask)] = c0;
            const int topa = top + __popc(mI0);
            if (n1 && c1 >= 0) ws.stack[topa + __popc(mI1 & lt_mask)] = c1;
            top = topa + __popc(mI1);
            if (ng + __popc(mL0) + __popc(mL1) > HEAVY_LIST_LIMIT || n_out + __popc(mF0) + __popc(mF1) > dcap) {
                failed = true;   // the slab itself overflows the shared list (or the deferred list): fused kernel
                why = ng + __popc(mL0) + __popc(mL1) > HEAVY_LIST_LIMIT ? 1 : 2;
                break;
            }
            if (n0 && c0 < 0) ws.glist[ng + __popc(mL0 & lt_mask)] = ~c0;
            const int nga = ng + __popc(mL0);
            if (n1 && c1 < 0) ws.glist[nga + __popc(mL1 & lt_mask)] = ~c1;
            ng = nga + __popc(mL1);
            if (f0) {
                dout[n_out + __popc(mF0 & lt_mask)] = make_int2(c0, __float_as_int(z0));
                zmin = fminf(zmin, z0);
            }
            const int noa = n_out + __popc(mF0);
            if (f1) {
                dout[noa + __popc(mF1 & lt_mask)] = make_int2(c1, __float_as_int(z1));
                zmin = fminf(zmin, z1);
            }
            n_out = noa + __popc(mF1);
            __syncwarp();
            ST(S.st_max_stack = max(S.st_max_stack, (unsigned)top));
        }
    };

    // the candidates found since the last call against this lane's sample ray (float64: o' = W (o - p), d' = W d,
    // |o' + t d'|^2 = 3; gaussian.py:203-230)
    auto test_new = [&]() {
#pragma unroll 1
        for (int idx = ntested; idx < ng; ++idx) {
            const int s = ws.glist[idx];
            float4 g0, g1, g2, g3;
            ldg256(P.geo + (int64_t)s * 4 + 0, g0, g1);
            ldg256(P.geo + (int64_t)s * 4 + 2, g2, g3);
            const double vx = cam.o[0] - (double)g0.x, vy = cam.o[1] - (double)g0.y, vz = cam.o[2] - (double)g0.z;
            const double ox = g1.x * vx + g1.y * vy + g1.z * vz, oy = g1.w * vx + g2.x * vy + g2.y * vz,
                         oz = g2.z * vx + g2.w * vy + g3.x * vz;
            const double dx = g1.x * sd.x + g1.y * sd.y + g1.z * sd.z, dy = g1.w * sd.x + g2.x * sd.y + g2.y * sd.z,
                         dz = g2.z * sd.x + g2.w * sd.y + g3.x * sd.z;
            const double A = dx * dx + dy * dy + dz * dz, Bh = ox * dx + oy * dy + oz * dz,
                         C = ox * ox + oy * oy + oz * oz - 3.0;
            const double disc = Bh * Bh - A * C;
            if (disc > 0.0 && s_active) {
                const float t1 = (float)((-Bh - sqrt(disc)) / A);
                if (t1 > 0.0f && t1 < kmax) {
                    const int slot = kcnt < HEAVY_K ? kcnt : kslot;
                    kb[slot] = t1;
                    if (kcnt < HEAVY_K) ++kcnt;
                    if (kcnt == HEAVY_K) {
                        float mt = -INFINITY;
#pragma unroll
                        for (int k = 0; k < HEAVY_K; ++k)
                            if (kb[k] > mt) { mt = kb[k]; kslot = k; }
                        kmax = mt;
                    }
                }
            }
        }
        ST(HS.tested += (unsigned)(ng - ntested));
        ntested = ng;
    };

    // ---- the slab walk -------------------------------------------------------------------------------------------
    // pass 0: the root is expanded unconditionally with cap = -inf, i.e. its grandchildren are all deferred
    if (lane == 0) ws.stack[0] = 0;
    top = 1;
    __syncwarp();
    drain(-INFINITY);
    float step = 0.02f, cap_valid = INFINITY;
    bool complete = false;   // no deferred node left: the list is the whole frustum
    int iters = 0;
#pragma unroll 1
    while (!failed) {
        // what the pass deferred: count n_out, nearest depth
        ST(HS.max_deferred = max(HS.max_deferred, (unsigned)n_out));
        float zm = zmin;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) zm = fminf(zm, __shfl_xor_sync(FULL, zm, o));
        const int found = ng - ntested;
        test_new();
        bool full = !s_active || kcnt == HEAVY_K;
        full = __all_sync(FULL, full);
        float est = s_active ? kmax : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) est = fmaxf(est, __shfl_xor_sync(FULL, est, o));
        const float want = est * P.slab_margin;
        if (n_out == 0) {
            complete = true;
            break;
        }
        if (full && want <= zm) {   // nothing nearer than the wanted cap is deferred
            cap_valid = zm;
            break;
        }
        // adapt the step to the yield of the last pass
        // (nothing listed yet: the walk is still in front of the scene, large steps cost nothing - a pass that
        // jumps into the dense part overflows and is repeated with a quarter of the step)
        if (found < 24) step = fminf(step * (ng == 0 ? 2.0f : 1.5f), ng == 0 ? 0.5f : P.slab_step_max);
        else if (found > 96) step = fmaxf(step * 0.5f, 0.004f);
        if (ng > 0) step = fminf(step, P.slab_step_max);
        // the deferred list of the last pass becomes the input of this one; a pass that overflows the shared list is
        // repeated with a quarter of the step (its input is untouched)
        { int2* t = din; din = dout; dout = t; }
        n_in = n_out;
        const int ng0 = ng;
#pragma unroll 1
        for (;;) {
            if (++iters > HEAVY_MAX_ITERS) {
                failed = true;
                why = 3;
                break;
            }
            // geometric raise from the nearest deferred depth, never beyond what the sample rays ask for (their
            // 16th hits so far may lie far behind the slab - a huge blob whose box begins in it - and move in as
            // nearer candidates are found)
            float cap = zm > 0.0f ? zm * (1.0f + step) : 1e-30f;
            const bool last = full && want * 1.000001f <= cap;
            if (last) cap = want * 1.000001f;
            ng = ng0;
            n_out = 0;
            top = 0;
            zmin = INFINITY;
            failed = false;
            __syncwarp();
            // entries below the cap are expanded (internal nodes, 32 at a time, each batch drained depth first) or
            // listed (leaves), the rest is deferred again; the next batch is in flight while this one is drained
            int2 e_next = make_int2(0, 0);
            if (lane < n_in) e_next = __ldcg(din + lane);
#pragma unroll 1
            for (int base = 0; base < n_in && !failed; base += 32) {
                const int2 e = e_next;
                const bool have = base + lane < n_in;
                if (base + 32 + lane < n_in) e_next = __ldcg(din + base + 32 + lane);
                const float z = __int_as_float(e.y);
                const bool nearb = have && z < cap;
                const bool farb = have && !nearb;
                const unsigned mI = __ballot_sync(FULL, nearb && e.x >= 0), mL = __ballot_sync(FULL, nearb && e.x < 0),
                               mF = __ballot_sync(FULL, farb);
                if (ng + __popc(mL) > HEAVY_LIST_LIMIT || n_out + __popc(mF) > dcap) {
                    failed = true;
                    why = ng + __popc(mL) > HEAVY_LIST_LIMIT ? 1 : 2;
                    break;
                }
                if (nearb && e.x >= 0) ws.stack[top + __popc(mI & lt_mask)] = e.x;
                top += __popc(mI);
                if (nearb && e.x < 0) ws.glist[ng + __popc(mL & lt_mask)] = ~e.x;
                ng += __popc(mL);
                if (farb) {
                    dout[n_out + __popc(mF & lt_mask)] = e;
                    zmin = fminf(zmin, z);
                }
                n_out += __popc(mF);
                __syncwarp();
                // the activated nodes collect on the stack and are expanded 16 at a time; above the batch threshold
                // (<= 32 entries over it here: within the bound of tile_lists.cuh) the stack is drained first
                if (top > P.lists_single) drain(cap);
            }
            if (!failed) drain(cap);
            if (!failed) break;
            // overflow: with the cap the sample rays ask for there is no smaller step to try
            if (last || step < 0.0005f || why != 1) break;
            step *= 0.25f;
            ST(HS.retries += 1);
        }
    }
    ST(HS.iters += (unsigned)iters);
    ST(HS.groups += 1);
    ST(S.st_max_list = max(S.st_max_list, (unsigned)ng));

    // ---- pool chunks (as in lists_group) -------------------------------------------------------------------------
    auto alloc_chunk = [&]() -> int {
        if (S.slab_next == S.slab_end) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(P.counters + CTR_POOL, (unsigned)SLAB_CHUNKS);
            base = __shfl_sync(FULL, base, 0);
            if (base + (unsigned)SLAB_CHUNKS > (unsigned)P.pool_chunks) return -1;
            S.slab_next = (int)base;
            S.slab_end = S.slab_next + SLAB_CHUNKS;
        }
        return S.slab_next++;
    };
    auto write_chunk = [&](const int* q, int m, int& ncq, int& head, int& count) -> bool {
        const int chunk = alloc_chunk();
        if (chunk < 0) return false;
        int v = head;
        if (lane < m) v = q[ncq - m + lane];
        if (lane < m || lane == CHUNK_INTS - 1) P.pool[(int64_t)chunk * CHUNK_INTS + lane] = v;
        head = chunk;
        ncq -= m;
        count += m;
        __syncwarp();
        return true;
    };
    auto finish_tile = [&](int tile, int head, int count, float limit, bool ok) {
        if (lane == 0) {
            TileDesc d;
            d.head = head;
            d.count = !ok ? -1 : (limit < INFINITY ? (count | TILE_CAPPED) : count);
            P.desc[tile] = d;
            // (1 - 1e-5: the depths are float32 lower bounds, see the header of shade.cuh's completeness check)
            if (ok && limit < INFINITY) P.tile_cap[tile] = limit * (1.0f - 1e-5f);
            if (!ok) P.fallback_tiles[atomicAdd(P.counters + CTR_FALLBACK, 1u)] = tile;
        }
        ST(S.st_cands += ok ? (unsigned)count : 0u);
    };

    bool valid[TILES_PER_GROUP];
#pragma unroll
    for (int t = 0; t < TILES_PER_GROUP; ++t)
        valid[t] = gi0 + (t / GROUP_TJ) * TILE_I < xe && gj0 + (t % GROUP_TJ) * TILE_J < ye;

    if (failed) {
        ST(HS.failed += 1);
        ST(HS.fail_list += why == 1);
        ST(HS.fail_defer += why == 2);
        ST(HS.fail_passes += why == 3);
        if (lane == 0) atomicAdd(P.counters + CTR_HEAVY_FAILED, 1u);
#pragma unroll
        for (int t = 0; t < TILES_PER_GROUP; ++t)
            if (valid[t]) finish_tile(group * TILES_PER_GROUP + t, -1, 0, INFINITY, false);
        return;
    }

    // ---- per-tile depth limits: margin x the farthest 16th hit of the tile's own sample rays ------------------------
    float limit[TILES_PER_GROUP];
#pragma unroll
    for (int t = 0; t < TILES_PER_GROUP; ++t) {
        const bool mine = s_active && s_tile == t;
        float m = mine ? kmax : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
        if (!__any_sync(FULL, mine)) m = INFINITY;   // (a sliver of a tile at the image border: no sample ray, no guess)
        limit[t] = fminf(m * P.slab_margin * 1.000001f, complete ? INFINITY : cap_valid);
    }

    // ---- fused four-tile filter (lists_group), plus depth < the tile's limit ---------------------------------------
    float pnx[6], pny[6], pnz[6], pd[6];
    {
        const float ox = (float)cam.o[0], oy = (float)cam.o[1], oz = (float)cam.o[2];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            d3 n;
            if (k < 3) n = cam_rot(cam, 1.0, 0.0, ((double)(gi0 + k * TILE_I) - 0.5 * cam.W) * cam.ifx);
            else n = cam_rot(cam, 0.0, 1.0, ((double)(gj0 + (k - 3) * TILE_J) - 0.5 * cam.H) * cam.ify);
            pnx[k] = (float)n.x; pny[k] = (float)n.y; pnz[k] = (float)n.z;
            pd[k] = pnx[k] * ox + pny[k] * oy + pnz[k] * oz;
        }
    }
    int head[TILES_PER_GROUP], count[TILES_PER_GROUP], ncq[TILES_PER_GROUP];
    bool ok[TILES_PER_GROUP];
#pragma unroll
    for (int t = 0; t < TILES_PER_GROUP; ++t) {
        head[t] = -1; count[t] = 0; ncq[t] = 0; ok[t] = true;
    }
#pragma unroll 1
    for (int gpos = 0; gpos < ng; gpos += 32) {
        const int idx = gpos + lane;
        int s = 0;
        float z = INFINITY;
        bool lo[6], hi[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) lo[k] = hi[k] = false;
        if (idx < ng) {
            s = ws.glist[idx];
            float4 a, b;
            ldg256(P.leafbox + (int64_t)s * 2, a, b);
            const float2 r01 = __half22float2(*reinterpret_cast<const __half2*>(&b.z));
            const float2 r2 = __half22float2(*reinterpret_cast<const __half2*>(&b.w));
#pragma unroll
            for (int k = 0; k < 6; ++k) plane_side(pnx[k], pny[k], pnz[k], pd[k], a, b, r01, r2, lo[k], hi[k]);
            z = box_depth(a.x, a.y, a.z, a.w, b.x, b.y);
        }
        ST(S.st_nodes += (unsigned)min(32, ng - gpos));
        ST(S.st_steps += 1);
#pragma unroll
        for (int t = 0; t < TILES_PER_GROUP; ++t) {
            const int ta = t / GROUP_TJ, tb = 3 + t % GROUP_TJ;
            const bool h = valid[t] && ok[t] && lo[ta] && hi[ta + 1] && lo[tb] && hi[tb + 1] && z < limit[t];
            const unsigned mh = __ballot_sync(FULL, h);
            if (h) ws.cq[t * CQ_TILE + ncq[t] + __popc(mh & lt_mask)] = s;
            ncq[t] += __popc(mh);
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < TILES_PER_GROUP; ++t)
            if (ncq[t] >= CHUNK_IDS) ok[t] = write_chunk(ws.cq + t * CQ_TILE, CHUNK_IDS, ncq[t], head[t], count[t]);
    }
#pragma unroll
    for (int t = 0; t < TILES_PER_GROUP; ++t) {
        if (!valid[t]) continue;
        if (ok[t] && ncq[t] > 0) ok[t] = write_chunk(ws.cq + t * CQ_TILE, ncq[t], ncq[t], head[t], count[t]);
        finish_tile(group * TILES_PER_GROUP + t, head[t], count[t], limit[t], ok[t]);
    }
#undef ST
}

}  // namespace rtgs_dev
