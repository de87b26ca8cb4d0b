// tile_lists.cu — k_tile_lists: the BVH traversal of the render path (scene.py:406-450), done once per
// frame for whole pixel tiles instead of once per ray and per compositing step.
//
// All primary rays share the camera origin, so a rectangle of pixels is a thin pyramid bounded by 4
// planes through the origin, and "every Gaussian whose bound some ray of the rectangle hits" is a
// frustum query on the LBVH.  A warp owns one 8x16-pixel GROUP at a time (persistent threads, groups
// pulled from an atomic counter in 32x32-pixel macro-tile order):
//   1. group traversal - up to 16 nodes are popped from a shared-memory stack per step; a pair of lanes
//      shares a node and each lane tests the two boxes in the record of its child (two-level nodes,
//      lbvh.cu: one step descends two tree levels) against the 4 planes; survivors are compacted with
//      ballot/popc (internal -> stack, leaves -> the group's candidate list in shared memory);
//   2. the list is filtered for the group's four 4x8-pixel tiles at once - the 2x2 tiles are bounded by
//      3 + 3 planes, one 32-byte leaf record per lane, six exact ellipsoid-vs-plane support tests decide
//      all four tiles - and streamed to the global list pool in 128-byte chunks (render_common.cuh),
//      where k_shade_tiles picks it up.
// A group whose list would overflow shared memory traverses the LBVH per tile instead (leaves stream
// straight to the pool).  A tile whose list does not fit the pool is handed to the fused kernel
// (render.cu) through the fallback list; nothing is ever dropped.
//
// The kernel holds no per-ray state: 64 registers and 5.75 KB of shared memory per warp, 32 warps per SM.
// It is bound by instruction issue and the dependent node fetches.
#include "render_common.cuh"

using namespace rtgs_dev;

namespace {

#ifndef K1_CTAS
#define K1_CTAS 4
#endif
constexpr int STACK_CAP = 256;
constexpr int STACK_SINGLE = STACK_CAP - 144;   // above this pop one node at a time: growth/step <= 48, then <= +3 per
                                                // level over <= 31 two-level steps (tree depth <= 62)
constexpr int GLIST_CAP = 960;
constexpr int CQ_TILE = 64;                      // per-tile output queue of the fused four-tile filter (< 31 + 32)
constexpr int CQ_CAP = TILES_PER_GROUP * CQ_TILE;   // the per-tile traversal uses it as one queue (< 31 + 64)
static_assert(GLIST_CAP >= STACK_CAP, "per-tile traversal keeps its stack in the group list");

struct __align__(16) WarpShared {
    int stack[STACK_CAP];
    int glist[GLIST_CAP];
    int cq[CQ_CAP];
};

template <bool STATS>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, K1_CTAS) k_tile_lists(const __grid_constant__ RenderParams P) {
    __shared__ WarpShared smem[WARPS_PER_CTA];
    WarpShared& ws = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const CamD& cam = P.cam;
    const int xe = P.x0 + P.w, ye = P.y0 + P.h;

    unsigned long long st_nodes = 0, st_steps = 0, st_cands = 0;
#define ST(expr) do { if (STATS) { expr; } } while (0)

    int slab_next = 0, slab_end = 0;   // chunks this warp owns in the pool

    // one traversal step, two tree levels deep: pop <= 16 nodes from `stk`; a PAIR of lanes shares a node, lane
    // parity picks the record of its left or right child in the two-level node (lbvh.cu: k_pack_nodes4), i.e. each
    // lane tests two GRANDCHILD boxes against `fr`; survivors are compacted with ballot/popc: internal
    // grandchildren back onto `stk`, leaves into `dst`.  (The step is bound by the dependent node fetch, so
    // halving the number of levels is what counts; the extra box tests under a culled child are cheap.)
    auto traverse_step = [&](int* stk, int& top, int* dst, int& nd, const Frustum& fr) {
        const int take = top > STACK_SINGLE ? 1 : min(16, top);
        int node = -1;
        if ((lane >> 1) < take) node = stk[top - 1 - (lane >> 1)];
        top -= take;
        __syncwarp();
        bool h0 = false, h1 = false;
        int c0 = 0, c1 = 0;
        if (node >= 0) {
            const float4* rec = P.nodes4 + (int64_t)node * 8 + (lane & 1) * 4;
            float4 a, b, c, d;
            ldg256(rec + 0, a, b);
            ldg256(rec + 2, c, d);
            c0 = __float_as_int(d.x);
            c1 = __float_as_int(d.y);
            h0 = box_in_frustum(fr, a.x, a.y, a.z, a.w, b.x, b.y);
            h1 = box_in_frustum(fr, b.z, b.w, c.x, c.y, c.z, c.w);
        }
        ST(st_nodes += 4ull * (unsigned)take);
        ST(st_steps += 1);
        const unsigned mI0 = __ballot_sync(FULL, h0 && c0 >= 0), mI1 = __ballot_sync(FULL, h1 && c1 >= 0);
        const unsigned mL0 = __ballot_sync(FULL, h0 && c0 < 0), mL1 = __ballot_sync(FULL, h1 && c1 < 0);
        if (h0 && c0 >= 0) stk[top + __popc(mI0 & lt_mask)] = c0;
        const int topa = top + __popc(mI0);
        if (h1 && c1 >= 0) stk[topa + __popc(mI1 & lt_mask)] = c1;
        top = topa + __popc(mI1);
        if (h0 && c0 < 0) dst[nd + __popc(mL0 & lt_mask)] = ~c0;
        const int nda = nd + __popc(mL0);
        if (h1 && c1 < 0) dst[nda + __popc(mL1 & lt_mask)] = ~c1;
        nd = nda + __popc(mL1);
        __syncwarp();
    };

    // take one chunk from the pool (-1: exhausted)
    auto alloc_chunk = [&]() -> int {
        if (slab_next == slab_end) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(P.counters + CTR_POOL, (unsigned)SLAB_CHUNKS);
            base = __shfl_sync(FULL, base, 0);
            if (base + (unsigned)SLAB_CHUNKS > (unsigned)P.pool_chunks) return -1;
            slab_next = (int)base;
            slab_end = slab_next + SLAB_CHUNKS;
        }
        return slab_next++;
    };

    // write the last `m` entries of cq as one chunk linked in front of `head`
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

#pragma unroll 1
    for (;;) {
        int group = 0;
        if (lane == 0) group = (int)atomicAdd(P.counters + CTR_WORK, 1u);
        group = work_to_id(P, __shfl_sync(FULL, group, 0), P.macro_cols * GROUPS_PER_MACRO, P.ntiles / TILES_PER_GROUP);
        if (group * TILES_PER_GROUP >= P.ntiles) break;
        int gi0, gj0;
        if (!group_origin(P, group, gi0, gj0) || gi0 >= xe || gj0 >= ye) continue;

        // ---- group traversal: the LBVH once for the 8x16-pixel frustum ---------------------------
        int ng = 0;
        bool per_tile = false;
        {
            Frustum fg;
            make_frustum(cam, gi0, min(gi0 + GPX_I, xe), gj0, min(gj0 + GPX_J, ye), fg);
            int top = 1;
            if (lane == 0) ws.stack[0] = 0;
            __syncwarp();
#pragma unroll 1
            while (top > 0) {
                if (ng > min(GLIST_CAP - 64, P.heavy_limit)) {   // too many candidates for the shared list
                    per_tile = true;
                    break;
                }
                traverse_step(ws.stack, top, ws.glist, ng, fg);
            }
        }

        // a tile's list is complete: descriptor, or the fallback list if the pool ran out
        auto finish_tile = [&](int tile, int head, int count, bool ok) {
            if (lane == 0) {
                TileDesc d;
                d.head = head;
                d.count = ok ? count : -1;
                P.desc[tile] = d;
                if (!ok) P.fallback_tiles[atomicAdd(P.counters + CTR_FALLBACK, 1u)] = tile;
            }
            ST(st_cands += ok ? (unsigned)count : 0u);
        };

        if (!per_tile) {
            // ---- fused filter of the group's candidates for its four tiles ------------------------------
            // The 2x2 tiles are bounded by 3 + 3 planes through the camera origin (pixel edges gi0, +4, +8 and
            // gj0, +8, +16); a plane serves the tile on either side, so six exact support tests per Gaussian
            // (render_common.cuh: plane_side) decide all four tiles, and every leaf record is fetched once.
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
            bool ok[TILES_PER_GROUP], valid[TILES_PER_GROUP];
#pragma unroll
            for (int t = 0; t < TILES_PER_GROUP; ++t) {
                head[t] = -1; count[t] = 0; ncq[t] = 0; ok[t] = true;
                valid[t] = gi0 + (t / GROUP_TJ) * TILE_I < xe && gj0 + (t % GROUP_TJ) * TILE_J < ye;
            }
#pragma unroll 1
            for (int gpos = 0; gpos < ng; gpos += 32) {
                const int idx = gpos + lane;
                int s = 0;
                bool lo[6], hi[6];   // the ellipsoid reaches the >= side / the <= side of plane k
#pragma unroll
                for (int k = 0; k < 6; ++k) lo[k] = hi[k] = false;
                if (idx < ng) {
                    s = ws.glist[idx];
                    float4 a, b;
                    ldg256(P.leafbox + (int64_t)s * 2, a, b);
                    const float2 r01 = __half22float2(*reinterpret_cast<const __half2*>(&b.z));
                    const float2 r2 = __half22float2(*reinterpret_cast<const __half2*>(&b.w));
#pragma unroll
                    for (int k = 0; k < 6; ++k)
                        plane_side(pnx[k], pny[k], pnz[k], pd[k], a, b, r01, r2, lo[k], hi[k]);
                }
                ST(st_nodes += (unsigned)min(32, ng - gpos));
                ST(st_steps += 1);
#pragma unroll
                for (int t = 0; t < TILES_PER_GROUP; ++t) {
                    const int ta = t / GROUP_TJ, tb = 3 + t % GROUP_TJ;
                    const bool h = valid[t] && ok[t] && lo[ta] && hi[ta + 1] && lo[tb] && hi[tb + 1];
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
                finish_tile(group * TILES_PER_GROUP + t, head[t], count[t], ok[t]);
            }
            __syncwarp();
        } else if (P.heavy_fused) {
            // ---- the group's list did not fit shared memory: its frustum holds ~1000 Gaussians or more, typically
            // because it looks along a surface.  Listing them all is the wrong plan - the rays will have their K
            // nearest hits after a small part of them - so the tiles go to the fused kernel, which traverses near
            // first and prunes by distance as its hit buffers fill (render.cu)
#pragma unroll 1
            for (int sub = 0; sub < TILES_PER_GROUP; ++sub) {
                const int i0 = gi0 + (sub / GROUP_TJ) * TILE_I, j0 = gj0 + (sub % GROUP_TJ) * TILE_J;
                if (i0 >= xe || j0 >= ye) continue;
                finish_tile(group * TILES_PER_GROUP + sub, -1, 0, false);
            }
        } else {
            // ---- (RTGS_HEAVY_FUSED=0) one traversal per tile, leaves stream out -------
#pragma unroll 1
            for (int sub = 0; sub < TILES_PER_GROUP; ++sub) {
                const int i0 = gi0 + (sub / GROUP_TJ) * TILE_I, j0 = gj0 + (sub % GROUP_TJ) * TILE_J;
                if (i0 >= xe || j0 >= ye) continue;
                Frustum fr;
                make_frustum(cam, i0, i0 + TILE_I, j0, j0 + TILE_J, fr);
                int head = -1, count = 0, ncq = 0;
                bool ok = true;
                int top = 1;
                if (lane == 0) ws.glist[0] = 0;
                __syncwarp();
#pragma unroll 1
                while (top > 0 && ok) {
                    traverse_step(ws.glist, top, ws.cq, ncq, fr);
#pragma unroll 1
                    while (ncq >= CHUNK_IDS && ok) ok = write_chunk(ws.cq, CHUNK_IDS, ncq, head, count);
                }
                if (ok && ncq > 0) ok = write_chunk(ws.cq, ncq, ncq, head, count);
                finish_tile(group * TILES_PER_GROUP + sub, head, count, ok);
                __syncwarp();
            }
        }
    }

    if (STATS && P.stats && lane == 0) {
        if (st_nodes) atomicAdd(P.stats + ST_NODES, st_nodes);
        if (st_steps) atomicAdd(P.stats + ST_STEPS, st_steps);
        if (st_cands) atomicAdd(P.stats + ST_CANDS, st_cands);
    }
#undef ST
}

template <bool STATS>
int launch(rtgs_scene* s, const RenderParams& P, cudaStream_t stream) {
    static int blocks_per_sm[16] = {0};
    int dev = s->device;
    if (dev < 0 || dev >= 16) dev = 0;
    if (blocks_per_sm[dev] == 0) {
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tile_lists<STATS>, WARPS_PER_CTA * 32, 0));
        if (nb < 1) {
            rtgs_set_error("k_tile_lists does not fit on an SM");
            return RTGS_ERR_CUDA;
        }
        blocks_per_sm[dev] = nb;
    }
    int grid = s->sm_count * blocks_per_sm[dev];
    const int need = (P.ntiles / TILES_PER_GROUP + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    k_tile_lists<STATS><<<grid, WARPS_PER_CTA * 32, 0, stream>>>(P);
    CUDA_TRY(cudaGetLastError());
    return RTGS_OK;
}

}  // namespace

int rtgs_launch_tile_lists(rtgs_scene* s, const RenderParams& P, cudaStream_t stream, bool want_stats) {
    return want_stats ? launch<true>(s, P, stream) : launch<false>(s, P, stream);
}
