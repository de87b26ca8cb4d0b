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
// ballot/popc (internal children -> stack, leaves -> candidate queue).  Candidates are staged 32 at
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
#include "common.cuh"
#include "gsmath.cuh"

namespace {

constexpr int TILE_I = 4, TILE_J = 8;    // pixels per warp tile; lane = li * TILE_J + lj
constexpr int MACRO_I = 8, MACRO_J = 4;  // tiles per 32x32-pixel macro tile (scheduling locality)
constexpr int WARPS_PER_CTA = 8;
#ifndef RTGS_STACK_CAP
#define RTGS_STACK_CAP 512
#endif
#ifndef RTGS_BATCH
#define RTGS_BATCH 32
#endif
constexpr int STACK_CAP = RTGS_STACK_CAP;
constexpr int STACK_SINGLE = STACK_CAP - 100;   // growth/step <= 32, then DFS depth <= 62  // above this, pop one node at a time (DFS bound)
constexpr int CQ_CAP = 96;
constexpr int BATCH = RTGS_BATCH;
constexpr int REC_Q = 5;                 // quads per staged record (80-byte stride: conflict-free gathers)
constexpr unsigned FULL = 0xffffffffu;

struct RenderParams {
    const float4* nodes;
    const float4* geo;
    const float4* shp;
    const float4* raw;
    CamD cam;
    int x0, y0, w, h;
    int macro_cols;  // macro tiles along j
    int ntiles;
    int depth;
    float t_cut;
    int accumulate, full_pitch, has_sh;
    float* out_rgb;
    float* out_T;
    unsigned int* tile_counter;
    unsigned long long* stats;
};

struct __align__(16) TraversalScratch {
    // precise record: {W00 W01 W02 W10} {W11 W12 W20 W21} {W22 e0.xyz} {g0.xyz t_c} {opacity, s, band, -}
    float4 rec[BATCH][REC_Q];
    // coarse test: S(a,b) = c0 + a (c1 + a c3 + b c4) + b (c2 + b c5) < 0  <=>  possibly q < 3 + band
    float4 poly[BATCH][2];
    int stack[STACK_CAP];
    int cq[CQ_CAP];
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

enum { ST_RAYS = 0, ST_RAYS_HIT, ST_LAYERS, ST_NODES, ST_CANDS, ST_PAIRS, ST_F64, ST_TILES, ST_STEPS, ST_INSERTS, ST_COUNT = 12 };

// ---- float64 exact evaluation from raw parameters (rare path) ---------------------------------
__device__ __noinline__ ExactHit exact_eval(const float4* __restrict__ raw, const CamD& cam, int s,
                                            int pi, int pj) {
    float4 a = __ldg(raw + (int64_t)s * 3 + 0), b = __ldg(raw + (int64_t)s * 3 + 1),
           c = __ldg(raw + (int64_t)s * 3 + 2);
    double p[3] = {a.x, a.y, a.z};
    double q[4] = {a.w, b.x, b.y, b.z};
    double sc[3] = {b.w, c.x, c.y};
    d3 d = cam_dir(cam, (double)pi + 0.5, (double)pj + 0.5);
    return exact_intersect(p, q, sc, cam.o, d);
}

// Exact ordering of two hits of one ray whose float32 entry distances are within rounding of each
// other: float64 t1 from the raw parameters, ties broken by sorted position (rare path).
__device__ __noinline__ bool exact_less(const float4* __restrict__ raw, const CamD& cam, int sa, int sb, int pi,
                                        int pj) {
    const ExactHit a = exact_eval(raw, cam, sa, pi, pj);
    const ExactHit b = exact_eval(raw, cam, sb, pi, pj);
    return a.t1 < b.t1 || (a.t1 == b.t1 && sa < sb);
}

// ---- SH basis, gaussian.py:149-163 (incl. the `5z^2 - 3z` term at :160 exactly as coded) ------
__device__ __forceinline__ void sh_basis(float x, float y, float z, float (&Y)[15]) {
    const float c0 = 0.9772050238058398f;   // sqrt(3/pi)
    const float c1 = 2.1850968611841584f;   // sqrt(15/pi)
    const float c2 = 1.2615662610100802f;   // sqrt(5/pi)
    const float c3 = 2.360174359706574f;   // sqrt(35/(2pi))
    const float c4 = 5.781222885281108f;   // sqrt(105/pi)
    const float c5 = 1.828183197857863f;   // sqrt(21/(2pi))
    const float c6 = 1.4927053303604616f;   // sqrt(7/pi)
    const float xx = x * x, yy = y * y, zz = z * z;
    Y[0] = 0.5f * c0 * y;
    Y[1] = 0.5f * c0 * z;
    Y[2] = 0.5f * c0 * x;
    Y[3] = 0.5f * c1 * x * y;
    Y[4] = 0.5f * c1 * y * z;
    Y[5] = 0.25f * c2 * (3.0f * zz - 1.0f);
    Y[6] = 0.5f * c1 * x * z;
    Y[7] = 0.25f * c1 * (xx - yy);
    Y[8] = 0.25f * c3 * y * (3.0f * xx - yy);
    Y[9] = 0.5f * c4 * x * y * z;
    Y[10] = 0.25f * c5 * y * (5.0f * zz - 1.0f);
    Y[11] = 0.25f * c6 * (5.0f * zz - 3.0f * z);
    Y[12] = 0.25f * c5 * x * (5.0f * zz - 1.0f);
    Y[13] = 0.25f * c4 * (xx - yy) * z;
    Y[14] = 0.25f * c3 * x * (xx - 3.0f * yy);
}

struct Frustum {
    // 4 planes through the camera origin o, inward normals n[k]; a box (centre c, half size h) is
    // outside plane k iff n.c + |n|.h - n.o < 0.
    float nx[4], ny[4], nz[4];
    float ax[4], ay[4], az[4];
    float d[4];   // n.o
};

__device__ __forceinline__ bool box_in_frustum(const Frustum& f, float cx, float cy, float cz, float hx,
                                               float hy, float hz) {
    bool in = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float v = f.nx[k] * cx + f.ny[k] * cy + f.nz[k] * cz + f.ax[k] * hx + f.ay[k] * hy + f.az[k] * hz;
        in = in && (v >= f.d[k]);  // NaN / -inf (empty box) -> false
    }
    return in;
}

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// STATS = true compiles the per-render counters in (rtgs_render with a stats pointer); the timed path
// uses STATS = false so that the ten 64-bit counters do not occupy registers.
template <int K, bool STATS>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, (K <= 16 ? 2 : 1)) k_render(const __grid_constant__ RenderParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpShared<K>& ws = reinterpret_cast<WarpShared<K>*>(smem_raw)[threadIdx.x >> 5];
    TraversalScratch& tr = ws.t;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const CamD& cam = P.cam;
    const float ox = (float)cam.o[0], oy = (float)cam.o[1], oz = (float)cam.o[2];

    unsigned long long st_nodes = 0, st_cands = 0, st_pairs = 0, st_f64 = 0, st_layers = 0, st_hit = 0,
                       st_rays = 0, st_tiles = 0, st_steps = 0, st_ins = 0;
#define ST(expr) do { if (STATS) { expr; } } while (0)

#pragma unroll 1
    for (;;) {
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(P.tile_counter, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= P.ntiles) break;
        const int macro = tile / (MACRO_I * MACRO_J), local = tile % (MACRO_I * MACRO_J);
        const int mi = macro / P.macro_cols, mj = macro % P.macro_cols;
        const int ti = mi * MACRO_I + local / MACRO_J, tj = mj * MACRO_J + local % MACRO_J;
        const int i0 = P.x0 + ti * TILE_I, j0 = P.y0 + tj * TILE_J;
        if (i0 >= P.x0 + P.w || j0 >= P.y0 + P.h) continue;
        const int pi = i0 + lane / TILE_J, pj = j0 + lane % TILE_J;
        const bool active = pi < P.x0 + P.w && pj < P.y0 + P.h;

        // ---- rays (float64 setup, camera.py:46-52) ---------------------------------------------
        // Image-plane coordinates: px = (i + 0.5 - W/2)/fx.  Tile centre (px0, py0); own offset
        // (a, b) = (px - px0, py - py0).  UNNORMALISED directions are linear in (a, b):
        //   D(a,b) = R (px0 + a, py0 + b, -1) = D0 + a Rx + b Ry,
        // the reference's direction is d = D / sqrt(px^2 + py^2 + 1); d0 likewise, d = d0 + delta.
        const double px0 = ((double)i0 + 0.5 * TILE_I - 0.5 * cam.W) * cam.ifx;
        const double py0 = ((double)j0 + 0.5 * TILE_J - 0.5 * cam.H) * cam.ify;
        const d3 D0 = cam_rot(cam, px0, py0, -1.0);
        const double n0 = rsqrt(px0 * px0 + py0 * py0 + 1.0);
        const d3 d0 = d3make(D0.x * n0, D0.y * n0, D0.z * n0);
        const double inv_d0d0 = 1.0 / d3dot(d0, d0);
        const double ad = active ? ((double)(pi - i0) + 0.5 - 0.5 * TILE_I) * cam.ifx : 0.0;
        const double bd = active ? ((double)(pj - j0) + 0.5 - 0.5 * TILE_J) * cam.ify : 0.0;
        float dlx, dly, dlz;          // delta = d - d0
        float dnx, dny, dnz;          // normalize(d) for the SH basis (gaussian.py:200)
        {
            const double px = px0 + ad, py = py0 + bd;
            const double nn = rsqrt(px * px + py * py + 1.0);
            const d3 dw = d3make((D0.x + ad * cam.R[0] + bd * cam.R[1]) * nn, (D0.y + ad * cam.R[3] + bd * cam.R[4]) * nn,
                                 (D0.z + ad * cam.R[6] + bd * cam.R[7]) * nn);
            dlx = (float)(dw.x - d0.x); dly = (float)(dw.y - d0.y); dlz = (float)(dw.z - d0.z);
            const double il = rsqrt(d3dot(dw, dw));
            dnx = (float)(dw.x * il); dny = (float)(dw.y * il); dnz = (float)(dw.z * il);
        }
        const float pa = (float)ad, pb = (float)bd;
        const float a_max = (float)(0.5 * TILE_I * fabs(cam.ifx)), b_max = (float)(0.5 * TILE_J * fabs(cam.ify));
        float dl_max = sqrtf(dlx * dlx + dly * dly + dlz * dlz);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dl_max = fmaxf(dl_max, __shfl_xor_sync(FULL, dl_max, o));

        // ---- tile frustum: planes through the origin along the tile's pixel EDGES -------------
        Frustum fr;
        {
            const double pxl = ((double)i0 - 0.5 * cam.W) * cam.ifx, pxh = ((double)(i0 + TILE_I) - 0.5 * cam.W) * cam.ifx;
            const double pyl = ((double)j0 - 0.5 * cam.H) * cam.ify, pyh = ((double)(j0 + TILE_J) - 0.5 * cam.H) * cam.ify;
            d3 n[4];
            n[0] = cam_rot(cam, 1.0, 0.0, pxl);     // px >= pxl
            n[1] = cam_rot(cam, -1.0, 0.0, -pxh);   // px <= pxh
            n[2] = cam_rot(cam, 0.0, 1.0, pyl);     // py >= pyl
            n[3] = cam_rot(cam, 0.0, -1.0, -pyh);   // py <= pyh
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                fr.nx[k] = (float)n[k].x; fr.ny[k] = (float)n[k].y; fr.nz[k] = (float)n[k].z;
                fr.ax[k] = fabsf(fr.nx[k]); fr.ay[k] = fabsf(fr.ny[k]); fr.az[k] = fabsf(fr.nz[k]);
                fr.d[k] = fr.nx[k] * ox + fr.ny[k] * oy + fr.nz[k] * oz;
            }
        }

        // ---- per-lane hit buffer (shared memory, unsorted; replace-max once K entries are held) --
        int cnt = 0;
        float kmax_t = INFINITY;
        int kmax_slot = 0;

        // one pending candidate per lane; the precise test and the hit-only work (entry distance,
        // alpha, float64 refinement, buffer append) run in warp-wide rounds
        bool pend = false;
        int pend_c = 0;

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
                    const float4 a = __ldg(P.nodes + (int64_t)node * 4 + 0);
                    const float4 b = __ldg(P.nodes + (int64_t)node * 4 + 1);
                    const float4 c = __ldg(P.nodes + (int64_t)node * 4 + 2);
                    const float4 d = __ldg(P.nodes + (int64_t)node * 4 + 3);
                    c0 = __float_as_int(d.x);
                    c1 = __float_as_int(d.y);
                    h0 = box_in_frustum(fr, a.x, a.y, a.z, a.w, b.x, b.y);
                    h1 = box_in_frustum(fr, b.z, b.w, c.x, c.y, c.z, c.w);
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
                    const int s = tr.cq[ncq + lane];
                    const float4 g0 = __ldg(P.geo + (int64_t)s * 4 + 0), g1 = __ldg(P.geo + (int64_t)s * 4 + 1),
                                 g2 = __ldg(P.geo + (int64_t)s * 4 + 2), g3 = __ldg(P.geo + (int64_t)s * 4 + 3);
                    const double W00 = g1.x, W01 = g1.y, W02 = g1.z, W10 = g1.w, W11 = g2.x, W12 = g2.y,
                                 W20 = g2.z, W21 = g2.w, W22 = g3.x;
                    auto Wmul = [&](const d3& v) {
                        return d3make(W00 * v.x + W01 * v.y + W02 * v.z, W10 * v.x + W11 * v.y + W12 * v.z,
                                      W20 * v.x + W21 * v.y + W22 * v.z);
                    };
                    const d3 v = d3make((double)g0.x - cam.o[0], (double)g0.y - cam.o[1], (double)g0.z - cam.o[2]);
                    // precise record: origin shifted to the closest point of the tile-centre ray
                    const double tc = d3dot(v, d0) * inv_d0d0;
                    const d3 e0 = Wmul(d3make(tc * d0.x - v.x, tc * d0.y - v.y, tc * d0.z - v.z));
                    const d3 gd = Wmul(d0);
                    const float wn = sqrtf(g1.x * g1.x + g1.y * g1.y + g1.z * g1.z + g1.w * g1.w + g2.x * g2.x +
                                           g2.y * g2.y + g2.z * g2.z + g2.w * g2.w + g3.x * g3.x);
                    const float eb = (float)sqrt(d3dot(e0, e0)) + fabsf((float)tc) * dl_max * wn;
                    const float band = 4e-6f * (3.0f + eb * eb);
                    tr.rec[lane][0] = g1;
                    tr.rec[lane][1] = g2;
                    tr.rec[lane][2] = make_float4(g3.x, (float)e0.x, (float)e0.y, (float)e0.z);
                    tr.rec[lane][3] = make_float4((float)gd.x, (float)gd.y, (float)gd.z, (float)tc);
                    tr.rec[lane][4] = make_float4(g0.w, __int_as_float(s), band, 0.0f);
                    // coarse quadratic: with o' = W (o - p) = -W v and G(a,b) = W D(a,b) = G0 + a Gx + b Gy,
                    //   q(a,b) = |o' x G|^2 / |G|^2 = N/Dn,  m = o' x G = M0 + a Mx + b My.
                    // S = N - (3 + band) Dn - margin, margin bounding the float32 evaluation error.
                    const d3 op = Wmul(d3make(-v.x, -v.y, -v.z));
                    const d3 G0 = Wmul(D0);
                    const d3 Gx = Wmul(d3make(cam.R[0], cam.R[3], cam.R[6]));
                    const d3 Gy = Wmul(d3make(cam.R[1], cam.R[4], cam.R[7]));
                    const d3 M0 = d3cross(op, G0), Mx = d3cross(op, Gx), My = d3cross(op, Gy);
                    const double lim = 3.0 + (double)band;
                    const double m00 = d3dot(M0, M0), m0x = d3dot(M0, Mx), m0y = d3dot(M0, My), mxx = d3dot(Mx, Mx),
                                 mxy = d3dot(Mx, My), myy = d3dot(My, My);
                    const double g00 = lim * d3dot(G0, G0), g0x = lim * d3dot(G0, Gx), g0y = lim * d3dot(G0, Gy),
                                 gxx = lim * d3dot(Gx, Gx), gxy = lim * d3dot(Gx, Gy), gyy = lim * d3dot(Gy, Gy);
                    const double am = a_max, bm = b_max;
                    // magnitude of the terms BEFORE cancellation (N and Dn parts separately)
                    const double E = m00 + g00 + 2.0 * am * (fabs(m0x) + fabs(g0x)) + 2.0 * bm * (fabs(m0y) + fabs(g0y)) +
                                     am * am * (mxx + gxx) + 2.0 * am * bm * (fabs(mxy) + fabs(gxy)) + bm * bm * (myy + gyy);
                    tr.poly[lane][0] = make_float4((float)(m00 - g00 - 2e-6 * E), (float)(2.0 * (m0x - g0x)),
                                                   (float)(2.0 * (m0y - g0y)), (float)(mxx - gxx));
                    tr.poly[lane][1] = make_float4((float)(2.0 * (mxy - gxy)), (float)(myy - gyy), 0.0f, 0.0f);
                }
                __syncwarp();
                ST(st_cands += (unsigned)m);
                ST(st_pairs += 32ull * (unsigned)m);
#pragma unroll 1
                for (int c = 0; c <= m; ++c) {
                    bool cand = false;
                    if (c < m) {
                        const float4 p0 = tr.poly[c][0];
                        const float2 p1 = *reinterpret_cast<const float2*>(&tr.poly[c][1]);
                        const float ta = fmaf(pa, p0.w, fmaf(pb, p1.x, p0.y));   // c1 + a c3 + b c4
                        const float tb = fmaf(pb, p1.y, p0.z);                   // c2 + b c5
                        const float S = fmaf(pa, ta, fmaf(pb, tb, p0.x));
                        cand = active && (S < 0.0f);
                    }
                    // flush when a lane gets a second candidate, and once at the end of the batch
                    // (the staged records are about to be overwritten)
                    if (__any_sync(FULL, pend && (cand || c == m))) {
                        if (pend) {
                            const float4* rc = tr.rec[pend_c];
                            const float4 r0 = rc[0], r1 = rc[1], r2 = rc[2], r3 = rc[3], ax = rc[4];
                            const float wx = r0.x * dlx + r0.y * dly + r0.z * dlz;
                            const float wy = r0.w * dlx + r1.x * dly + r1.y * dlz;
                            const float wz = r1.z * dlx + r1.w * dly + r2.x * dlz;
                            const float tc = r3.w;
                            const float dx = r3.x + wx, dy = r3.y + wy, dz = r3.z + wz;        // d' = W d
                            const float ex = r2.y + tc * wx, ey = r2.z + tc * wy, ez = r2.w + tc * wz;  // W (r(tc) - p)
                            const float A = dx * dx + dy * dy + dz * dz;
                            const float Bh = ex * dx + ey * dy + ez * dz;
                            const float mx = ey * dz - ez * dy, my = ez * dx - ex * dz, mz = ex * dy - ey * dx;
                            const float iA = rcp_approx(A);
                            float q = (mx * mx + my * my + mz * mz) * iA;   // min Mahalanobis^2 along the ray
                            // tau = -Bh/A - sqrt((3 - q)/A)  (near root relative to tc)
                            const float tau = -Bh * iA - sqrt_approx(fmaxf(3.0f - q, 0.0f) * iA);
                            float t1 = tc + tau;
                            bool hit = (q < 3.0f) && (t1 > 0.0f);
                            const bool near_q = fabsf(q - 3.0f) < ax.z;
                            const bool near_t = (q < 3.0f + ax.z) && fabsf(t1) <= 2e-6f * (fabsf(tc) + fabsf(tau));
                            if (near_q || near_t) {
                                const ExactHit e = exact_eval(P.raw, cam, __float_as_int(ax.y), pi, pj);
                                hit = e.hit && (e.t1 > 0.0);
                                q = (float)e.q;
                                t1 = (float)e.t1;
                                ST(st_f64 += 1);
                            }
                            if (hit) {
                                const float alpha = ax.x * __expf(-q);   // opacity * exp(-q)  (gaussian.py:197-198)
                                int slot = -1;
                                if (cnt < K) slot = cnt++;
                                else if (t1 < kmax_t) slot = kmax_slot;
                                if (slot >= 0) {
                                    ws.kb_t[slot][lane] = t1;
                                    ws.kb_i[slot][lane] = __float_as_int(ax.y);
                                    ws.kb_a[slot][lane] = alpha;
                                    if (cnt == K) {   // buffer full: track the farthest entry
                                        float mt = -INFINITY;
                                        int ms = 0;
#pragma unroll 4
                                        for (int k = 0; k < K; ++k) {
                                            const float t = ws.kb_t[k][lane];
                                            if (t > mt) { mt = t; ms = k; }
                                        }
                                        kmax_t = mt;
                                        kmax_slot = ms;
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
            sh_basis(dnx, dny, dnz, Y);
            const int nmine = min(cnt, P.depth);
            const int nloop = min(maxcnt, P.depth);
#pragma unroll 1
            for (int k = 0; k < nloop; ++k) {
                if (k < nmine && T >= P.t_cut) {
                    const int s = ws.c.so_i[k][lane];
                    const float alpha = ws.c.so_a[k][lane];
                    const float4 g3 = __ldg(P.geo + (int64_t)s * 4 + 3);
                    float r = g3.y, g = g3.z, b = g3.w;
                    if (P.has_sh) {
                        const float4* sp = P.shp + (int64_t)s * 12;
                        float v[48];
#pragma unroll
                        for (int f = 0; f < 12; ++f) {
                            const float4 x = __ldg(sp + f);
                            v[4 * f] = x.x; v[4 * f + 1] = x.y; v[4 * f + 2] = x.z; v[4 * f + 3] = x.w;
                        }
#pragma unroll
                        for (int j = 0; j < 15; ++j) {
                            r = fmaf(Y[j], v[3 * j + 0], r);
                            g = fmaf(Y[j], v[3 * j + 1], g);
                            b = fmaf(Y[j], v[3 * j + 2], b);
                        }
                    }
                    const float wgt = T * alpha;
                    cr = fmaf(wgt, r, cr);
                    cg = fmaf(wgt, g, cg);
                    cb = fmaf(wgt, b, cb);
                    T *= 1.0f - alpha;
                    ++nl;
                }
            }
        }
        // ---- framebuffer write: stage the tile in shared memory so that every store instruction
        // covers whole 32-byte sectors (each tile column is 8 pixels = 96 contiguous bytes) --------
        {
            float* ob = reinterpret_cast<float*>(&ws.kb_t[0][0]);
            __syncwarp();
            ob[lane * 3 + 0] = cr;
            ob[lane * 3 + 1] = cg;
            ob[lane * 3 + 2] = cb;
            __syncwarp();
            const int pitch = P.full_pitch ? cam.H : P.h;
            const int bi = P.full_pitch ? 0 : P.x0, bj = P.full_pitch ? 0 : P.y0;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int f = r * 32 + lane;
                const int row = f / (3 * TILE_J), col = f % (3 * TILE_J);
                const int qi = i0 + row, qj = j0 + col / 3;
                if (qi < P.x0 + P.w && qj < P.y0 + P.h) {
                    float* o = P.out_rgb + ((int64_t)(qi - bi) * pitch + (j0 - bj)) * 3 + col;
                    if (P.accumulate) *o += ob[f];
                    else *o = ob[f];
                }
            }
            __syncwarp();
        }
        if (active) {
            const int64_t idx = P.full_pitch ? ((int64_t)pi * cam.H + pj)
                                             : ((int64_t)(pi - P.x0) * P.h + (pj - P.y0));
            if (P.out_T) P.out_T[idx] = T;
            ST(st_rays += 1);
            ST(st_hit += nl > 0);
            ST(st_layers += (unsigned)nl);
        }
        ST(if (lane == 0) st_tiles += 1);
    }

    if (STATS && P.stats) {
        unsigned long long v[10] = {st_rays, st_hit, st_layers, 0, 0, 0, st_f64, st_tiles, 0, 0};
        // warp-uniform counters are taken from lane 0 only
        if (lane == 0) {
            v[ST_NODES] = st_nodes;
            v[ST_CANDS] = st_cands;
            v[ST_PAIRS] = st_pairs;
            v[ST_STEPS] = st_steps;
            v[ST_INSERTS] = st_ins;
        }
#pragma unroll
        for (int k = 0; k < 10; ++k) {
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
    if (grid < 1) grid = 1;
    k_render<K, STATS><<<grid, WARPS_PER_CTA * 32, smem, stream>>>(P);
    CUDA_TRY(cudaGetLastError());
    return RTGS_OK;
}

}  // namespace

int rtgs_launch_render(rtgs_scene* s, const rtgs_camera* cam, int x0, int y0, int w, int h, int depth,
                       float t_cut, int accumulate, int full_pitch, float* out_rgb, float* out_T,
                       cudaStream_t stream, bool want_stats) {
    RenderParams P;
    P.nodes = s->nodes;
    P.geo = s->geo;
    P.shp = s->shp;
    P.raw = s->raw;
    P.cam = make_camd(cam);
    P.x0 = x0; P.y0 = y0; P.w = w; P.h = h;
    const int mrows = (w + TILE_I * MACRO_I - 1) / (TILE_I * MACRO_I);
    const int mcols = (h + TILE_J * MACRO_J - 1) / (TILE_J * MACRO_J);
    P.macro_cols = mcols;
    P.ntiles = mrows * mcols * MACRO_I * MACRO_J;
    P.depth = depth;
    P.t_cut = t_cut;
    P.accumulate = accumulate;
    P.full_pitch = full_pitch;
    P.has_sh = s->has_sh ? 1 : 0;
    P.out_rgb = out_rgb;
    P.out_T = out_T;
    P.tile_counter = s->tile_counter;
    P.stats = want_stats ? s->stats_dev : nullptr;
    CUDA_TRY(cudaMemsetAsync(s->tile_counter, 0, sizeof(unsigned int), stream));
    if (want_stats) CUDA_TRY(cudaMemsetAsync(s->stats_dev, 0, 12 * sizeof(unsigned long long), stream));
    if (depth <= 16) return want_stats ? launch_render_k<16, true>(s, P, stream) : launch_render_k<16, false>(s, P, stream);
    return want_stats ? launch_render_k<32, true>(s, P, stream) : launch_render_k<32, false>(s, P, stream);
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
