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
// shared-memory reads.  Hits enter a per-lane sorted k-buffer held in registers (K = 16 or 32
// nearest entry distances); after the traversal the lane composites its k-buffer front to back.
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
constexpr int STACK_CAP = 1024;
constexpr int STACK_SINGLE = STACK_CAP - 160;  // above this, pop one node at a time (DFS bound)
constexpr int CQ_CAP = 96;
constexpr int BATCH = 32;
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

struct __align__(16) WarpShared {
    float4 rec[BATCH][4];  // {W00 W01 W02 W10} {W11 W12 W20 W21} {W22 e0.xyz} {g0.xyz t_c}
    float4 aux[BATCH];     // {opacity, sorted position (int bits), q band, unused}
    int stack[STACK_CAP];
    int cq[CQ_CAP];
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

// ---- register k-buffer: sorted ascending by entry distance, ties keep arrival order -----------
template <int K>
__device__ __forceinline__ void kb_insert(float (&kt)[K], int (&ki)[K], float (&ka)[K], float t, int id,
                                          float a) {
    if (!(t < kt[K - 1])) return;
#pragma unroll
    for (int s = K - 1; s >= 1; --s) {
        const bool up = t < kt[s - 1];
        const bool here = (!up) && (t < kt[s]);
        kt[s] = up ? kt[s - 1] : (here ? t : kt[s]);
        ki[s] = up ? ki[s - 1] : (here ? id : ki[s]);
        ka[s] = up ? ka[s - 1] : (here ? a : ka[s]);
    }
    if (t < kt[0]) {
        kt[0] = t;
        ki[0] = id;
        ka[0] = a;
    }
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

template <int K>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, (K <= 16 ? 2 : 1)) k_render(const __grid_constant__ RenderParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpShared& ws = reinterpret_cast<WarpShared*>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const CamD& cam = P.cam;
    const float ox = (float)cam.o[0], oy = (float)cam.o[1], oz = (float)cam.o[2];

    unsigned long long st_nodes = 0, st_cands = 0, st_pairs = 0, st_f64 = 0, st_layers = 0, st_hit = 0,
                       st_rays = 0, st_tiles = 0, st_steps = 0, st_ins = 0;

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

        // ---- rays (float64 setup, camera.py:46-52): centre ray d0, own ray d = d0 + delta -----
        const d3 d0 = cam_dir(cam, (double)i0 + 0.5 * TILE_I, (double)j0 + 0.5 * TILE_J);
        const double inv_d0d0 = 1.0 / d3dot(d0, d0);
        d3 dw = active ? cam_dir(cam, (double)pi + 0.5, (double)pj + 0.5) : d0;
        const float dlx = (float)(dw.x - d0.x), dly = (float)(dw.y - d0.y), dlz = (float)(dw.z - d0.z);
        const float dlen = sqrtf(dlx * dlx + dly * dly + dlz * dlz);
        float dl_max = dlen;
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

        // ---- k-buffer ------------------------------------------------------------------------
        float kt[K], ka[K];
        int ki[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            kt[k] = INFINITY;
            ka[k] = 0.0f;
            ki[k] = -1;
        }
        // one pending candidate hit per lane; the hit-only work (entry distance, alpha, float64
        // refinement, k-buffer insertion) runs in warp-wide rounds when some lane gets a second one
        bool pend = false;
        float pend_q = 0.0f, pend_iA = 0.0f, pend_Bh = 0.0f;
        int pend_c = 0;
        auto flush = [&]() {
            if (pend) {
                const float4 ax = ws.aux[pend_c];
                const float tc = ws.rec[pend_c][3].w;
                float q = pend_q;
                // disc/A^2 = (3 - q)/A  ->  tau = -Bh/A - sqrt((3 - q)/A)
                const float tau = -pend_Bh * pend_iA - sqrt_approx(fmaxf(3.0f - q, 0.0f) * pend_iA);
                float t1 = tc + tau;
                bool hit = (q < 3.0f) && (t1 > 0.0f);
                const bool near_q = fabsf(q - 3.0f) < ax.z;
                const bool near_t = fabsf(t1) <= 2e-6f * (fabsf(tc) + fabsf(tau));
                if (near_q || near_t) {
                    const ExactHit e = exact_eval(P.raw, cam, __float_as_int(ax.y), pi, pj);
                    hit = e.hit && (e.t1 > 0.0);
                    q = (float)e.q;
                    t1 = (float)e.t1;
                    st_f64 += 1;
                }
                // alpha = opacity * exp(-q)  (gaussian.py:197-198)
                if (hit) kb_insert<K>(kt, ki, ka, t1, __float_as_int(ax.y), ax.x * __expf(-q));
                pend = false;
            }
            st_ins += 1;
        };

        int top = 1, ncq = 0;
        if (lane == 0) ws.stack[0] = 0;
        __syncwarp();

        // ================================ traversal ==========================================
        while (top > 0 || ncq > 0) {
            if (top > 0) {
                const int take = top > STACK_SINGLE ? 1 : min(32, top);
                int node = -1;
                if (lane < take) node = ws.stack[top - 1 - lane];
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
                st_nodes += 2ull * (unsigned)take;
                st_steps += 1;
                const unsigned mI0 = __ballot_sync(FULL, h0 && c0 >= 0), mI1 = __ballot_sync(FULL, h1 && c1 >= 0);
                const unsigned mL0 = __ballot_sync(FULL, h0 && c0 < 0), mL1 = __ballot_sync(FULL, h1 && c1 < 0);
                if (h0 && c0 >= 0) ws.stack[top + __popc(mI0 & lt_mask)] = c0;
                const int topa = top + __popc(mI0);
                if (h1 && c1 >= 0) ws.stack[topa + __popc(mI1 & lt_mask)] = c1;
                top = topa + __popc(mI1);
                if (h0 && c0 < 0) ws.cq[ncq + __popc(mL0 & lt_mask)] = ~c0;
                const int ncqa = ncq + __popc(mL0);
                if (h1 && c1 < 0) ws.cq[ncqa + __popc(mL1 & lt_mask)] = ~c1;
                ncq = ncqa + __popc(mL1);
                __syncwarp();
            }
            // -------- candidate batch: stage (one lane each, float64) then test (all lanes) --
            while (ncq >= BATCH || (top == 0 && ncq > 0)) {
                const int m = min(BATCH, ncq);
                ncq -= m;
                if (lane < m) {
                    const int s = ws.cq[ncq + lane];
                    const float4 g0 = __ldg(P.geo + (int64_t)s * 4 + 0), g1 = __ldg(P.geo + (int64_t)s * 4 + 1),
                                 g2 = __ldg(P.geo + (int64_t)s * 4 + 2), g3 = __ldg(P.geo + (int64_t)s * 4 + 3);
                    const double W00 = g1.x, W01 = g1.y, W02 = g1.z, W10 = g1.w, W11 = g2.x, W12 = g2.y,
                                 W20 = g2.z, W21 = g2.w, W22 = g3.x;
                    const d3 v = d3make((double)g0.x - cam.o[0], (double)g0.y - cam.o[1], (double)g0.z - cam.o[2]);
                    const double tc = d3dot(v, d0) * inv_d0d0;
                    const d3 u = d3make(tc * d0.x - v.x, tc * d0.y - v.y, tc * d0.z - v.z);
                    const double e0x = W00 * u.x + W01 * u.y + W02 * u.z, e0y = W10 * u.x + W11 * u.y + W12 * u.z,
                                 e0z = W20 * u.x + W21 * u.y + W22 * u.z;
                    const double gx = W00 * d0.x + W01 * d0.y + W02 * d0.z, gy = W10 * d0.x + W11 * d0.y + W12 * d0.z,
                                 gz = W20 * d0.x + W21 * d0.y + W22 * d0.z;
                    // conservative bound on |e| over the tile for the float32 error band of q
                    const float wn = sqrtf(g1.x * g1.x + g1.y * g1.y + g1.z * g1.z + g1.w * g1.w + g2.x * g2.x +
                                           g2.y * g2.y + g2.z * g2.z + g2.w * g2.w + g3.x * g3.x);
                    const float eb = (float)sqrt(e0x * e0x + e0y * e0y + e0z * e0z) + fabsf((float)tc) * dl_max * wn;
                    const float band = 4e-6f * (3.0f + eb * eb);
                    ws.rec[lane][0] = g1;
                    ws.rec[lane][1] = g2;
                    ws.rec[lane][2] = make_float4(g3.x, (float)e0x, (float)e0y, (float)e0z);
                    ws.rec[lane][3] = make_float4((float)gx, (float)gy, (float)gz, (float)tc);
                    ws.aux[lane] = make_float4(g0.w, __int_as_float(s), band, 0.0f);
                }
                __syncwarp();
                st_cands += (unsigned)m;
                st_pairs += 32ull * (unsigned)m;
                for (int c = 0; c < m; ++c) {
                    const float4 r0 = ws.rec[c][0], r1 = ws.rec[c][1], r2 = ws.rec[c][2], r3 = ws.rec[c][3];
                    const float band = ws.aux[c].z;
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
                    const float q = (mx * mx + my * my + mz * mz) * iA;   // min Mahalanobis^2 along the ray
                    const bool cand = active && (q < 3.0f + band);
                    if (__any_sync(FULL, cand && pend)) flush();
                    if (cand) {
                        pend = true;
                        pend_q = q;
                        pend_iA = iA;
                        pend_Bh = Bh;
                        pend_c = c;
                    }
                }
                if (__any_sync(FULL, pend)) flush();   // the staged records are about to be overwritten
                __syncwarp();
            }
        }

        // ---- near-tie resolution: adjacent entries closer than a few ulp are ordered in f64 ---
#pragma unroll
        for (int k = 0; k + 1 < K; ++k) {
            if (kt[k + 1] < INFINITY && (kt[k + 1] - kt[k]) <= 2e-6f * fabsf(kt[k + 1])) {
                const ExactHit ea = exact_eval(P.raw, cam, ki[k], pi, pj);
                const ExactHit eb = exact_eval(P.raw, cam, ki[k + 1], pi, pj);
                st_f64 += 2;
                if (eb.t1 < ea.t1 || (eb.t1 == ea.t1 && ki[k + 1] < ki[k])) {
                    const float tt = kt[k]; kt[k] = kt[k + 1]; kt[k + 1] = tt;
                    const float ta = ka[k]; ka[k] = ka[k + 1]; ka[k + 1] = ta;
                    const int tid = ki[k]; ki[k] = ki[k + 1]; ki[k + 1] = tid;
                }
            }
        }

        // ================================ compositing ========================================
        // accum += T * alpha * rgb ; T *= 1 - alpha   (ray_tracer.py:96-98), rgb = color +
        // eval_sh(normalize(dir)) (gaussian.py:199-200).
        float Y[15];
        {
            const double il = 1.0 / sqrt(d3dot(dw, dw));
            sh_basis((float)(dw.x * il), (float)(dw.y * il), (float)(dw.z * il), Y);
        }
        float T = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;
        int nl = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (k < P.depth && ki[k] >= 0 && T >= P.t_cut) {
                const int s = ki[k];
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
                const float wgt = T * ka[k];
                cr = fmaf(wgt, r, cr);
                cg = fmaf(wgt, g, cg);
                cb = fmaf(wgt, b, cb);
                T *= 1.0f - ka[k];
                ++nl;
            }
        }
        // ---- framebuffer write: stage the tile in shared memory so that every store instruction
        // covers whole 32-byte sectors (each tile column is 8 pixels = 96 contiguous bytes) --------
        {
            float* ob = reinterpret_cast<float*>(&ws.rec[0][0]);
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
            st_rays += 1;
            st_hit += nl > 0;
            st_layers += (unsigned)nl;
        }
        if (lane == 0) st_tiles += 1;
    }

    if (P.stats) {
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

template <int K>
int launch_render_k(rtgs_scene* s, const RenderParams& P, cudaStream_t stream) {
    static int blocks_per_sm[16] = {0};
    const size_t smem = sizeof(WarpShared) * WARPS_PER_CTA;
    int dev = s->device;
    if (dev < 0 || dev >= 16) dev = 0;
    if (blocks_per_sm[dev] == 0) {
        CUDA_TRY(cudaFuncSetAttribute(k_render<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_render<K>, WARPS_PER_CTA * 32, smem));
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
    k_render<K><<<grid, WARPS_PER_CTA * 32, smem, stream>>>(P);
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
    if (depth <= 16) return launch_render_k<16>(s, P, stream);
    return launch_render_k<32>(s, P, stream);
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
