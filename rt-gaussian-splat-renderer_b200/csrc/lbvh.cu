// lbvh.cu — GPU LBVH build for the rtgs render path (sm_100a).
//
// Replaces the reference's host-driven binned-SAH builder (scene.py:162-404: >= 10 kernel
// launches + 4 device->host syncs per node, ~120 nodes/s) with a fully device-side build:
//   scene bounds -> 30-bit (or, optionally, 63-bit) Morton codes -> LSD radix sort of (code<<32 | index) ->
//   Karras 2012 hierarchy -> per-Gaussian preprocessing + packing in sorted order ->
//   bottom-up AABB refit -> 64-byte two-child traversal nodes.
// The integer spec (codes, keys, hierarchy ids) is oracle/lbvh_ref.py; it must match bit for bit.
// The split key is the Gaussian centre, as in the reference (scene.py:263-266).
#include "common.cuh"
#include "gsmath.cuh"
#include <cuda_fp16.h>

namespace {

// ------------------------------------------------------------------ scene bounds (exact min/max)
__global__ void k_bounds(const float* __restrict__ pos, int64_t n, unsigned int* __restrict__ out) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float v = pos[i * 3 + a];
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&out[a], float_to_ordered(lo[a]));
            atomicMax(&out[3 + a], float_to_ordered(hi[a]));
        }
    }
}

// ------------------------------------------------------------------ Morton codes + sort keys
__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// x = (p - lo) * inv  (float32 subtract, then float32 multiply: not an FMA pattern),
// u = uint(min(max(x*1024, 0), 1023)).  inv = 1/(hi-lo) correctly rounded, 0 when hi == lo.
__global__ void k_morton(const float* __restrict__ pos, int64_t n, const unsigned int* __restrict__ bnd,
                         uint32_t* __restrict__ morton, uint64_t* __restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t u[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float lo = ordered_to_float(bnd[a]), hi = ordered_to_float(bnd[3 + a]);
        float ext = __fsub_rn(hi, lo);
        float inv = ext > 0.0f ? __fdiv_rn(1.0f, ext) : 0.0f;
        float x = __fmul_rn(__fsub_rn(pos[i * 3 + a], lo), inv);
        float q = fminf(fmaxf(__fmul_rn(x, 1024.0f), 0.0f), 1023.0f);
        u[a] = (uint32_t)q;
    }
    uint32_t code = (expand_bits(u[0]) << 2) | (expand_bits(u[1]) << 1) | expand_bits(u[2]);
    morton[i] = code;
    keys[i] = ((uint64_t)code << 32) | (uint64_t)(uint32_t)i;
}

// 63-bit variant (RTGS_OPT_MORTON_BITS = 63; spec: oracle/lbvh_ref.py morton63): 21 bits per axis, so that a few
// far outliers - common in trained 3DGS scenes - do not collapse the rest of the scene into a handful of cells.
//   u = uint(min(max(x * 2^21, 0), 2^21 - 1)),  code = spread(ux) << 2 | spread(uy) << 1 | spread(uz)
// Codes are not unique any more and do not leave room for the index in one 64-bit key, so the order
// (code, index) is produced by two stable 32-bit sorts: on the low word, then on the high word.
__device__ __forceinline__ uint64_t expand_bits21(uint32_t u) {
    uint64_t v = u & 0x1fffffull;
    v = (v | v << 32) & 0x001f00000000ffffull;
    v = (v | v << 16) & 0x001f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void k_morton63(const float* __restrict__ pos, int64_t n, const unsigned int* __restrict__ bnd,
                           uint64_t* __restrict__ code64, uint64_t* __restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t u[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float lo = ordered_to_float(bnd[a]), hi = ordered_to_float(bnd[3 + a]);
        float ext = __fsub_rn(hi, lo);
        float inv = ext > 0.0f ? __fdiv_rn(1.0f, ext) : 0.0f;
        float x = __fmul_rn(__fsub_rn(pos[i * 3 + a], lo), inv);
        float q = fminf(fmaxf(__fmul_rn(x, 2097152.0f), 0.0f), 2097151.0f);
        u[a] = (uint32_t)q;
    }
    const uint64_t code = (expand_bits21(u[0]) << 2) | (expand_bits21(u[1]) << 1) | expand_bits21(u[2]);
    code64[i] = code;
    keys[i] = (code << 32) | (uint64_t)(uint32_t)i;   // first sort: low word of the code
}

// number of distinct 30-bit codes among the sorted keys (upper word), for the automatic choice of the code width
__global__ void k_count_distinct(const uint64_t* __restrict__ sorted_keys, int64_t n, unsigned long long* __restrict__ out) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool first = j < n && (j == 0 || (sorted_keys[j] >> 32) != (sorted_keys[j - 1] >> 32));
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

// second sort key: (high word of the code << 32) | index, in the order the first sort produced
__global__ void k_keys_high(const uint64_t* __restrict__ sorted_low, const uint64_t* __restrict__ code64, int64_t n,
                            uint64_t* __restrict__ keys) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint64_t idx = sorted_low[j] & 0xFFFFFFFFull;
    keys[j] = (code64[idx] & 0xFFFFFFFF00000000ull) | idx;
}

__global__ void k_gather_codes(const uint64_t* __restrict__ sorted_keys, const uint64_t* __restrict__ code64,
                               int64_t n, uint64_t* __restrict__ sorted_codes) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    sorted_codes[j] = code64[sorted_keys[j] & 0xFFFFFFFFull];
}

// ------------------------------------------------------------------ LSD radix sort (8-bit digits)
// Keys are unique 64-bit (code<<32 | index) and arrive in ascending index order, so a STABLE sort
// on the 30 code bits alone (bits 32..61, four 8-bit passes) yields the full-key order.
// Per pass: (1) per-tile digit histograms, (2) exclusive scan over (digit, tile), (3) stable
// scatter with warp-level match_any ranking.  Tile = RS_WARPS warps x 32 lanes x RS_ITEMS keys;
// a warp owns RS_ITEMS consecutive 32-key rows of the tile so that in-tile order is preserved.
constexpr int RS_WARPS = 8;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_WARPS * 32 * RS_ITEMS;  // 2048 keys per block
constexpr int RS_BINS = 256;

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rs_hist(const uint64_t* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ hist /*[bin][tile]*/,
          int ntiles) {
    __shared__ uint32_t sh[RS_BINS];
    for (int b = threadIdx.x; b < RS_BINS; b += blockDim.x) sh[b] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int k = threadIdx.x; k < RS_TILE; k += blockDim.x) {
        int64_t i = base + k;
        if (i < n) atomicAdd(&sh[(uint32_t)(keys[i] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < RS_BINS; b += blockDim.x) hist[(int64_t)b * ntiles + blockIdx.x] = sh[b];
}

// single-block exclusive scan over `len` uint32 (len = 256 * ntiles): every thread scans SCAN_ITEMS consecutive
// values serially, the block scans the per-thread totals (8192 values per iteration)
constexpr int SCAN_ITEMS = 8;
__global__ void __launch_bounds__(1024) k_rs_scan(uint32_t* __restrict__ data, int64_t len) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_sh;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t base = 0; base < len; base += 1024 * SCAN_ITEMS) {
        const int64_t i0 = base + (int64_t)threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS];
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            v[k] = i0 + k < len ? data[i0 + k] : 0u;
            sum += v[k];
        }
        uint32_t x = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warp_tot[lane];
            uint32_t s = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            warp_tot[lane] = s - w;  // exclusive warp offsets
        }
        __syncthreads();
        const uint32_t incl = x + warp_tot[wid] + carry_sh;   // inclusive over this thread's values
        uint32_t run = incl - sum;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (i0 + k < len) data[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_sh = incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_WARPS * 32)
k_rs_scatter(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, int64_t n, int shift,
             const uint32_t* __restrict__ hist_scanned, int ntiles) {
    __shared__ uint32_t wcount[RS_WARPS][RS_BINS];  // per-warp digit counts -> exclusive bases
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < RS_WARPS * RS_BINS; k += blockDim.x) (&wcount[0][0])[k] = 0;
    __syncthreads();

    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)wid * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];  // rank within this warp among equal digits (stable)
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t i = wbase + r * 32 + lane;
        bool valid = i < n;
        key[r] = valid ? in[i] : ~0ull;
        uint32_t dig = (uint32_t)(key[r] >> shift) & 0xFFu;
        // lanes with the same digit (invalid lanes form their own group via bit 8)
        uint32_t peers = __match_any_sync(0xffffffffu, valid ? dig : 0x100u);
        uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint32_t prev = 0;
        if (valid) prev = wcount[wid][dig];
        __syncwarp();
        rank[r] = prev + before;
        if (valid && before == 0) wcount[wid][dig] = prev + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // exclusive scan over warps for each digit, plus the global base of (digit, tile)
    for (int b = threadIdx.x; b < RS_BINS; b += blockDim.x) {
        uint32_t run = hist_scanned[(int64_t)b * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = wcount[w][b];
            wcount[w][b] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            uint32_t dig = (uint32_t)(key[r] >> shift) & 0xFFu;
            out[wcount[wid][dig] + rank[r]] = key[r];
        }
    }
}

// ------------------------------------------------------------------ Karras 2012 hierarchy
// DUP = false: unique keys (code << 32 | index), delta = clz(key_i ^ key_j).  DUP = true: sorted 63-bit codes that
// may repeat; equal codes are told apart by their sorted positions, delta = 64 + clz(i ^ j) (Karras 2012, sec. 4).
template <bool DUP>
__device__ __forceinline__ int delta_fn(const uint64_t* __restrict__ keys, int64_t n, uint64_t ki, int64_t i,
                                        int64_t j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t x = ki ^ keys[j];
    if (DUP && x == 0) return 64 + __clzll((long long)((uint64_t)i ^ (uint64_t)j));
    return __clzll((long long)x);
}

template <bool DUP>
__global__ void k_karras(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ child,
                         int32_t* __restrict__ parent) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    uint64_t ki = keys[i];
    auto delta = [&](int64_t j) { return delta_fn<DUP>(keys, n, ki, i, j); };
    int d = (delta(i + 1) - delta(i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(i - d);
    int64_t lmax = 2;
    while (delta(i + lmax * d) > dmin) lmax <<= 1;
    int64_t l = 0;
    for (int64_t t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(i + (l + t) * d) > dmin) l += t;
    int64_t j = i + l * d;
    int dnode = delta(j);
    int64_t s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int64_t gamma = i + s * d + (d < 0 ? -1 : 0);
    int64_t first = i < j ? i : j, last = i < j ? j : i;
    int32_t left = (int32_t)(first == gamma ? (n - 1) + gamma : gamma);
    int32_t right = (int32_t)(last == gamma + 1 ? (n - 1) + gamma + 1 : gamma + 1);
    child[i * 2 + 0] = left;
    child[i * 2 + 1] = right;
    parent[left] = (int32_t)i;
    parent[right] = (int32_t)i;
    if (i == 0) parent[0] = -1;
}

// ------------------------------------------------------------------ preprocessing + packing
// One thread per SORTED position: gather the stored parameters of Gaussian sorted_idx[s] and emit
// the packed render records (layout in common.cuh) and the leaf AABB.
//   R = as_rotation_mat3(q)         utils/quaternion.py:99-121   (float64)
//   W = S^-1 R^T / |q|^4            Sigma^-1 = W^T W, Sigma = R S S^T R^T  (gaussian.py:86-102)
//   AABB = p +- sqrt(3 Sigma_ii)    tight box of the sqrt(3)-sigma ellipsoid; the reference's box
//          (gaussian.py:104-138) is a conservative superset, and the image does not depend on it.
// Directed rounding keeps the float32 box a superset of the exact ellipsoid extent.
__global__ void k_pack(int64_t n, const uint64_t* __restrict__ keys, const float* __restrict__ pos,
                       const float* __restrict__ rot, const float* __restrict__ scale,
                       const float* __restrict__ color, const float* __restrict__ opacity,
                       const float* __restrict__ sh, uint32_t* __restrict__ sorted_idx,
                       float4* __restrict__ geo, float4* __restrict__ shp, float4* __restrict__ raw,
                       float* __restrict__ aabb, float4* __restrict__ leafbox) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t g = (uint32_t)(keys[s] & 0xFFFFFFFFull);
    sorted_idx[s] = g;
    float p[3] = {pos[g * 3 + 0], pos[g * 3 + 1], pos[g * 3 + 2]};
    float qf[4] = {rot[g * 4 + 0], rot[g * 4 + 1], rot[g * 4 + 2], rot[g * 4 + 3]};
    float sf[3] = {scale[g * 3 + 0], scale[g * 3 + 1], scale[g * 3 + 2]};
    double q[4] = {qf[0], qf[1], qf[2], qf[3]};
    double sc[3] = {sf[0], sf[1], sf[2]};
    double Wm[3][3], Rm[3][3];
    local_frame(q, sc, Wm);
    quat_to_mat(q, Rm);
    geo[s * 4 + 0] = make_float4(p[0], p[1], p[2], opacity[g]);
    geo[s * 4 + 1] = make_float4((float)Wm[0][0], (float)Wm[0][1], (float)Wm[0][2], (float)Wm[1][0]);
    geo[s * 4 + 2] = make_float4((float)Wm[1][1], (float)Wm[1][2], (float)Wm[2][0], (float)Wm[2][1]);
    geo[s * 4 + 3] = make_float4((float)Wm[2][2], color[g * 3 + 0], color[g * 3 + 1], color[g * 3 + 2]);
    raw[s * 3 + 0] = make_float4(p[0], p[1], p[2], qf[0]);
    raw[s * 3 + 1] = make_float4(qf[1], qf[2], qf[3], sf[0]);
    raw[s * 3 + 2] = make_float4(sf[1], sf[2], __int_as_float((int)g), 0.0f);
    if (shp != nullptr) {
        const float* src = sh + (int64_t)g * 45;
        float v[48];
#pragma unroll
        for (int k = 0; k < 45; ++k) v[k] = src[k];
        v[45] = color[g * 3 + 0]; v[46] = color[g * 3 + 1]; v[47] = color[g * 3 + 2];   // DC colour rides in the pad
#pragma unroll
        for (int k = 0; k < 12; ++k)
            shp[s * 12 + k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    }
    // tight AABB: h_i = sqrt(3 * Sigma_ii), Sigma_ii = sum_k R_ik^2 s_k^2
    float* bb = aabb + (int64_t)((n - 1) + s) * 6;
    double var[3];
    float hf[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        var[a] = Rm[a][0] * Rm[a][0] * sc[0] * sc[0] + Rm[a][1] * Rm[a][1] * sc[1] * sc[1] +
                 Rm[a][2] * Rm[a][2] * sc[2] * sc[2];
        double h = sqrt(RTGS_BOUNDING_THRESHOLD * var[a]) * (1.0 + 1e-6);
        bb[a] = __double2float_rd((double)p[a] - h);
        bb[3 + a] = __double2float_ru((double)p[a] + h);
        hf[a] = __double2float_ru(h);
    }
    // leaf record of the per-tile filter (k_tile_lists): exact centre, half extents (rounded up) and the
    // correlations rho_ij = Sigma_ij / sqrt(Sigma_ii Sigma_jj) in fp16, from which the support of the sqrt(3)-sigma
    // ellipsoid along any plane normal n is  s^2 = (n.h)^T C (n.h)  (render_common.cuh: ellipsoid_in_frustum)
    auto cov = [&](int a, int b) {
        return Rm[a][0] * Rm[b][0] * sc[0] * sc[0] + Rm[a][1] * Rm[b][1] * sc[1] * sc[1] +
               Rm[a][2] * Rm[b][2] * sc[2] * sc[2];
    };
    auto rho = [&](int a, int b) {
        const double d = sqrt(var[a] * var[b]);
        double r = d > 0.0 ? cov(a, b) / d : 0.0;
        return (float)fmin(1.0, fmax(-1.0, r));
    };
    const __half2 r01 = __floats2half2_rn(2.0f * rho(0, 1), 2.0f * rho(0, 2)), r2 = __floats2half2_rn(2.0f * rho(1, 2), 0.0f);   // stored doubled
    leafbox[s * 2 + 0] = make_float4(p[0], p[1], p[2], hf[0]);
    leafbox[s * 2 + 1] = make_float4(hf[1], hf[2], __uint_as_float(*reinterpret_cast<const unsigned*>(&r01)),
                                     __uint_as_float(*reinterpret_cast<const unsigned*>(&r2)));
}

// ------------------------------------------------------------------ bottom-up refit
__global__ void k_refit(int64_t n, const int32_t* __restrict__ child, const int32_t* __restrict__ parent,
                        float* __restrict__ aabb, unsigned int* __restrict__ visit) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n) return;
    int32_t node = parent[(n - 1) + s];
    while (node >= 0) {
        // the first arriver stops; the second sees both children complete
        __threadfence();
        if (atomicAdd(&visit[node], 1u) == 0u) return;
        __threadfence();
        int32_t l = child[node * 2 + 0], r = child[node * 2 + 1];
        const volatile float* bl = aabb + (int64_t)l * 6;
        const volatile float* br = aabb + (int64_t)r * 6;
        float* bo = aabb + (int64_t)node * 6;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            bo[a] = fminf(bl[a], br[a]);
            bo[3 + a] = fmaxf(bl[3 + a], br[3 + a]);
        }
        node = parent[node];
    }
}

// ------------------------------------------------------------------ tree depth
// Depth of the deepest leaf (root = depth 0, so a leaf at depth d has d ancestors).  The traversal kernels size
// their stacks from the bound RTGS_MAX_TREE_DEPTH: unique 62-bit keys give <= 62 levels, 63-bit codes with repeats
// 63 + 30 (one level per differing bit of the sorted positions of equal codes); the build verifies it.
__global__ void k_max_depth(int64_t n, const int32_t* __restrict__ parent, unsigned int* __restrict__ out) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned d = 0;
    if (s < n) {
        int32_t node = parent[(n - 1) + s];
        while (node >= 0 && d < 4096u) {
            ++d;
            node = parent[node];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d) atomicMax(out, d);
}

// ------------------------------------------------------------------ traversal nodes (64 B)
// Child boxes are stored as (centre, half extent) with the half extent rounded UP so that
// [c-h, c+h] contains the exact float32 [min, max] box.
__device__ __forceinline__ void box_ch(const float* __restrict__ b, float c[3], float h[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        c[a] = 0.5f * (b[a] + b[3 + a]);
        h[a] = fmaxf(__fsub_ru(b[3 + a], c[a]), __fsub_ru(c[a], b[a]));
    }
}

__global__ void k_pack_nodes(int64_t n, const int32_t* __restrict__ child, const float* __restrict__ aabb,
                             float4* __restrict__ nodes) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    float lc[3], lh[3], rc[3], rh[3];
    if (n == 1) {
        if (i == 0) {  // single Gaussian: synthetic root, left = leaf 0, right = empty box (h = -inf)
            box_ch(aabb, lc, lh);
            nodes[0] = make_float4(lc[0], lc[1], lc[2], lh[0]);
            nodes[1] = make_float4(lh[1], lh[2], 0.0f, 0.0f);
            nodes[2] = make_float4(0.0f, -INFINITY, -INFINITY, -INFINITY);
            nodes[3] = make_float4(__int_as_float(~0), __int_as_float(~0), 0.0f, 0.0f);
        }
        return;
    }
    if (i >= n - 1) return;
    int32_t l = child[i * 2 + 0], r = child[i * 2 + 1];
    box_ch(aabb + (int64_t)l * 6, lc, lh);
    box_ch(aabb + (int64_t)r * 6, rc, rh);
    int32_t le = l >= n - 1 ? ~(int32_t)(l - (n - 1)) : l;
    int32_t re = r >= n - 1 ? ~(int32_t)(r - (n - 1)) : r;
    nodes[i * 4 + 0] = make_float4(lc[0], lc[1], lc[2], lh[0]);
    nodes[i * 4 + 1] = make_float4(lh[1], lh[2], rc[0], rc[1]);
    nodes[i * 4 + 2] = make_float4(rc[2], rh[0], rh[1], rh[2]);
    nodes[i * 4 + 3] = make_float4(__int_as_float(le), __int_as_float(re), 0.0f, 0.0f);
}

// Two-level nodes for k_tile_lists (128 B per internal node): the record of the LEFT child followed by the
// record of the RIGHT child, each in the 64-byte layout of `nodes` (two boxes + two references), so that one
// traversal step descends two levels.  A child that is a leaf becomes {its box, empty box, ~position, none}.
__global__ void k_pack_nodes4(int64_t n, const float4* __restrict__ nodes, const float4* __restrict__ leafbox,
                              float4* __restrict__ nodes4) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (n > 1 ? n - 1 : 1)) return;
    const float4 d = nodes[i * 4 + 3];
    const int ref[2] = {__float_as_int(d.x), __float_as_int(d.y)};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        float4* o = nodes4 + i * 8 + k * 4;
        if (ref[k] >= 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = nodes[(int64_t)ref[k] * 4 + q];
        } else if (k == 1 && n == 1) {   // single Gaussian: the synthetic root has no right child
            o[0] = make_float4(0.0f, 0.0f, 0.0f, -INFINITY);
            o[1] = make_float4(-INFINITY, -INFINITY, 0.0f, 0.0f);
            o[2] = make_float4(0.0f, -INFINITY, -INFINITY, -INFINITY);
            o[3] = make_float4(__int_as_float(~0), __int_as_float(~0), 0.0f, 0.0f);
        } else {
            const float4 a = leafbox[(int64_t)(~ref[k]) * 2 + 0], b = leafbox[(int64_t)(~ref[k]) * 2 + 1];
            o[0] = a;                                               // c.xyz, h.x
            o[1] = make_float4(b.x, b.y, 0.0f, 0.0f);               // h.yz, empty second box: centre 0,
            o[2] = make_float4(0.0f, -INFINITY, -INFINITY, -INFINITY);   // half extent -inf (never inside)
            o[3] = make_float4(__int_as_float(ref[k]), __int_as_float(~0), 0.0f, 0.0f);
        }
    }
}

__global__ void k_init_bounds(unsigned int* b) {
    if (threadIdx.x < 3) b[threadIdx.x] = 0xFFFFFFFFu;
    else if (threadIdx.x < 6) b[threadIdx.x] = 0u;
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
};

}  // namespace

int rtgs_lbvh_build(rtgs_scene* s) {
    const int64_t n = s->n;
    cudaStream_t st = s->own_stream;
    DevBuf<unsigned int> bnd, visit;
    DevBuf<uint64_t> keys_a, keys_b;
    DevBuf<uint32_t> hist;
    CUDA_TRY(cudaMalloc(&bnd.p, 6 * sizeof(unsigned int)));
    CUDA_TRY(cudaMalloc(&keys_a.p, n * sizeof(uint64_t)));
    CUDA_TRY(cudaMalloc(&keys_b.p, n * sizeof(uint64_t)));
    const int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
    CUDA_TRY(cudaMalloc(&hist.p, (size_t)RS_BINS * ntiles * sizeof(uint32_t)));
    CUDA_TRY(cudaMalloc(&visit.p, (size_t)(n > 1 ? n - 1 : 1) * sizeof(unsigned int)));

    if (s->morton_bits_used == 63 && !s->morton64) CUDA_TRY(cudaMalloc(&s->morton64, n * sizeof(uint64_t)));
    cudaEvent_t e0 = nullptr, e1 = nullptr;   // device time of the build proper (rtgs_scene_build_ms)
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    struct EventGuard { cudaEvent_t &a, &b; ~EventGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } eg{e0, e1};
    CUDA_TRY(cudaEventRecord(e0, st));
    const int TB = 256;
    const int nb = (int)((n + TB - 1) / TB);
    const bool wide = s->morton_bits_used == 63;
    k_init_bounds<<<1, 32, 0, st>>>(bnd.p);
    k_bounds<<<min(nb, s->sm_count * 8), TB, 0, st>>>(s->pos, n, bnd.p);
    if (wide) k_morton63<<<nb, TB, 0, st>>>(s->pos, n, bnd.p, s->morton64, keys_a.p);
    else k_morton<<<nb, TB, 0, st>>>(s->pos, n, bnd.p, s->morton, keys_a.p);
    CUDA_TRY(cudaGetLastError());

    uint64_t* src = keys_a.p;
    uint64_t* dst = keys_b.p;
    // stable sort of `src` on its upper 32 bits; afterwards `src` holds the sorted keys
    auto sort_upper_word = [&]() {
        for (int pass = 0; pass < 4; ++pass) {
            int shift = 32 + 8 * pass;
            k_rs_hist<<<ntiles, RS_WARPS * 32, 0, st>>>(src, n, shift, hist.p, ntiles);
            k_rs_scan<<<1, 1024, 0, st>>>(hist.p, (int64_t)RS_BINS * ntiles);
            k_rs_scatter<<<ntiles, RS_WARPS * 32, 0, st>>>(src, dst, n, shift, hist.p, ntiles);
            uint64_t* t = src;
            src = dst;
            dst = t;
        }
    };
    sort_upper_word();
    DevBuf<unsigned long long> distinct;
    if (!wide) {
        CUDA_TRY(cudaMalloc(&distinct.p, sizeof(unsigned long long)));
        CUDA_TRY(cudaMemsetAsync(distinct.p, 0, sizeof(unsigned long long), st));
        k_count_distinct<<<nb, TB, 0, st>>>(src, n, distinct.p);
    }
    if (wide) {
        k_keys_high<<<nb, TB, 0, st>>>(src, s->morton64, n, dst);
        uint64_t* t = src;
        src = dst;
        dst = t;
        sort_upper_word();
    }
    CUDA_TRY(cudaGetLastError());
    // src now holds the sorted keys (low word = original index)
    if (n > 1 && wide) {
        k_gather_codes<<<nb, TB, 0, st>>>(src, s->morton64, n, dst);
        k_karras<true><<<(int)((n - 1 + TB - 1) / TB), TB, 0, st>>>(dst, n, s->child, s->parent);
    } else if (n > 1) {
        k_karras<false><<<(int)((n - 1 + TB - 1) / TB), TB, 0, st>>>(src, n, s->child, s->parent);
    } else {
        int32_t m1 = -1;
        CUDA_TRY(cudaMemcpyAsync(s->parent, &m1, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    k_pack<<<(int)((n + 127) / 128), 128, 0, st>>>(n, src, s->pos, s->rot, s->scale, s->color, s->opacity, s->sh,
                               s->sorted_idx, s->geo, s->shp, s->raw, s->aabb, s->leafbox);
    CUDA_TRY(cudaGetLastError());
    if (n > 1) {
        CUDA_TRY(cudaMemsetAsync(visit.p, 0, (size_t)(n - 1) * sizeof(unsigned int), st));
        k_refit<<<nb, TB, 0, st>>>(n, s->child, s->parent, s->aabb, visit.p);
    }
    DevBuf<unsigned int> depth;
    CUDA_TRY(cudaMalloc(&depth.p, sizeof(unsigned int)));
    CUDA_TRY(cudaMemsetAsync(depth.p, 0, sizeof(unsigned int), st));
    if (n > 1) k_max_depth<<<nb, TB, 0, st>>>(n, s->parent, depth.p);
    k_pack_nodes<<<(int)((s->num_nodes + TB - 1) / TB), TB, 0, st>>>(n, s->child, s->aabb, s->nodes);
    k_pack_nodes4<<<(int)((s->num_nodes + TB - 1) / TB), TB, 0, st>>>(n, s->nodes, s->leafbox, s->nodes4);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(e1, st));
    unsigned int hb[6];
    CUDA_TRY(cudaMemcpyAsync(hb, bnd.p, sizeof(hb), cudaMemcpyDeviceToHost, st));
    unsigned long long hd = 0;
    if (!wide) CUDA_TRY(cudaMemcpyAsync(&hd, distinct.p, sizeof(hd), cudaMemcpyDeviceToHost, st));
    unsigned int hdepth = 0;
    CUDA_TRY(cudaMemcpyAsync(&hdepth, depth.p, sizeof(hdepth), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    s->max_depth = (int)hdepth;
    if (s->max_depth > RTGS_MAX_TREE_DEPTH) {   // cannot happen for < 2^30 Gaussians (see k_max_depth); never traverse it
        rtgs_set_error("LBVH depth %d exceeds the traversal stacks' bound %d", s->max_depth, RTGS_MAX_TREE_DEPTH);
        return RTGS_ERR_STATE;
    }
    s->distinct_codes = wide ? -1 : (int64_t)hd;
    CUDA_TRY(cudaEventElapsedTime(&s->build_ms, e0, e1));
    for (int a = 0; a < 6; ++a) s->bounds[a] = ordered_to_float(hb[a]);
    return RTGS_OK;
}
