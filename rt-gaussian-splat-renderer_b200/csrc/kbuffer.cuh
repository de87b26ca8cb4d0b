// kbuffer.cuh — ordering of a ray's k-buffer (shared by shade.cuh and fused.cuh).
//
// A lane's hits live unsorted in shared memory (entry distance kb_t[slot][lane], Gaussian kb_i[slot][lane]).  They
// are ordered by a bitonic network over register-resident keys = entry-distance bits (positive floats order like
// their bit patterns) with the slot in the 4 low bits; the result is a permutation, the slot of every rank packed
// 4 bits each.  Neighbours within float32 rounding (and the 4 truncated bits) of each other are then ordered by
// their float64 entry distances from the raw parameters (exact_less; rare) - the reference's order is the order of
// the exact entry distances (ray_tracer.py:100-102 with scene.py:433).
#pragma once
#include "render_common.cuh"

namespace rtgs_dev {

// Bitonic sorting network over the first N (8 or 16) of 16 register-resident keys, ascending.
template <int N>
__device__ __forceinline__ void sort_keys(unsigned (&key)[16]) {
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned a = key[i], b = key[l];
                    const bool up = (i & k) == 0;
                    key[i] = up ? min(a, b) : max(a, b);
                    key[l] = up ? max(a, b) : min(a, b);
                }
            }
        }
    }
}

// Could the entry distances behind two sorted keys be within 4e-6 relative of each other?  (The keys
// carry the distances with 4 truncated bits, hence the wider 6e-6 screen; unused keys are 0xffffffff.)
__device__ __forceinline__ bool keys_near(unsigned ka, unsigned kb) {
    const float ta = __uint_as_float(ka & ~15u), tb = __uint_as_float(kb & ~15u);
    return kb != 0xffffffffu && (tb - ta) <= 6e-6f * tb;
}

// Slots of the lane's `cnt` (<= 16) hits by ascending entry distance; `maxcnt` = the warp's largest cnt (selects the
// 8- or 16-key network for the whole warp).  n_exact counts the float64 evaluations spent on near ties.
__device__ __forceinline__ unsigned long long order_hits16(const RenderParams& P, const float (&kb_t)[16][32],
                                                           const int (&kb_i)[16][32], int cnt, int maxcnt, int lane,
                                                           int pi, int pj, unsigned long long& n_exact) {
    unsigned long long perm = 0;
    if (maxcnt <= 0) return perm;
    unsigned key[16];
    bool near = false;
    if (maxcnt <= 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            key[k] = k < cnt ? ((__float_as_uint(kb_t[k][lane]) & ~15u) | (unsigned)k) : 0xffffffffu;
        sort_keys<8>(key);
        unsigned lo = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) lo |= (key[r] & 15u) << (4 * r);
        perm = lo;
#pragma unroll
        for (int r = 0; r + 1 < 8; ++r) near = near || keys_near(key[r], key[r + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            key[k] = k < cnt ? ((__float_as_uint(kb_t[k][lane]) & ~15u) | (unsigned)k) : 0xffffffffu;
        sort_keys<16>(key);
        unsigned lo = 0, hi = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            lo |= (key[r] & 15u) << (4 * r);
            hi |= (key[r + 8] & 15u) << (4 * r);
        }
        perm = ((unsigned long long)hi << 32) | lo;
#pragma unroll
        for (int r = 0; r + 1 < 16; ++r) near = near || keys_near(key[r], key[r + 1]);
    }
    if (near) {
        float tp = kb_t[(int)(perm & 15u)][lane];
#pragma unroll 1
        for (int k = 1; k < cnt; ++k) {
            const int sk = (int)((perm >> (4 * k)) & 15u);
            const float tk = kb_t[sk][lane];
            if (fabsf(tk - tp) <= 4e-6f * fabsf(tk)) {
                // insertion among the near-tied predecessors
                int j = k;
                while (j > 0) {
                    const int sa = (int)((perm >> (4 * (j - 1))) & 15u), sb = (int)((perm >> (4 * j)) & 15u);
                    const float ta = kb_t[sa][lane], tb = kb_t[sb][lane];
                    if (fabsf(tb - ta) > 4e-6f * fabsf(tb)) break;
                    n_exact += 2;
                    if (!exact_less(P.raw, P.cam, kb_i[sb][lane], kb_i[sa][lane], pi, pj)) break;
                    const unsigned long long ma = 15ull << (4 * (j - 1)), mb = 15ull << (4 * j);
                    perm = (perm & ~(ma | mb)) | ((unsigned long long)sb << (4 * (j - 1))) |
                           ((unsigned long long)sa << (4 * j));
                    --j;
                }
            }
            tp = kb_t[(int)((perm >> (4 * k)) & 15u)][lane];
        }
    }
    return perm;
}

}  // namespace rtgs_dev
