// ref_cpu.cpp — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
//
// C++/OpenMP CPU restatement of the reference's per-ray render path WITH THE REFERENCE'S ALGORITHM
// SHAPE, used (a) as the large-scene float64 oracle and (b) as the timed CPU baseline
// (`bench.py` cpu_baseline / `--impl reference`; kind = "port": the reference itself is Python +
// Taichi and cannot be compiled).  Nothing in the product links or calls this file.
//
// Reference code being restated (file:line under /root/reference/src/rtgs):
//   ray_tracer.py:39-104   per pixel: `depth` successive closest-hit queries, ray.start advanced to
//                          the previous entry point (+EPS), accum += T*alpha*rgb, T *= 1-alpha
//   scene.py:406-450       closest hit: 32-entry stack DFS, node box tested when popped, far-node
//                          pruning (t_min > best), both child boxes tested to push the far one first
//   bounding_box.py:50-89  slab test with per-axis sign select and unguarded division
//   gaussian.py:203-230    A = d^T M d, B = 2 d^T M (o-p), C = (o-p)^T M (o-p) - 3, delta = B^2-4AC
//   gaussian.py:183-201    rho = exp(-d^T M d) at the mid point, alpha = opacity*rho, SH colour
//   gaussian.py:140-181    15 SH basis values (incl. the 5z^2 - 3z term as coded)
//   camera.py:31-71        pinhole rays
// Differences from the reference, none of which changes the image: M = Sigma^-1 is precomputed
// once per Gaussian instead of per test (gaussian.py:215 recomputes it), and the tree is the LBVH
// of oracle/lbvh_ref.py (the image does not depend on BVH topology).
//
// Scalar = float: the reference's own arithmetic type (timed baseline).
// Scalar = double: float64 evaluation of the same maths (oracle); validated against
//                  oracle/ref_numpy.py (brute force) in tests/test_ref_cpu.py.
#include <math.h>
#ifndef RTGS_NO_OMP
#include <omp.h>
#else
static inline int omp_get_max_threads() { return 1; }
static inline void omp_set_num_threads(int) {}
#endif
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <vector>

namespace {

struct Node {
    float bmin[3], bmax[3];
    int32_t left, right;  // unified ids, -1 = none (leaf)
    int32_t prim;         // sorted position for leaves, -1 for internal
};

struct Scene {
    int64_t n = 0;
    bool has_sh = false;
    std::vector<float> pos, rot, scale, color, opacity, sh;  // original order
    std::vector<uint32_t> morton, sorted_idx;
    std::vector<int32_t> child, parent;
    std::vector<Node> nodes;        // 2n-1, internal [0,n-1), leaves [n-1,2n-1)
    std::vector<double> Minv;       // n*6 (xx,xy,xz,yy,yz,zz), sorted order, float64
    std::vector<float> Minv_f;      // same rounded to float
};

inline uint32_t expand_bits(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// q (v,0) q*  (utils/quaternion.py:84-96), general (non-unit) q
inline void quat_rot(const double q[4], const double v[3], double out[3]) {
    const double ux = q[0], uy = q[1], uz = q[2], w = q[3];
    const double uu = ux * ux + uy * uy + uz * uz, uv = ux * v[0] + uy * v[1] + uz * v[2];
    const double cx = uy * v[2] - uz * v[1], cy = uz * v[0] - ux * v[2], cz = ux * v[1] - uy * v[0];
    const double a = w * w - uu;
    out[0] = a * v[0] + 2.0 * (uv * ux + w * cx);
    out[1] = a * v[1] + 2.0 * (uv * uy + w * cy);
    out[2] = a * v[2] + 2.0 * (uv * uz + w * cz);
}

inline int clz64(uint64_t x) { return __builtin_clzll(x); }

void build_lbvh(Scene& s) {
    const int64_t n = s.n;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], s.pos[i * 3 + a]);
            hi[a] = std::max(hi[a], s.pos[i * 3 + a]);
        }
    float inv[3];
    for (int a = 0; a < 3; ++a) {
        volatile float ext = hi[a] - lo[a];
        inv[a] = ext > 0.0f ? 1.0f / ext : 0.0f;
    }
    s.morton.resize(n);
    std::vector<uint64_t> keys(n);
    for (int64_t i = 0; i < n; ++i) {
        uint32_t u[3];
        for (int a = 0; a < 3; ++a) {
            volatile float d = s.pos[i * 3 + a] - lo[a];   // volatile: forbid contraction / reassociation
            volatile float x = d * inv[a];
            volatile float q = x * 1024.0f;
            float c = std::min(std::max((float)q, 0.0f), 1023.0f);
            u[a] = (uint32_t)c;
        }
        uint32_t code = (expand_bits(u[0]) << 2) | (expand_bits(u[1]) << 1) | expand_bits(u[2]);
        s.morton[i] = code;
        keys[i] = ((uint64_t)code << 32) | (uint64_t)(uint32_t)i;
    }
    std::sort(keys.begin(), keys.end());
    s.sorted_idx.resize(n);
    for (int64_t i = 0; i < n; ++i) s.sorted_idx[i] = (uint32_t)(keys[i] & 0xFFFFFFFFull);
    s.child.assign((size_t)std::max<int64_t>(n - 1, 0) * 2, -1);
    s.parent.assign(2 * n - 1, -1);
    auto delta = [&](int64_t i, int64_t j) -> int {
        if (j < 0 || j >= n) return -1;
        return clz64(keys[i] ^ keys[j]);
    };
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n - 1; ++i) {
        int d = (delta(i, i + 1) - delta(i, i - 1)) >= 0 ? 1 : -1;
        int dmin = delta(i, i - d);
        int64_t lmax = 2;
        while (delta(i, i + lmax * d) > dmin) lmax <<= 1;
        int64_t l = 0;
        for (int64_t t = lmax >> 1; t >= 1; t >>= 1)
            if (delta(i, i + (l + t) * d) > dmin) l += t;
        int64_t j = i + l * d;
        int dnode = delta(i, j);
        int64_t sp = 0, t = l;
        do {
            t = (t + 1) >> 1;
            if (delta(i, i + (sp + t) * d) > dnode) sp += t;
        } while (t > 1);
        int64_t gamma = i + sp * d + (d < 0 ? -1 : 0);
        int64_t first = std::min(i, j), last = std::max(i, j);
        s.child[i * 2 + 0] = (int32_t)(first == gamma ? (n - 1) + gamma : gamma);
        s.child[i * 2 + 1] = (int32_t)(last == gamma + 1 ? (n - 1) + gamma + 1 : gamma + 1);
    }
    for (int64_t i = 0; i < n - 1; ++i) {
        s.parent[s.child[i * 2 + 0]] = (int32_t)i;
        s.parent[s.child[i * 2 + 1]] = (int32_t)i;
    }
    // per-Gaussian Sigma^-1 and tight sqrt(3)-sigma leaf boxes, sorted order
    s.nodes.resize(2 * n - 1);
    s.Minv.resize(n * 6);
    s.Minv_f.resize(n * 6);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        const uint32_t g = s.sorted_idx[k];
        const double q[4] = {s.rot[g * 4], s.rot[g * 4 + 1], s.rot[g * 4 + 2], s.rot[g * 4 + 3]};
        const double sc[3] = {s.scale[g * 3], s.scale[g * 3 + 1], s.scale[g * 3 + 2]};
        double R[3][3];
        for (int c = 0; c < 3; ++c) {
            double e[3] = {0, 0, 0}, col[3];
            e[c] = 1.0;
            quat_rot(q, e, col);
            R[0][c] = col[0];
            R[1][c] = col[1];
            R[2][c] = col[2];
        }
        const double cq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
        const double ic4 = 1.0 / (cq * cq * cq * cq);
        // Sigma = R S^2 R^T (gaussian.py:86-102); Sigma^-1 = R S^-2 R^T / |q|^8 since R R^T = |q|^4 I
        double M[3][3];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                double v = 0;
                for (int c = 0; c < 3; ++c) v += R[a][c] * R[b][c] / (sc[c] * sc[c]);
                M[a][b] = v * ic4;
            }
        double* m = &s.Minv[k * 6];
        m[0] = M[0][0]; m[1] = M[0][1]; m[2] = M[0][2]; m[3] = M[1][1]; m[4] = M[1][2]; m[5] = M[2][2];
        for (int c = 0; c < 6; ++c) s.Minv_f[k * 6 + c] = (float)m[c];
        Node& nd = s.nodes[(n - 1) + k];
        for (int a = 0; a < 3; ++a) {
            double v = 0;
            for (int c = 0; c < 3; ++c) v += R[a][c] * R[a][c] * sc[c] * sc[c];
            double h = sqrt(3.0 * v) * (1.0 + 1e-6);
            nd.bmin[a] = nextafterf((float)((double)s.pos[g * 3 + a] - h), -INFINITY);
            nd.bmax[a] = nextafterf((float)((double)s.pos[g * 3 + a] + h), INFINITY);
        }
        nd.left = nd.right = -1;
        nd.prim = (int32_t)k;
    }
    // refit: internal nodes by decreasing covered-range length is not available; do a post-order walk
    if (n > 1) {
        std::vector<int32_t> order;
        order.reserve(n - 1);
        std::vector<int32_t> st;
        st.push_back(0);
        while (!st.empty()) {
            int32_t i = st.back();
            st.pop_back();
            order.push_back(i);
            for (int c = 0; c < 2; ++c) {
                int32_t ch = s.child[(int64_t)i * 2 + c];
                if (ch < n - 1) st.push_back(ch);
            }
        }
        for (auto it = order.rbegin(); it != order.rend(); ++it) {
            const int32_t i = *it;
            Node& nd = s.nodes[i];
            const Node& a = s.nodes[s.child[(int64_t)i * 2]];
            const Node& b = s.nodes[s.child[(int64_t)i * 2 + 1]];
            for (int c = 0; c < 3; ++c) {
                nd.bmin[c] = std::min(a.bmin[c], b.bmin[c]);
                nd.bmax[c] = std::max(a.bmax[c], b.bmax[c]);
            }
            nd.left = s.child[(int64_t)i * 2];
            nd.right = s.child[(int64_t)i * 2 + 1];
            nd.prim = -1;
        }
    }
}

struct Cam {
    float position[3];
    float rotation[4];
    float focal[2];
    int32_t width, height;
};

template <typename T>
struct RayT {
    T o[3], d[3], start, end;
};

// camera.py:46-52,68-70
template <typename T>
void gen_ray(const Cam& c, int i, int j, RayT<T>& r) {
    const T W = (T)c.width, H = (T)c.height;
    const T u = ((T)i + (T)0.5) / W, v = ((T)j + (T)0.5) / H;
    const T px = (W * u - (T)0.5 * W) / (T)c.focal[0], py = (H * v - (T)0.5 * H) / (T)c.focal[1];
    const T inv = (T)1 / std::sqrt(px * px + py * py + (T)1);
    const T dc[3] = {px * inv, py * inv, -inv};
    const T q[4] = {(T)c.rotation[0], (T)c.rotation[1], (T)c.rotation[2], (T)c.rotation[3]};
    const T uu = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], uv = q[0] * dc[0] + q[1] * dc[1] + q[2] * dc[2];
    const T cx = q[1] * dc[2] - q[2] * dc[1], cy = q[2] * dc[0] - q[0] * dc[2], cz = q[0] * dc[1] - q[1] * dc[0];
    const T a = q[3] * q[3] - uu;
    r.d[0] = a * dc[0] + (T)2 * (uv * q[0] + q[3] * cx);
    r.d[1] = a * dc[1] + (T)2 * (uv * q[1] + q[3] * cy);
    r.d[2] = a * dc[2] + (T)2 * (uv * q[2] + q[3] * cz);
    for (int k = 0; k < 3; ++k) r.o[k] = (T)c.position[k];
    r.start = 0;
    r.end = std::numeric_limits<T>::infinity();
}

// bounding_box.py:50-89
template <typename T>
inline void bound_hit(const Node& nd, const RayT<T>& r, T& tmin, T& tmax) {
    tmin = -std::numeric_limits<T>::infinity();
    tmax = std::numeric_limits<T>::infinity();
    for (int a = 0; a < 3; ++a) {
        const T pmin = r.d[a] < 0 ? (T)nd.bmax[a] : (T)nd.bmin[a];
        const T pmax = r.d[a] < 0 ? (T)nd.bmin[a] : (T)nd.bmax[a];
        const T t0 = (pmin - r.o[a]) / r.d[a], t1 = (pmax - r.o[a]) / r.d[a];
        tmin = std::max(tmin, t0);   // ti.math.max / min semantics (NaN-free inputs assumed)
        tmax = std::min(tmax, t1);
    }
}

template <typename T>
inline const T* minv(const Scene& s, int64_t k);
template <>
inline const float* minv<float>(const Scene& s, int64_t k) { return &s.Minv_f[k * 6]; }
template <>
inline const double* minv<double>(const Scene& s, int64_t k) { return &s.Minv[k * 6]; }

// gaussian.py:203-230
template <typename T>
inline bool gaussian_hit(const Scene& s, int64_t k, const RayT<T>& r, T& t1, T& t2) {
    const T* m = minv<T>(s, k);
    const uint32_t g = s.sorted_idx[k];
    const T v[3] = {r.o[0] - (T)s.pos[g * 3], r.o[1] - (T)s.pos[g * 3 + 1], r.o[2] - (T)s.pos[g * 3 + 2]};
    const T Md[3] = {m[0] * r.d[0] + m[1] * r.d[1] + m[2] * r.d[2], m[1] * r.d[0] + m[3] * r.d[1] + m[4] * r.d[2],
                     m[2] * r.d[0] + m[4] * r.d[1] + m[5] * r.d[2]};
    const T Mv[3] = {m[0] * v[0] + m[1] * v[1] + m[2] * v[2], m[1] * v[0] + m[3] * v[1] + m[4] * v[2],
                     m[2] * v[0] + m[4] * v[1] + m[5] * v[2]};
    const T A = r.d[0] * Md[0] + r.d[1] * Md[1] + r.d[2] * Md[2];
    const T B = (T)2 * (r.d[0] * Mv[0] + r.d[1] * Mv[1] + r.d[2] * Mv[2]);
    const T C = v[0] * Mv[0] + v[1] * Mv[1] + v[2] * Mv[2] - (T)3;
    const T delta = B * B - (T)4 * A * C;
    if (delta > 0) {
        const T sq = std::sqrt(delta);
        t1 = (-B - sq) / ((T)2 * A);
        t2 = (-B + sq) / ((T)2 * A);
        return true;
    }
    if (delta == 0) {
        t1 = -B / ((T)2 * A);
        t2 = std::numeric_limits<T>::infinity();
        return true;
    }
    return false;
}

// scene.py:406-450
template <typename T>
inline int64_t scene_hit(const Scene& s, const RayT<T>& r, T& t1o, T& t2o, uint64_t& nvisit, uint64_t& ntest) {
    int64_t hit = -1;
    T hit_t = std::numeric_limits<T>::infinity();
    int32_t stack[128];
    int top = 0;
    stack[top++] = s.n > 1 ? 0 : 0;
    const Node* nodes = s.nodes.data();
    while (top != 0) {
        const Node& nd = nodes[stack[--top]];
        T tmin, tmax;
        bound_hit(nd, r, tmin, tmax);
        ++nvisit;
        if (tmin > hit_t) continue;                       // prune far node (scene.py:425)
        if (tmin < tmax) {
            if (nd.left == -1 && nd.right == -1) {        // leaf (scene.py:429-437)
                T a, b;
                ++ntest;
                if (gaussian_hit(s, nd.prim, r, a, b)) {
                    if (a < r.end && a > r.start && a < hit_t) {
                        hit = nd.prim;
                        hit_t = a;
                        t1o = a;
                        t2o = b;
                    }
                }
            } else {                                      // scene.py:439-448: push close child last
                T lmin, lmax, rmin, rmax;
                bound_hit(nodes[nd.left], r, lmin, lmax);
                bound_hit(nodes[nd.right], r, rmin, rmax);
                if (lmin < rmin) {
                    stack[top++] = nd.right;
                    stack[top++] = nd.left;
                } else {
                    stack[top++] = nd.left;
                    stack[top++] = nd.right;
                }
            }
        }
    }
    return hit;
}

template <typename T>
inline void sh_basis(const T d[3], T Y[15]) {
    const T c0 = (T)0.9772050238058398, c1 = (T)2.1850968611841584, c2 = (T)1.2615662610100802,
            c3 = (T)2.360174359706574, c4 = (T)5.781222885281108, c5 = (T)1.828183197857863,
            c6 = (T)1.4927053303604616;
    const T x = d[0], y = d[1], z = d[2];
    Y[0] = (T)0.5 * c0 * y; Y[1] = (T)0.5 * c0 * z; Y[2] = (T)0.5 * c0 * x;
    Y[3] = (T)0.5 * c1 * x * y; Y[4] = (T)0.5 * c1 * y * z; Y[5] = (T)0.25 * c2 * ((T)3 * z * z - (T)1);
    Y[6] = (T)0.5 * c1 * x * z; Y[7] = (T)0.25 * c1 * (x * x - y * y);
    Y[8] = (T)0.25 * c3 * y * ((T)3 * x * x - y * y); Y[9] = (T)0.5 * c4 * x * y * z;
    Y[10] = (T)0.25 * c5 * y * ((T)5 * z * z - (T)1); Y[11] = (T)0.25 * c6 * ((T)5 * z * z - (T)3 * z);
    Y[12] = (T)0.25 * c5 * x * ((T)5 * z * z - (T)1); Y[13] = (T)0.25 * c4 * (x * x - y * y) * z;
    Y[14] = (T)0.25 * c3 * x * (x * x - (T)3 * y * y);
}

// gaussian.py:183-201 + ray_tracer.py:88-98 for one hit
template <typename T>
inline void shade(const Scene& s, int64_t k, const RayT<T>& r, T t1, T t2, T& T_att, T acc[3]) {
    const uint32_t g = s.sorted_idx[k];
    const T tm = (t1 + t2) / (T)2;
    const T dv[3] = {r.o[0] + tm * r.d[0] - (T)s.pos[g * 3], r.o[1] + tm * r.d[1] - (T)s.pos[g * 3 + 1],
                     r.o[2] + tm * r.d[2] - (T)s.pos[g * 3 + 2]};
    const T* m = minv<T>(s, k);
    const T q = dv[0] * (m[0] * dv[0] + m[1] * dv[1] + m[2] * dv[2]) + dv[1] * (m[1] * dv[0] + m[3] * dv[1] + m[4] * dv[2]) +
                dv[2] * (m[2] * dv[0] + m[4] * dv[1] + m[5] * dv[2]);
    const T alpha = (T)s.opacity[g] * std::exp(-q);
    T col[3] = {(T)s.color[g * 3], (T)s.color[g * 3 + 1], (T)s.color[g * 3 + 2]};
    if (s.has_sh) {
        const T il = (T)1 / std::sqrt(r.d[0] * r.d[0] + r.d[1] * r.d[1] + r.d[2] * r.d[2]);
        const T dn[3] = {r.d[0] * il, r.d[1] * il, r.d[2] * il};
        T Y[15];
        sh_basis(dn, Y);
        const float* c = &s.sh[(size_t)g * 45];
        for (int j = 0; j < 15; ++j)
            for (int ch = 0; ch < 3; ++ch) col[ch] += Y[j] * (T)c[j * 3 + ch];
    }
    for (int ch = 0; ch < 3; ++ch) acc[ch] += T_att * alpha * col[ch];
    T_att *= (T)1 - alpha;
}

// The step the reference adds to ray.start after every layer (ray_tracer.py:100-102: `t1 + 1e-8`).  In the
// reference's own float32 arithmetic it is below half an ulp for t1 >= 0.25 and does nothing there; evaluated in
// float64 it skips a layer that starts within 1e-8 of the previous one (about 3 pixels of a 1080p / 1 M frame).
// The image oracle therefore runs with 0 (every distinct crossing counts, as in the NumPy brute force); the
// timed float32 baseline and `restart_eps=1e-8` keep the literal value.
static double g_restart_eps = 1e-8;

template <typename T>
void render_t(const Scene& s, const Cam& cam, int depth, int64_t npix, const int32_t* pix, double* rgb, double* Tout,
              int32_t* nlayers, uint64_t* counters) {
    uint64_t tot_visit = 0, tot_test = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : tot_visit, tot_test)
    for (int64_t p = 0; p < npix; ++p) {
        RayT<T> r;
        gen_ray<T>(cam, pix[p * 2], pix[p * 2 + 1], r);
        T T_att = 1, acc[3] = {0, 0, 0};
        int nl = 0;
        uint64_t nv = 0, nt = 0;
        for (int step = 0; step < depth; ++step) {          // ray_tracer.py:39-54: `depth` sample_steps
            T t1 = 0, t2 = 0;
            const int64_t k = scene_hit<T>(s, r, t1, t2, nv, nt);
            if (k < 0) break;                                // ray.start = inf: later steps cannot hit
            shade<T>(s, k, r, t1, t2, T_att, acc);
            ++nl;
            r.start = t1 + (T)g_restart_eps;                 // ray_tracer.py:100-102
        }
        rgb[p * 3] = (double)acc[0];
        rgb[p * 3 + 1] = (double)acc[1];
        rgb[p * 3 + 2] = (double)acc[2];
        if (Tout) Tout[p] = (double)T_att;
        if (nlayers) nlayers[p] = nl;
        tot_visit += nv;
        tot_test += nt;
    }
    if (counters) {
        counters[0] = tot_visit;
        counters[1] = tot_test;
    }
}

// brute force (no BVH): all crossings, sorted by entry distance, first `depth` composited
template <typename T>
void render_brute_t(const Scene& s, const Cam& cam, int depth, int64_t npix, const int32_t* pix, double* rgb,
                    double* Tout, int32_t* nhit) {
#pragma omp parallel
    {
        std::vector<std::pair<T, std::pair<T, int64_t>>> hits;
#pragma omp for schedule(dynamic, 16)
        for (int64_t p = 0; p < npix; ++p) {
            RayT<T> r;
            gen_ray<T>(cam, pix[p * 2], pix[p * 2 + 1], r);
            hits.clear();
            for (int64_t k = 0; k < s.n; ++k) {
                T a, b;
                if (gaussian_hit<T>(s, k, r, a, b) && a > 0 && b != std::numeric_limits<T>::infinity())
                    hits.push_back({a, {b, k}});
            }
            const size_t keep = std::min<size_t>(hits.size(), (size_t)depth);
            std::partial_sort(hits.begin(), hits.begin() + keep, hits.end(), [&](const auto& x, const auto& y) {
                if (x.first != y.first) return x.first < y.first;
                return s.sorted_idx[x.second.second] < s.sorted_idx[y.second.second];
            });
            T T_att = 1, acc[3] = {0, 0, 0};
            for (size_t h = 0; h < keep; ++h) shade<T>(s, hits[h].second.second, r, hits[h].first, hits[h].second.first, T_att, acc);
            rgb[p * 3] = (double)acc[0];
            rgb[p * 3 + 1] = (double)acc[1];
            rgb[p * 3 + 2] = (double)acc[2];
            if (Tout) Tout[p] = (double)T_att;
            if (nhit) nhit[p] = (int32_t)hits.size();
        }
    }
}

}  // namespace

extern "C" {

void* rc_build(int64_t n, const float* pos, const float* rot, const float* scale, const float* color,
               const float* opacity, const float* sh) {
    if (n < 1) return nullptr;
    Scene* s = new Scene();
    s->n = n;
    s->pos.assign(pos, pos + n * 3);
    s->rot.assign(rot, rot + n * 4);
    s->scale.assign(scale, scale + n * 3);
    s->color.assign(color, color + n * 3);
    s->opacity.assign(opacity, opacity + n);
    s->has_sh = sh != nullptr;
    if (sh) s->sh.assign(sh, sh + n * 45);
    build_lbvh(*s);
    return s;
}

void rc_free(void* h) { delete (Scene*)h; }

void rc_read_lbvh(void* h, uint32_t* morton, uint32_t* sorted_idx, int32_t* child, int32_t* parent, float* aabb) {
    Scene* s = (Scene*)h;
    if (morton) memcpy(morton, s->morton.data(), s->n * 4);
    if (sorted_idx) memcpy(sorted_idx, s->sorted_idx.data(), s->n * 4);
    if (child && s->n > 1) memcpy(child, s->child.data(), (s->n - 1) * 8);
    if (parent) memcpy(parent, s->parent.data(), (2 * s->n - 1) * 4);
    if (aabb)
        for (int64_t i = 0; i < 2 * s->n - 1; ++i) {
            memcpy(aabb + i * 6, s->nodes[i].bmin, 12);
            memcpy(aabb + i * 6 + 3, s->nodes[i].bmax, 12);
        }
}

int rc_max_threads(void) { return omp_get_max_threads(); }

// precision: 0 = float (reference arithmetic), 1 = double.  pix = npix (i,j) pairs.
// counters (optional, 2 x uint64): node visits, Gaussian tests.
void rc_set_restart_eps(double eps) { g_restart_eps = eps; }

int rc_render(void* h, const void* cam, int depth, int precision, int64_t npix, const int32_t* pix, double* rgb,
              double* Tout, int32_t* nlayers, uint64_t* counters, int nthreads) {
    Scene* s = (Scene*)h;
    if (!s || !cam || !pix || !rgb) return -1;
    if (nthreads > 0) omp_set_num_threads(nthreads);
    if (precision == 0) render_t<float>(*s, *(const Cam*)cam, depth, npix, pix, rgb, Tout, nlayers, counters);
    else render_t<double>(*s, *(const Cam*)cam, depth, npix, pix, rgb, Tout, nlayers, counters);
    return 0;
}

int rc_render_brute(void* h, const void* cam, int depth, int precision, int64_t npix, const int32_t* pix, double* rgb,
                    double* Tout, int32_t* nhit, int nthreads) {
    Scene* s = (Scene*)h;
    if (!s || !cam || !pix || !rgb) return -1;
    if (nthreads > 0) omp_set_num_threads(nthreads);
    if (precision == 0) render_brute_t<float>(*s, *(const Cam*)cam, depth, npix, pix, rgb, Tout, nhit);
    else render_brute_t<double>(*s, *(const Cam*)cam, depth, npix, pix, rgb, Tout, nhit);
    return 0;
}

}  // extern "C"
