// gsmath.cuh — device maths shared by the build and render kernels.
// Reference arithmetic being restated (file:line under /root/reference):
//   utils/quaternion.py:8-23,26-35,84-96,99-121  (Hamilton product, q v q*, rotation matrix)
//   gaussian.py:86-102 (covariance), :203-230 (ray-ellipsoid quadratic), :183-201 (response)
//   camera.py:31-71 (pinhole ray generation)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#define RTGS_BOUNDING_THRESHOLD 3.0   // gaussian.py:13 (sqrt(3)-sigma surface)

struct d3 {
    double x, y, z;
};

__host__ __device__ inline d3 d3make(double x, double y, double z) {
    d3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
__host__ __device__ inline double d3dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ inline d3 d3cross(d3 a, d3 b) {
    return d3make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// Vector part of q (v,0) q*  for a (not necessarily unit) scalar-last quaternion
// (utils/quaternion.py:84-96): (w^2 - |u|^2) v + 2 (u.v) u + 2 w (u x v); scales by |q|^2.
__host__ __device__ inline d3 quat_rot(const double q[4], d3 v) {
    d3 u = d3make(q[0], q[1], q[2]);
    double w = q[3];
    double uu = d3dot(u, u), uv = d3dot(u, v);
    d3 c = d3cross(u, v);
    double a = w * w - uu;
    return d3make(a * v.x + 2.0 * (uv * u.x + w * c.x), a * v.y + 2.0 * (uv * u.y + w * c.y),
                  a * v.z + 2.0 * (uv * u.z + w * c.z));
}

// R = as_rotation_mat3(q) (utils/quaternion.py:99-121): columns are the rotated basis vectors.
// Rm[r][c].
__host__ __device__ inline void quat_to_mat(const double q[4], double Rm[3][3]) {
    d3 cx = quat_rot(q, d3make(1, 0, 0));
    d3 cy = quat_rot(q, d3make(0, 1, 0));
    d3 cz = quat_rot(q, d3make(0, 0, 1));
    Rm[0][0] = cx.x; Rm[1][0] = cx.y; Rm[2][0] = cx.z;
    Rm[0][1] = cy.x; Rm[1][1] = cy.y; Rm[2][1] = cy.z;
    Rm[0][2] = cz.x; Rm[1][2] = cz.y; Rm[2][2] = cz.z;
}

// Local-frame matrix W with Sigma^-1 = W^T W where Sigma = R S S^T R^T (gaussian.py:86-102).
// R = c Rhat (c = |q|^2)  =>  Sigma^-1 = Rhat S^-2 Rhat^T / c^2,  W = S^-1 R^T / c^2.
__host__ __device__ inline void local_frame(const double q[4], const double s[3], double Wm[3][3]) {
    double Rm[3][3];
    quat_to_mat(q, Rm);
    double c = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    double ic2 = 1.0 / (c * c);
    for (int k = 0; k < 3; ++k) {
        double f = ic2 / s[k];
        Wm[k][0] = Rm[0][k] * f;
        Wm[k][1] = Rm[1][k] * f;
        Wm[k][2] = Rm[2][k] * f;
    }
}

// Camera ray direction (camera.py:46-52,68-70) in float64 for the continuous pixel coordinate
// (ci, cj) in pixel units (pixel centre = i + 0.5).  Not re-normalised after the rotation,
// exactly like the reference: |dir| = |q_cam|^2.
struct CamD {
    double o[3];
    double q[4];
    double R[9];     // as_rotation_mat3(q), row-major, float64 (host-computed once per launch)
    double fx, fy;
    double ifx, ify; // 1/fx, 1/fy
    int W, H;
};

// dir = R normalize(px, py, -1) with px = (ci - W/2)/fx  (camera.py:46-52; W*u - W/2 with u = ci/W).
__host__ __device__ inline d3 cam_dir(const CamD& c, double ci, double cj) {
    const double px = (ci - 0.5 * (double)c.W) * c.ifx;
    const double py = (cj - 0.5 * (double)c.H) * c.ify;
#ifdef __CUDA_ARCH__
    const double inv = rsqrt(px * px + py * py + 1.0);
#else
    const double inv = 1.0 / sqrt(px * px + py * py + 1.0);
#endif
    const double x = px * inv, y = py * inv, z = -inv;
    return d3make(c.R[0] * x + c.R[1] * y + c.R[2] * z, c.R[3] * x + c.R[4] * y + c.R[5] * z,
                  c.R[6] * x + c.R[7] * y + c.R[8] * z);
}

// plane-normal helper: R v
__host__ __device__ inline d3 cam_rot(const CamD& c, double x, double y, double z) {
    return d3make(c.R[0] * x + c.R[1] * y + c.R[2] * z, c.R[3] * x + c.R[4] * y + c.R[5] * z,
                  c.R[6] * x + c.R[7] * y + c.R[8] * z);
}

// Exact (float64) ray / sqrt(3)-sigma ellipsoid evaluation from the RAW stored parameters
// (gaussian.py:203-230 in the well-conditioned local frame):
//   o' = W (o - p), d' = W d, q_min = |o' x d'|^2 / |d'|^2, hit iff q_min < 3,
//   t1 = (-o'.d' - sqrt(|d'|^2 (3 - q_min))) / |d'|^2   (near root), t2 the far root.
struct ExactHit {
    double q;    // minimum Mahalanobis^2 along the ray (d^T Sigma^-1 d at the max-response point)
    double t1, t2;
    bool hit;    // delta > 0
};

__host__ __device__ inline ExactHit exact_intersect(const double p[3], const double q[4],
                                                    const double s[3], const double o[3], d3 d) {
    double Wm[3][3];
    local_frame(q, s, Wm);
    d3 v = d3make(o[0] - p[0], o[1] - p[1], o[2] - p[2]);
    d3 op = d3make(Wm[0][0] * v.x + Wm[0][1] * v.y + Wm[0][2] * v.z,
                   Wm[1][0] * v.x + Wm[1][1] * v.y + Wm[1][2] * v.z,
                   Wm[2][0] * v.x + Wm[2][1] * v.y + Wm[2][2] * v.z);
    d3 dp = d3make(Wm[0][0] * d.x + Wm[0][1] * d.y + Wm[0][2] * d.z,
                   Wm[1][0] * d.x + Wm[1][1] * d.y + Wm[1][2] * d.z,
                   Wm[2][0] * d.x + Wm[2][1] * d.y + Wm[2][2] * d.z);
    double A = d3dot(dp, dp);
    double Bh = d3dot(op, dp);
    d3 m = d3cross(op, dp);
    ExactHit r;
    r.q = d3dot(m, m) / A;
    double disc = A * (RTGS_BOUNDING_THRESHOLD - r.q);
    r.hit = disc > 0.0;
    double sq = sqrt(disc > 0.0 ? disc : 0.0);
    r.t1 = (-Bh - sq) / A;
    r.t2 = (-Bh + sq) / A;
    return r;
}

// float <-> order-preserving unsigned (for atomicMin/Max reductions on floats)
__device__ inline unsigned int float_to_ordered(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ordered_to_float(unsigned int u) {
    unsigned int v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}
