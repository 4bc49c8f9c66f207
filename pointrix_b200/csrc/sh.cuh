// Real spherical harmonics, 3DGS sign convention (Condon-Shortley phase kept).
// Restates msplat/msplat/src/compute_sh.cu:17-35 (constants), :116-164 (degree
// <= 3 homogeneous forms), :166-503 (degrees 4..10, which factor as
// c_nm * Q_n^m(z) * {A_m,B_m}(x,y) with Q_n^m = d^m P_n / dz^m and
// A_m + i B_m = (x + i y)^m -- evaluated here by recurrence instead of the
// reference's ~100 expanded polynomials).
#pragma once
#include "common.cuh"

namespace pxb {

#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
#define SH_C2A 1.0925484305920792f
#define SH_C2C 0.31539156525252005f
#define SH_C2E 0.5462742152960396f
#define SH_C3A 0.5900435899266435f
#define SH_C3B 2.890611442640554f
#define SH_C3C 0.4570457994644658f
#define SH_C3D 0.3731763325901154f
#define SH_C3E 1.445305721320277f

// basis values for the first K (<=16) functions
template <int K>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float B[K]) {
    B[0] = SH_C0;
    if (K > 1) {
        B[1] = -SH_C1 * y;
        B[2] = SH_C1 * z;
        B[3] = -SH_C1 * x;
    }
    if (K > 4) {
        const float xx = x * x, yy = y * y, zz = z * z;
        B[4] = SH_C2A * x * y;
        B[5] = -SH_C2A * y * z;
        B[6] = SH_C2C * (2.f * zz - xx - yy);
        B[7] = -SH_C2A * x * z;
        B[8] = SH_C2E * (xx - yy);
        if (K > 9) {
            B[9] = -SH_C3A * y * (3.f * xx - yy);
            B[10] = SH_C3B * x * y * z;
            B[11] = -SH_C3C * y * (4.f * zz - xx - yy);
            B[12] = SH_C3D * z * (2.f * zz - 3.f * xx - 3.f * yy);
            B[13] = -SH_C3C * x * (4.f * zz - xx - yy);
            B[14] = SH_C3E * z * (xx - yy);
            B[15] = -SH_C3A * x * (xx - 3.f * yy);
        }
    }
}

// gradient of basis k (k < 16) wrt (x,y,z): unconstrained partials of the
// homogeneous forms above
template <int K>
__device__ __forceinline__ void sh_basis_grad(float x, float y, float z, float gx[K], float gy[K], float gz[K]) {
    gx[0] = gy[0] = gz[0] = 0.f;
    if (K > 1) {
        gx[1] = 0.f; gy[1] = -SH_C1; gz[1] = 0.f;
        gx[2] = 0.f; gy[2] = 0.f; gz[2] = SH_C1;
        gx[3] = -SH_C1; gy[3] = 0.f; gz[3] = 0.f;
    }
    if (K > 4) {
        const float xx = x * x, yy = y * y, zz = z * z;
        gx[4] = SH_C2A * y; gy[4] = SH_C2A * x; gz[4] = 0.f;
        gx[5] = 0.f; gy[5] = -SH_C2A * z; gz[5] = -SH_C2A * y;
        gx[6] = -2.f * SH_C2C * x; gy[6] = -2.f * SH_C2C * y; gz[6] = 4.f * SH_C2C * z;
        gx[7] = -SH_C2A * z; gy[7] = 0.f; gz[7] = -SH_C2A * x;
        gx[8] = 2.f * SH_C2E * x; gy[8] = -2.f * SH_C2E * y; gz[8] = 0.f;
        if (K > 9) {
            const float xy = x * y, xz = x * z, yz = y * z;
            gx[9] = -SH_C3A * 6.f * xy; gy[9] = -SH_C3A * 3.f * (xx - yy); gz[9] = 0.f;
            gx[10] = SH_C3B * yz; gy[10] = SH_C3B * xz; gz[10] = SH_C3B * xy;
            gx[11] = SH_C3C * 2.f * xy; gy[11] = -SH_C3C * (4.f * zz - xx - 3.f * yy); gz[11] = -SH_C3C * 8.f * yz;
            gx[12] = -SH_C3D * 6.f * xz; gy[12] = -SH_C3D * 6.f * yz; gz[12] = SH_C3D * (6.f * zz - 3.f * xx - 3.f * yy);
            gx[13] = -SH_C3C * (4.f * zz - 3.f * xx - yy); gy[13] = SH_C3C * 2.f * xy; gz[13] = -SH_C3C * 8.f * xz;
            gx[14] = SH_C3E * 2.f * xz; gy[14] = -SH_C3E * 2.f * yz; gz[14] = SH_C3E * (xx - yy);
            gx[15] = -SH_C3A * 3.f * (xx - yy); gy[15] = SH_C3A * 6.f * xy; gz[15] = 0.f;
        }
    }
}

// Normalisation c_nm for degrees 4..10, filled by the host (pxb_init):
// c_n0 = N_n^0, c_nm = (-1)^m sqrt(2) N_n^m.
__constant__ float c_sh_norm[11][11];

// Visit every basis function of degree 4..deg: f(index, value, dvalue/dx, dy, dz).
template <typename F>
__device__ __forceinline__ void sh_high_visit(int deg, float x, float y, float z, F&& f) {
    float Am = 1.f, Bm = 0.f;        // A_m, B_m
    float Am1 = 0.f, Bm1 = 0.f;      // A_{m-1}, B_{m-1}
    float dfact = 1.f;               // (2m-1)!!
    for (int m = 0; m <= deg; m++) {
        if (m > 0) {
            const float a = x * Am - y * Bm, b = x * Bm + y * Am;
            Am1 = Am; Bm1 = Bm; Am = a; Bm = b;
            dfact *= (float)(2 * m - 1);
        }
        // Q_n^m and Q_n^{m+1} (= dQ_n^m/dz) by upward recurrence in n
        float q2 = 0.f, q1 = 0.f;      // Q_{n-2}^m, Q_{n-1}^m
        float d2 = 0.f, d1 = 0.f;      // Q_{n-2}^{m+1}, Q_{n-1}^{m+1}
        const float dfact1 = dfact * (float)(2 * m + 1);  // (2m+1)!!
        for (int n = m; n <= deg; n++) {
            float q, d;
            if (n == m) { q = dfact; d = 0.f; }
            else if (n == m + 1) { q = (float)(2 * m + 1) * z * q1; d = dfact1; }
            else {
                q = ((float)(2 * n - 1) * z * q1 - (float)(n + m - 1) * q2) / (float)(n - m);
                d = (n == m + 2) ? (float)(2 * m + 3) * z * d1
                                 : ((float)(2 * n - 1) * z * d1 - (float)(n + m) * d2) / (float)(n - m - 1);
            }
            if (n >= 4) {
                const float c = c_sh_norm[n][m];
                const int base = n * n + n;
                if (m == 0) {
                    f(base, c * q, 0.f, 0.f, c * d);
                } else {
                    const float fm = (float)m;
                    // d A_m/dx = m A_{m-1}, d A_m/dy = -m B_{m-1}; d B_m/dx = m B_{m-1}, d B_m/dy = m A_{m-1}
                    f(base + m, c * q * Am, c * q * fm * Am1, -c * q * fm * Bm1, c * d * Am);
                    f(base - m, c * q * Bm, c * q * fm * Bm1, c * q * fm * Am1, c * d * Bm);
                }
            }
            q2 = q1; q1 = q; d2 = d1; d1 = d;
        }
    }
}

}  // namespace pxb
