// integrands.cuh -- device functors compiled into the library (the "fused" integrand plug-in).
//
// Each functor has: static constexpr int NF (number of integrand components) and
//   template<int D> void operator()(const double (&x)[D], int dim, double (&f)[NF]) const
// where dim <= D is the run-time dimension (D is the padded compile-time bound).  They stand in
// the place of the reference's user integrand call (VegasIntegrand.eval, _vegas.pyx:2103-2131);
// numpy twins with the same constants live in vegas_b200/integrands.py.
#pragma once
#include "common.cuh"

// ids shared with include/vegas_b200.h
#define VB200_F_POLY        0
#define VB200_F_GAUSS_MIX   1
#define VB200_F_RIDGE       2
#define VB200_F_GENZ_OSC    3
#define VB200_F_GENZ_PRODPEAK 4
#define VB200_F_GENZ_CORNER 5
#define VB200_F_GENZ_GAUSS  6
#define VB200_F_GENZ_C0     7
#define VB200_F_GENZ_DISC   8
#define VB200_F_PATHINT     9

// f = c0 + sum_d c[d] * x[d]**p[d]           (tests: constants, EPSILON clamp, polynomials)
struct FPoly {
    static constexpr int NF = 1;
    double c0;
    double c[VB_MAXD];
    int p[VB_MAXD];
    template <int D>
    __device__ __forceinline__ void operator()(const double (&x)[D], int dim, double (&f)[1]) const
    {
        double s = c0;
#pragma unroll
        for (int d = 0; d < D; ++d)
            if (d < dim) {
                double t = 1.0;
                for (int i = 0; i < p[d]; ++i) t *= x[d];
                s += c[d] * t;
            }
        f[0] = s;
    }
};

// f = norm * sum_p exp(-a * |x - c_p|^2)      (examples/simple.py, doc eg6.py three-peak)
struct FGaussMix {
    static constexpr int NF = 1;
    const double* centers;   // [npeak][dim] device
    int npeak;
    double a, norm;
    template <int D>
    __device__ __forceinline__ void operator()(const double (&x)[D], int dim, double (&f)[1]) const
    {
        double s = 0.0;
        for (int p = 0; p < npeak; ++p) {
            const double* c = centers + (size_t)p * dim;
            double dx2 = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (d < dim) { double t = x[d] - __ldg(c + d); dx2 += t * t; }
            s += exp(-a * dx2);
        }
        f[0] = s * norm;
    }
};

// f = norm * mean_k exp(-a * sum_d (x_d - x0_k)^2)     (examples/ridge.py:18-24)
// Same arithmetic as the numpy original: per term the D squared distances are summed in axis
// order, then exp(-a*dx2); the N terms are averaged, then scaled.
struct FRidge {
    static constexpr int NF = 1;
    const double* x0;   // [n] device
    int n;
    double a, norm;
    template <int D>
    __device__ __forceinline__ void operator()(const double (&x)[D], int dim, double (&f)[1]) const
    {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int k = 0;
        for (; k + 4 <= n; k += 4) {
            double c0 = __ldg(x0 + k), c1 = __ldg(x0 + k + 1), c2 = __ldg(x0 + k + 2), c3 = __ldg(x0 + k + 3);
            double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (d < dim) {
                    double t0 = x[d] - c0, t1 = x[d] - c1, t2 = x[d] - c2, t3 = x[d] - c3;
                    q0 += t0 * t0; q1 += t1 * t1; q2 += t2 * t2; q3 += t3 * t3;
                }
            s0 += exp(-a * q0); s1 += exp(-a * q1); s2 += exp(-a * q2); s3 += exp(-a * q3);
        }
        for (; k < n; ++k) {
            double c0 = __ldg(x0 + k), q0 = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (d < dim) { double t0 = x[d] - c0; q0 += t0 * t0; }
            s0 += exp(-a * q0);
        }
        f[0] = ((s0 + s1) + (s2 + s3)) / (double)n * norm;
    }
};

// ---- Genz (1984) test family on the unit cube; a = difficulty, u = shift ------------------------
struct FGenz {
    static constexpr int NF = 1;
    int kind;
    double a[VB_MAXD], u[VB_MAXD];
    template <int D>
    __device__ __forceinline__ void operator()(const double (&x)[D], int dim, double (&f)[1]) const
    {
        double r;
        if (kind == VB200_F_GENZ_OSC) {                 // cos(2 pi u_0 + sum a_d x_d)
            double s = 6.283185307179586 * u[0];
#pragma unroll
            for (int d = 0; d < D; ++d) if (d < dim) s += a[d] * x[d];
            r = cos(s);
        } else if (kind == VB200_F_GENZ_PRODPEAK) {     // prod 1 / (a_d^-2 + (x_d - u_d)^2)
            r = 1.0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (d < dim) { double t = x[d] - u[d]; r *= 1.0 / (1.0 / (a[d] * a[d]) + t * t); }
        } else if (kind == VB200_F_GENZ_CORNER) {       // (1 + sum a_d x_d)^-(dim+1)
            double s = 1.0;
#pragma unroll
            for (int d = 0; d < D; ++d) if (d < dim) s += a[d] * x[d];
            r = pow(s, -(double)(dim + 1));
        } else if (kind == VB200_F_GENZ_GAUSS) {        // exp(-sum a_d^2 (x_d - u_d)^2)
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (d < dim) { double t = x[d] - u[d]; s += a[d] * a[d] * (t * t); }
            r = exp(-s);
        } else if (kind == VB200_F_GENZ_C0) {           // exp(-sum a_d |x_d - u_d|)
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d) if (d < dim) s += a[d] * fabs(x[d] - u[d]);
            r = exp(-s);
        } else {                                        // discontinuous: 0 if x_0>u_0 or x_1>u_1
            double s = 0.0;
            bool zero = false;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (d < dim) { s += a[d] * x[d]; if (d < 2 && x[d] > u[d]) zero = true; }
            r = zero ? 0.0 : exp(s);
        }
        f[0] = r;
    }
};

// ---- lattice path integral, 1-d particle, periodic in time (examples/path_integrand.pyx:88-142)
// theta[D] in (-pi/2, pi/2), x = xscale*tan(theta), V(x) = c2*x^2 + c4*x^4.
// f[0] = norm * prod_j jfac_j * exp(-S(x_0 free));  f[1+i] = norm/pi * prod_{j>0} jfac_j *
// exp(-S(x_0 := x0list[i])).
template <int NX0>
struct FPathInt {
    static constexpr int NF = 1 + NX0;
    double T, m, xscale, c2, c4, norm, norm_x0;
    double x0list[NX0 > 0 ? NX0 : 1];
    __device__ __forceinline__ double V(double x) const { double x2 = x * x; return c2 * x2 + c4 * (x2 * x2); }
    template <int D>
    __device__ __forceinline__ void operator()(const double (&th)[D], int dim, double (&f)[NF]) const
    {
        double x[D], Vx[D];
        double a = T / dim, m_2a = m / 2. / a;
        double jac = norm, jac_x0 = norm_x0;
#pragma unroll
        for (int j = 0; j < D; ++j)
            if (j < dim) {
                x[j] = xscale * tan(th[j]);
                Vx[j] = V(x[j]);
                double jfac = xscale + x[j] * x[j] / xscale;
                jac *= jfac;
                if (j > 0) jac_x0 *= jfac;
            }
        double xl = x[0], Vl = Vx[0];           // x[dim-1], V(x[dim-1])
#pragma unroll
        for (int j = 1; j < D; ++j) if (j == dim - 1) { xl = x[j]; Vl = Vx[j]; }
        double Smid = a * Vl;
#pragma unroll
        for (int j = 1; j < D - 1; ++j)
            if (j < dim - 1) { double t = x[j + 1] - x[j]; Smid += m_2a * (t * t) + a * Vx[j]; }
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            double e = (i == 0) ? x[0] : x0list[i > 0 ? i - 1 : 0];
            double Ve = (i == 0) ? Vx[0] : V(e);
            double t1 = x[1] - e, t2 = e - xl;
            double S = Smid + (m_2a * (t1 * t1 + t2 * t2) + a * Ve);
            f[i] = (i == 0 ? jac : jac_x0) * exp(-S);
        }
    }
};
