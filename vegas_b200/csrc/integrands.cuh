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
            s += vb_exp(-a * dx2);
        }
        f[0] = s * norm;
    }
};

// f = norm * mean_k exp(-a * sum_d (x_d - x0_k)^2)     (examples/ridge.py:18-24)
// mode 0: the numpy original's arithmetic -- per term the D squared distances are summed in axis
//         order, then exp(-a*dx2); the N terms are averaged, then scaled.
// mode 1: the same value through the exact identity sum_d (x_d-c)^2 = V + D (xbar-c)^2 with
//         xbar = mean_d x_d, V = sum_d (x_d-xbar)^2 (a sum of non-negative terms: no cancellation);
//         the argument is -(a V) - (sqrt(aD)(xbar-c))^2, two FP64 instructions per term instead of
//         2D+1.  xs[k] = sqrt(a D) x0[k] is precomputed on the host.
// W terms are evaluated in lock-step so their exp() chains interleave in the FP64 pipe.
//
// The centres travel in the kernel parameters (cpar: constant bank) when there are at most
// VB_RIDGE_PMAX of them: the term loop is warp-uniform, so each centre reaches the FP64
// instructions as a uniform-register operand -- no per-thread loads, no registers holding the next
// group's centres (round 1 read them through __ldg with an 8-deep register prefetch: 32 registers
// and 4e9 L1 sectors per launch).  cpar is padded with copies of the last centre up to a multiple of
// the lock-step width; the padded terms are evaluated and dropped.  Longer ridges read x0 / xs
// from global memory.
#ifndef VB_RIDGE_W
#define VB_RIDGE_W 8
#endif
#define VB_RIDGE_PMAX 1024
struct FRidge {
    static constexpr int NF = 1;
    const double* x0;   // [n] device
    const double* xs;   // [n] device: sqrt(a*dim) * x0[k]   (mode 1)
    int n, mode;
    double a, norm;
    double scale;       // norm / n
    int npar;           // n when the centres are in cpar, else 0
    double cpar[VB_RIDGE_PMAX + 8];   // x0 (mode 0) / xs (mode 1), padded

    // terms [0, n) from the parameter bank, W at a time
    template <int D, bool CHECK, int W>
    __device__ __forceinline__ double sum_axis_order_par(const double (&x)[D], int dim) const
    {
        double s0 = 0.0, s1 = 0.0;
        for (int k = 0; k < npar; k += W) {
            double q[W], e[W];
#pragma unroll
            for (int j = 0; j < W; ++j) q[j] = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (!CHECK || d < dim) {
#pragma unroll
                    for (int j = 0; j < W; ++j) { double t = x[d] - cpar[k + j]; q[j] = fma(t, t, q[j]); }
                }
#pragma unroll
            for (int j = 0; j < W; ++j) q[j] *= -a;
            vb_exp_n<W>(q, e);
            if (k + W <= npar) {
#pragma unroll
                for (int j = 0; j < W; j += 2) { s0 += e[j]; s1 += e[j + 1]; }
            } else {
#pragma unroll
                for (int j = 0; j < W; ++j) if (k + j < npar) s0 += e[j];
            }
        }
        return s0 + s1;
    }

    template <int D, bool CHECK>
    __device__ __forceinline__ double sum_axis_order(const double (&x)[D], int dim) const
    {
        constexpr int W = VB_RIDGE_W;
        double s0 = 0.0, s1 = 0.0;
        int k = 0;
        if (n >= W) {
            double cn[W];                                   // centres of the NEXT group (prefetch)
#pragma unroll
            for (int j = 0; j < W; ++j) cn[j] = __ldg(x0 + j);
            for (; k + W <= n; k += W) {
                double c[W], q[W], e[W];
#pragma unroll
                for (int j = 0; j < W; ++j) { c[j] = cn[j]; q[j] = 0.0; }
                if (k + 2 * W <= n) {
#pragma unroll
                    for (int j = 0; j < W; ++j) cn[j] = __ldg(x0 + k + W + j);
                }
#pragma unroll
                for (int d = 0; d < D; ++d)
                    if (!CHECK || d < dim) {
#pragma unroll
                        for (int j = 0; j < W; ++j) { double t = x[d] - c[j]; q[j] = fma(t, t, q[j]); }
                    }
#pragma unroll
                for (int j = 0; j < W; ++j) q[j] *= -a;
                vb_exp_n<W>(q, e);
#pragma unroll
                for (int j = 0; j < W; j += 2) { s0 += e[j]; s1 += e[j + 1]; }
            }
        }
        for (; k < n; ++k) {
            double c = __ldg(x0 + k), q = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (!CHECK || d < dim) { double t = x[d] - c; q = fma(t, t, q); }
            s0 += vb_exp(-a * q);
        }
        return s0 + s1;
    }

    template <int D>
    __device__ __forceinline__ double sum_shifted(const double (&x)[D], int dim) const
    {
        constexpr int W = VB_RIDGE_W;
        double xbar = 0.0, V = 0.0;
#pragma unroll
        for (int d = 0; d < D; ++d) if (d < dim) xbar += x[d];
        xbar /= (double)dim;
#pragma unroll
        for (int d = 0; d < D; ++d) if (d < dim) { double t = x[d] - xbar; V = fma(t, t, V); }
        const double mV = -a * V, xb = sqrt(a * (double)dim) * xbar;
        double s[W];
#pragma unroll
        for (int j = 0; j < W; ++j) s[j] = 0.0;
        if (npar > 0) {
            for (int k = 0; k < npar; k += W) {
                double q[W], e[W];
#pragma unroll
                for (int j = 0; j < W; ++j) { double t = xb - cpar[k + j]; q[j] = fma(-t, t, mV); }
                vb_exp_n<W>(q, e);
#pragma unroll
                for (int j = 0; j < W; ++j) if (k + W <= npar || k + j < npar) s[j] += e[j];
            }
        } else {
            int k = 0;
            for (; k + W <= n; k += W) {
                double q[W], e[W];
#pragma unroll
                for (int j = 0; j < W; ++j) { double t = xb - __ldg(xs + k + j); q[j] = fma(-t, t, mV); }
                vb_exp_n<W>(q, e);
#pragma unroll
                for (int j = 0; j < W; ++j) s[j] += e[j];
            }
            for (; k < n; ++k) { double t = xb - __ldg(xs + k); s[0] += vb_exp(fma(-t, t, mV)); }
        }
        double tot = s[0];
#pragma unroll
        for (int j = 1; j < W; ++j) tot += s[j];
        return tot;
    }

    template <int D>
    __device__ __forceinline__ void operator()(const double (&x)[D], int dim, double (&f)[1]) const
    {
        double tot;
        if (mode == 1) tot = sum_shifted<D>(x, dim);
        else if (npar > 0) tot = dim == D ? sum_axis_order_par<D, false, VB_RIDGE_W>(x, dim) : sum_axis_order_par<D, true, VB_RIDGE_W>(x, dim);
        else if (dim == D) tot = sum_axis_order<D, false>(x, dim);
        else tot = sum_axis_order<D, true>(x, dim);
        f[0] = tot * scale;
    }
};

// The same ridge (mode 0) for the light engine geometry, whose CTAs leave ~128 registers per thread:
// terms 4 at a time in lock-step (FRidge's 8-wide loop needs more).  Used for short ridges, where the
// sampler around the integrand is a large part of the work.
#ifndef VB_RIDGE_LW
#define VB_RIDGE_LW 4
#endif
struct FRidgeLight : FRidge {
    template <int D>
    __device__ __forceinline__ void operator()(const double (&x)[D], int dim, double (&f)[1]) const
    {
        constexpr int W = VB_RIDGE_LW;
        if (npar > 0) {
            const double tot = dim == D ? sum_axis_order_par<D, false, W>(x, dim) : sum_axis_order_par<D, true, W>(x, dim);
            f[0] = tot * scale;
            return;
        }
        double s0 = 0.0, s1 = 0.0;
        int k = 0;
        for (; k + W <= n; k += W) {
            double c[W], q[W], e[W];
#pragma unroll
            for (int j = 0; j < W; ++j) { c[j] = __ldg(x0 + k + j); q[j] = 0.0; }
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (d < dim) {
#pragma unroll
                    for (int j = 0; j < W; ++j) { double t = x[d] - c[j]; q[j] = fma(t, t, q[j]); }
                }
#pragma unroll
            for (int j = 0; j < W; ++j) q[j] *= -a;
            vb_exp_n<W>(q, e);
#pragma unroll
            for (int j = 0; j < W; j += 2) { s0 += e[j]; s1 += e[j + 1]; }
        }
        for (; k < n; ++k) {
            const double c = __ldg(x0 + k);
            double q = 0.0;
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (d < dim) { double t = x[d] - c; q = fma(t, t, q); }
            s0 += vb_exp(-a * q);
        }
        f[0] = (s0 + s1) * scale;
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
#pragma unroll 1
        for (int j = 0; j < D; ++j)            // rolled: D inlined copies of tan() made the kernel instruction-fetch bound
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
