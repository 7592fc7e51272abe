// fb_lss.cuh — batched LinearizedSS: N linear state-space models  ẋ = ẋ0 + A (x − x0) + B (u − u0),
// y = y0 + C (x − x0) + D (u − u0)  stepped in lock-step by the same fixed-step RK4.
// Behavioural spec: lib/FlightPhysics/src/linearization.jl:24-111 (LinearizedSS built from linearize()),
// :157-192 (f_ode!), README.md:142-204 (example 3: the linear model beside the nonlinear one).
//
// B200 mapping: one lane GROUP per member — a half warp for nx = 16, a whole warp for nx = 20. Lane r of the group owns
// row r of A and B in registers, the state is one component per lane, and a right-hand side is nx (shuffle, FMA)
// pairs — the Jacobians never leave the register file during the K fused steps of a launch. Inputs are constant over a
// launch, so B (u − u0) is formed once.
#pragma once
#include <cuda_runtime.h>

struct LssDev {
    double *A, *B, *C, *D, *x0, *u0, *y0, *xd0;   // layouts of fb_linearize: [col][row][pitch], vectors [row][pitch]
    double *x, *u;                                // state [nx][pitch], inputs [nu][pitch]
    long long pitch, n;
};

template <int NX, int NU>
__global__ void __launch_bounds__(256) lss_step_kernel(LssDev L, int n_steps, double dt) {
    constexpr int W = (NX <= 16) ? 16 : 32;                                          // lanes per member
    long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / W;           // member = global lane-group index
    const int r = threadIdx.x & (W - 1);
    const bool live = m < L.n;                  // a group past the end shadows the last member (its warp may hold a live group:
    if (!live) m = L.n - 1;                     // every lane must reach the full-mask shuffles) and stores nothing
    const bool act = live && r < NX;
    const int rr = act ? r : 0;
    const long long P = L.pitch;
    double Ar[NX], bu = 0.0;
#pragma unroll
    for (int j = 0; j < NX; j++) Ar[j] = L.A[((long long)j * NX + rr) * P + m];
#pragma unroll
    for (int k = 0; k < NU; k++) bu = fma(L.B[((long long)k * NX + rr) * P + m], L.u[(long long)k * P + m] - L.u0[(long long)k * P + m], bu);
    const double x0 = L.x0[(long long)rr * P + m], xd0 = L.xd0[(long long)rr * P + m];
    double x = L.x[(long long)rr * P + m];
    auto f = [&](double xs) {   // row r of f_ode!: (ẋ0 + A Δx) + B Δu
        const double dx = xs - x0;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < NX; j++) acc = fma(Ar[j], __shfl_sync(0xffffffffu, dx, j, W), acc);
        return (xd0 + acc) + bu;
    };
    const double h2 = 0.5 * dt, h6 = dt / 6.0;
    for (int s = 0; s < n_steps; s++) {
        const double k1 = f(x), k2 = f(fma(h2, k1, x)), k3 = f(fma(h2, k2, x)), k4 = f(fma(dt, k3, x));
        x = fma(h6, 2.0 * (k2 + k3) + (k1 + k4), x);
    }
    if (act) L.x[(long long)r * P + m] = x;
}

// y = y0 + C (x − x0) + D (u − u0): one thread per member (coalesced over members), not on the stepping path
__global__ void lss_output_kernel(LssDev L, int nx, int nu, int ny, double* y) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= L.n) return;
    const long long P = L.pitch;
    for (int r = 0; r < ny; r++) {
        double acc = 0.0;
        for (int j = 0; j < nx; j++) acc = fma(L.C[((long long)j * ny + r) * P + m], L.x[(long long)j * P + m] - L.x0[(long long)j * P + m], acc);
        double v = L.y0[(long long)r * P + m] + acc;
        acc = 0.0;
        for (int k = 0; k < nu; k++) acc = fma(L.D[((long long)k * ny + r) * P + m], L.u[(long long)k * P + m] - L.u0[(long long)k * P + m], acc);
        y[(long long)r * P + m] = v + acc;
    }
}
