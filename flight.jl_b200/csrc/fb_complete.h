// fb_complete.h — launchers of the complete integrator kernels (flightbatch_complete.cu) for the C ABI (flightbatch.cu)
#pragma once
#include "fb_kernels.cuh"

struct CompleteLaunch {
    int grid, block;
    cudaStream_t stream;
    EnsembleDev E;
    long long step0; int n_steps; double dt; int periodic_every; double Dt;
    unsigned long long* rhs_count; long long* own; int resume_mode; unsigned* ground_count;
    unsigned* queue;   // non-null: `grid` persistent CTAs whose warps draw member-warps from this counter (the cheap shape of a
                       // hand-over pass that is probably empty: 148 CTAs instead of one per 256 members)
};
cudaError_t fb_complete_prepare(int var_t, int kin);                 // opt in to the kernel's dynamic shared memory
int fb_complete_tmax(int var_t, int kin);                            // compile-time column stride = largest block
cudaError_t fb_complete_launch(int var_t, int kin, const CompleteLaunch& a);
