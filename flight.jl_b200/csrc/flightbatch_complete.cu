// flightbatch_complete.cu — the complete (ground-capable) integrator kernels in their own translation unit, with the
// strut code inlined in the leg loop (FB_LDG_INLINE, see fb_physics.cuh). The C ABI lives in flightbatch.cu and reaches
// these kernels through the three launcher functions declared in fb_complete.h.
#define FB_LDG_INLINE 1
#include "fb_kernels.cuh"
#include "fb_complete.h"

namespace {
template <int VAR, int KIN> struct Complete {
    static cudaError_t prepare() {
        return cudaFuncSetAttribute(step_kernel<VAR, KIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StepShape<VAR, KIN, false>::SMEM);
    }
    static int tmax() { return StepShape<VAR, KIN, false>::T; }
    static cudaError_t launch(const CompleteLaunch& a) {
        step_kernel<VAR, KIN, false><<<a.grid, a.block, StepShape<VAR, KIN, false>::SMEM, a.stream>>>(
            a.E, a.step0, a.n_steps, a.dt, a.periodic_every, a.Dt, a.rhs_count, a.own, a.resume_mode, a.ground_count,
            a.block, a.n_steps > 0 ? a.n_steps : 1,    // one thread per member, the whole launch in one chunk
            a.queue, a.queue ? reinterpret_cast<int*>(a.queue + 1) : nullptr);
        return cudaGetLastError();
    }
};
}  // namespace

#if defined(FB_ONLY_S_WA)
#define FB_C_DISPATCH(var_t, kin, CALL) (Complete<FB_C172S, FB_WA>::CALL)
#elif defined(FB_ONLY_X1_WA)
#define FB_C_DISPATCH(var_t, kin, CALL) (Complete<FB_C172X1, FB_WA>::CALL)
#else
#define FB_C_KIN(V, kin, CALL) ((kin) == FB_WA ? Complete<V, FB_WA>::CALL : ((kin) == FB_NED ? Complete<V, FB_NED>::CALL : Complete<V, FB_ECEF>::CALL))
#define FB_C_DISPATCH(var_t, kin, CALL) \
    ((var_t) == FB_C172S ? FB_C_KIN(FB_C172S, kin, CALL) : ((var_t) == FB_C172X0 ? FB_C_KIN(FB_C172X0, kin, CALL) : FB_C_KIN(FB_C172X1, kin, CALL)))
#endif

cudaError_t fb_complete_prepare(int var_t, int kin) { return FB_C_DISPATCH(var_t, kin, prepare()); }
int fb_complete_tmax(int var_t, int kin) { return FB_C_DISPATCH(var_t, kin, tmax()); }
cudaError_t fb_complete_launch(int var_t, int kin, const CompleteLaunch& a) { return FB_C_DISPATCH(var_t, kin, launch(a)); }
