// flightbatch.cu — C ABI (include/flightbatch.h) and host-side driver of the batched C172 engine.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// There is no CPU fallback anywhere in this file: without a CUDA device every computing call fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <numeric>
#include <dlfcn.h>
#include <nccl.h>
#include "fb_host.cuh"
#include "fb_lss.cuh"

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(FB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

#if defined(FB_ONLY_S_WA)   // development builds: one instantiation only (fast compile)
#define FB_DISPATCH(s, CALL) (Ops<FB_C172S, FB_WA>::CALL)
#elif defined(FB_ONLY_X1_WA)
#define FB_DISPATCH(s, CALL) (Ops<FB_C172X1, FB_WA>::CALL)
#else
#define FB_KIN_DISPATCH(s, V, CALL) \
    ((s)->cfg.kinematics == FB_WA ? Ops<V, FB_WA>::CALL : ((s)->cfg.kinematics == FB_NED ? Ops<V, FB_NED>::CALL : Ops<V, FB_ECEF>::CALL))
#define FB_DISPATCH(s, CALL) \
    ((s)->var_t == FB_C172S ? FB_KIN_DISPATCH(s, FB_C172S, CALL) : ((s)->var_t == FB_C172X0 ? FB_KIN_DISPATCH(s, FB_C172X0, CALL) : \
     ((s)->var_t == FB_C172XA2 ? fb_ext_ops((s)->cfg.kinematics)->CALL : FB_KIN_DISPATCH(s, FB_C172X1, CALL))))
#endif

// Join: everything that reads or writes the ensemble other than the stepping itself starts here. The launch stream waits for
// the hand-over pass still running on the side stream, and every member goes back to the airborne kernel (a member still
// near the terrain is handed over again by the first evaluation of the next launch: one right-hand side). After it the
// ensemble is at one step index and lives on one stream.
static int32_t join(fb_sim* s) {
    if (s->side_active) { CK(cudaStreamWaitEvent(s->stream, s->side_done, 0)); s->side_active = false; }
    if (s->own_dirty) {
        CK(cudaMemsetAsync(s->d_own, 0, sizeof(long long) * s->pitch, s->stream));
        CK(cudaMemsetAsync(s->d_side, 0, sizeof(unsigned), s->stream));
        s->own_dirty = false;
    }
    s->launches_since_join = 0; s->join_step = s->n_steps;
    return FB_OK;
}
#define JOIN(s) do { if ((s)->side_active || (s)->own_dirty) { CK(cudaSetDevice((s)->cfg.device)); if (int32_t rc_ = join(s)) return rc_; } } while (0)

static size_t hp(const fb_sim* s) { return sizeof(double) * (size_t)(s->host_pitch > 0 ? s->host_pitch : s->E.n); }
static int layout_nx(int variant, int kin) {
    int nact = (variant == FB_C172S) ? 0 : (variant == FB_C172XA2 ? 14 : 7);
    int nkin = (kin == FB_WA) ? 9 : (kin == FB_ECEF ? 8 : 6);
    return 12 + nact + nkin + 6;
}

// ------------------------------------------------------------------------------------ statistics kernels
// Moments of one row of values over the members with status RUNNING or DONE, folded into moment record `rec`: the same
// warp-shuffle reduction + one set of atomics per warp the integrator's log tap uses (moments_add, fb_kernels.cuh).
__global__ void moments_kernel(EnsembleDev E, const double* v, long long rec) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E.n && (E.status[i] == FB_ST_RUNNING || E.status[i] == FB_ST_DONE)) moments_add(E, rec, v[i]);
}
__global__ void moments_reset_kernel(EnsembleDev E, long long first, long long count) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const long long r = first + k;
    E.mom_cnt[r] = 0ull; E.mom_sum[2 * r] = 0.0; E.mom_sum[2 * r + 1] = 0.0; E.mom_min[r] = INFINITY; E.mom_max[r] = -INFINITY;
}
// members within `gate` metres of the terrain (h_o - elevation): the look at the initial state that picks the first launches' shape
__global__ void count_near_kernel(const double* h_o, const double* elev, const int* status, long long n, double gate, unsigned* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool near = i < n && status[i] == FB_ST_RUNNING && !(h_o[i] - elev[i] > gate);
    const unsigned m = __ballot_sync(0xffffffffu, near);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned)__popc(m));
}
// row[i] = Ranged-saturated staged value (types.jl:15-20), the device half of fb_set_input
__global__ void saturate_row_kernel(double* row, const double* staged, long long n, int field) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) row[i] = saturate_input(field, staged[i]);
}
__host__ __device__ inline double default_input(int f) {
    switch (f) {
        case FB_U_MIXTURE: return 0.5; case FB_U_MIXTURE_AUTO: return 1.0; case FB_U_STEERING_ENGAGED: return 1.0;
        case FB_U_M_PILOT: case FB_U_M_COPILOT: return 75.0; case FB_U_M_BAGGAGE: return 50.0;
        case FB_U_T_SL: return 288.15; case FB_U_P_SL: return 101325.0;
        case FB_U_LON_EAS_REF: return 50.0;
        case FB_U_GDC_LAT2: return 1e-3;   // Segment() default target, c172x_gdc.jl:90
        default: return 0.0;
    }
}
__global__ void fill_defaults_kernel(double* u, long long pitch) {   // every input row = the constructed model's default
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < pitch * FB_U__COUNT) u[k] = default_input((int)(k / pitch));
}
__global__ void fill_kernel(double* p, long long n, double v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// FP64 FMA-pipe peak probe: 8 independent dependent-FMA chains per thread, all SMs, no memory traffic
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double r0 = threadIdx.x * 1e-3, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
        r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
    double s = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
    if (s == 123.456) out[0] = s;   // never true; keeps the chains alive
}

// ------------------------------------------------------------------------------------ table blob
struct BlobSec { std::string name; uint32_t dtype, ndim, dims[4]; uint64_t off, nbytes; };
static bool parse_blob(const uint8_t* b, size_t n, std::vector<BlobSec>& out) {
    if (n < 16 || memcmp(b, "FBTABLE1", 8) != 0) return false;
    uint32_t ns; memcpy(&ns, b + 8, 4);
    if (16 + 64ull * ns > n) return false;
    for (uint32_t i = 0; i < ns; i++) {
        const uint8_t* r = b + 16 + 64ull * i;
        BlobSec s; char nm[25]; memcpy(nm, r, 24); nm[24] = 0; s.name = nm;
        memcpy(&s.dtype, r + 24, 4); memcpy(&s.ndim, r + 28, 4); memcpy(s.dims, r + 32, 16);
        memcpy(&s.off, r + 48, 8); memcpy(&s.nbytes, r + 56, 8);
        if (s.off > n || s.nbytes > n - s.off) return false;   // (written so that a huge offset cannot wrap the sum)
        out.push_back(s);
    }
    return true;
}

// ------------------------------------------------------------------------------------ NCCL, loaded at run time
// The library has no link-time dependency on NCCL: a single-GPU host never needs it, and a process that already holds an
// NCCL (torch's, CUDA.jl's) must get that one, which dlopen by soname does.
struct NcclApi {
    void* dl = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi* nccl_api() {
    static NcclApi api; static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {getenv("FLIGHTBATCH_NCCL"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { if (nm && (api.dl = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break; }
        if (api.dl) {
            auto sym = [&](const char* n) { return dlsym(api.dl, n); };
            api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId"); api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
            api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll"); api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce"); api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd"); api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
            if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.AllReduce || !api.GroupStart || !api.GroupEnd || !api.GetErrorString) api.dl = nullptr;
        }
    }
    return api.dl ? &api : nullptr;
}
#define NK(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) return fail(FB_ERR_CUDA, std::string(#call) + ": " + nccl_api()->GetErrorString(r_)); \
    } while (0)

extern "C" {

int32_t fb_api_version(void) { return FB_API_VERSION; }
const char* fb_last_error(void) { return g_err.c_str(); }

static int32_t create_body(fb_sim* s);
int32_t fb_create(const fb_config* cfg, fb_handle* out) {
    if (!cfg || !out) return fail(FB_ERR_ARG, "fb_create: null argument");
    *out = nullptr;
    if (cfg->variant < FB_C172S || cfg->variant > FB_C172XA2) return fail(FB_ERR_ARG, "fb_create: unknown variant");
    if (cfg->variant == FB_C172XA2) {
        if (cfg->prop_pitch_n > 64 || cfg->prop_pitch_n < 0) return fail(FB_ERR_ARG, "fb_create: prop_pitch_n out of range");
        if (cfg->prop_pitch_n > 1 && !(cfg->prop_pitch_hi > cfg->prop_pitch_lo)) return fail(FB_ERR_ARG, "fb_create: VariablePitch needs prop_pitch_hi > prop_pitch_lo");
        if (cfg->act2_omega_n < 0.0 || cfg->act2_zeta < 0.0) return fail(FB_ERR_ARG, "fb_create: negative Actuator2 parameter");
    }
    if (cfg->kinematics != FB_WA && cfg->kinematics != FB_NED && cfg->kinematics != FB_ECEF) return fail(FB_ERR_ARG, "fb_create: kinematics must be FB_WA, FB_ECEF or FB_NED");
    if (cfg->n_members <= 0) return fail(FB_ERR_ARG, "fb_create: n_members must be positive");
    if (cfg->member_offset < 0) return fail(FB_ERR_ARG, "fb_create: member_offset must not be negative");
    if (!(cfg->dt > 0.0)) return fail(FB_ERR_ARG, "fb_create: dt must be positive");
    double Dt = cfg->Dt > 0.0 ? cfg->Dt : cfg->dt;
    long long pe = llround(Dt / cfg->dt);
    if (pe < 1 || fabs(pe * cfg->dt - Dt) > 1e-9 * Dt) return fail(FB_ERR_ARG, "fb_create: Dt must be a positive integer multiple of dt");
    if (cfg->log_every > 0 && cfg->log_n_fields > 0) {
        if (cfg->log_n_fields > 32) return fail(FB_ERR_ARG, "fb_create: at most 32 log fields");
        for (int f = 0; f < cfg->log_n_fields; f++)
            if (cfg->log_fields[f] < 0 || cfg->log_fields[f] >= FB_Y__COUNT) return fail(FB_ERR_ARG, "fb_create: unknown output field in log_fields");
        if (cfg->t_end < cfg->t_start) return fail(FB_ERR_ARG, "fb_create: t_end < t_start");
        if (cfg->log_flags & ~(FB_LOG_MEMBERS | FB_LOG_MOMENTS)) return fail(FB_ERR_ARG, "fb_create: unknown bit in log_flags");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(FB_ERR_CUDA, "fb_create: no CUDA device available (this engine has no CPU path)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(FB_ERR_ARG, "fb_create: bad device ordinal");
    CK(cudaSetDevice(cfg->device));
    fb_sim* s = new fb_sim();
    s->cfg = *cfg; s->cfg.Dt = Dt;
    s->periodic_every = (int)pe;
    int32_t rc = create_body(s);
    if (rc != FB_OK) { std::string msg = g_err; fb_destroy(s); g_err = msg; return rc; }   // nothing half-built survives a failure
    *out = s;
    return FB_OK;
}

static int32_t create_body(fb_sim* s) {
    const fb_config* cfg = &s->cfg;
    s->var_t = (cfg->variant == FB_C172X1 || cfg->variant == FB_C172X2) ? FB_C172X1 : cfg->variant;
    s->nx = layout_nx(cfg->variant, cfg->kinematics);
    long long N = cfg->n_members;
    s->pitch = (N + 31) / 32 * 32;
    if (s->cfg.steps_per_launch <= 0) s->cfg.steps_per_launch = 250;
    CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s->snap_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&s->copy_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->snap_ready, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->snap_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->mom_ready, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s->mom_done, cudaEventDisableTiming));
    for (int f = 0; f < FB_U__COUNT; f++) CK(cudaEventCreateWithFlags(&s->stage_consumed[f], cudaEventDisableTiming));
    EnsembleDev& E = s->E;
    E.pitch = s->pitch; E.n = N;
    CK(cudaMalloc(&E.x, sizeof(double) * s->nx * s->pitch));
    CK(cudaMalloc(&E.u, sizeof(double) * FB_U__COUNT * s->pitch));
    CK(cudaMalloc(&E.stall, s->pitch)); CK(cudaMalloc(&E.eng, s->pitch));
    CK(cudaMalloc(&E.status, sizeof(int) * s->pitch));
    CK(cudaMalloc(&E.ctl, sizeof(double) * FB_CTL_STATE_COUNT * s->pitch));
    CK(cudaMalloc(&s->d_ev_cursor, sizeof(long long) * s->pitch));
    CK(cudaMalloc(&s->d_scratch, sizeof(double) * std::max(s->nx, 8) * s->pitch));
    CK(cudaMalloc(&s->d_rhs_count, sizeof(unsigned long long)));
    CK(cudaMemsetAsync(s->d_rhs_count, 0, sizeof(unsigned long long), s->stream));
    CK(cudaMemsetAsync(E.x, 0, sizeof(double) * s->nx * s->pitch, s->stream));
    CK(cudaMemsetAsync(E.ctl, 0, sizeof(double) * FB_CTL_STATE_COUNT * s->pitch, s->stream));
    fill_kernel<<<(int)((s->pitch + 255) / 256), 256, 0, s->stream>>>(E.ctl + (long long)FB_CTL_ROW_H_STATE * s->pitch, s->pitch, 1.0);   // AltTrackingState.hold
    CK(cudaMemsetAsync(E.stall, 0, s->pitch, s->stream)); CK(cudaMemsetAsync(E.eng, 0, s->pitch, s->stream));
    CK(cudaMemsetAsync(E.status, 0, sizeof(int) * s->pitch, s->stream));
    CK(cudaMemsetAsync(s->d_ev_cursor, 0, sizeof(long long) * s->pitch, s->stream));
    // one launch fills every input row with its default (the rows are contiguous)
    fill_defaults_kernel<<<(int)((s->pitch * FB_U__COUNT + 255) / 256), 256, 0, s->stream>>>(E.u, s->pitch);
    CK(cudaGetLastError());
    E.ev_cursor = s->d_ev_cursor;
    // decimated log: every member's samples (FB_LOG_MEMBERS) and / or the ensemble moments of each sample (FB_LOG_MOMENTS)
    E.log_every = cfg->log_every > 0 ? cfg->log_every : 0;
    E.log_n_fields = 0; E.log_capacity = 0;
    const int flags = cfg->log_flags ? cfg->log_flags : (FB_LOG_MEMBERS | FB_LOG_MOMENTS);
    if (E.log_every > 0 && cfg->log_n_fields > 0) {
        long long total = llround((cfg->t_end - cfg->t_start) / cfg->dt);
        E.log_n_fields = cfg->log_n_fields;
        E.log_capacity = total / E.log_every + 1;
        if (flags & FB_LOG_MEMBERS) {
            double bytes = (double)E.log_capacity * E.log_n_fields * s->pitch * 8.0;
            if (bytes > 64e9) return fail(FB_ERR_ALLOC, "fb_create: log would exceed 64 GB; raise log_every, log fewer fields or keep the moments only (FB_LOG_MOMENTS)");
            if (cudaMalloc(&E.log, (size_t)bytes) != cudaSuccess) { (void)cudaGetLastError(); return fail(FB_ERR_ALLOC, "fb_create: out of device memory for the log"); }
        }
        CK(cudaMalloc(&s->d_log_fields, sizeof(int) * 32));
        CK(cudaMemcpyAsync(s->d_log_fields, s->cfg.log_fields, sizeof(int) * cfg->log_n_fields, cudaMemcpyHostToDevice, s->stream));
        E.log_fields = s->d_log_fields;
    }
    {   // moment records: [capacity][fields] for the log samples (when kept) + one scratch record for fb_stats
        s->mom_records = ((flags & FB_LOG_MOMENTS) ? E.log_capacity * E.log_n_fields : 0) + 1;
        const long long R = s->mom_records;
        CK(cudaMalloc(&s->d_mom, (size_t)R * 40));
        s->m_cnt = (unsigned long long*)s->d_mom; s->m_sum = (double*)((char*)s->d_mom + R * 8);
        s->m_min = (double*)((char*)s->d_mom + R * 24); s->m_max = (double*)((char*)s->d_mom + R * 32);
        if (R > 1) { E.mom_cnt = s->m_cnt; E.mom_sum = s->m_sum; E.mom_min = s->m_min; E.mom_max = s->m_max; }
    }
    // Integrator launch shape: one CTA per SM holding the table pack plus two NX x T column blocks in shared
    // memory. T is the largest multiple of 32 that fits 227 KB and the register budget, then shrunk while the
    // number of waves over the SMs stays the same (balances the last wave).
    {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, cfg->device));
        int sms = prop.multiProcessorCount;
        // largest block <= tmax that keeps the number of CTA waves (balances the last wave); returns the wave count
        auto shape = [&](int tmax, int& T) {
            T = (int)std::min<long long>(tmax, (N + 31) / 32 * 32);
            long long waves = ((N + T - 1) / T + sms - 1) / sms;
            while (T > 32 && (((N + (T - 32) - 1) / (T - 32) + sms - 1) / sms) == waves) T -= 32;
            return waves;
        };
        shape(FB_DISPATCH(s, step_tmax)(0), s->step_block);
        int Tn, Tw;
        long long waves_n = shape(FB_DISPATCH(s, step_tmax)(1), Tn), waves_w = shape(FB_DISPATCH(s, step_tmax)(2), Tw);
        // cost model: a launch of the wide (128-register) shape takes ~1.45x as long as one of the 168-register shape
        s->air_wide = (double)waves_w * 1.45 < (double)waves_n;
        s->step_block_air = s->air_wide ? Tw : Tn;
        // Third option, for ensembles of more than one and at most FB_MW_MAX member-warps per SM: ONE persistent CTA
        // per SM of the 168-register shape that time-slices its MW member-warps over its 12 warps in chunks of steps:
        // ceil(MW * C / 12) / C waves (+ ~3 % for the k1 re-evaluated at every chunk start).
        {
            const int tmax = FB_DISPATCH(s, step_tmax)(1), hw = tmax / 32;
            const long long per_sm = (N + sms - 1) / sms;
            const int mw = (int)((per_sm + 31) / 32), C = 10;
            const double cost_plain = s->air_wide ? (double)waves_w * 1.45 : (double)waves_n;
            const double cost_sliced = 1.03 * (double)((mw * C + hw - 1) / hw) / C;
            const char* env = getenv("FLIGHTBATCH_SLICED");   // diagnostic switch: 0 = never, 1 = whenever possible
            bool ok = mw > hw && mw <= FB_MW_MAX;
            bool use = ok && cost_sliced < cost_plain;
            if (env && ok) use = atoi(env) != 0;
            if (use) { s->air_wide = false; s->step_block_air = tmax; s->air_mpc = mw * 32; }
        }
    }
    {   // Fourth option: one persistent CTA per SM of the full-register shape, (member-warp, chunk) tasks drawn from a global
        // queue (fb_kernels.cuh). Measured on one B200 (same box, interleaved): fly-by-wire kernels (8 warps per SM) +3.3 % on a
        // 131,072-member shard (3.5 tasks per hardware warp), -0.7 % at 2^20 members (27.7 CTA waves balance themselves);
        // C172S (12 warps per SM, whose static time-slicing is already an exact number of rounds) -1 %. Hence: the variants
        // with control laws, below eight CTA waves. FLIGHTBATCH_DYNAMIC=1/0 forces it on/off, FLIGHTBATCH_SLICED=0/1 selects
        // the static shapes above (diagnostic switches).
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, cfg->device));
        const int tmax = FB_DISPATCH(s, step_tmax)(1), hw = tmax / 32;
        const long long mwt = (N + 31) / 32;
        bool use = s->var_t == FB_C172X1 && mwt < 8ll * prop.multiProcessorCount * hw;
        if (const char* dynv = getenv("FLIGHTBATCH_DYNAMIC")) use = atoi(dynv) != 0;
        if (getenv("FLIGHTBATCH_SLICED")) use = false;
        if (use) {
            s->air_dyn = true; s->air_wide = false; s->air_mpc = 0; s->step_block_air = tmax;
            s->air_grid = (int)std::min<long long>(prop.multiProcessorCount, (mwt + hw - 1) / hw);
            CK(cudaMalloc(&s->d_queue, sizeof(unsigned) * (size_t)(1 + mwt)));
        }
    }
    // FLIGHTBATCH_KERNEL=complete pins the complete kernel (no airborne instantiation, no hand-over): a diagnostic switch used
    // by the tests to check that the two launch paths produce the same trajectories
    if (const char* k = getenv("FLIGHTBATCH_KERNEL")) s->force_complete = (strcmp(k, "complete") == 0);
    if (cfg->variant == FB_C172XA2) s->force_complete = true;   // the non-default components exist in the complete kernel only
    if (s->force_complete) s->air_mode = false;
    CK(cudaMalloc(&s->E.handover_rel, sizeof(int) * s->pitch));
    CK(cudaMalloc(&s->d_own, sizeof(long long) * s->pitch));
    CK(cudaMemsetAsync(s->d_own, 0, sizeof(long long) * s->pitch, s->stream));
    CK(cudaMalloc(&s->d_side, sizeof(unsigned)));
    CK(cudaMalloc(&s->d_queue_c, sizeof(unsigned) * 2));
    { cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, cfg->device)); s->n_sms = prop.multiProcessorCount; }
    CK(cudaMemsetAsync(s->d_side, 0, sizeof(unsigned), s->stream));
    // hand-over passes on a stream of their own, one launch behind the airborne launches (FLIGHTBATCH_PIPELINE=0: on the launch
    // stream, right behind each airborne launch: diagnostic)
    if (const char* v = getenv("FLIGHTBATCH_PIPELINE")) s->pipeline = atoi(v) != 0;
    {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&s->side_stream, cudaStreamNonBlocking, hi));   // its few CTAs go first when an SM frees up
        CK(cudaEventCreateWithFlags(&s->air_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s->side_done, cudaEventDisableTiming));
    }
    CK(cudaMalloc(&s->d_ground, sizeof(unsigned)));
    CK(cudaMallocHost(&s->h_ground, sizeof(unsigned)));
    *s->h_ground = 0;
    CK(cudaEventCreateWithFlags(&s->ground_ev, cudaEventDisableTiming));
    CK(FB_DISPATCH(s, prepare)(s));
    CK(cudaStreamSynchronize(s->stream));
    return FB_OK;
}

int32_t fb_destroy(fb_handle s) {
    if (!s) return FB_OK;
    cudaSetDevice(s->cfg.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->side_stream) { cudaStreamSynchronize(s->side_stream); cudaStreamDestroy(s->side_stream); }   // a hand-over pass may still be running
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    if (s->snap_stream) { cudaStreamSynchronize(s->snap_stream); cudaStreamDestroy(s->snap_stream); }
    cudaFree(s->d_snap); if (s->snap_ready) cudaEventDestroy(s->snap_ready); if (s->snap_done) cudaEventDestroy(s->snap_done);
    EnsembleDev& E = s->E;
    cudaFree(E.x); cudaFree(E.u); cudaFree(E.stall); cudaFree(E.eng); cudaFree(E.status); cudaFree(E.ctl); cudaFree(E.log);
    cudaFree(s->d_pack); cudaFree(s->d_geoid); cudaFree(s->d_prop3); cudaFree(s->d_upload);
    cudaFree(s->d_ev_off); cudaFree(s->d_ev_step); cudaFree(s->d_ev_field); cudaFree(s->d_ev_value); cudaFree(s->d_ev_cursor);
    cudaFree(s->d_log_fields); cudaFree(s->d_scratch); cudaFree(s->d_rhs_count);
    if (s->air_done) cudaEventDestroy(s->air_done); if (s->side_done) cudaEventDestroy(s->side_done);
    cudaFree(E.handover_rel); cudaFree(s->d_own); cudaFree(s->d_side); cudaFree(s->d_queue_c); cudaFree(s->d_hint); if (s->h_hint) cudaFreeHost(s->h_hint); if (s->hint_ev) cudaEventDestroy(s->hint_ev); cudaFree(s->d_ground); cudaFree(s->d_queue); if (s->h_ground) cudaFreeHost(s->h_ground); if (s->ground_ev) cudaEventDestroy(s->ground_ev);
    if (s->comm_stream) cudaStreamSynchronize(s->comm_stream);
    if (s->comm && s->comm_owned && nccl_api()) nccl_api()->CommDestroy(s->comm);
    for (auto& t : s->comm_timers) { cudaEventDestroy(t.first); cudaEventDestroy(t.second); }
    if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
    if (s->comm_in) cudaEventDestroy(s->comm_in); if (s->comm_out) cudaEventDestroy(s->comm_out);
    cudaFree(s->d_mom); if (s->h_mom) cudaFreeHost(s->h_mom);
    if (s->mom_ready) cudaEventDestroy(s->mom_ready); if (s->mom_done) cudaEventDestroy(s->mom_done);
    for (int f = 0; f < FB_U__COUNT; f++) { cudaFree(s->d_stage[f]); if (s->stage_consumed[f]) cudaEventDestroy(s->stage_consumed[f]); }
    if (s->copy_done) cudaEventDestroy(s->copy_done);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->stream) cudaStreamDestroy(s->stream);
    (void)cudaGetLastError();
    delete s;
    return FB_OK;
}

int32_t fb_wait_state(fb_handle s);
int32_t fb_sync(fb_handle s) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamSynchronize(s->stream));
    if (s->comm_stream) CK(cudaStreamSynchronize(s->comm_stream));
    return fb_wait_state(s);
}

int32_t fb_load_tables(fb_handle s, const void* blob, size_t nbytes) {
    if (!s || !blob) return fail(FB_ERR_ARG, "fb_load_tables: null argument");
    JOIN(s);
    CK(cudaSetDevice(s->cfg.device));
    std::vector<BlobSec> secs;
    const uint8_t* b = (const uint8_t*)blob;
    if (!parse_blob(b, nbytes, secs)) return fail(FB_ERR_ARG, "fb_load_tables: not a table blob");
    auto find = [&](const char* name) -> const BlobSec* { for (auto& x : secs) if (x.name == name) return &x; return nullptr; };
    std::vector<double> pack(FB_TB_BYTES / 8, 0.0);
    struct Map { const char* name; int off; int count; };
    const Map maps[] = {
        {"aero.ge_x", TB_GE_X, 13}, {"aero.CD_ge", TB_CD_GE, 13}, {"aero.CL_ge", TB_CL_GE, 13}, {"aero.df_x", TB_DF_X, 4},
        {"aero.CD_df", TB_CD_DF, 4}, {"aero.CL_df", TB_CL_DF, 4}, {"aero.Cm_df", TB_CM_DF, 4}, {"aero.CD_a_x", TB_CD_A_X, 26},
        {"aero.CD_a_df", TB_CD_A_DF, 104}, {"aero.CL_a_x", TB_CL_A_X, 17}, {"aero.CL_a", TB_CL_A, 34}, {"aero.CY_b_df", TB_CY_B_DF, 6},
        {"aero.CY_p", TB_CY_P, 4}, {"aero.CY_r", TB_CY_R, 4}, {"aero.Cl_r", TB_CL_R, 4},
        {"eng.d_wot", TB_D_WOT, 18}, {"eng.mu_wot", TB_MU_WOT, 18}, {"eng.n_std", TB_N_STD, 13}, {"eng.pi_std", TB_PI_STD, 39},
        {"eng.n_wot", TB_N_WOT, 5}, {"eng.pi_wot", TB_PI_WOT, 15}, {"eng.f_x", TB_F_X, 11}, {"eng.pi_ratio", TB_PI_RATIO, 11},
        {"eng.sfc_ratio", TB_SFC_RATIO, 11}, {"eng.n_sfc", TB_N_SFC, 5}, {"eng.pi_sfc", TB_PI_SFC, 8}, {"eng.sfc_pow", TB_SFC_POW, 40}};
    for (const Map& m : maps) {
        const BlobSec* sec = find(m.name);
        if (!sec || sec->dtype != 0 || sec->nbytes != (uint64_t)m.count * 8) return fail(FB_ERR_ARG, std::string("fb_load_tables: bad or missing section ") + m.name);
        memcpy(pack.data() + m.off, b + sec->off, sec->nbytes);
    }
    {
        const BlobSec* sec = find("prop.tables");
        if (!sec || sec->nbytes != 6ull * 441 * 8) return fail(FB_ERR_ARG, "fb_load_tables: bad or missing section prop.tables");
        const double* p = (const double*)(b + sec->off);
        for (int c = 0; c < 4; c++) for (int ij = 0; ij < 441; ij++) pack[TB_PROP + ij * 4 + c] = p[c * 441 + ij];
    }
    {   // reciprocal knot spacings
        struct Inv { int knots, inv, n; };
        const Inv invs[] = {{TB_GE_X, TB_I_GE, 13}, {TB_DF_X, TB_I_DF, 4}, {TB_CD_A_X, TB_I_CD_A, 26}, {TB_CL_A_X, TB_I_CL_A, 17},
                            {TB_N_STD, TB_I_N_STD, 13}, {TB_N_WOT, TB_I_N_WOT, 5}, {TB_F_X, TB_I_F, 11}, {TB_N_SFC, TB_I_N_SFC, 5},
                            {TB_PI_SFC, TB_I_PI_SFC, 8}};
        for (const Inv& v : invs) for (int i = 0; i + 1 < v.n; i++) pack[v.inv + i] = 1.0 / (pack[v.knots + i + 1] - pack[v.knots + i]);
        // bucketed cell finders (fb_tables.h: TB_BK), in the order of the FB_AX_* ids
        const int axes[FB_AX_COUNT] = {0, 1, 2, 3, 4, 5, 6, 7, 8};
        for (int a = 0; a < FB_AX_COUNT; a++) {
            const Inv& v = invs[axes[a]];
            const double* k = pack.data() + v.knots;
            const double k0 = k[0], h = (k[v.n - 1] - k0) / FB_BK;
            pack[TB_BK_IH + a] = 1.0 / h;
            for (int bkt = 0; bkt < FB_BK; bkt++) {
                double left = k0 + bkt * h, right = k0 + (bkt + 1) * h;
                int c0 = 0;
                while (c0 < v.n - 2 && k[c0 + 1] <= left) c0++;
                double kb = 1e300; int inside = 0;
                for (int j = 1; j <= v.n - 2; j++) if (k[j] > left && k[j] < right) { kb = k[j]; inside++; }
                if (inside > 1) return fail(FB_ERR_ARG, "fb_load_tables: an interpolation axis has two knots in one cell-finder bucket");
                double* rec = pack.data() + TB_BK + (a * FB_BK + bkt) * 2;
                rec[0] = kb;
                int64_t word = (int64_t)c0;
                memcpy(&rec[1], &word, 8);
            }
            // check the finder against a plain search: any disagreement must sit on a knot (weight 0 or 1 up to rounding)
            for (int q = 0; q <= 8192; q++) {
                double x = k0 + (k[v.n - 1] - k0) * q / 8192.0;
                if (q % 3 == 1) x = k[(q / 3) % v.n];
                int bkt = std::min(std::max((int)((x - k0) * pack[TB_BK_IH + a]), 0), FB_BK - 1);
                const double* rec = pack.data() + TB_BK + (a * FB_BK + bkt) * 2;
                int64_t word; memcpy(&word, &rec[1], 8);
                int c = (int)word + (x >= rec[0] ? 1 : 0);
                double w = (x - k[c]) / (k[c + 1] - k[c]);
                if (c < 0 || c > v.n - 2 || w < -1e-9 || w > 1.0 + 1e-9) return fail(FB_ERR_ARG, "fb_load_tables: cell finder self-check failed");
            }
        }
    }
    const BlobSec* g = find("geoid");
    if (!g || g->dtype != 1 || g->nbytes != (uint64_t)FB_GEOID_NLAT * FB_GEOID_NLON * 4) return fail(FB_ERR_ARG, "fb_load_tables: bad or missing geoid section");
    if (!s->d_pack) CK(cudaMalloc(&s->d_pack, FB_TB_BYTES + FB_GS_BYTES));
    if (!s->d_geoid) CK(cudaMalloc(&s->d_geoid, FB_GEOID_BYTES));
    CK(cudaMemcpyAsync(s->d_pack, pack.data(), FB_TB_BYTES, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(s->d_geoid, b + g->off, g->nbytes, cudaMemcpyHostToDevice, s->stream));
    // tail of the geoid buffer: parameters and tables of the non-default components (FB_C172XA2; zero otherwise)
    FbTablesTail tail; memset(&tail, 0, sizeof(tail));
    std::vector<double> p3;   // staging must outlive the asynchronous copy
    if (s->cfg.variant == FB_C172XA2) {
        const double wn = s->cfg.act2_omega_n > 0.0 ? s->cfg.act2_omega_n : 5.0 * 2.0 * FB_PI, zt = s->cfg.act2_zeta > 0.0 ? s->cfg.act2_zeta : 0.6;   // c172x.jl:71
        tail.act2_wn2 = wn * wn; tail.act2_2zw = 2.0 * zt * wn;
        if (s->cfg.prop_pitch_n > 1) {
            // Lookup of Propeller(VariablePitch(lo, hi; length = n)): "prop.tables3" [21][21][n][6] with its range in "prop.pitch3",
            // written by tools/build_tables.py for exactly this VariablePitch (the reference builds the Lookup when the
            // Propeller is constructed, propellers.jl:356-359)
            const BlobSec* t3 = find("prop.tables3"); const BlobSec* r3 = find("prop.pitch3");
            const int nb = s->cfg.prop_pitch_n;
            if (!t3 || !r3 || t3->dtype != 0 || r3->nbytes != 16 || t3->nbytes != 441ull * nb * 6 * 8)
                return fail(FB_ERR_ARG, "fb_load_tables: the table blob holds no prop.tables3 / prop.pitch3 section for this prop_pitch_n");
            double rng[2]; memcpy(rng, b + r3->off, 16);
            if (rng[0] != s->cfg.prop_pitch_lo || rng[1] != s->cfg.prop_pitch_hi)
                return fail(FB_ERR_ARG, "fb_load_tables: prop.pitch3 of the table blob differs from (prop_pitch_lo, prop_pitch_hi); rebuild the blob for this VariablePitch");
            const double* src = reinterpret_cast<const double*>(b + t3->off);
            p3.resize(441ull * nb * 4);
            for (size_t o = 0; o < 441ull * nb; o++) for (int c = 0; c < 4; c++) p3[o * 4 + c] = src[o * 6 + c];   // C_Fx, C_Mx, C_Fz_a, C_Mz_a (C_P, eta_p are not on the path)
            if (!s->d_prop3) CK(cudaMalloc(&s->d_prop3, p3.size() * 8));
            CK(cudaMemcpyAsync(s->d_prop3, p3.data(), p3.size() * 8, cudaMemcpyHostToDevice, s->stream));
            tail.prop3 = s->d_prop3; tail.prop3_lo = rng[0]; tail.prop3_hi = rng[1]; tail.prop3_n = nb; tail.prop3_is = (nb - 1) / (rng[1] - rng[0]);
        }
    }
    CK(cudaMemcpyAsync((char*)s->d_geoid + FB_GEOID_TAIL_OFF, &tail, sizeof(tail), cudaMemcpyHostToDevice, s->stream));
    static const char* gnames[GS_COUNT] = {"gain.te2te", "gain.tv2te", "gain.vh2te", "gain.ar2ar", "gain.phibeta2ar",
                                           "gain.q2e", "gain.c2theta", "gain.v2t", "gain.p2phi", "gain.chi2phi"};
    std::vector<double> gpack(FB_GS_BYTES / 8, 0.0);   // staging must outlive the asynchronous copy
    for (int i = 0; i < GS_COUNT; i++) {
        const BlobSec* sec = find(gnames[i]);
        if (!sec || sec->dtype != 0 || sec->nbytes != (uint64_t)28 * gs_NV(i) * 8) return fail(FB_ERR_ARG, std::string("fb_load_tables: bad or missing section ") + gnames[i]);
        for (int pt = 0; pt < 28; pt++) memcpy(gpack.data() + gs_off(i) + pt * gs_P(i), b + sec->off + (size_t)pt * gs_NV(i) * 8, (size_t)gs_NV(i) * 8);
    }
    CK(cudaMemcpyAsync((char*)s->d_pack + FB_TB_BYTES, gpack.data(), FB_GS_BYTES, cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->E.tables.pack = s->d_pack; s->E.tables.gains = s->d_pack + FB_TB_BYTES / 8; s->E.tables.geoid = s->d_geoid;
    s->tables_loaded = true;
    return FB_OK;
}

// upload an optional host column [N] into a device row of the scratch arena; returns device pointer or nullptr
// (rows of ONE arena per handle, allocated at the first initialisation and kept: an allocation per column was up to 19
// cudaMalloc / cudaFree pairs, each a device-wide synchronisation, per init call)
struct Uploader {
    fb_sim* s; int used = 0; cudaError_t err = cudaSuccess;
    static constexpr int ROWS = 40;      // fb_init_soa has 32 columns; fb_trim_params_soa 17 + 7 trim-state rows + the cost
    explicit Uploader(fb_sim* s_) : s(s_) {
        if (!s->d_upload) { err = cudaMalloc(&s->d_upload, sizeof(double) * ROWS * s->pitch); if (err != cudaSuccess) s->d_upload = nullptr; }
    }
    double* row(int n_rows = 1) {        // n_rows consecutive rows of the arena (device memory only)
        if (!s->d_upload || used + n_rows > ROWS) { if (err == cudaSuccess) err = cudaErrorMemoryAllocation; return nullptr; }
        double* d = s->d_upload + (long long)used * s->pitch; used += n_rows; return d;
    }
    const double* up(const double* h) {
        if (!h) return nullptr;
        double* d = row();
        if (!d) return nullptr;
        cudaError_t e = cudaMemcpyAsync(d, h, sizeof(double) * s->E.n, cudaMemcpyHostToDevice, s->stream);
        if (e != cudaSuccess) err = e;
        return d;
    }
    ~Uploader() { cudaStreamSynchronize(s->stream); }   // the caller's arrays and the arena are free again on return
};

static EnsembleDev mom_view(const fb_sim* s) {
    EnsembleDev E = s->E;
    E.mom_cnt = s->m_cnt; E.mom_sum = s->m_sum; E.mom_min = s->m_min; E.mom_max = s->m_max;
    return E;
}
static int32_t after_init(fb_sim* s, bool init_avionics) {
    s->n_steps = 0; s->join_step = 0; s->initialised = true;
    if (init_avionics && s->var_t == FB_C172X1) CK(FB_DISPATCH(s, avionics_init)(s));   // f_init!(avionics, vehicle)
    CK(cudaMemsetAsync(s->d_ev_cursor, 0, sizeof(long long) * s->pitch, s->stream));
    // A member that stops stepping (crash, altitude bounds, failed trim) writes no further samples: every sample starts
    // out as NaN, so that what a frozen member never wrote reads as missing rather than as stale memory; the moments
    // count the members that did write.
    if (s->E.log) {
        const long long total = s->E.log_capacity * s->E.log_n_fields * s->pitch;
        fill_kernel<<<(int)std::min<long long>((total + 255) / 256, 1 << 20), 256, 0, s->stream>>>(s->E.log, total, nan(""));
        CK(cudaGetLastError());
    }
    moments_reset_kernel<<<(int)((s->mom_records + 255) / 256), 256, 0, s->stream>>>(mom_view(s), 0, s->mom_records);
    CK(cudaGetLastError());
    s->reduced_upto = 0;
    // log sample 0 = y left by f_init! (sim.jl:405-408)
    for (int f = 0; f < s->E.log_n_fields; f++) {
        double* row = s->E.log ? s->E.log + (long long)f * s->pitch : s->d_scratch;
        CK(FB_DISPATCH(s, output)(s, s->cfg.log_fields[f], row));
        if (s->E.mom_cnt) {
            moments_kernel<<<(int)((s->E.n + 255) / 256), 256, 0, s->stream>>>(s->E, row, f);
            CK(cudaGetLastError());
        }
    }
    // A look at the initial state: how many members start within the height gate of the terrain? The answer reaches the host
    // while it is still busy elsewhere (never waited for) and picks the shape of the first hand-over passes (Ops::side_pass).
    s->side_seen = false; s->hint_pending = false;
    if (s->var_t != FB_C172XA2 && !s->force_complete) {
        if (!s->d_hint) { CK(cudaMalloc(&s->d_hint, sizeof(unsigned))); CK(cudaMallocHost(&s->h_hint, sizeof(unsigned))); CK(cudaEventCreateWithFlags(&s->hint_ev, cudaEventDisableTiming)); }
        *s->h_hint = 0;
        CK(cudaMemsetAsync(s->d_hint, 0, sizeof(unsigned), s->stream));
        CK(FB_DISPATCH(s, output)(s, FB_Y_H_O, s->d_scratch));
        count_near_kernel<<<(int)((s->pitch + 255) / 256), 256, 0, s->stream>>>(s->d_scratch, s->E.u + (long long)FB_U_TRN_ELEVATION * s->pitch, s->E.status, s->E.n, 4.0, s->d_hint);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(s->h_hint, s->d_hint, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaEventRecord(s->hint_ev, s->stream));
        s->hint_pending = true;
    }
    return FB_OK;
}

int32_t fb_init_trim(fb_handle s, const fb_trim_params_soa* p, double* trim_state_out, double* cost_out) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    if (!s->tables_loaded) return fail(FB_ERR_STATE, "fb_init_trim: call fb_load_tables first");
    CK(cudaSetDevice(s->cfg.device));
    fb_trim_params_soa z; memset(&z, 0, sizeof(z));
    if (!p) p = &z;
    double* d_ts = nullptr; double* d_cost = nullptr;
    {
        Uploader U{s};
        TrimSoA P;
        P.lat = U.up(p->lat); P.lon = U.up(p->lon); P.h_e = U.up(p->h_e); P.psi_nb = U.up(p->psi_nb); P.EAS = U.up(p->EAS);
        P.gamma_wb_n = U.up(p->gamma_wb_n); P.psi_wb_dot = U.up(p->psi_wb_dot); P.theta_wb_dot = U.up(p->theta_wb_dot);
        P.beta_a = U.up(p->beta_a); P.fuel_load = U.up(p->fuel_load); P.mixture = U.up(p->mixture); P.flaps = U.up(p->flaps);
        for (int j = 0; j < 5; j++) P.payload[j] = U.up(p->payload[j]);
        d_ts = U.row(7); d_cost = U.row();
        if (U.err != cudaSuccess) return fail(FB_ERR_CUDA, cudaGetErrorString(U.err));
        CK(FB_DISPATCH(s, trim)(s, P, d_ts, d_cost));
        if (trim_state_out) CK(cudaMemcpy2DAsync(trim_state_out, hp(s), d_ts, sizeof(double) * s->pitch, sizeof(double) * s->E.n, 7, cudaMemcpyDeviceToHost, s->stream));
        if (cost_out) CK(cudaMemcpyAsync(cost_out, d_cost, sizeof(double) * s->E.n, cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
    }
    return after_init(s, true);
}

__global__ void release_trim_failed_kernel(int* status, long long n, unsigned* count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && status[i] == FB_ST_TRIM_FAILED) { status[i] = FB_ST_RUNNING; atomicAdd(count, 1u); }
}
int32_t fb_release_trim_failed(fb_handle s, int64_t* n_released) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    if (!s->initialised) return fail(FB_ERR_STATE, "fb_release_trim_failed: not initialised");
    CK(cudaSetDevice(s->cfg.device));
    unsigned* d_n = reinterpret_cast<unsigned*>(s->d_scratch);
    CK(cudaMemsetAsync(d_n, 0, sizeof(unsigned), s->stream));
    release_trim_failed_kernel<<<(int)((s->E.n + 255) / 256), 256, 0, s->stream>>>(s->E.status, s->E.n, d_n);
    CK(cudaGetLastError());
    unsigned h_n = 0;
    CK(cudaMemcpyAsync(&h_n, d_n, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    if (n_released) *n_released = (int64_t)h_n;
    return FB_OK;
}

int32_t fb_init(fb_handle s, const fb_init_soa* ic) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    if (!s->tables_loaded) return fail(FB_ERR_STATE, "fb_init: call fb_load_tables first");
    CK(cudaSetDevice(s->cfg.device));
    fb_init_soa z; memset(&z, 0, sizeof(z));
    if (!ic) ic = &z;
    {
        Uploader U{s};
        InitSoA I;
        for (int j = 0; j < 4; j++) I.q_nb[j] = U.up(ic->q_nb[j]);
        I.lat = U.up(ic->lat); I.lon = U.up(ic->lon); I.h_e = U.up(ic->h_e);
        for (int j = 0; j < 3; j++) { I.w_wb_b[j] = U.up(ic->w_wb_b[j]); I.v_eb_n[j] = U.up(ic->v_eb_n[j]); }
        I.n_eng = U.up(ic->n_eng); I.engine_state = U.up(ic->engine_state); I.throttle = U.up(ic->throttle); I.mixture = U.up(ic->mixture);
        I.elevator = U.up(ic->elevator); I.aileron = U.up(ic->aileron); I.rudder = U.up(ic->rudder); I.flaps = U.up(ic->flaps);
        I.brake_left = U.up(ic->brake_left); I.brake_right = U.up(ic->brake_right); I.fuel_load = U.up(ic->fuel_load);
        for (int j = 0; j < 5; j++) I.payload[j] = U.up(ic->payload[j]);
        I.stall = U.up(ic->stall); I.alpha_filt = U.up(ic->alpha_filt); I.beta_filt = U.up(ic->beta_filt);
        if (U.err != cudaSuccess) return fail(FB_ERR_CUDA, cudaGetErrorString(U.err));
        CK(FB_DISPATCH(s, init)(s, I));
        CK(cudaStreamSynchronize(s->stream));
    }
    return after_init(s, true);
}

int32_t fb_init_state(fb_handle s, const double* x, const uint8_t* stall, const uint8_t* eng_state) {
    if (!s || !x) return fail(FB_ERR_ARG, "fb_init_state: null argument");
    JOIN(s);
    if (!s->tables_loaded) return fail(FB_ERR_STATE, "fb_init_state: call fb_load_tables first");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpy2DAsync(s->E.x, sizeof(double) * s->pitch, x, sizeof(double) * s->E.n, sizeof(double) * s->E.n, s->nx, cudaMemcpyHostToDevice, s->stream));
    if (stall) CK(cudaMemcpyAsync(s->E.stall, stall, s->E.n, cudaMemcpyHostToDevice, s->stream)); else CK(cudaMemsetAsync(s->E.stall, 0, s->pitch, s->stream));
    if (eng_state) CK(cudaMemcpyAsync(s->E.eng, eng_state, s->E.n, cudaMemcpyHostToDevice, s->stream)); else CK(cudaMemsetAsync(s->E.eng, FB_ENG_RUNNING, s->pitch, s->stream));
    CK(cudaMemsetAsync(s->E.status, 0, sizeof(int) * s->pitch, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return after_init(s, false);   // raw state injection keeps the controller memory as it is
}

int32_t fb_set_input(fb_handle s, int32_t field, const double* values, int64_t stride) {
    if (!s || !values) return fail(FB_ERR_ARG, "fb_set_input: null argument");
    JOIN(s);
    if (field < 0 || field >= FB_U__COUNT) return fail(FB_ERR_ARG, "fb_set_input: unknown field");
    if (stride != 0 && stride != 1) return fail(FB_ERR_ARG, "fb_set_input: stride must be 0 (broadcast) or 1");
    if (field >= FB_U_GDC_MODE_REQ && field <= FB_U_GDC_H2 && s->cfg.variant != FB_C172X2)
        return fail(FB_ERR_ARG, "fb_set_input: guidance inputs exist only on Cessna172Xv2");
    CK(cudaSetDevice(s->cfg.device));
    double* row = s->E.u + (long long)field * s->pitch;
    if (stride == 0) {
        int grid = (int)((s->pitch + 255) / 256);
        fill_kernel<<<grid, 256, 0, s->stream>>>(row, s->pitch, saturate_input(field, values[0]));
        CK(cudaGetLastError());
    } else {
        // The caller's array goes straight to a per-field staging row on the copy stream (a DMA at PCIe speed when the
        // array is pinned), which is all this call waits for: the caller may reuse the array on return. Saturation and
        // the write into the ensemble happen on the launch stream, ordered after whatever is still stepping there.
        if (!s->d_stage[field]) CK(cudaMalloc(&s->d_stage[field], sizeof(double) * s->pitch));
        CK(cudaStreamWaitEvent(s->copy_stream, s->stage_consumed[field], 0));      // previous value of this field consumed
        CK(cudaMemcpyAsync(s->d_stage[field], values, sizeof(double) * s->E.n, cudaMemcpyHostToDevice, s->copy_stream));
        CK(cudaEventRecord(s->copy_done, s->copy_stream));
        CK(cudaEventSynchronize(s->copy_done));
        int grid = (int)((s->E.n + 255) / 256);
        saturate_row_kernel<<<grid, 256, 0, s->stream>>>(row, s->d_stage[field], s->E.n, field);
        CK(cudaGetLastError());
        CK(cudaEventRecord(s->stage_consumed[field], s->stream));
    }
    return FB_OK;
}

int32_t fb_get_input(fb_handle s, int32_t field, double* out) {
    if (!s || !out) return fail(FB_ERR_ARG, "fb_get_input: null argument");
    JOIN(s);
    if (field < 0 || field >= FB_U__COUNT) return fail(FB_ERR_ARG, "fb_get_input: unknown field");
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpyAsync(out, s->E.u + (long long)field * s->pitch, sizeof(double) * s->E.n, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return FB_OK;
}

int32_t fb_schedule_inputs(fb_handle s, int64_t n_events, const int64_t* member, const int64_t* step, const int32_t* field, const double* value) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaStreamSynchronize(s->stream));
    cudaFree(s->d_ev_off); cudaFree(s->d_ev_step); cudaFree(s->d_ev_field); cudaFree(s->d_ev_value);
    s->d_ev_off = nullptr; s->d_ev_step = nullptr; s->d_ev_field = nullptr; s->d_ev_value = nullptr;
    s->E.ev_off = nullptr; s->E.ev_step = nullptr; s->E.ev_field = nullptr; s->E.ev_value = nullptr;
    CK(cudaMemsetAsync(s->d_ev_cursor, 0, sizeof(long long) * s->pitch, s->stream));
    if (n_events <= 0) return FB_OK;
    if (!member || !step || !field || !value) return fail(FB_ERR_ARG, "fb_schedule_inputs: null array");
    long long N = s->E.n;
    std::vector<long long> order(n_events);
    std::iota(order.begin(), order.end(), 0LL);
    for (long long k = 0; k < n_events; k++) {
        if (member[k] < 0 || member[k] >= N) return fail(FB_ERR_ARG, "fb_schedule_inputs: member out of range");
        if (field[k] < 0 || field[k] >= FB_U__COUNT) return fail(FB_ERR_ARG, "fb_schedule_inputs: unknown field");
    }
    std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) { return member[a] != member[b] ? member[a] < member[b] : step[a] < step[b]; });
    std::vector<long long> off(N + 1, 0), st(n_events); std::vector<int> fl(n_events); std::vector<double> vl(n_events);
    for (long long k = 0; k < n_events; k++) { long long o = order[k]; off[member[o] + 1]++; st[k] = step[o]; fl[k] = field[o]; vl[k] = value[o]; }
    for (long long i = 0; i < N; i++) off[i + 1] += off[i];
    CK(cudaMalloc(&s->d_ev_off, sizeof(long long) * (N + 1))); CK(cudaMalloc(&s->d_ev_step, sizeof(long long) * n_events));
    CK(cudaMalloc(&s->d_ev_field, sizeof(int) * n_events)); CK(cudaMalloc(&s->d_ev_value, sizeof(double) * n_events));
    CK(cudaMemcpy(s->d_ev_off, off.data(), sizeof(long long) * (N + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->d_ev_step, st.data(), sizeof(long long) * n_events, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->d_ev_field, fl.data(), sizeof(int) * n_events, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(s->d_ev_value, vl.data(), sizeof(double) * n_events, cudaMemcpyHostToDevice));
    s->E.ev_off = s->d_ev_off; s->E.ev_step = s->d_ev_step; s->E.ev_field = s->d_ev_field; s->E.ev_value = s->d_ev_value;
    return FB_OK;
}

static int32_t advance(fb_sim* s, long long n) {
    CK(cudaSetDevice(s->cfg.device));
    while (n > 0) {
        int k = (int)std::min<long long>(n, s->cfg.steps_per_launch);
        // Pick the launch shape from the most recent count of members near the terrain that has already arrived on
        // the host (never waits for the GPU): mostly airborne -> AIR kernel + hand-over pass, otherwise complete kernel.
        if (s->ground_ev_pending) {
            if (cudaEventQuery(s->ground_ev) == cudaSuccess) {
                s->ground_ev_pending = false;
                const bool air = !s->force_complete && (long long)*s->h_ground * 8 <= s->E.n;
                if (s->air_mode) s->side_seen = *s->h_ground > 0;
                if (air != s->air_mode) { if (int32_t rc = join(s)) return rc; }   // the complete-only launches take every member
                s->air_mode = air;
            } else (void)cudaGetLastError();   // cudaErrorNotReady is not an error: keep it out of the next launch check
        }
        if (s->hint_pending && cudaEventQuery(s->hint_ev) == cudaSuccess) {
            // the look at the initial state has arrived: it stands in for the first hand-over counts (same 1/8 rule: an ensemble
            // that starts on the runway goes to the complete kernel alone from its first launch — cfg 4: 7.5e8 against 7.1e8)
            s->hint_pending = false;
            if (*s->h_hint > 0) s->side_seen = true;
            const bool air = !s->force_complete && (long long)*s->h_hint * 8 <= s->E.n;
            if (air != s->air_mode && !s->ground_ev_pending) { if (int32_t rc = join(s)) return rc; s->air_mode = air; }
        } else if (s->hint_pending) (void)cudaGetLastError();
        // a member that has climbed away stays with the (slower) complete kernel until a join returns it: inside a long run
        // without reads, join every eighth launch
        if (s->air_mode && s->launches_since_join >= 8) { if (int32_t rc = join(s)) return rc; }
        int launched = 0;
        CK(FB_DISPATCH(s, step)(s, s->n_steps, k, &launched));
        s->launches_since_join++;
        if (!s->ground_ev_pending) {
            CK(cudaMemcpyAsync(s->h_ground, s->air_mode ? s->d_side : s->d_ground, sizeof(unsigned), cudaMemcpyDeviceToHost, s->stream));
            CK(cudaEventRecord(s->ground_ev, s->stream));
            s->ground_ev_pending = true;
        }
        s->launches += launched;
        s->n_steps += k;
        n -= k;
    }
    return FB_OK;
}

int32_t fb_step(fb_handle s, double t_advance) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    if (!s->initialised) return fail(FB_ERR_STATE, "fb_step: not initialised (call fb_init_trim / fb_init / fb_init_state)");
    if (t_advance < 0) return fail(FB_ERR_ARG, "fb_step: negative advance");
    return advance(s, llround(t_advance / s->cfg.dt));
}

int32_t fb_run(fb_handle s) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    if (!s->initialised) return fail(FB_ERR_STATE, "fb_run: not initialised");
    long long total = llround((s->cfg.t_end - s->cfg.t_start) / s->cfg.dt);
    return advance(s, std::max(0LL, total - s->n_steps));
}

int32_t fb_time(fb_handle s, double* t, int64_t* n_steps) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    if (t) *t = s->cfg.t_start + s->n_steps * s->cfg.dt;
    if (n_steps) *n_steps = s->n_steps;
    return FB_OK;
}

int32_t fb_nx(fb_handle s, int32_t* nx) { if (!s || !nx) return fail(FB_ERR_ARG, "null argument"); *nx = s->nx; return FB_OK; }

int32_t fb_get_state(fb_handle s, double* x) {
    if (!s || !x) return fail(FB_ERR_ARG, "fb_get_state: null argument");
    JOIN(s);
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpy2DAsync(x, hp(s), s->E.x, sizeof(double) * s->pitch, sizeof(double) * s->E.n, s->nx, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return FB_OK;
}

int32_t fb_wait_state(fb_handle s) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    if (s->snap_pending) { CK(cudaEventSynchronize(s->snap_done)); s->snap_pending = false; }
    return FB_OK;
}

int32_t fb_get_state_async(fb_handle s, double* x) {
    if (!s || !x) return fail(FB_ERR_ARG, "fb_get_state_async: null argument");
    JOIN(s);
    CK(cudaSetDevice(s->cfg.device));
    int32_t rc = fb_wait_state(s);   // the snapshot buffer is free again once the previous copy has landed
    if (rc != FB_OK) return rc;
    if (!s->d_snap) {
        CK(cudaMalloc(&s->d_snap, sizeof(double) * s->pitch * s->nx));
        CK(cudaEventCreateWithFlags(&s->snap_ready, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s->snap_done, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&s->snap_stream, cudaStreamNonBlocking));
    }
    CK(cudaMemcpyAsync(s->d_snap, s->E.x, sizeof(double) * s->pitch * s->nx, cudaMemcpyDeviceToDevice, s->stream));
    CK(cudaEventRecord(s->snap_ready, s->stream));
    CK(cudaStreamWaitEvent(s->snap_stream, s->snap_ready, 0));
    CK(cudaMemcpy2DAsync(x, sizeof(double) * s->E.n, s->d_snap, sizeof(double) * s->pitch, sizeof(double) * s->E.n, s->nx, cudaMemcpyDeviceToHost, s->snap_stream));
    CK(cudaEventRecord(s->snap_done, s->snap_stream));
    s->snap_pending = true;
    return FB_OK;
}

int32_t fb_get_discrete(fb_handle s, uint8_t* stall, uint8_t* eng_state) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    CK(cudaSetDevice(s->cfg.device));
    if (stall) CK(cudaMemcpyAsync(stall, s->E.stall, s->E.n, cudaMemcpyDeviceToHost, s->stream));
    if (eng_state) CK(cudaMemcpyAsync(eng_state, s->E.eng, s->E.n, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return FB_OK;
}

int32_t fb_get_output(fb_handle s, int32_t field, double* out) {
    if (!s || !out) return fail(FB_ERR_ARG, "fb_get_output: null argument");
    JOIN(s);
    if (field < 0 || field >= FB_Y__COUNT) return fail(FB_ERR_ARG, "fb_get_output: unknown field");
    if (!s->initialised) return fail(FB_ERR_STATE, "fb_get_output: not initialised");
    CK(cudaSetDevice(s->cfg.device));
    CK(FB_DISPATCH(s, output)(s, field, s->d_scratch));
    CK(cudaMemcpyAsync(out, s->d_scratch, sizeof(double) * s->E.n, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return FB_OK;
}

int32_t fb_status(fb_handle s, int32_t* status) {
    if (!s || !status) return fail(FB_ERR_ARG, "fb_status: null argument");
    JOIN(s);
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaMemcpyAsync(status, s->E.status, sizeof(int) * s->E.n, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    long long total = llround((s->cfg.t_end - s->cfg.t_start) / s->cfg.dt);
    if (s->n_steps >= total && total > 0)
        for (long long i = 0; i < s->E.n; i++) if (status[i] == FB_ST_RUNNING) status[i] = FB_ST_DONE;
    return FB_OK;
}

int32_t fb_log_info(fb_handle s, int64_t* n_samples, int32_t* n_fields) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    long long ns = 0;
    if (s->E.log_every > 0 && s->E.log_n_fields > 0 && s->initialised) ns = std::min<long long>(s->E.log_capacity, s->n_steps / s->E.log_every + 1);
    if (n_samples) *n_samples = ns;
    if (n_fields) *n_fields = s->E.log_n_fields;
    return FB_OK;
}

int32_t fb_log_times(fb_handle s, double* t) {
    if (!s || !t) return fail(FB_ERR_ARG, "fb_log_times: null argument");
    int64_t ns; int32_t nf;
    fb_log_info(s, &ns, &nf);
    for (int64_t k = 0; k < ns; k++) t[k] = s->cfg.t_start + (double)(k * s->E.log_every) * s->cfg.dt;
    return FB_OK;
}

int32_t fb_log_fetch(fb_handle s, int32_t field, double* out) {
    if (!s || !out) return fail(FB_ERR_ARG, "fb_log_fetch: null argument");
    JOIN(s);
    int64_t ns; int32_t nf;
    fb_log_info(s, &ns, &nf);
    int fi = -1;
    for (int f = 0; f < nf; f++) if (s->cfg.log_fields[f] == field) fi = f;
    if (fi < 0) return fail(FB_ERR_ARG, "fb_log_fetch: field is not logged");
    if (!s->E.log) return fail(FB_ERR_STATE, "fb_log_fetch: the per-member samples are not kept (log_flags = FB_LOG_MOMENTS); use fb_log_stats");
    CK(cudaSetDevice(s->cfg.device));
    // sample k of field fi lives at log[(k * nf + fi) * pitch]
    CK(cudaMemcpy2DAsync(out, sizeof(double) * s->E.n, s->E.log + (long long)fi * s->pitch, sizeof(double) * s->pitch * nf,
                         sizeof(double) * s->E.n, ns, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return FB_OK;
}

static void moments_from_raw(const char* raw, long long R, long long r, fb_moments* out) {
    unsigned long long c; double sm[2], mn, mx;
    memcpy(&c, raw + r * 8, 8); memcpy(sm, raw + R * 8 + r * 16, 16); memcpy(&mn, raw + R * 24 + r * 8, 8); memcpy(&mx, raw + R * 32 + r * 8, 8);
    out->count = (int64_t)c; out->sum = sm[0]; out->sum_sq = sm[1]; out->min = mn; out->max = mx;
}

int32_t fb_stats(fb_handle s, int32_t field, fb_moments* out) {
    if (!s || !out) return fail(FB_ERR_ARG, "fb_stats: null argument");
    JOIN(s);
    if (field < 0 || field >= FB_Y__COUNT) return fail(FB_ERR_ARG, "fb_stats: unknown field");
    if (!s->initialised) return fail(FB_ERR_STATE, "fb_stats: not initialised");
    CK(cudaSetDevice(s->cfg.device));
    const long long R = s->mom_records, rec = R - 1;   // the scratch record
    CK(FB_DISPATCH(s, output)(s, field, s->d_scratch));
    moments_reset_kernel<<<1, 32, 0, s->stream>>>(mom_view(s), rec, 1);
    moments_kernel<<<(int)((s->E.n + 255) / 256), 256, 0, s->stream>>>(mom_view(s), s->d_scratch, rec);
    CK(cudaGetLastError());
    char raw[40];   // one record gathered from the four arrays
    CK(cudaMemcpyAsync(raw, s->m_cnt + rec, 8, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(raw + 8, s->m_sum + 2 * rec, 16, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(raw + 24, s->m_min + rec, 8, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(raw + 32, s->m_max + rec, 8, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    moments_from_raw(raw, 1, 0, out);
    return FB_OK;
}

int32_t fb_log_stats_begin(fb_handle s, int64_t first, int64_t count) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    if (!s->E.mom_cnt) return fail(FB_ERR_STATE, "fb_log_stats: no moments are kept (no log configured, or log_flags without FB_LOG_MOMENTS)");
    int64_t ns; int32_t nf;
    fb_log_info(s, &ns, &nf);
    if (first < 0 || count <= 0 || first + count > ns) return fail(FB_ERR_ARG, "fb_log_stats: sample range outside the samples taken so far");
    CK(cudaSetDevice(s->cfg.device));
    if (s->mom_pending) { CK(cudaEventSynchronize(s->mom_done)); s->mom_pending = false; }
    const size_t need = (size_t)count * nf * 40;
    if (need > s->h_mom_bytes) {
        if (s->h_mom) cudaFreeHost(s->h_mom);
        s->h_mom = nullptr; s->h_mom_bytes = 0;
        CK(cudaMallocHost(&s->h_mom, need));
        s->h_mom_bytes = need;
    }
    // behind everything submitted so far on the launch stream and behind a pending all-reduce
    CK(cudaEventRecord(s->mom_ready, s->stream));
    CK(cudaStreamWaitEvent(s->snap_stream, s->mom_ready, 0));
    if (s->comm_busy) CK(cudaStreamWaitEvent(s->snap_stream, s->comm_out, 0));
    const long long r0 = first * nf, n = count * nf;
    CK(cudaMemcpyAsync(s->h_mom, s->m_cnt + r0, n * 8, cudaMemcpyDeviceToHost, s->snap_stream));
    CK(cudaMemcpyAsync(s->h_mom + n * 8, s->m_sum + 2 * r0, n * 16, cudaMemcpyDeviceToHost, s->snap_stream));
    CK(cudaMemcpyAsync(s->h_mom + n * 24, s->m_min + r0, n * 8, cudaMemcpyDeviceToHost, s->snap_stream));
    CK(cudaMemcpyAsync(s->h_mom + n * 32, s->m_max + r0, n * 8, cudaMemcpyDeviceToHost, s->snap_stream));
    CK(cudaEventRecord(s->mom_done, s->snap_stream));
    s->mom_pending = true; s->mom_first = first; s->mom_count = count;
    return FB_OK;
}

int32_t fb_log_stats_end(fb_handle s, int32_t field, fb_moments* out) {
    if (!s || !out) return fail(FB_ERR_ARG, "fb_log_stats_end: null argument");
    if (s->mom_count <= 0) return fail(FB_ERR_STATE, "fb_log_stats_end: no fb_log_stats_begin before");
    int fi = -1;
    for (int f = 0; f < s->E.log_n_fields; f++) if (s->cfg.log_fields[f] == field) fi = f;
    if (fi < 0) return fail(FB_ERR_ARG, "fb_log_stats: field is not logged");
    CK(cudaSetDevice(s->cfg.device));
    if (s->mom_pending) { CK(cudaEventSynchronize(s->mom_done)); s->mom_pending = false; }
    const long long nf = s->E.log_n_fields, n = s->mom_count * nf;
    for (long long k = 0; k < s->mom_count; k++) moments_from_raw(s->h_mom, n, k * nf + fi, out + k);
    return FB_OK;
}

int32_t fb_log_stats(fb_handle s, int32_t field, int64_t first, int64_t count, fb_moments* out) {
    if (!out) return fail(FB_ERR_ARG, "fb_log_stats: null argument");
    int32_t rc = fb_log_stats_begin(s, first, count);
    return rc != FB_OK ? rc : fb_log_stats_end(s, field, out);
}

// ------------------------------------------------------------------------------------------------ sharded ensembles
int32_t fb_comm_unique_id(void* id128) {
    if (!id128) return fail(FB_ERR_ARG, "fb_comm_unique_id: null argument");
    NcclApi* N = nccl_api();
    if (!N) return fail(FB_ERR_STATE, "fb_comm_unique_id: NCCL (libnccl.so.2) could not be loaded");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NK(N->GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return FB_OK;
}

static int32_t comm_prepare(fb_sim* s) {   // side stream + events of the all-reduce
    if (!s->comm_stream) {
        // Highest priority: the integrator's CTAs fill every SM (one per SM, the whole register file), and the work distributor
        // would otherwise dispatch the all-reduce only when the NEXT step's launch runs out of CTAs. With priority its CTA takes
        // the first slot a retiring integrator CTA frees (measured at N = 2: latency 54 ms -> a few ms; the stepping rate does
        // not change either way, the reduction is a few KB).
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CK(cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, greatest));
        CK(cudaEventCreateWithFlags(&s->comm_in, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s->comm_out, cudaEventDisableTiming));
    }
    return FB_OK;
}

int32_t fb_comm_init(fb_handle s, int32_t n_ranks, int32_t rank, const void* id128) {
    if (!s || !id128) return fail(FB_ERR_ARG, "fb_comm_init: null argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(FB_ERR_ARG, "fb_comm_init: bad rank / n_ranks");
    if (s->comm) return fail(FB_ERR_STATE, "fb_comm_init: the handle already has a communicator");
    NcclApi* N = nccl_api();
    if (!N) return fail(FB_ERR_STATE, "fb_comm_init: NCCL (libnccl.so.2) could not be loaded");
    CK(cudaSetDevice(s->cfg.device));
    ncclUniqueId id; memcpy(&id, id128, 128);
    NK(N->CommInitRank(&s->comm, n_ranks, id, rank));
    s->comm_owned = true; s->comm_ranks = n_ranks;
    return comm_prepare(s);
}

// enqueue the all-reduce of the log samples [reduced_upto, upto) on the side stream (inside an NCCL group when the caller
// drives several communicators from one thread)
static int32_t stats_allreduce_enqueue(fb_sim* s, long long upto) {
    NcclApi* N = nccl_api();
    const long long nf = s->E.log_n_fields, r0 = s->reduced_upto * nf, n = (upto - s->reduced_upto) * nf;
    NK(N->AllReduce(s->m_cnt + r0, s->m_cnt + r0, (size_t)n, ncclUint64, ncclSum, s->comm, s->comm_stream));
    NK(N->AllReduce(s->m_sum + 2 * r0, s->m_sum + 2 * r0, (size_t)(2 * n), ncclFloat64, ncclSum, s->comm, s->comm_stream));
    NK(N->AllReduce(s->m_min + r0, s->m_min + r0, (size_t)n, ncclFloat64, ncclMin, s->comm, s->comm_stream));
    NK(N->AllReduce(s->m_max + r0, s->m_max + r0, (size_t)n, ncclFloat64, ncclMax, s->comm, s->comm_stream));
    return FB_OK;
}
static int32_t stats_allreduce_pre(fb_sim* s, long long* upto) {
    if (!s->comm) return fail(FB_ERR_STATE, "fb_log_stats_allreduce: no communicator (fb_comm_init / fb_group_create)");
    if (!s->E.mom_cnt) return fail(FB_ERR_STATE, "fb_log_stats_allreduce: no moments are kept");
    int64_t ns; int32_t nf;
    fb_log_info(s, &ns, &nf);
    *upto = ns;
    if (ns <= s->reduced_upto) return FB_OK;
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaEventRecord(s->comm_in, s->stream));                 // the samples are final once the launches so far have run
    CK(cudaStreamWaitEvent(s->comm_stream, s->comm_in, 0));
    if (s->side_active) CK(cudaStreamWaitEvent(s->comm_stream, s->side_done, 0));   // ... and the hand-over passes so far (no join: stepping goes on)
    if (s->mom_pending) CK(cudaStreamWaitEvent(s->comm_stream, s->mom_done, 0));   // a fetch in flight reads the same arrays
    cudaEvent_t t0, t1;
    CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
    s->comm_timers.push_back({t0, t1});
    CK(cudaEventRecord(t0, s->comm_stream));
    return FB_OK;
}
static int32_t stats_allreduce_post(fb_sim* s, long long upto) {
    if (upto <= s->reduced_upto) return FB_OK;
    CK(cudaSetDevice(s->cfg.device));
    CK(cudaEventRecord(s->comm_timers.back().second, s->comm_stream));
    CK(cudaEventRecord(s->comm_out, s->comm_stream));
    s->comm_busy = true; s->reduced_upto = upto; s->comm_calls++;
    return FB_OK;
}

int32_t fb_log_stats_allreduce(fb_handle s) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    long long upto = 0;
    int32_t rc = stats_allreduce_pre(s, &upto);
    if (rc != FB_OK || upto <= s->reduced_upto) return rc;
    NcclApi* N = nccl_api();
    NK(N->GroupStart());
    rc = stats_allreduce_enqueue(s, upto);
    NK(N->GroupEnd());
    return rc != FB_OK ? rc : stats_allreduce_post(s, upto);
}

int32_t fb_comm_time(fb_handle s, double* ms_total, int64_t* n_calls) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    CK(cudaSetDevice(s->cfg.device));
    if (s->comm_stream) CK(cudaStreamSynchronize(s->comm_stream));
    for (auto& t : s->comm_timers) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, t.first, t.second) == cudaSuccess) s->comm_ms += ms; else (void)cudaGetLastError();
        cudaEventDestroy(t.first); cudaEventDestroy(t.second);
    }
    s->comm_timers.clear();
    if (ms_total) *ms_total = s->comm_ms;
    if (n_calls) *n_calls = s->comm_calls;
    return FB_OK;
}

int32_t fb_rhs(fb_handle s, double* xdot) {
    if (!s || !xdot) return fail(FB_ERR_ARG, "fb_rhs: null argument");
    JOIN(s);
    if (!s->initialised) return fail(FB_ERR_STATE, "fb_rhs: not initialised");
    CK(cudaSetDevice(s->cfg.device));
    CK(FB_DISPATCH(s, rhs)(s, s->d_scratch));
    CK(cudaMemcpy2DAsync(xdot, sizeof(double) * s->E.n, s->d_scratch, sizeof(double) * s->pitch, sizeof(double) * s->E.n, s->nx, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return FB_OK;
}

int32_t fb_fstep(fb_handle s, int32_t* changed) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    if (!s->initialised) return fail(FB_ERR_STATE, "fb_fstep: not initialised");
    CK(cudaSetDevice(s->cfg.device));
    int* d_changed = (int*)s->d_scratch;   // [pitch] ints fit in the scratch rows
    CK(FB_DISPATCH(s, fstep)(s, changed ? d_changed : nullptr));
    if (changed) CK(cudaMemcpyAsync(changed, d_changed, sizeof(int) * s->E.n, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return FB_OK;
}

int32_t fb_fperiodic(fb_handle s) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    if (!s->initialised) return fail(FB_ERR_STATE, "fb_fperiodic: not initialised");
    CK(cudaSetDevice(s->cfg.device));
    CK(FB_DISPATCH(s, fperiodic)(s));
    CK(cudaStreamSynchronize(s->stream));
    return FB_OK;
}

int32_t fb_counters(fb_handle s, int64_t* kernel_launches, int64_t* rhs_evals) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    if (kernel_launches) *kernel_launches = s->launches;
    if (rhs_evals) {
        CK(cudaSetDevice(s->cfg.device));
        unsigned long long c = 0;
        CK(cudaMemcpyAsync(&c, s->d_rhs_count, sizeof(c), cudaMemcpyDeviceToHost, s->stream));
        CK(cudaStreamSynchronize(s->stream));
        *rhs_evals = (int64_t)c;
    }
    return FB_OK;
}

int32_t fb_linearize_dims(fb_handle s, int32_t* dims) {
    if (!s || !dims) return fail(FB_ERR_ARG, "fb_linearize_dims: null argument");
    if (s->cfg.variant == FB_C172XA2) return fail(FB_ERR_ARG, "fb_linearize: not offered for FB_C172XA2 (the reference's XStateSpace knows first-order actuators only, c172x.jl:332-498)");
    dims[0] = s->cfg.variant == FB_C172S ? 16 : 20; dims[1] = 4; dims[2] = s->cfg.variant == FB_C172S ? 33 : 38;
    return FB_OK;
}

int32_t fb_linearize(fb_handle s, double* A, double* B, double* C, double* D, double* x0, double* u0, double* y0, double* xdot0) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    if (!s->initialised) return fail(FB_ERR_STATE, "fb_linearize: not initialised (trim first)");
    if (s->cfg.kinematics != FB_NED) return fail(FB_ERR_ARG, "fb_linearize: needs NED kinematics (the reference defines XStateSpace for Vehicle{NED} only)");
    CK(cudaSetDevice(s->cfg.device));
    int32_t d[3]; if (int32_t rc = fb_linearize_dims(s, d)) return rc;
    const long long nxs = d[0], nus = d[1], nys = d[2], P = s->pitch, N = s->E.n;
    const long long counts[8] = {nxs * nxs, nxs * nus, nys * nxs, nys * nus, nxs, nus, nys, nxs};
    double* host[8] = {A, B, C, D, x0, u0, y0, xdot0};
    long long total = 0; for (int k = 0; k < 8; k++) total += counts[k];
    double* dbuf = nullptr;
    CK(cudaMalloc(&dbuf, sizeof(double) * total * P));
    double* dp[8]; long long off = 0;
    for (int k = 0; k < 8; k++) { dp[k] = dbuf + off * P; off += counts[k]; }
    cudaError_t e = FB_DISPATCH(s, linearize)(s, dp[0], dp[1], dp[2], dp[3], dp[4], dp[5], dp[6], dp[7]);
    for (int k = 0; k < 8 && e == cudaSuccess; k++)
        if (host[k]) e = cudaMemcpy2DAsync(host[k], sizeof(double) * N, dp[k], sizeof(double) * P, sizeof(double) * N, counts[k], cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(dbuf);
    if (e != cudaSuccess) return fail(FB_ERR_CUDA, std::string("fb_linearize: ") + cudaGetErrorString(e));
    return FB_OK;
}

// ------------------------------------------------------------------------------------------------ LinearizedSS
struct fb_lss {
    int device = 0, nx = 0, nu = 0, ny = 0;
    long long n = 0, pitch = 0, n_steps = 0;
    double dt = 0.0;
    double* buf = nullptr;     // one allocation: A B C D x0 u0 y0 xd0 x u y
    double* y = nullptr;
    LssDev L{};
    cudaStream_t stream = nullptr;
    bool has_model = false;
};

int32_t fb_lss_create(int32_t device, int64_t n, int32_t nx, int32_t nu, int32_t ny, double dt, fb_lss_handle* out) {
    if (!out) return fail(FB_ERR_ARG, "fb_lss_create: null argument");
    if (n <= 0 || !(dt > 0.0)) return fail(FB_ERR_ARG, "fb_lss_create: n and dt must be positive");
    if (!((nx == 16 && ny == 33) || (nx == 20 && ny == 38)) || nu != 4)
        return fail(FB_ERR_ARG, "fb_lss_create: (nx, nu, ny) must be (16, 4, 33) [C172S] or (20, 4, 38) [C172X], the XStateSpace / UStateSpace / YStateSpace sizes");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(FB_ERR_CUDA, "fb_lss_create: no CUDA device (there is no CPU fallback)");
    CK(cudaSetDevice(device));
    fb_lss* l = new fb_lss();
    l->device = device; l->nx = nx; l->nu = nu; l->ny = ny; l->n = n; l->pitch = (n + 31) / 32 * 32; l->dt = dt;
    const long long P = l->pitch;
    const long long counts[11] = {(long long)nx * nx, (long long)nx * nu, (long long)ny * nx, (long long)ny * nu, nx, nu, ny, nx, nx, nu, ny};
    long long total = 0; for (long long c : counts) total += c;
    if (cudaMalloc(&l->buf, sizeof(double) * total * P) != cudaSuccess) { delete l; return fail(FB_ERR_ALLOC, "fb_lss_create: out of device memory"); }
    CK(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking));
    CK(cudaMemsetAsync(l->buf, 0, sizeof(double) * total * P, l->stream));
    double* p[11]; long long off = 0;
    for (int k = 0; k < 11; k++) { p[k] = l->buf + off * P; off += counts[k]; }
    l->L.A = p[0]; l->L.B = p[1]; l->L.C = p[2]; l->L.D = p[3]; l->L.x0 = p[4]; l->L.u0 = p[5]; l->L.y0 = p[6]; l->L.xd0 = p[7];
    l->L.x = p[8]; l->L.u = p[9]; l->y = p[10]; l->L.pitch = P; l->L.n = n;
    *out = l;
    return FB_OK;
}

int32_t fb_lss_destroy(fb_lss_handle l) {
    if (!l) return FB_OK;
    cudaSetDevice(l->device);
    cudaStreamSynchronize(l->stream);
    cudaFree(l->buf); cudaStreamDestroy(l->stream);
    delete l;
    return FB_OK;
}

static cudaError_t lss_rows_h2d(fb_lss* l, double* dst, const double* src, long long rows) {
    return cudaMemcpy2DAsync(dst, sizeof(double) * l->pitch, src, sizeof(double) * l->n, sizeof(double) * l->n, rows, cudaMemcpyHostToDevice, l->stream);
}
static void lss_reset_to_operating_point(fb_lss* l) {   // X(lss) = copy(x0), U(lss) = copy(u0), linearization.jl:152-153
    cudaMemcpyAsync(l->L.x, l->L.x0, sizeof(double) * l->nx * l->pitch, cudaMemcpyDeviceToDevice, l->stream);
    cudaMemcpyAsync(l->L.u, l->L.u0, sizeof(double) * l->nu * l->pitch, cudaMemcpyDeviceToDevice, l->stream);
    l->n_steps = 0;
}

int32_t fb_lss_set_model(fb_lss_handle l, const double* A, const double* B, const double* C, const double* D,
                         const double* x0, const double* u0, const double* y0, const double* xdot0) {
    if (!l || !A || !B || !C || !D || !x0 || !u0 || !y0 || !xdot0) return fail(FB_ERR_ARG, "fb_lss_set_model: null argument");
    CK(cudaSetDevice(l->device));
    const long long nx = l->nx, nu = l->nu, ny = l->ny;
    CK(lss_rows_h2d(l, l->L.A, A, nx * nx)); CK(lss_rows_h2d(l, l->L.B, B, nx * nu)); CK(lss_rows_h2d(l, l->L.C, C, ny * nx));
    CK(lss_rows_h2d(l, l->L.D, D, ny * nu)); CK(lss_rows_h2d(l, l->L.x0, x0, nx)); CK(lss_rows_h2d(l, l->L.u0, u0, nu));
    CK(lss_rows_h2d(l, l->L.y0, y0, ny)); CK(lss_rows_h2d(l, l->L.xd0, xdot0, nx));
    lss_reset_to_operating_point(l);
    CK(cudaStreamSynchronize(l->stream));
    l->has_model = true;
    return FB_OK;
}

int32_t fb_lss_from_linearize(fb_lss_handle l, fb_handle s) {
    if (!l || !s) return fail(FB_ERR_ARG, "fb_lss_from_linearize: null argument");
    JOIN(s);
    if (!s->initialised) return fail(FB_ERR_STATE, "fb_lss_from_linearize: the simulation is not initialised (trim first)");
    if (s->cfg.kinematics != FB_NED) return fail(FB_ERR_ARG, "fb_lss_from_linearize: needs NED kinematics");
    int32_t d[3]; if (int32_t rc = fb_linearize_dims(s, d)) return rc;
    if (d[0] != l->nx || d[1] != l->nu || d[2] != l->ny || s->E.n != l->n || s->cfg.device != l->device)
        return fail(FB_ERR_ARG, "fb_lss_from_linearize: sizes / device of the two handles differ");
    CK(cudaSetDevice(l->device));
    CK(cudaStreamSynchronize(l->stream));
    CK(FB_DISPATCH(s, linearize)(s, l->L.A, l->L.B, l->L.C, l->L.D, l->L.x0, l->L.u0, l->L.y0, l->L.xd0));   // device to device
    CK(cudaStreamSynchronize(s->stream));
    lss_reset_to_operating_point(l);
    CK(cudaStreamSynchronize(l->stream));
    l->has_model = true;
    return FB_OK;
}

int32_t fb_lss_set_state(fb_lss_handle l, const double* x) {
    if (!l || !x) return fail(FB_ERR_ARG, "fb_lss_set_state: null argument");
    CK(cudaSetDevice(l->device));
    CK(lss_rows_h2d(l, l->L.x, x, l->nx));
    CK(cudaStreamSynchronize(l->stream));
    return FB_OK;
}
int32_t fb_lss_set_input(fb_lss_handle l, const double* u) {
    if (!l || !u) return fail(FB_ERR_ARG, "fb_lss_set_input: null argument");
    CK(cudaSetDevice(l->device));
    CK(lss_rows_h2d(l, l->L.u, u, l->nu));
    CK(cudaStreamSynchronize(l->stream));
    return FB_OK;
}
int32_t fb_lss_get_state(fb_lss_handle l, double* x) {
    if (!l || !x) return fail(FB_ERR_ARG, "fb_lss_get_state: null argument");
    CK(cudaSetDevice(l->device));
    CK(cudaMemcpy2DAsync(x, sizeof(double) * l->n, l->L.x, sizeof(double) * l->pitch, sizeof(double) * l->n, l->nx, cudaMemcpyDeviceToHost, l->stream));
    CK(cudaStreamSynchronize(l->stream));
    return FB_OK;
}
int32_t fb_lss_get_input(fb_lss_handle l, double* u) {
    if (!l || !u) return fail(FB_ERR_ARG, "fb_lss_get_input: null argument");
    CK(cudaSetDevice(l->device));
    CK(cudaMemcpy2DAsync(u, sizeof(double) * l->n, l->L.u, sizeof(double) * l->pitch, sizeof(double) * l->n, l->nu, cudaMemcpyDeviceToHost, l->stream));
    CK(cudaStreamSynchronize(l->stream));
    return FB_OK;
}
int32_t fb_lss_get_output(fb_lss_handle l, double* y) {
    if (!l || !y) return fail(FB_ERR_ARG, "fb_lss_get_output: null argument");
    if (!l->has_model) return fail(FB_ERR_STATE, "fb_lss_get_output: no model (fb_lss_set_model / fb_lss_from_linearize)");
    CK(cudaSetDevice(l->device));
    lss_output_kernel<<<(int)((l->n + 127) / 128), 128, 0, l->stream>>>(l->L, l->nx, l->nu, l->ny, l->y);
    CK(cudaGetLastError());
    CK(cudaMemcpy2DAsync(y, sizeof(double) * l->n, l->y, sizeof(double) * l->pitch, sizeof(double) * l->n, l->ny, cudaMemcpyDeviceToHost, l->stream));
    CK(cudaStreamSynchronize(l->stream));
    return FB_OK;
}
int32_t fb_lss_step(fb_lss_handle l, double t_advance) {
    if (!l) return fail(FB_ERR_ARG, "null handle");
    if (!l->has_model) return fail(FB_ERR_STATE, "fb_lss_step: no model (fb_lss_set_model / fb_lss_from_linearize)");
    if (t_advance < 0) return fail(FB_ERR_ARG, "fb_lss_step: negative advance");
    CK(cudaSetDevice(l->device));
    long long n = llround(t_advance / l->dt);
    const int wpb = 8;                                            // warps per CTA; a warp holds two 16-state members or one 20-state member
    const long long mpb = (l->nx == 16) ? 2 * wpb : wpb;
    int grid = (int)((l->n + mpb - 1) / mpb);
    while (n > 0) {
        int k = (int)std::min<long long>(n, 1 << 20);
        if (l->nx == 16) lss_step_kernel<16, 4><<<grid, wpb * 32, 0, l->stream>>>(l->L, k, l->dt);
        else lss_step_kernel<20, 4><<<grid, wpb * 32, 0, l->stream>>>(l->L, k, l->dt);
        CK(cudaGetLastError());
        l->n_steps += k; n -= k;
    }
    return FB_OK;
}
int32_t fb_lss_sync(fb_lss_handle l) {
    if (!l) return fail(FB_ERR_ARG, "null handle");
    CK(cudaSetDevice(l->device));
    CK(cudaStreamSynchronize(l->stream));
    return FB_OK;
}
int32_t fb_lss_stream(fb_lss_handle l, void** stream) {
    if (!l || !stream) return fail(FB_ERR_ARG, "fb_lss_stream: null argument");
    *stream = (void*)l->stream;
    return FB_OK;
}

// measured FP64 FMA throughput of the device (TFLOP/s, FMA = 2 flop): the roofline denominator of this path
int32_t fb_measure_fp64_peak(int32_t device, double* tflops) {
    if (!tflops) return fail(FB_ERR_ARG, "null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(FB_ERR_CUDA, "no CUDA device");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double* d = nullptr;
    CK(cudaMalloc(&d, 64));
    const int iters = 1 << 15, blocks = prop.multiProcessorCount * 8, threads = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(e0));
        fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) * 1e-12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best;
    return FB_OK;
}

int32_t fb_device_state(fb_handle s, void** d_x, int64_t* pitch_elems) {
    if (!s) return fail(FB_ERR_ARG, "null handle");
    JOIN(s);
    if (d_x) *d_x = s->E.x;
    if (pitch_elems) *pitch_elems = s->pitch;
    return FB_OK;
}

int32_t fb_stream(fb_handle s, void** stream) {
    if (!s || !stream) return fail(FB_ERR_ARG, "null argument");
    *stream = (void*)s->stream;
    return FB_OK;
}

// ------------------------------------------------------------------------------------------------ one process, several GPUs
struct fb_group {
    std::vector<fb_sim*> sims;
    std::vector<long long> off;      // off[k] .. off[k + 1]: global members of shard k
    long long n = 0;
};
#define GK(call) do { int32_t rc_ = (call); if (rc_ != FB_OK) return rc_; } while (0)

int32_t fb_group_destroy(fb_group_handle g) {
    if (!g) return FB_OK;
    for (fb_sim* s : g->sims) fb_destroy(s);
    delete g;
    return FB_OK;
}

int32_t fb_group_create(const fb_config* cfg, const int32_t* devices, int32_t n_dev, fb_group_handle* out) {
    if (!cfg || !devices || !out) return fail(FB_ERR_ARG, "fb_group_create: null argument");
    *out = nullptr;
    if (n_dev < 1 || cfg->n_members < n_dev) return fail(FB_ERR_ARG, "fb_group_create: need at least one device and one member per device");
    for (int a = 0; a < n_dev; a++) for (int b = 0; b < a; b++) if (devices[a] == devices[b]) return fail(FB_ERR_ARG, "fb_group_create: a device is listed twice");
    NcclApi* N = n_dev > 1 ? nccl_api() : nullptr;
    if (n_dev > 1 && !N) return fail(FB_ERR_STATE, "fb_group_create: NCCL (libnccl.so.2) could not be loaded");
    fb_group* g = new fb_group();
    g->n = cfg->n_members;
    g->off.resize(n_dev + 1);
    for (int k = 0; k <= n_dev; k++) g->off[k] = (k == n_dev) ? g->n : (g->n * k + n_dev - 1) / n_dev;   // contiguous blocks
    for (int k = 0; k < n_dev; k++) {
        fb_config c = *cfg;
        c.device = devices[k]; c.member_offset = g->off[k]; c.n_members = g->off[k + 1] - g->off[k];
        fb_sim* s = nullptr;
        int32_t rc = fb_create(&c, &s);
        if (rc != FB_OK) { std::string msg = g_err; fb_group_destroy(g); g_err = msg; return rc; }
        g->sims.push_back(s);
    }
    if (n_dev > 1) {
        std::vector<ncclComm_t> comms(n_dev);
        std::vector<int> devs(devices, devices + n_dev);
        ncclResult_t r = N->CommInitAll(comms.data(), n_dev, devs.data());
        if (r != ncclSuccess) { fb_group_destroy(g); return fail(FB_ERR_CUDA, std::string("ncclCommInitAll: ") + N->GetErrorString(r)); }
        for (int k = 0; k < n_dev; k++) {
            g->sims[k]->comm = comms[k]; g->sims[k]->comm_owned = true; g->sims[k]->comm_ranks = n_dev;
            cudaSetDevice(devices[k]);
            int32_t rc = comm_prepare(g->sims[k]);
            if (rc != FB_OK) { std::string msg = g_err; fb_group_destroy(g); g_err = msg; return rc; }
        }
    }
    *out = g;
    return FB_OK;
}

int32_t fb_group_size(fb_group_handle g, int32_t* n_dev) {
    if (!g || !n_dev) return fail(FB_ERR_ARG, "fb_group_size: null argument");
    *n_dev = (int32_t)g->sims.size();
    return FB_OK;
}
int32_t fb_group_shard(fb_group_handle g, int32_t k, fb_handle* h, int64_t* offset, int64_t* count) {
    if (!g) return fail(FB_ERR_ARG, "null group");
    if (k < 0 || k >= (int)g->sims.size()) return fail(FB_ERR_ARG, "fb_group_shard: no such shard");
    if (h) *h = g->sims[k];
    if (offset) *offset = g->off[k];
    if (count) *count = g->off[k + 1] - g->off[k];
    return FB_OK;
}
int32_t fb_group_load_tables(fb_group_handle g, const void* blob, size_t nbytes) {
    if (!g) return fail(FB_ERR_ARG, "null group");
    for (fb_sim* s : g->sims) GK(fb_load_tables(s, blob, nbytes));
    return FB_OK;
}
int32_t fb_group_init_trim(fb_group_handle g, const fb_trim_params_soa* p, double* trim_state_out, double* cost_out) {
    if (!g) return fail(FB_ERR_ARG, "null group");
    fb_trim_params_soa z; memset(&z, 0, sizeof(z));
    if (!p) p = &z;
    for (size_t k = 0; k < g->sims.size(); k++) {
        const long long o = g->off[k];
        auto sh = [&](const double* a) { return a ? a + o : nullptr; };   // every field is its own [N] array: a shard is a pointer offset
        fb_trim_params_soa q;
        q.lat = sh(p->lat); q.lon = sh(p->lon); q.h_e = sh(p->h_e); q.psi_nb = sh(p->psi_nb); q.EAS = sh(p->EAS); q.gamma_wb_n = sh(p->gamma_wb_n);
        q.psi_wb_dot = sh(p->psi_wb_dot); q.theta_wb_dot = sh(p->theta_wb_dot); q.beta_a = sh(p->beta_a); q.fuel_load = sh(p->fuel_load);
        q.mixture = sh(p->mixture); q.flaps = sh(p->flaps);
        for (int j = 0; j < 5; j++) q.payload[j] = sh(p->payload[j]);
        fb_sim* s = g->sims[k];
        s->host_pitch = g->n;
        int32_t rc = fb_init_trim(s, &q, trim_state_out ? trim_state_out + o : nullptr, cost_out ? cost_out + o : nullptr);
        s->host_pitch = 0;
        GK(rc);
    }
    return FB_OK;
}
int32_t fb_group_set_input(fb_group_handle g, int32_t field, const double* values, int64_t stride) {
    if (!g || !values) return fail(FB_ERR_ARG, "fb_group_set_input: null argument");
    for (size_t k = 0; k < g->sims.size(); k++) GK(fb_set_input(g->sims[k], field, values + (stride ? g->off[k] : 0), stride));
    return FB_OK;
}
int32_t fb_group_step(fb_group_handle g, double t_advance) {   // launches are asynchronous: the shards advance concurrently
    if (!g) return fail(FB_ERR_ARG, "null group");
    for (fb_sim* s : g->sims) GK(fb_step(s, t_advance));
    return FB_OK;
}
int32_t fb_group_run(fb_group_handle g) {
    if (!g) return fail(FB_ERR_ARG, "null group");
    for (fb_sim* s : g->sims) GK(fb_run(s));
    return FB_OK;
}
int32_t fb_group_sync(fb_group_handle g) {
    if (!g) return fail(FB_ERR_ARG, "null group");
    for (fb_sim* s : g->sims) GK(fb_sync(s));
    return FB_OK;
}
int32_t fb_group_get_state(fb_group_handle g, double* x) {
    if (!g || !x) return fail(FB_ERR_ARG, "fb_group_get_state: null argument");
    for (size_t k = 0; k < g->sims.size(); k++) {
        fb_sim* s = g->sims[k];
        s->host_pitch = g->n;
        int32_t rc = fb_get_state(s, x + g->off[k]);
        s->host_pitch = 0;
        GK(rc);
    }
    return FB_OK;
}
int32_t fb_group_status(fb_group_handle g, int32_t* status) {
    if (!g || !status) return fail(FB_ERR_ARG, "fb_group_status: null argument");
    for (size_t k = 0; k < g->sims.size(); k++) GK(fb_status(g->sims[k], status + g->off[k]));
    return FB_OK;
}
int32_t fb_group_log_stats(fb_group_handle g, int32_t field, int64_t first, int64_t count, fb_moments* out) {
    if (!g || !out) return fail(FB_ERR_ARG, "fb_group_log_stats: null argument");
    if (g->sims.size() > 1) {
        NcclApi* N = nccl_api();
        std::vector<long long> upto(g->sims.size(), 0);
        for (size_t k = 0; k < g->sims.size(); k++) GK(stats_allreduce_pre(g->sims[k], &upto[k]));
        if (upto[0] > g->sims[0]->reduced_upto) {
            NK(N->GroupStart());   // one thread drives every communicator of the clique
            for (size_t k = 0; k < g->sims.size(); k++) { int32_t rc = stats_allreduce_enqueue(g->sims[k], upto[k]); if (rc != FB_OK) { N->GroupEnd(); return rc; } }
            NK(N->GroupEnd());
            for (size_t k = 0; k < g->sims.size(); k++) GK(stats_allreduce_post(g->sims[k], upto[k]));
        }
    }
    return fb_log_stats(g->sims[0], field, first, count, out);
}

}  // extern "C"
