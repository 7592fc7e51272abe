// fb_host.cuh — the per-handle host record (fb_sim) and the kernel dispatch (Ops<VAR, KIN>) shared by the translation units
// that hold kernels: flightbatch.cu (C ABI; Cessna172Sv0 / Xv0 / Xv1 / Xv2) and flightbatch_ext.cu (FB_C172XA2, the variant
// assembled from the non-default components — reached through the FbOpsTable below, so that its kernels are compiled
// beside the tuned ones, not among them).
#pragma once
#include <vector>
#include <utility>
#include <nccl.h>
#include "fb_init.cuh"
#include "fb_complete.h"

struct fb_sim {
    fb_config cfg;
    int var_t = 0;   // template variant actually instantiated (X2 shares X1's kernels)
    int nx = 0;
    long long pitch = 0;
    long long host_pitch = 0;        // row pitch (elements) of the caller's [C][N] host arrays; 0 = n_members (an fb_group passes global arrays)
    EnsembleDev E{};
    cudaStream_t stream = nullptr;
    double* d_snap = nullptr;                          // device snapshot of x for fb_get_state_async
    cudaEvent_t snap_ready = nullptr, snap_done = nullptr;
    cudaStream_t snap_stream = nullptr;                // its own stream: the input copy stream must not queue behind a read-back
    bool snap_pending = false;
    cudaStream_t copy_stream = nullptr;                // host-to-device staging of per-member inputs (fb_set_input)
    double* d_stage[FB_U__COUNT] = {};                 // lazily allocated staging rows
    cudaEvent_t stage_consumed[FB_U__COUNT] = {};
    cudaEvent_t copy_done = nullptr;
    double* d_pack = nullptr; float* d_geoid = nullptr;   // d_geoid: FB_GEOID_BYTES (the table + FbTablesTail)
    double* d_prop3 = nullptr;                             // VariablePitch propeller table (FB_C172XA2)
    bool tables_loaded = false, initialised = false;
    long long n_steps = 0;
    long long* d_ev_off = nullptr; long long* d_ev_step = nullptr; int* d_ev_field = nullptr; double* d_ev_value = nullptr;
    long long* d_ev_cursor = nullptr;
    int* d_log_fields = nullptr;
    double* d_scratch = nullptr;      // [max(nx, 8)][pitch]
    double* d_upload = nullptr;       // [Uploader::ROWS][pitch]: staging rows of the initialisation calls
    unsigned long long* d_rhs_count = nullptr;
    long long launches = 0;
    int periodic_every = 1;
    int step_block = 128;            // threads per CTA of the complete integrator kernel (one CTA per SM)
    int step_block_air = 128;        // ... of the AIR (clear-of-terrain) instantiation
    // hand-over between the two instantiations (fb_kernels.cuh): per-member resume record, device counter of members that
    // were (AIR launch) or would have been (complete launch) handed over, and its pinned host copy read one launch late
    long long* d_own = nullptr;      // [pitch] ownership words (FB_OWN_*, fb_kernels.cuh)
    unsigned* d_side = nullptr;      // members handed over since the last join
    bool own_dirty = false;          // an airborne launch has run since the last join (ownership words may be set)
    bool pipeline = true;            // hand-over passes run on side_stream, one launch behind the airborne launches
    cudaStream_t side_stream = nullptr;
    cudaEvent_t air_done = nullptr, side_done = nullptr;
    bool side_active = false;        // a hand-over pass may still be running on side_stream
    int launches_since_join = 0;
    bool side_seen = false;          // members are believed to be near the terrain (shape of the hand-over pass): the most recent count the
                                     // host has seen, from the launches or from the look at the initial state, was above zero
    unsigned* d_hint = nullptr; unsigned* h_hint = nullptr; cudaEvent_t hint_ev = nullptr; bool hint_pending = false;
    int n_sms = 148;
    unsigned* d_queue_c = nullptr;   // [0]: next member-warp of a queue-shaped hand-over pass
    long long join_step = 0;         // step index of the last join (every ownership word counts from at least here)
    unsigned* d_ground = nullptr;
    unsigned* h_ground = nullptr;
    cudaEvent_t ground_ev = nullptr;
    bool ground_ev_pending = false;
    bool air_mode = true;
    bool force_complete = false;
    bool air_wide = false;           // airborne kernel shape: FB_STEP_BLOCK_WIDE threads (one-wave ensembles) or FB_STEP_BLOCK
    int air_mpc = 0;                 // > step_block_air: time-sliced airborne launches, this many members per CTA (fb_kernels.cuh)
    bool air_dyn = false;            // airborne launches draw (member-warp, chunk) tasks from a global queue: one persistent CTA per SM
    int air_grid = 0;                // ... this many CTAs
    unsigned* d_queue = nullptr;     // [0]: next task, [1 .. 1 + member-warps): chunks published per member-warp
    // ensemble moments of the log samples (EnsembleDev::mom_*): one device allocation, a pinned staging copy for the
    // asynchronous fetch, and the side stream + communicator of the all-reduce over the shards
    void* d_mom = nullptr;
    long long mom_records = 0;                         // [log_capacity][log_n_fields] + 1 scratch record (fb_stats)
    unsigned long long* m_cnt = nullptr; double* m_sum = nullptr; double* m_min = nullptr; double* m_max = nullptr;
    char* h_mom = nullptr; size_t h_mom_bytes = 0;
    long long mom_first = 0, mom_count = 0;            // sample range of the fetch in flight
    bool mom_pending = false;
    cudaEvent_t mom_ready = nullptr, mom_done = nullptr;
    ncclComm_t comm = nullptr;
    bool comm_owned = false;                           // false: the communicator belongs to an fb_group
    int comm_ranks = 1;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t comm_in = nullptr, comm_out = nullptr; // launch stream -> side stream, side stream -> readers
    bool comm_busy = false;
    long long reduced_upto = 0;                        // log samples [0, reduced_upto) have been all-reduced
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> comm_timers;
    double comm_ms = 0.0; long long comm_calls = 0;
};

// ------------------------------------------------------------------------------------ kernel dispatch
template <int VAR, int KIN> struct Ops {
    // One advance of n steps: the AIR instantiation for everybody followed by the complete kernel for the members it
    // handed over (it returns at once when there are none), or — while a large share of the ensemble is near the
    // terrain — the complete kernel alone. Returns the number of kernels launched through *launched.
    static cudaError_t step(fb_sim* s, long long step0, int n, int* launched) {
        cudaError_t e = cudaMemsetAsync(s->d_ground, 0, sizeof(unsigned), s->stream);   // (the count of the complete-only launches; the
        if (e != cudaSuccess) return e;                                                 //  hand-over count d_side lives until the next join)
        int T = s->step_block;
        int grid = (int)((s->E.n + T - 1) / T);
        if constexpr (VAR == FB_C172XA2) {   // complete kernel only (instantiated in this translation unit): no airborne layout, no hand-over
            step_kernel<VAR, KIN, false><<<grid, T, StepShape<VAR, KIN, false>::SMEM, s->stream>>>(
                s->E, step0, n, s->cfg.dt, s->periodic_every, s->cfg.Dt, s->d_rhs_count, nullptr, 0, s->d_ground, T, n > 0 ? n : 1);
            *launched = 1;
            return cudaGetLastError();
        } else
        if (s->air_mode) {
            int Ta = s->step_block_air;
            // time-sliced shape (fb_kernels.cuh): mpc members per CTA on Ta threads, in chunks of air_chunk steps — only
            // when the launch is long enough to be cut up; otherwise (and always for the plain shapes) one thread per member
            int mpc = s->air_mpc > Ta ? s->air_mpc : Ta, chunk = n > 0 ? n : 1;
            if (s->air_mpc > Ta && n >= 8) {
                // cut the launch into C chunks so that the MW * C tasks fill whole rounds of the HW warps; a chunk start
                // costs one extra right-hand side (no k1 carried over), i.e. a quarter of a step
                const int MW = s->air_mpc / 32, HW = Ta / 32;
                double best = 1e300;
                for (int C = 2; C <= 16 && C * 4 <= n; C++) {
                    const int ch = (n + C - 1) / C, Cr = (n + ch - 1) / ch;
                    const double cost = (double)((MW * Cr + HW - 1) / HW) * (ch + 0.25);
                    if (cost < best) { best = cost; chunk = ch; }
                }
            }
            int grid_a = (int)((s->E.n + mpc - 1) / mpc);
            if (s->air_dyn) {
                // dynamic distribution: cut the launch into C chunks so that the MWT * C tasks keep all HWT hardware warps
                // of the GPU busy to the end (cost in step units: rounds x (chunk + the k1 a chunk start re-evaluates +
                // the state round trip of a task))
                const long long MWT = (s->E.n + 31) / 32, HWT = (long long)s->air_grid * (Ta / 32);
                double best = 1e300;
                chunk = n > 0 ? n : 1;
                for (int C = 1; C <= 16 && (C == 1 || C * 4 <= n); C++) {
                    const int ch = (n + C - 1) / C, Cr = (n + ch - 1) / ch;
                    const double cost = (double)((MWT * Cr + HWT - 1) / HWT) * (ch + (Cr > 1 ? 0.25 : 0.0) + 0.3);
                    if (cost < best) { best = cost; chunk = ch; }
                }
                e = cudaMemsetAsync(s->d_queue, 0, sizeof(unsigned) * (size_t)(1 + MWT), s->stream);
                if (e != cudaSuccess) return e;
                step_kernel<VAR, KIN, true><<<s->air_grid, Ta, StepShape<VAR, KIN, true>::SMEM, s->stream>>>(
                    s->E, step0, n, s->cfg.dt, s->periodic_every, s->cfg.Dt, s->d_rhs_count, s->d_own, 0, s->d_side, Ta, chunk,
                    s->d_queue, reinterpret_cast<int*>(s->d_queue + 1));
            } else if (s->air_wide)
                step_kernel<VAR, KIN, true, FB_STEP_BLOCK_WIDE><<<grid_a, Ta, StepShape<VAR, KIN, true, FB_STEP_BLOCK_WIDE>::SMEM, s->stream>>>(
                    s->E, step0, n, s->cfg.dt, s->periodic_every, s->cfg.Dt, s->d_rhs_count, s->d_own, 0, s->d_side, mpc, chunk);
            else
                step_kernel<VAR, KIN, true><<<grid_a, Ta, StepShape<VAR, KIN, true>::SMEM, s->stream>>>(
                    s->E, step0, n, s->cfg.dt, s->periodic_every, s->cfg.Dt, s->d_rhs_count, s->d_own, 0, s->d_side, mpc, chunk);
            // Hand-over pass: the members the airborne kernel has handed over since the last join, each from the step it has
            // reached to the end of this launch. It runs on the side stream behind this airborne launch, i.e. beside the NEXT one instead of in front of it
            // (a handful of members near the terrain are a latency-bound 4-10 ms per 100 steps; half an ensemble on the runway is
            // most of a step). The ownership words keep the two kernels on disjoint members; join() brings everybody level.
            {
                cudaStream_t st = s->stream;
                if (s->pipeline) {
                    e = cudaEventRecord(s->air_done, s->stream); if (e != cudaSuccess) return e;
                    e = cudaStreamWaitEvent(s->side_stream, s->air_done, 0); if (e != cudaSuccess) return e;
                    st = s->side_stream;
                }
                e = side_pass(s, st, step0 + n); if (e != cudaSuccess) return e;
                if (s->pipeline) { e = cudaEventRecord(s->side_done, s->side_stream); if (e != cudaSuccess) return e; s->side_active = true; }
                *launched = 2;
            }
            s->own_dirty = true;
        } else {
            CompleteLaunch a{grid, T, s->stream, s->E, step0, n, s->cfg.dt, s->periodic_every, s->cfg.Dt, s->d_rhs_count, nullptr, 0, s->d_ground, nullptr};
            e = fb_complete_launch(VAR, s->cfg.kinematics, a); if (e != cudaSuccess) return e;
            *launched = 1;
        }
        return cudaGetLastError();
    }
    // the complete kernel in hand-over mode over [join_step, target): every member the airborne kernel has handed over, from
    // the step it has reached (its ownership word) to `target`; returns at once when nobody has been handed over
    static cudaError_t side_pass(fb_sim* s, cudaStream_t st, long long target) {
        if constexpr (VAR == FB_C172XA2) return cudaSuccess;
        else {
            const int T = s->step_block, grid = (int)((s->E.n + T - 1) / T);
            if (target <= s->join_step) return cudaSuccess;
            CompleteLaunch a{grid, T, st, s->E, s->join_step, (int)(target - s->join_step), s->cfg.dt, s->periodic_every, s->cfg.Dt, s->d_rhs_count, s->d_own, 1, s->d_side, nullptr};
            if (!s->side_seen && grid > s->n_sms) {
                // Nobody is known to have been handed over: the pass is probably empty, and one CTA per 256 members, each taking an
                // SM slot from the airborne launch it runs beside just to leave again, cost cfg 3 1.5 %. One persistent CTA per SM
                // whose warps draw member-warps from a counter instead (5 % slower than the static grid when there IS work —
                // its warps drift out of phase — which is why the static grid takes over once a count has been seen).
                cudaError_t e = cudaMemsetAsync(s->d_queue_c, 0, sizeof(unsigned) * 2, st);
                if (e != cudaSuccess) return e;
                a.grid = s->n_sms; a.block = step_tmax(0); a.queue = s->d_queue_c;
            }
            return fb_complete_launch(VAR, s->cfg.kinematics, a);
        }
    }
    // shape 0: complete kernel, 1: airborne, 2: airborne wide
    static int step_tmax(int shape) {
        if constexpr (VAR == FB_C172XA2) return StepShape<VAR, KIN, false>::T;
        else return shape == 0 ? fb_complete_tmax(VAR, KIN) : (shape == 1 ? StepShape<VAR, KIN, true>::T : StepShape<VAR, KIN, true, FB_STEP_BLOCK_WIDE>::T);
    }
    static cudaError_t prepare(fb_sim* s) {
        cudaError_t e;
        if constexpr (VAR == FB_C172XA2) {
            e = cudaFuncSetAttribute(step_kernel<VAR, KIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StepShape<VAR, KIN, false>::SMEM);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(rhs_kernel<VAR, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TB_BYTES);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(output_kernel<VAR, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TB_BYTES);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(fstep_kernel<VAR, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TB_BYTES);
            if (e != cudaSuccess) return e;
            return cudaFuncSetAttribute(trim_kernel<VAR, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TB_BYTES);
        } else {
        e = fb_complete_prepare(VAR, KIN);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(step_kernel<VAR, KIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StepShape<VAR, KIN, true>::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(step_kernel<VAR, KIN, true, FB_STEP_BLOCK_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)StepShape<VAR, KIN, true, FB_STEP_BLOCK_WIDE>::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(rhs_kernel<VAR, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TB_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(output_kernel<VAR, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TB_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(avionics_init_kernel<VAR, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TB_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(fstep_kernel<VAR, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TB_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(fperiodic_kernel<VAR, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TB_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(linearize_kernel<VAR, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TB_BYTES);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(trim_kernel<VAR, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TB_BYTES);
        }
    }
    static cudaError_t rhs(fb_sim* s, double* xdot) {
        int grid = (int)((s->E.n + FB_BLOCK - 1) / FB_BLOCK);
        rhs_kernel<VAR, KIN><<<grid, FB_BLOCK, FB_TB_BYTES, s->stream>>>(s->E, xdot, nullptr, nullptr, nullptr);
        return cudaGetLastError();
    }
    static cudaError_t output(fb_sim* s, int field, double* out) {
        int grid = (int)((s->E.n + FB_BLOCK - 1) / FB_BLOCK);
        output_kernel<VAR, KIN><<<grid, FB_BLOCK, FB_TB_BYTES, s->stream>>>(s->E, field, out);
        return cudaGetLastError();
    }
    static cudaError_t fstep(fb_sim* s, int* changed) {
        int grid = (int)((s->E.n + FB_BLOCK - 1) / FB_BLOCK);
        fstep_kernel<VAR, KIN><<<grid, FB_BLOCK, FB_TB_BYTES, s->stream>>>(s->E, changed);
        return cudaGetLastError();
    }
    static cudaError_t fperiodic(fb_sim* s) {
        if constexpr (VAR == FB_C172XA2) return cudaSuccess;   // NoAvionics
        else {
        int grid = (int)((s->E.n + FB_BLOCK - 1) / FB_BLOCK);
        fperiodic_kernel<VAR, KIN><<<grid, FB_BLOCK, FB_TB_BYTES, s->stream>>>(s->E, s->cfg.Dt);
        return cudaGetLastError();
        }
    }
    static cudaError_t init(fb_sim* s, const InitSoA& I) {
        int grid = (int)((s->E.n + FB_BLOCK - 1) / FB_BLOCK);
        init_kernel<VAR, KIN><<<grid, FB_BLOCK, 0, s->stream>>>(s->E, I);
        return cudaGetLastError();
    }
    static cudaError_t linearize(fb_sim* s, double* A, double* B, double* C, double* D, double* x0, double* u0, double* y0, double* xd0) {
        if constexpr (VAR == FB_C172XA2) return cudaErrorNotSupported;   // the reference's XStateSpace knows first-order actuators only (c172x.jl:332-498)
        else {
        int grid = (int)((s->E.n + FB_BLOCK - 1) / FB_BLOCK);
        linearize_kernel<VAR, KIN><<<grid, FB_BLOCK, FB_TB_BYTES, s->stream>>>(s->E, A, B, C, D, x0, u0, y0, xd0);
        return cudaGetLastError();
        }
    }
    static cudaError_t avionics_init(fb_sim* s) {
        if constexpr (VAR == FB_C172XA2) return cudaSuccess;
        else {
        int grid = (int)((s->E.n + FB_BLOCK - 1) / FB_BLOCK);
        avionics_init_kernel<VAR, KIN><<<grid, FB_BLOCK, FB_TB_BYTES, s->stream>>>(s->E, s->cfg.Dt);
        return cudaGetLastError();
        }
    }
    static cudaError_t trim(fb_sim* s, const TrimSoA& P, double* ts, double* cost) {
        int grid = (int)((s->E.n + FB_BLOCK - 1) / FB_BLOCK);
        trim_kernel<VAR, KIN><<<grid, FB_BLOCK, FB_TB_BYTES, s->stream>>>(s->E, P, ts, cost);
        return cudaGetLastError();
    }
};

// Entry points of one (variant, kinematics) instantiation as plain function pointers: how flightbatch.cu reaches the kernels
// compiled in flightbatch_ext.cu.
struct FbOpsTable {
    cudaError_t (*step)(fb_sim*, long long, int, int*);
    int (*step_tmax)(int);
    cudaError_t (*prepare)(fb_sim*);
    cudaError_t (*rhs)(fb_sim*, double*);
    cudaError_t (*output)(fb_sim*, int, double*);
    cudaError_t (*fstep)(fb_sim*, int*);
    cudaError_t (*fperiodic)(fb_sim*);
    cudaError_t (*init)(fb_sim*, const InitSoA&);
    cudaError_t (*linearize)(fb_sim*, double*, double*, double*, double*, double*, double*, double*, double*);
    cudaError_t (*avionics_init)(fb_sim*);
    cudaError_t (*trim)(fb_sim*, const TrimSoA&, double*, double*);
    cudaError_t (*side_pass)(fb_sim*, cudaStream_t, long long);
};
template <int VAR, int KIN> constexpr FbOpsTable ops_table() {
    typedef Ops<VAR, KIN> O;
    return FbOpsTable{O::step, O::step_tmax, O::prepare, O::rhs, O::output, O::fstep, O::fperiodic, O::init, O::linearize, O::avionics_init, O::trim, O::side_pass};
}
const FbOpsTable* fb_ext_ops(int kin);   // flightbatch_ext.cu: FB_C172XA2 over the three kinematic mechanisations
