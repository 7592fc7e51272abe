// fb_kernels.cuh — ensemble kernels: fused K-step RK4 integrator, single RHS (test hook), output pass.
// One thread per aircraft. The integrator keeps the step-start state x and the Runge-Kutta accumulator in
// shared memory ([component][thread], conflict-free 64-bit columns), the stage state in registers, and folds
// every derivative into both as soon as the RHS produces it; ensemble state in HBM is struct-of-arrays
// (component-major, member fastest) and is touched once per launch of K fused steps.
#pragma once
#include <climits>
#include <type_traits>
#include "fb_physics.cuh"
#include "fb_avionics.cuh"

#define FB_BLOCK 128          // block size of the light kernels (rhs / output / init / trim)
// Threads per CTA of the integrator (one CTA per SM). The register file gives 168 registers per thread to 12 warps and
// 128 to 13..16 warps (a sub-partition then holds 4 warps); measured on B200, a 100-step launch of the airborne C172S
// kernel takes 2.9 ms at 384 threads and 4.3 ms at 448, so 384 wins whenever the ensemble spans several waves of CTAs
// and 448 only when it lets the whole ensemble fit ONE wave (cfg 2: 65,536 members = 443 per SM). The host picks.
#ifndef FB_STEP_BLOCK
#define FB_STEP_BLOCK 384       // complete kernel, and the default airborne shape
#endif
#ifndef FB_STEP_BLOCK_WIDE
#define FB_STEP_BLOCK_WIDE 448  // second airborne shape
#endif
#ifndef FB_STEP_BLOCK_CTL
#define FB_STEP_BLOCK_CTL 256   // variants with control laws (C172X1/X2) and every complete kernel: 8 warps at 255 registers
#endif

struct EnsembleDev {
    double* x;            // [NX][pitch]
    double* u;            // [FB_U__COUNT][pitch]
    unsigned char* stall; // [pitch]
    unsigned char* eng;   // [pitch]
    int* status;          // [pitch]
    double* ctl;          // [FB_CTL_STATE_COUNT][pitch] avionics discrete state (C172X1/X2)
    long long pitch;
    long long n;
    // input schedule, CSR by member: events of member i are ev_*[ev_off[i] .. ev_off[i+1]) sorted by step
    const long long* ev_off;
    const long long* ev_step;
    const int* ev_field;
    const double* ev_value;
    long long* ev_cursor; // [pitch]
    int* handover_rel;    // [pitch] launch-relative hand-over record of the airborne kernels (-1: none). Functionally the ownership
                          // words (step_kernel's `own`) carry everything it holds; it is kept, written and read exactly as before
                          // them, because ptxas' register allocation of the whole airborne fly-by-wire kernel follows the shape of
                          // this code: without it cfg 3 runs at 1.33e9 aircraft-steps/s, with it at 1.35e9 (measured, eight forms). Used by the
                          // fly-by-wire airborne kernels only (LEGACY_REC in step_kernel)
    // decimated log: log[(sample * n_fields + f) * pitch + i]
    double* log;
    const int* log_fields;
    int log_n_fields;
    int log_every;
    long long log_capacity;
    // ensemble moments of the logged fields, one record per (sample, field), accumulated as the samples are taken:
    // count [cap * nf], {sum, sum of squares} [cap * nf][2], min [cap * nf], max [cap * nf] (separate arrays: each is one
    // typed NCCL all-reduce over a contiguous range of samples). Null = moments not kept.
    unsigned long long* mom_cnt;
    double* mom_sum;
    double* mom_min;
    double* mom_max;
    FbDeviceTables tables;
};

// ---- ensemble moments ---------------------------------------------------------------------------------------------
// min / max of doubles by compare-and-swap; the first read may be stale, which is harmless for a monotone update
FB_DEV void atomic_min_f64(double* p, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = *a, assumed;
    do { assumed = old; if (__longlong_as_double((long long)assumed) <= v) break; old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v)); } while (assumed != old);
}
FB_DEV void atomic_max_f64(double* p, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = *a, assumed;
    do { assumed = old; if (__longlong_as_double((long long)assumed) >= v) break; old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v)); } while (assumed != old);
}
// One value per calling thread folded into moment record `rec`: the threads of the warp that arrive here together
// (members of a warp can run one stage out of phase, so the group is whatever is converged) reduce by shuffles and the
// group's first lane issues one set of atomics.
FB_DEV void moments_add(const EnsembleDev& E, long long rec, double v) {
    // Threads that sit at this instruction together need not be in the same iteration of the caller's loop over the logged
    // fields (the lane that issued the previous record's atomics rejoins the others by program counter, one iteration
    // behind): the group is the set of converged threads that hold the SAME record.
    const unsigned m = __match_any_sync(__activemask(), (unsigned long long)rec);
    const int lane = threadIdx.x & 31;
    double s = v, s2 = v * v, mn = v, mx = v;
    if (m == 0xffffffffu) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ts = __shfl_xor_sync(m, s, o), ts2 = __shfl_xor_sync(m, s2, o), tmn = __shfl_xor_sync(m, mn, o), tmx = __shfl_xor_sync(m, mx, o);
            s += ts; s2 += ts2; mn = (tmn < mn) ? tmn : mn; mx = (tmx > mx) ? tmx : mx;
        }
    } else {   // a partial group: every member of it gathers the values of the lanes present, in lane order
        s = 0.0; s2 = 0.0;
        for (unsigned rest = m; rest; rest &= rest - 1) {
            const double t = __shfl_sync(m, v, __ffs(rest) - 1);
            s += t; s2 += t * t; mn = (t < mn) ? t : mn; mx = (t > mx) ? t : mx;
        }
    }
    if (lane == __ffs(m) - 1) {
        atomicAdd(E.mom_cnt + rec, (unsigned long long)__popc(m));
        atomicAdd(E.mom_sum + 2 * rec, s); atomicAdd(E.mom_sum + 2 * rec + 1, s2);
        atomic_min_f64(E.mom_min + rec, mn); atomic_max_f64(E.mom_max + rec, mx);
    }
}

// Stage the lookup-table pack into shared memory with one bulk asynchronous copy (TMA unit,
// cp.async.bulk + mbarrier complete_tx); every thread of the CTA then waits on the barrier phase.
FB_DEV void stage_tables(double* smem_tb, const double* __restrict__ gsrc, unsigned nbytes = FB_TB_BYTES) {
    __shared__ __align__(8) unsigned long long bar;
    unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
    unsigned dst_a = (unsigned)__cvta_generic_to_shared(smem_tb);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(nbytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst_a), "l"(gsrc), "r"(nbytes), "r"(bar_a) : "memory");
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a) : "memory");
    }
}

__host__ __device__ inline double saturate_input(int f, double v) {   // Ranged{T,Min,Max} assignment (types.jl:15-20)
    if (f <= FB_U_RUDDER_OFFSET) return clampd(v, -1.0, 1.0);
    if (f <= FB_U_MIXTURE) return clampd(v, 0.0, 1.0);
    if (f >= FB_U_M_PILOT && f <= FB_U_M_BAGGAGE) return clampd(v, 0.0, 100.0);
    if (f == FB_U_PROP_PITCH) return clampd(v, 0.0, 1.0);   // Ranged(0, 0, 1), propellers.jl:385
    if (f == FB_U_T_SL) return clampd(v, isa::T_std - 50.0, isa::T_std + 50.0);
    if (f == FB_U_P_SL) return clampd(v, isa::p_std - 10000.0, isa::p_std + 10000.0);
    if (f >= FB_U_ACT_THROTTLE && f <= FB_U_ACT_BRAKE_RIGHT) {
        int i = f - FB_U_ACT_THROTTLE;
        return clampd(v, (i == 0 || i >= 4) ? 0.0 : -1.0, 1.0);
    }
    return avionics_saturate_input(f, v);
}

// Runge-Kutta sink: folds derivative component c into the accumulators the moment the RHS produces it.
//   stage 0: acc = k            xs = x + dt/2 k
//   stage 1: acc += 2k          xs = x + dt/2 k
//   stage 2: acc += 2k          xs = x + dt k
//   stage 3: acc += k           (xs is rebuilt by the combine pass: x* = x + dt/6 acc)
// with per-thread coefficients (cb, cx) so that one code path serves all four stages. The step-start column is
// never written by a stage, so an evaluation that has to be abandoned half-way leaves the step-start state intact;
// the combine pass runs only after a complete, error-free stage-3 evaluation.
//
// Actuator states (C172X: rows ACT .. ACT+NACT-1) get neither a stage register nor an accumulator column. Each is a
// first-order lag x' = (cmd - x) / tau (c172x.jl:39-52) that nothing feeds back into, with cmd frozen over the step, so
// its Runge-Kutta stage states and step result are closed-form in the step-start value: with e = x - cmd, z = dt / tau,
//   stage 0: e   1: e (1 - z/2)   2: e (1 - z/2 + z^2/4)   3: e (1 - z + z^2/2 - z^3/4)
//   x* = cmd + e (1 - z + z^2/2 - z^3/6 + z^4/24)
// which is what the tableau produces on that equation (agreement to rounding). The stage factor is the per-thread `pa`.
template <int T, int ACT = 0, int NACT = 0>   // T: column stride in doubles (compile-time -> immediate shared-memory offsets)
struct RkSink {
    double* xs;          // stage state, registers (actuator entries unused)
    double* sx;          // shared: step-start state column of this thread  (stride T doubles)
    double* sacc;        // shared: accumulator column of this thread (actuator rows squeezed out)
    double cb, cx, pa;
    int stage;           // 0..3; 4 = stage 0 of the next step with this step's callbacks still pending
    __host__ __device__ static constexpr int acol(int c) { return c >= ACT + NACT ? c - NACT : c; }
    FB_DEV void put(int c, double k) {
        double a = 0.0;
        if (stage & 3) a = sacc[acol(c) * T];    // stage 0: acc = k, whatever the column holds
        double xv = sx[c * T];
        sacc[acol(c) * T] = fma(cb, k, a);
        xs[c] = fma(cx, k, xv);
    }
    FB_DEV double act(int, int c, const double*, double cmd) const {   // stage 0 (and 4): the step-start value itself, exactly
        const double xv = sx[c * T];
        return (stage & 3) ? fma(pa, xv - cmd, cmd) : xv;
    }
    FB_DEV void put_act(int, double, double) {}
    template <int NX, class Cmd>
    FB_DEV void combine(double h6, double act_r, const Cmd& cmd) {   // x* = x + dt/6 (k1 + 2 k2 + 2 k3 + k4); xs = x*
#pragma unroll
        for (int c = 0; c < NX; c++) {
            if (c >= ACT && c < ACT + NACT) { const double u = cmd(c - ACT); sx[c * T] = fma(act_r, sx[c * T] - u, u); }
            else { double xn = fma(h6, sacc[acol(c) * T], sx[c * T]); sx[c * T] = xn; xs[c] = xn; }
        }
    }
};
template <int T>
struct SmemState {       // f_step! view of the step-start state column
    double* sx;
    FB_DEV SmemState(double* p) : sx(p) {}
    FB_DEV double get(int c) const { return sx[c * T]; }
    FB_DEV void set(int c, double v) { sx[c * T] = v; }
};

// ------------------------------------------------------------------------------------------------
// Fused integrator: K reference steps per launch. Per step:
//   k1..k4 (inputs and discrete state frozen), x* = x + dt/6 (k1 + 2k2 + 2k3 + k4),
//   y* = outputs of f(x*), f_step!, f_periodic! (every Dt), user-callback slot (scheduled inputs), save.
// The derivative evaluated at x* doubles as the next step's k1 unless a callback changed something the
// RHS reads, in which case it is evaluated again (as the reference always does). The loop below has ONE
// call site of the RHS; each thread carries its own stage counter, so members that need the extra
// evaluation simply run one iteration out of phase with their warp neighbours.
// Dynamic shared memory: [table pack | x columns NX*T | acc columns NX*T], T = StepShape::T (a compile-time column
// stride, so every column access is base + immediate); the CTA may be launched with fewer than T threads.
//
// Two instantiations per model. AIR = true is the fast path for members clear of the terrain: compact layout without
// the six gear-regulator states (fewer columns -> more resident members per SM), no strut code. A member that comes
// within the height gate of the terrain (or starts there) is handed over: its state is written back at the last
// consistent point and `own[i]` records how many steps it has completed and whether the step's callbacks are still
// pending, i.e. the state is x* (FB_OWN_*). AIR = false is the complete kernel; with `resume_mode` it takes exactly those
// members, from that point to the end of the pass, so that every member ends up advanced by the same number of steps.
template <int VAR, int KIN, bool AIR, int CAP = FB_STEP_BLOCK>
struct StepShape {
    static constexpr int NX = Lay<VAR, KIN, AIR>::NX;
    // shared-memory tables: the lookup pack, plus the gain schedules for the variants that run control laws
    static constexpr int TBL = FB_TB_BYTES + (VAR == FB_C172X1 ? FB_GS_BYTES : 0);
    static constexpr int BUDGET = 232448 - TBL - 1024;                                         // 227 KB per CTA
    static constexpr int NCOL = 2 * NX - Lay<VAR, KIN, AIR>::NACT_CF;                             // state columns + accumulator columns (none for actuators)
    static constexpr int FIT = (int)(BUDGET / (sizeof(double) * NCOL)) / 32 * 32;
    // The variants with control laws run 8 warps per CTA at up to 255 registers per thread instead of 12 at 168 (the
    // register file is split over four sub-partitions: 2 warps each get 255 registers, 3 get 168): their right-hand side
    // plus the control laws spill ~130 doubles per step at 168 registers, every spill an L2 round trip beside 229 KB of
    // shared memory, and the 8-warp shape leaves room for the whole input cache and the output tap in shared memory
    // (measured on cfg 3: 1.11e9 -> 1.17e9 aircraft-steps/s before anything else changed; C172S loses 11 % at 256).
    // The complete (ground-capable) kernels of every variant take the same shape: at 168 registers the C172S one spilled 880
    // bytes per thread (172 at 255) and twelve warps at twelve different program counters overran the instruction cache
    // (cfg 4, 2^18 members on the ground: 5.78e8 -> 6.41e8 aircraft-steps/s). Drawing the member-warps of the hand-over pass
    // from a global queue (persistent CTAs, as the airborne fly-by-wire launches do) was measured on top and is 5 % SLOWER
    // here (6.07e8): warps that pick up work one by one drift apart in program counter, and this kernel lives on the
    // instruction-cache lines its warps share (`no_instruction` is its second-largest stall). For the same reason two 4-warp
    // CTAs per SM (a table pack each fits; it would hide the 6 % the early warps of a CTA spend at its closing barrier) are 6 %
    // slower as well (6.03e8): the two CTAs run out of phase.
    static constexpr int VCAP = ((VAR == FB_C172X1 || !AIR) && CAP > FB_STEP_BLOCK_CTL) ? FB_STEP_BLOCK_CTL : CAP;
    static constexpr int T = FIT < VCAP ? FIT : VCAP;
    static constexpr int SPARE = (int)(BUDGET / (sizeof(double) * T)) - NCOL;                  // columns left for the input cache and the tap
    // Spare columns, in this order: the anchor of the slowly varying transcendentals (SmemAnchor: five columns, taken when
    // at least FB_ANC_MIN_SLOTS input slots remain beside it), the input cache, and what is left holds the leading entries
    // of the output tap of the f(x*) evaluation (control-law feedback); the rest of the tap goes to the member's column of
    // E.ctl (written and read back once per step, against inputs that every evaluation reads).
#ifndef FB_ANCHORS
#define FB_ANCHORS 1
#endif
#ifndef FB_ANC_MIN_SLOTS
#define FB_ANC_MIN_SLOTS 9
#endif
    static constexpr int NANC = (FB_ANCHORS && KIN != FB_NED && SPARE >= ANC_COUNT + FB_ANC_MIN_SLOTS) ? ANC_COUNT : 0;
    // input slots wanted: 9 common + 6 payload moments + wind 3 + sea-level pressure; the complete layout adds the brakes
    static constexpr int NSLOT_MAX = AIR ? 19 : 21;
    static constexpr int NSLOT = (SPARE - NANC) < NSLOT_MAX ? (SPARE - NANC) : NSLOT_MAX;
    static constexpr int TAP_WANT = (VAR == FB_C172X1) ? (NANC ? TAP_LAT : TAP_COUNT) : 0;     // latitude / longitude are the anchor's
    static constexpr int NTAP = (SPARE - NANC - NSLOT) < TAP_WANT ? (SPARE - NANC - NSLOT) : TAP_WANT;
    static constexpr size_t SMEM = TBL + sizeof(double) * (NCOL + NSLOT + NTAP + NANC) * T;
};
#define FB_RESUME_PENDING (1 << 30)
// Ownership word of a member, own[i] (64 bits, written with one store): bit 0 = the member is with the complete kernel (handed
// over by an airborne launch; it stays there, launch after launch, until the host returns everybody to the airborne kernel at
// its next join), bit 1 = its current step's callbacks are still pending (the state is x*), bits 2.. = steps completed since
// init. The airborne kernel skips a member whose bit 0 is set; the hand-over pass over [S, T) takes the members whose bit 0
// is set and whose step count is below T, from that count — so it may run on a second stream while the next airborne launch
// is already under way (a member that launch hands over carries a count >= T and is left alone).
#define FB_OWN_SIDE 1ll
#define FB_OWN_PENDING 2ll
// Both out of line on purpose: they run once per hand-over, and inlined their 64-bit arithmetic changed ptxas' register
// allocation of the whole airborne kernel around them (cfg 3 -1.5 %).
// Steps (relative to S, FB_RESUME_PENDING or-ed in) from which a hand-over pass over [S, S + n) takes member i; -1: not its member.
static __device__ __noinline__ int own_resume_point(const long long* own, long long i, long long S, int n) {
    const long long r = own[i];
    if (!(r & FB_OWN_SIDE)) return -1;                     // flying: the airborne kernel has it
    const long long sd = r >> 2;
    if (sd >= S + n) return -1;                            // handed over by a LATER launch, already running on the main stream
    return (int)max(sd - S, 0ll) | ((r & FB_OWN_PENDING) ? FB_RESUME_PENDING : 0);
}
// Member i is with the complete kernel from step S + (rel & ~PENDING) on (rel as `handover` in step_kernel)
static __device__ __noinline__ void own_publish(long long* own, long long i, long long S, int rel, unsigned* count) {
    own[i] = ((S + (long long)(rel & (FB_RESUME_PENDING - 1))) << 2) | ((rel & FB_RESUME_PENDING) ? FB_OWN_PENDING : 0) | FB_OWN_SIDE;
    if (count) atomicAdd(count, 1u);
}
#define FB_MW_MAX 32            // member-warps a CTA can time-slice over its hardware warps (step_kernel)
template <int VAR, int KIN, bool AIR, int CAP = FB_STEP_BLOCK>
__global__ void __launch_bounds__((StepShape<VAR, KIN, AIR, CAP>::T), 1)
step_kernel(EnsembleDev E, long long step0_launch, int n_steps_launch, double dt, int periodic_every, double Dt,
            unsigned long long* rhs_count, long long* own, int resume_mode, unsigned* ground_count, int mpc, int chunk,
            unsigned* queue = nullptr, int* done_g = nullptr) {
    typedef Lay<VAR, KIN, AIR> L;
    typedef Lay<VAR, KIN, false> LF;
    constexpr int NX = L::NX;
    constexpr bool LEGACY_REC = AIR && VAR == FB_C172X1;   // see EnsembleDev::handover_rel (the C172S kernels are faster without it: cfg 2 2.03e9 against 1.91e9)
    extern __shared__ __align__(16) double smem[];
    double* smem_tb = smem;
    constexpr int T = StepShape<VAR, KIN, AIR, CAP>::T;
    constexpr int T_ = T;
    if (resume_mode && *ground_count == 0) return;        // nobody was handed over by the AIR launch (CTA-uniform)
    constexpr int TBL = StepShape<VAR, KIN, AIR, CAP>::TBL;
    __shared__ int chunk_done_s[FB_MW_MAX];
    if (threadIdx.x < FB_MW_MAX) chunk_done_s[threadIdx.x] = 0;
    __shared__ unsigned long long rhs_cta;
    if (threadIdx.x == 0) rhs_cta = 0ull;
    // Member index of lane 0 of each hardware warp's current task. Everything indexed by the member inside the step loop (the
    // member's input column, its tap column, the log) re-derives its address from this word where it is used: those 64-bit
    // pointers are touched once per step, and carried across the right-hand side they were spilled and re-read from local
    // memory — an L2 round trip each beside 229 KB of shared memory — on every evaluation (16 % of all long-scoreboard stalls;
    // cfg 3 +3.3 %). (Moving the loop's own counters to shared-memory words the same way was measured too: 3 % slower — the
    // columns they take are worth more as tap entries, and ptxas spills more around the volatile accesses.)
    __shared__ long long task_i0_s[32];
    volatile long long* const task_i0 = task_i0_s + (threadIdx.x >> 5);
    stage_tables(smem_tb, E.tables.pack, TBL);   // (contains the __syncthreads that orders the two initialisations above)
    GainsShared gains; gains.base = nullptr; gains.sbase = (unsigned)__cvta_generic_to_shared(smem) + FB_TB_BYTES;
    // Work distribution. A task = (member-warp, chunk of `chunk` steps): the whole per-member body below — state from HBM, the
    // chunk's steps, state back to HBM. Tasks are numbered chunk-major (all member-warps of chunk 0, then chunk 1, ...), and a
    // task may only start when the same member-warp's previous chunk has been published.
    //  * queue == nullptr (static): the CTA owns `mpc` consecutive members = MW member-warps; hardware warp w runs tasks w,
    //    w + HW, ...; the predecessor of a task always sits in an earlier round of the same CTA, so the wait on the
    //    shared-memory counter cannot deadlock. With mpc = block size and one chunk this is the plain one-thread-per-member launch.
    //  * queue != nullptr (dynamic, the airborne launches): ONE persistent CTA per SM, and every hardware warp of the GPU
    //    draws its next task from a global counter, so that neither a slow warp nor a slow SM holds the launch up (a
    //    131,072-member shard is 3.46 tasks per hardware warp: statically distributed, 6.5 % of the launch was SMs waiting for
    //    the slowest one). A predecessor has a smaller number than its successor, so it has been drawn by a resident warp
    //    before the successor can wait for it: no deadlock. State handed from one SM to another goes through L2: the writer
    //    fences before it publishes `done_g[mw]`, and this translation unit is compiled with -dlcm=cg (every global load is
    //    an L2 load), so a stale line in the other SM's L1 can never be read.
    const int lane = threadIdx.x & 31, HW = blockDim.x >> 5;
    const bool dyn = queue != nullptr;
    const long long cta_base = dyn ? 0ll : (long long)blockIdx.x * mpc;
    const int MW = (int)((min(dyn ? E.n : (long long)mpc, E.n - cta_base) + 31) / 32);
    const int C = (n_steps_launch + chunk - 1) / chunk;
    volatile int* chunk_done = dyn ? done_g : chunk_done_s;
    // right-hand-side evaluations of this thread's tasks; folded into one shared-memory counter and ONE global atomic per CTA
    // at the end (a global atomic per thread and task put 2.6e5 same-address atomics behind every time-sliced cfg-3 launch of a
    // 131,072-member shard: ~0.3 ms of an 11.5 ms launch)
    unsigned rhs_mine = 0;
    for (int task = threadIdx.x >> 5;; task += HW) {
    if (dyn) {
        int t = 0;
        if (lane == 0) t = (int)atomicAdd(queue, 1u);
        task = __shfl_sync(0xffffffffu, t, 0);
    }
    if (task >= MW * C) break;
    const int mw = task % MW, ck = task / MW;
    if (C > 1) {
        while (chunk_done[mw] < ck) __nanosleep(dyn ? 200 : 64);
        if (dyn) __threadfence(); else __threadfence_block();
        __syncwarp();
    }
    do {   // one task; `break` leaves it
    long long step0 = step0_launch + (long long)ck * chunk;
    int n_steps = min(chunk, n_steps_launch - ck * chunk);
    __syncwarp();
    if (lane == 0) *task_i0 = cta_base + (long long)mw * 32;
    __syncwarp();
    const long long i = cta_base + (long long)mw * 32 + lane;
    if (i >= E.n) break;
    int status = E.status[i];
    bool pending = false;
    if (resume_mode) {
        if (status != FB_ST_RUNNING) break;                              // frozen while with this kernel
        const int d0 = own_resume_point(own, i, step0_launch, n_steps_launch);
        if (d0 < 0) break;                                               // not handed over, or by a later launch already under way
        pending = (d0 & FB_RESUME_PENDING) != 0;
        step0 += d0 & (FB_RESUME_PENDING - 1); n_steps -= d0 & (FB_RESUME_PENDING - 1);
    } else if (status != FB_ST_RUNNING) {
        if (LEGACY_REC) E.handover_rel[i] = -1;
        break;
    } else if (LEGACY_REC && ck > 0 && E.handover_rel[i] >= 0) break;   // handed over in an earlier chunk of this launch
    else if (AIR && (own[i] & FB_OWN_SIDE)) break;               // with the complete kernel since an earlier launch
    MemberIn ug; ug.u = E.u + i; ug.pitch = E.pitch;
    double* sx = smem + TBL / 8 + threadIdx.x;
    double* sacc = sx + NX * T;
    CachedIn<VAR, T, StepShape<VAR, KIN, AIR, CAP>::NSLOT> u; u.g = ug; u.sc = sx + StepShape<VAR, KIN, AIR, CAP>::NCOL * T;
    constexpr int NTAP = StepShape<VAR, KIN, AIR, CAP>::NTAP, NANC = StepShape<VAR, KIN, AIR, CAP>::NANC;
    double* const stap = u.sc + StepShape<VAR, KIN, AIR, CAP>::NSLOT * T;   // this thread's tap columns (NTAP of them)
    double* const sanc = stap + NTAP * T;                                   // ... and its anchor columns (NANC of them)
    double* const gtap0 = E.ctl ? E.ctl + (long long)FB_CTL_TAP_ROW * E.pitch : nullptr;   // tap rows of member 0
    double xs[NX];
#pragma unroll
    for (int c = 0; c < NX; c++) { double v = E.x[(long long)L::row(c) * E.pitch + i]; xs[c] = v; sx[c * T] = v; }
    if (AIR) {   // regulator states must be zero to be left out
        double nz = 0.0;
#pragma unroll
        for (int c = 0; c < 6; c++) nz += fabs(E.x[(long long)(LF::LDG + c) * E.pitch + i]);
        if (nz != 0.0) { own[i] = (step0 << 2) | FB_OWN_SIDE; atomicAdd(ground_count, 1u); break; }
    }
    int stall = E.stall[i], eng = E.eng[i];
    long long ev_cur = E.ev_off ? E.ev_off[i] + E.ev_cursor[i] : 0;
    const long long ev_end = E.ev_off ? E.ev_off[i + 1] : 0;
    CtlState cs;
    if (VAR == FB_C172X1) cs.load(E.ctl + i, E.pitch);
    // inputs scheduled at or before the current step index are in force from the start of this launch
    while (ev_cur < ev_end && E.ev_step[ev_cur] <= step0) {
        int f = E.ev_field[ev_cur];
        ug.set(f, saturate_input(f, E.ev_value[ev_cur]));
        ev_cur++;
    }
    // launch-relative step count at which the next scheduled input fires (checked once per step with one compare)
    int next_ev = (ev_cur < ev_end) ? (int)min(E.ev_step[ev_cur] - step0, (long long)INT_MAX) : INT_MAX;
    if (VAR == FB_C172X1) avionics_assign(cs, ug);   // assign!(vehicle, avionics) precedes every f_ode! (aircraftbase.jl:221-230)
    u.refresh();

    int done = 0, n_reeval = 0;
    int gnd_seen = 0, handover = -1;
    int ctl_in = periodic_every - (int)(step0 % periodic_every);                               // steps to the next control update
    int log_in = (E.log_every > 0 && E.log_n_fields > 0) ? E.log_every - (int)(step0 % E.log_every) : 0;   // ... log sample (0: no log)
    const double h2 = 0.5 * dt, h6 = dt * (1.0 / 6.0);
    RkSink<T, L::ACT, L::NACT_CF> sink; sink.xs = xs; sink.sx = sx; sink.sacc = sacc;
    // stage factors and step factor of the closed-form actuator propagation (RkSink)
    const double az = dt * (1.0 / 0.05);
    const double ap1 = 1.0 - 0.5 * az, ap2 = 1.0 - 0.5 * az * ap1, ap3 = 1.0 - az * ap2;
    const double act_r = 1.0 - az * (1.0 / 6.0) * (1.0 + 2.0 * ap1 + 2.0 * ap2 + ap3);
    sink.stage = pending ? 4 : 0;
    SmemState<T> xstate(sink.sx);
    if (n_steps > 0) for (;;) {
        // coefficients of this thread's current stage
        const int st = sink.stage;
        sink.cb = (st == 1 || st == 2) ? 2.0 : 1.0;
        sink.cx = (st == 2) ? dt : h2;
        sink.pa = (st == 1) ? ap1 : ((st == 2) ? ap2 : ((st == 3) ? ap3 : 1.0));
        RhsOut y;
        typedef typename std::conditional<(NANC > 0), SmemAnchor<T, KIN, L::KINX>, NoAnchor>::type Anc;
        Anc anc;
        if constexpr (NANC > 0) { anc.p = sanc; anc.sx = sx; anc.ex = (st & 3) == 0; }
        if (VAR == FB_C172X1) {
            SplitTap<T, NTAP, (NANC > 0)> tap; tap.ps = stap; tap.gbase = gtap0; tap.i0 = task_i0; tap.lane = lane; tap.pitch = E.pitch; tap.on = (st == 4);
            rhs<VAR, KIN, false, AIR>(xs, sink, stall, eng, u, smem_tb, E.tables.geoid, y, nullptr, tap, anc);
        } else rhs<VAR, KIN, false, AIR>(xs, sink, stall, eng, u, smem_tb, E.tables.geoid, y, nullptr, NoTap(), anc);
        if (!AIR) gnd_seen |= y.gnd;
        if (y.err) {
            if (AIR && y.err == FB_NEED_COMPLETE) {
                // sx is the step-start state (stages 1..3 and re-evaluations) or x* with this step's callbacks pending
                handover = (ck * chunk + done) | ((sink.stage == 4) ? FB_RESUME_PENDING : 0);
            } else status = y.err;
            break;
        }
        if (sink.stage == 3) { sink.template combine<NX>(h6, act_r, [&](int a) { return u(FB_U_ACT_THROTTLE + a); }); sink.stage = 4; continue; }   // xs == x == x*: next evaluation is the FSAL-last one
        if (sink.stage != 4) { sink.stage++; continue; }
        // ---- stage 0 with a pending step end: this evaluation was the reference's f(x*) and y = y*.
        // (It also produced the tentative k1 / first stage state of the next step.)
        done++;
        const long long step_idx = step0 + done;
        const long long im = *task_i0 + lane;          // this member (see task_i0_s)
        ug.u = E.u + im; u.g = ug;
        const double* const gtap = gtap0 + im;
        // countdowns instead of 64-bit remainders: steps left to the next control update / log sample
        bool ctl_due = false, log_due = false;
        if (VAR == FB_C172X1) { if (--ctl_in == 0) { ctl_due = true; ctl_in = periodic_every; } }
        if (log_in > 0 && --log_in == 0) { log_in = E.log_every; log_due = (step_idx / E.log_every) < E.log_capacity; }
        YFull yf;
        const int stall_y = stall, eng_y = eng;
        if (ctl_due && !log_due) {
            // Control-law inputs from the tapped intermediates of the f(x*) evaluation just made and from x* itself. The
            // angles come from the fast primitives of the right-hand side (a few ulp from the library functions; the
            // control laws are continuous in them, and the parity tests hold the commands to 1e-9).
            auto tapv = [&](int k) { return k < NTAP ? stap[k * T_] : ((NANC && k >= TAP_LAT) ? sanc[(k - TAP_LAT + ANC_LAT) * T_] : gtap[(long long)k * E.pitch]); };
            q4 q = mkq(tapv(TAP_Q_NB), tapv(TAP_Q_NB + 1), tapv(TAP_Q_NB + 2), tapv(TAP_Q_NB + 3));
            if (KIN == FB_NED) { yf.theta = sink.sx[(L::KINX + 1) * T_]; yf.phi = sink.sx[(L::KINX + 2) * T_]; }
            else {   // REuler(q_nb), attitude.jl:382-390: asin(s) = atan2(s, sqrt((1 - s)(1 + s)))
                const double st_ = clampd(2.0 * (q.w * q.y - q.x * q.z), -1.0, 1.0);
                yf.theta = fb_atan2(st_, fb_sqrt(fb_max((1.0 - st_) * (1.0 + st_), 1e-280)));
                yf.phi = fb_atan2(2.0 * (q.w * q.x + q.y * q.z), 1.0 - 2.0 * (q.x * q.x + q.y * q.y));
            }
            yf.v_eb_n = mk3(tapv(TAP_V_EB_N), tapv(TAP_V_EB_N + 1), tapv(TAP_V_EB_N + 2));
            yf.w_wb_b = mk3(tapv(TAP_W_WB_B), tapv(TAP_W_WB_B + 1), tapv(TAP_W_WB_B + 2));
            yf.w_eb_b = mk3(sink.sx[L::DYN * T_], sink.sx[(L::DYN + 1) * T_], sink.sx[(L::DYN + 2) * T_]);
            yf.v_gnd = fb_sqrt(fb_max(dot(yf.v_eb_n, yf.v_eb_n), 1e-280));
            const bool valid = (KIN == FB_NED) || (yf.v_gnd > 0.1);   // kinematics.jl:212-213 vs :393-394
            yf.chi_gnd = valid ? fb_atan2(yf.v_eb_n.y, yf.v_eb_n.x) : 0.0;
            yf.EAS = tapv(TAP_TAS) * fb_sqrt(tapv(TAP_RHO) * (1.0 / isa::rho_std));
            yf.alpha = tapv(TAP_ALPHA); yf.beta = tapv(TAP_BETA);
            yf.lat = tapv(TAP_LAT); yf.lon = tapv(TAP_LON);
            yf.alpha_filt = sink.sx[L::AERO * T_]; yf.beta_filt = sink.sx[(L::AERO + 1) * T_];
            yf.eng_n = sink.sx[L::ENG * T_] * (1.0 / (2700 * FB_PI / 30));
            yf.h_e = sink.sx[(L::KINX + (KIN == FB_WA ? 8 : (KIN == FB_ECEF ? 7 : 5))) * T_];
#pragma unroll
            for (int a = 0; a < L::NACT_CF; a++) {
                yf.act_pos[a] = clampd(sink.sx[(L::ACT + a) * T_], (a == 0 || a >= 4) ? 0.0 : -1.0, 1.0);
                yf.act_cmd[a] = ug(FB_U_ACT_THROTTLE + a);
            }
        } else if (ctl_due || log_due) {   // the saved / fed-back y is the one left by f(x*), before any callback
            double xl[LF::NX], kd[LF::NX];
#pragma unroll
            for (int c = 0; c < LF::NX; c++) xl[c] = 0.0;
#pragma unroll
            for (int c = 0; c < NX; c++) xl[L::row(c)] = sink.sx[c * T];
            RhsOut y2;
            rhs_with_outputs<VAR, KIN>(xl, kd, stall, eng, ug, smem_tb, E.tables.geoid, &y2, &yf);
        }
        bool changed = f_step<VAR, KIN, AIR>(xstate, stall, eng, u, y, status);
        if (status != FB_ST_RUNNING) break;
        // the new commands go to the ensemble and straight into their input-cache slots (no refresh)
        if (ctl_due) {
            CtlIn ci; ci.g = ug; ci.load();   // one L2 round trip for all control-law inputs
            avionics_periodic<VAR, KIN>(yf, y, ci, u, cs, gains, Dt);
        }
        if (done >= next_ev) {
            while (ev_cur < ev_end && E.ev_step[ev_cur] <= step_idx) {
                int f = E.ev_field[ev_cur];
                ug.set(f, saturate_input(f, E.ev_value[ev_cur]));
                ev_cur++;
            }
            next_ev = (ev_cur < ev_end) ? (int)min(E.ev_step[ev_cur] - step0, (long long)INT_MAX) : INT_MAX;
            changed = true;
            if (VAR == FB_C172X1) avionics_assign(cs, ug);   // an event cannot override a command the control laws own
        } else if (VAR == FB_C172X1 && changed) avionics_assign(cs, ug);
        if (changed) u.refresh();
        if (log_due) {
            const long long smp = step_idx / E.log_every;
            for (int f = 0; f < E.log_n_fields; f++) {
                const double v = pick_output(yf, stall_y, eng_y, cs, E.log_fields[f]);
                if (E.log) E.log[(smp * E.log_n_fields + f) * E.pitch + im] = v;
                if (E.mom_cnt) moments_add(E, smp * E.log_n_fields + f, v);
            }
        }
        if (done == n_steps) break;
        if (changed) {            // re-evaluate k1 from the post-callback x / s / u
#pragma unroll
            for (int c = 0; c < NX; c++) xs[c] = sink.sx[c * T];
            sink.stage = 0;       // not pending
            n_reeval++;
            continue;
        }
        // New actuator commands alone need no re-evaluation: of the derivative just evaluated (the next step's k1) only
        // x'.act depends on them (c172x.jl:44-50), and the actuator states are propagated in closed form (RkSink).
        // Variants with control laws: the first stage state of the next step is rebuilt from the two shared-memory columns
        // (this evaluation ran as stage 0: acc = k1, xs = x + dt/2 k1) rather than carried in registers across the
        // callbacks above. (Without control laws carrying it is cheaper: measured on cfg 4.)
        if (VAR == FB_C172X1) {
#pragma unroll
            for (int c = 0; c < NX; c++)
                if (c < L::ACT || c >= L::ACT + L::NACT_CF) xs[c] = fma(h2, sink.sacc[sink.acol(c) * T], sink.sx[c * T]);
        }
        sink.stage = 1;
    }
    {
    const long long ie = *task_i0 + lane;              // (== i; re-derived so that i does not live across the step loop)
#pragma unroll
    for (int c = 0; c < NX; c++) E.x[(long long)L::row(c) * E.pitch + ie] = sink.sx[c * T];
    E.stall[ie] = (unsigned char)stall; E.eng[ie] = (unsigned char)eng; E.status[ie] = status;
    if (E.ev_off) E.ev_cursor[ie] = ev_cur - E.ev_off[ie];
    if (VAR == FB_C172X1) cs.store(E.ctl + ie, E.pitch);
    // (the state above is written before the ownership word: whoever reads the word as handed over finds the state it describes)
    if (AIR) {
        if (LEGACY_REC) E.handover_rel[ie] = handover;
        if (handover >= 0) own_publish(own, ie, step0_launch, handover, ground_count);
    } else if (resume_mode) { if (status == FB_ST_RUNNING) own_publish(own, ie, step0_launch, n_steps_launch, nullptr); }
    else if (gnd_seen) atomicAdd(ground_count, 1u);
    }
    // RHS evaluations of this launch: k1 of the first step, then k2..k4 and f(x*) per step, plus re-evaluations
    rhs_mine += max(0, 4 * done + n_reeval + (pending ? -3 : 1) + (handover >= 0 ? (sink.stage & 3) : 0));
    } while (0);
    if (C > 1) {   // publish the chunk: the next task on these members may run on another warp (of this CTA, or of any SM)
        if (dyn) __threadfence(); else __threadfence_block();
        __syncwarp();
        if (lane == 0) chunk_done[mw] = ck + 1;
    }
    }
    if (rhs_count) {   // every thread of the CTA arrives here (tasks are left by `break`, never by `return`)
        if (rhs_mine) atomicAdd(&rhs_cta, (unsigned long long)rhs_mine);
        __syncthreads();
        if (threadIdx.x == 0 && rhs_cta) atomicAdd(rhs_count, rhs_cta);
    }
}

// f_ode!(world) at the current state: test hook + reinit! evaluation
template <int VAR, int KIN>
__global__ void __launch_bounds__(FB_BLOCK, 2)
rhs_kernel(EnsembleDev E, double* xdot, double* alpha_out, int* wow_out, int* err_out) {
    typedef Lay<VAR, KIN> L;
    constexpr int NX = L::NX;
    extern __shared__ __align__(16) double smem_tb[];
    stage_tables(smem_tb, E.tables.pack);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E.n) return;
    MemberIn u; u.u = E.u + i; u.pitch = E.pitch;
    double x[NX], k[NX];
#pragma unroll
    for (int c = 0; c < NX; c++) x[c] = E.x[(long long)c * E.pitch + i];
    if (VAR == FB_C172X1) { CtlState cs; cs.load(E.ctl + i, E.pitch); avionics_assign(cs, u); }
    RhsOut y;
    ArraySink sink; sink.xd = k;
    rhs<VAR, KIN, false>(x, sink, E.stall[i], E.eng[i], u, smem_tb, E.tables.geoid, y, nullptr);
#pragma unroll
    for (int c = 0; c < NX; c++) xdot[(long long)c * E.pitch + i] = k[c];
    if (alpha_out) alpha_out[i] = y.alpha;
    if (wow_out) wow_out[i] = y.wow;
    if (err_out) err_out[i] = y.err;
}

// output pass: one fb_output field of every member from the current state (the y the reference would hold
// after an f_ode! at this state)
template <int VAR, int KIN>
__global__ void __launch_bounds__(FB_BLOCK, 2)
output_kernel(EnsembleDev E, int field, double* out) {
    typedef Lay<VAR, KIN> L;
    constexpr int NX = L::NX;
    extern __shared__ __align__(16) double smem_tb[];
    stage_tables(smem_tb, E.tables.pack);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E.n) return;
    MemberIn u; u.u = E.u + i; u.pitch = E.pitch;
    double x[NX];
#pragma unroll
    for (int c = 0; c < NX; c++) x[c] = E.x[(long long)c * E.pitch + i];
    CtlState cs;
    if (VAR == FB_C172X1) { cs.load(E.ctl + i, E.pitch); avionics_assign(cs, u); }
    YFull yf; RhsOut y; double k[NX];
    ArraySink sink; sink.xd = k;
    rhs<VAR, KIN, true>(x, sink, E.stall[i], E.eng[i], u, smem_tb, E.tables.geoid, y, &yf);
    out[i] = pick_output(yf, E.stall[i], E.eng[i], cs, field);
}

// f_step!(world) at the current state (test hook): y of an f_ode! at x, then the discrete updates of world.jl:34-39
// (quaternion renormalisation, stall hysteresis, crash checks, contact reset, engine state machine)
template <int VAR, int KIN>
__global__ void __launch_bounds__(FB_BLOCK, 2)
fstep_kernel(EnsembleDev E, int* changed_out) {
    typedef Lay<VAR, KIN> L;
    constexpr int NX = L::NX;
    extern __shared__ __align__(16) double smem_tb[];
    stage_tables(smem_tb, E.tables.pack);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E.n) return;
    int status = E.status[i];
    if (status != FB_ST_RUNNING) { if (changed_out) changed_out[i] = 0; return; }
    MemberIn u; u.u = E.u + i; u.pitch = E.pitch;
    double x[NX], k[NX];
#pragma unroll
    for (int c = 0; c < NX; c++) x[c] = E.x[(long long)c * E.pitch + i];
    if (VAR == FB_C172X1) { CtlState cs; cs.load(E.ctl + i, E.pitch); avionics_assign(cs, u); }
    int stall = E.stall[i], eng = E.eng[i];
    RhsOut y;
    ArraySink sink; sink.xd = k;
    rhs<VAR, KIN, false>(x, sink, stall, eng, u, smem_tb, E.tables.geoid, y, nullptr);
    bool changed = false;
    if (y.err) status = y.err;
    else { ArrayState xs; xs.x = x; changed = f_step<VAR, KIN>(xs, stall, eng, u, y, status); }
#pragma unroll
    for (int c = 0; c < NX; c++) E.x[(long long)c * E.pitch + i] = x[c];
    E.stall[i] = (unsigned char)stall; E.eng[i] = (unsigned char)eng; E.status[i] = status;
    if (changed_out) changed_out[i] = changed ? 1 : 0;
}

// f_periodic!(avionics, vehicle) followed by assign! at the current state (test hook): guidance, ControlLawsLon,
// ControlLawsLat from the outputs of an f_ode! at x (c172x2.jl:28-37, c172x_ctl.jl:286-446,880-983)
template <int VAR, int KIN>
__global__ void __launch_bounds__(FB_BLOCK, 2)
fperiodic_kernel(EnsembleDev E, double Dt) {
    typedef Lay<VAR, KIN> L;
    constexpr int NX = L::NX;
    extern __shared__ __align__(16) double smem_tb[];
    stage_tables(smem_tb, E.tables.pack);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E.n) return;
    if (VAR < FB_C172X1) return;
    if (E.status[i] != FB_ST_RUNNING) return;
    MemberIn u; u.u = E.u + i; u.pitch = E.pitch;
    double x[NX], k[NX];
#pragma unroll
    for (int c = 0; c < NX; c++) x[c] = E.x[(long long)c * E.pitch + i];
    CtlState cs;
    cs.load(E.ctl + i, E.pitch);
    avionics_assign(cs, u);
    YFull yf; RhsOut y;
    rhs_with_outputs<VAR, KIN>(x, k, E.stall[i], E.eng[i], u, smem_tb, E.tables.geoid, &y, &yf);
    CtlIn ci; ci.g = u; ci.load();
    avionics_periodic<VAR, KIN>(yf, y, ci, u, cs, gains_global(E.tables), Dt);
    cs.store(E.ctl + i, E.pitch);
}

// f_init!(avionics, vehicle) for every member, from the vehicle outputs at the initial state
// (c172x_ctl.jl:1246-1250); leaves the controller memory in E.ctl and the commands assigned.
template <int VAR, int KIN>
__global__ void __launch_bounds__(FB_BLOCK, 2)
avionics_init_kernel(EnsembleDev E, double Dt) {
    typedef Lay<VAR, KIN> L;
    constexpr int NX = L::NX;
    extern __shared__ __align__(16) double smem_tb[];
    stage_tables(smem_tb, E.tables.pack);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E.n) return;
    if (VAR < FB_C172X1) return;
    MemberIn u; u.u = E.u + i; u.pitch = E.pitch;
    double x[NX], k[NX];
#pragma unroll
    for (int c = 0; c < NX; c++) x[c] = E.x[(long long)c * E.pitch + i];
    YFull yf; RhsOut y;
    rhs_with_outputs<VAR, KIN>(x, k, E.stall[i], E.eng[i], u, smem_tb, E.tables.geoid, &y, &yf);
    CtlState cs;
    cs.load(E.ctl + i, E.pitch);
    cs.lon_mode = LON_DIRECT; cs.lat_mode = LAT_DIRECT; cs.h_state = 1;
    cs.thr_cmd = 0.0; cs.ele_cmd = 0.0; cs.ail_cmd = 0.0; cs.rud_cmd = 0.0;
    avionics_init(yf, y.wow, u, cs, E.tables, Dt);
    avionics_assign(cs, u);
    cs.store(E.ctl + i, E.pitch);
}
