/* abi_driver.c — the C ABI of include/flightbatch.h exercised from plain C (no Python, no torch): what a Julia `ccall`
 * host, or any other FFI, sees. Built by tests/test_abi_and_host.py (CPU: `layout`, `null`) and tests/test_gpu_abi_c.py
 * (GPU: `run`) with `gcc -std=c11 -Wall -Werror`, linked against libflightbatch.so, so every call below is type-checked
 * against the header's prototypes.
 *   abi_driver layout            sizeof / offsetof of the boundary structs as JSON (compared with the Julia and ctypes mirrors)
 *   abi_driver null              every entry point with NULL handles: FB_ERR_ARG (or the documented no-op), never a crash
 *   abi_driver run <tables.bin>  create -> load tables -> trim -> schedule -> step -> read back -> statistics -> linearize
 *                                -> LinearizedSS -> destroy on cuda:0, with sanity checks on the numbers                   */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "flightbatch.h"

_Static_assert(sizeof(fb_moments) == 40, "fb_moments: int64 + 4 doubles");
_Static_assert(offsetof(fb_config, n_members) == 8 && offsetof(fb_config, dt) == 24, "fb_config head");
_Static_assert(offsetof(fb_config, log_fields) == 68 && offsetof(fb_config, steps_per_launch) == 196, "fb_config log block");
_Static_assert(offsetof(fb_config, log_flags) == 200 && offsetof(fb_config, prop_pitch_n) == 204 && offsetof(fb_config, prop_pitch_lo) == 216 && sizeof(fb_config) == 248, "fb_config tail");
_Static_assert(sizeof(fb_trim_params_soa) == 17 * sizeof(void*), "fb_trim_params_soa: 12 columns + payload[5]");
_Static_assert(sizeof(fb_init_soa) == 32 * sizeof(void*), "fb_init_soa: 32 columns");

static int failures = 0;
#define EXPECT(cond, ...) do { if (!(cond)) { failures++; fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "  [%s]\n", fb_last_error()); } } while (0)

static int layout(void) {
    printf("{\"fb_config\": {\"size\": %zu, \"variant\": %zu, \"kinematics\": %zu, \"n_members\": %zu, \"member_offset\": %zu, \"dt\": %zu, \"Dt\": %zu, "
           "\"t_start\": %zu, \"t_end\": %zu, \"device\": %zu, \"log_every\": %zu, \"log_n_fields\": %zu, \"log_fields\": %zu, \"steps_per_launch\": %zu, \"log_flags\": %zu, "
           "\"prop_pitch_n\": %zu, \"reserved0\": %zu, \"prop_pitch_lo\": %zu, \"prop_pitch_hi\": %zu, \"act2_omega_n\": %zu, \"act2_zeta\": %zu}, ",
           sizeof(fb_config), offsetof(fb_config, variant), offsetof(fb_config, kinematics), offsetof(fb_config, n_members), offsetof(fb_config, member_offset),
           offsetof(fb_config, dt), offsetof(fb_config, Dt), offsetof(fb_config, t_start), offsetof(fb_config, t_end), offsetof(fb_config, device),
           offsetof(fb_config, log_every), offsetof(fb_config, log_n_fields), offsetof(fb_config, log_fields), offsetof(fb_config, steps_per_launch),
           offsetof(fb_config, log_flags), offsetof(fb_config, prop_pitch_n), offsetof(fb_config, reserved0), offsetof(fb_config, prop_pitch_lo),
           offsetof(fb_config, prop_pitch_hi), offsetof(fb_config, act2_omega_n), offsetof(fb_config, act2_zeta));
    printf("\"fb_moments\": {\"size\": %zu, \"count\": %zu, \"sum\": %zu, \"sum_sq\": %zu, \"min\": %zu, \"max\": %zu}, ",
           sizeof(fb_moments), offsetof(fb_moments, count), offsetof(fb_moments, sum), offsetof(fb_moments, sum_sq), offsetof(fb_moments, min), offsetof(fb_moments, max));
    printf("\"fb_trim_params_soa\": {\"size\": %zu, \"lat\": %zu, \"flaps\": %zu, \"payload\": %zu}, ",
           sizeof(fb_trim_params_soa), offsetof(fb_trim_params_soa, lat), offsetof(fb_trim_params_soa, flaps), offsetof(fb_trim_params_soa, payload));
    printf("\"fb_init_soa\": {\"size\": %zu, \"q_nb\": %zu, \"lat\": %zu, \"n_eng\": %zu, \"payload\": %zu, \"beta_filt\": %zu}, ",
           sizeof(fb_init_soa), offsetof(fb_init_soa, q_nb), offsetof(fb_init_soa, lat), offsetof(fb_init_soa, n_eng), offsetof(fb_init_soa, payload), offsetof(fb_init_soa, beta_filt));
    printf("\"api_version\": %d, \"FB_U__COUNT\": %d, \"FB_Y__COUNT\": %d}\n", FB_API_VERSION, (int)FB_U__COUNT, (int)FB_Y__COUNT);
    return 0;
}

static int null_calls(void) {
    double d = 0.0, buf[64]; int32_t i32 = 0, dims[3]; int64_t i64 = 0; uint8_t u8[4]; void* p = NULL; fb_moments m; fb_handle h = NULL; fb_lss_handle l = NULL;
    fb_group_handle g = NULL; fb_config cfg; memset(&cfg, 0, sizeof(cfg)); memset(buf, 0, sizeof(buf));
    EXPECT(fb_api_version() == FB_API_VERSION, "api version");
    EXPECT(fb_last_error() != NULL, "fb_last_error");
    EXPECT(fb_create(NULL, &h) == FB_ERR_ARG, "fb_create(NULL cfg)");
    EXPECT(fb_create(&cfg, NULL) == FB_ERR_ARG, "fb_create(NULL out)");
    cfg.variant = 7;
    EXPECT(fb_create(&cfg, &h) == FB_ERR_ARG && h == NULL && strstr(fb_last_error(), "variant"), "fb_create(bad variant)");
    cfg.variant = FB_C172S; cfg.n_members = 4; cfg.dt = 0.01; cfg.log_every = 1; cfg.log_n_fields = 1; cfg.log_fields[0] = 9999;
    EXPECT(fb_create(&cfg, &h) == FB_ERR_ARG && strstr(fb_last_error(), "log_fields"), "fb_create(bad log field)");
    EXPECT(fb_load_tables(NULL, buf, 8) == FB_ERR_ARG, "fb_load_tables");
    EXPECT(fb_destroy(NULL) == FB_OK, "fb_destroy(NULL) is a no-op");
    EXPECT(fb_sync(NULL) == FB_ERR_ARG, "fb_sync");
    EXPECT(fb_init_trim(NULL, NULL, NULL, NULL) == FB_ERR_ARG, "fb_init_trim");
    EXPECT(fb_init(NULL, NULL) == FB_ERR_ARG, "fb_init");
    EXPECT(fb_init_state(NULL, buf, NULL, NULL) == FB_ERR_ARG, "fb_init_state");
    EXPECT(fb_set_input(NULL, 0, buf, 0) == FB_ERR_ARG, "fb_set_input");
    EXPECT(fb_get_input(NULL, 0, buf) == FB_ERR_ARG, "fb_get_input");
    EXPECT(fb_schedule_inputs(NULL, 0, NULL, NULL, NULL, NULL) == FB_ERR_ARG, "fb_schedule_inputs");
    EXPECT(fb_step(NULL, 1.0) == FB_ERR_ARG, "fb_step");
    EXPECT(fb_run(NULL) == FB_ERR_ARG, "fb_run");
    EXPECT(fb_time(NULL, &d, &i64) == FB_ERR_ARG, "fb_time");
    EXPECT(fb_nx(NULL, &i32) == FB_ERR_ARG, "fb_nx");
    EXPECT(fb_get_state(NULL, buf) == FB_ERR_ARG, "fb_get_state");
    EXPECT(fb_get_state_async(NULL, buf) == FB_ERR_ARG, "fb_get_state_async");
    EXPECT(fb_wait_state(NULL) == FB_ERR_ARG, "fb_wait_state");
    EXPECT(fb_get_discrete(NULL, u8, u8) == FB_ERR_ARG, "fb_get_discrete");
    EXPECT(fb_get_output(NULL, 0, buf) == FB_ERR_ARG, "fb_get_output");
    EXPECT(fb_status(NULL, &i32) == FB_ERR_ARG, "fb_status");
    EXPECT(fb_release_trim_failed(NULL, NULL) == FB_ERR_ARG, "fb_release_trim_failed");
    EXPECT(fb_log_info(NULL, &i64, &i32) == FB_ERR_ARG, "fb_log_info");
    EXPECT(fb_log_times(NULL, buf) == FB_ERR_ARG, "fb_log_times");
    EXPECT(fb_log_fetch(NULL, 0, buf) == FB_ERR_ARG, "fb_log_fetch");
    EXPECT(fb_stats(NULL, 0, &m) == FB_ERR_ARG, "fb_stats");
    EXPECT(fb_log_stats_begin(NULL, 0, 1) == FB_ERR_ARG, "fb_log_stats_begin");
    EXPECT(fb_log_stats_end(NULL, 0, &m) == FB_ERR_ARG, "fb_log_stats_end");
    EXPECT(fb_log_stats(NULL, 0, 0, 1, &m) == FB_ERR_ARG, "fb_log_stats");
    EXPECT(fb_comm_unique_id(NULL) == FB_ERR_ARG, "fb_comm_unique_id");
    EXPECT(fb_comm_init(NULL, 1, 0, buf) == FB_ERR_ARG, "fb_comm_init");
    EXPECT(fb_log_stats_allreduce(NULL) == FB_ERR_ARG, "fb_log_stats_allreduce");
    EXPECT(fb_comm_time(NULL, &d, &i64) == FB_ERR_ARG, "fb_comm_time");
    EXPECT(fb_group_create(NULL, NULL, 0, &g) == FB_ERR_ARG, "fb_group_create");
    EXPECT(fb_group_destroy(NULL) == FB_OK, "fb_group_destroy(NULL) is a no-op");
    EXPECT(fb_group_size(NULL, &i32) == FB_ERR_ARG, "fb_group_size");
    EXPECT(fb_group_shard(NULL, 0, &h, &i64, &i64) == FB_ERR_ARG, "fb_group_shard");
    EXPECT(fb_group_load_tables(NULL, buf, 8) == FB_ERR_ARG, "fb_group_load_tables");
    EXPECT(fb_group_init_trim(NULL, NULL, NULL, NULL) == FB_ERR_ARG, "fb_group_init_trim");
    EXPECT(fb_group_set_input(NULL, 0, buf, 0) == FB_ERR_ARG, "fb_group_set_input");
    EXPECT(fb_group_step(NULL, 1.0) == FB_ERR_ARG, "fb_group_step");
    EXPECT(fb_group_run(NULL) == FB_ERR_ARG, "fb_group_run");
    EXPECT(fb_group_sync(NULL) == FB_ERR_ARG, "fb_group_sync");
    EXPECT(fb_group_get_state(NULL, buf) == FB_ERR_ARG, "fb_group_get_state");
    EXPECT(fb_group_status(NULL, &i32) == FB_ERR_ARG, "fb_group_status");
    EXPECT(fb_group_log_stats(NULL, 0, 0, 1, &m) == FB_ERR_ARG, "fb_group_log_stats");
    EXPECT(fb_linearize_dims(NULL, dims) == FB_ERR_ARG, "fb_linearize_dims");
    EXPECT(fb_linearize(NULL, buf, buf, buf, buf, buf, buf, buf, buf) == FB_ERR_ARG, "fb_linearize");
    EXPECT(fb_lss_create(0, 4, 16, 4, 33, 0.01, NULL) == FB_ERR_ARG, "fb_lss_create(NULL out)");
    EXPECT(fb_lss_create(0, 4, 17, 4, 33, 0.01, &l) == FB_ERR_ARG, "fb_lss_create(bad sizes)");
    EXPECT(fb_lss_destroy(NULL) == FB_OK, "fb_lss_destroy(NULL) is a no-op");
    EXPECT(fb_lss_set_model(NULL, buf, buf, buf, buf, buf, buf, buf, buf) == FB_ERR_ARG, "fb_lss_set_model");
    EXPECT(fb_lss_from_linearize(NULL, NULL) == FB_ERR_ARG, "fb_lss_from_linearize");
    EXPECT(fb_lss_set_state(NULL, buf) == FB_ERR_ARG, "fb_lss_set_state");
    EXPECT(fb_lss_get_state(NULL, buf) == FB_ERR_ARG, "fb_lss_get_state");
    EXPECT(fb_lss_set_input(NULL, buf) == FB_ERR_ARG, "fb_lss_set_input");
    EXPECT(fb_lss_get_input(NULL, buf) == FB_ERR_ARG, "fb_lss_get_input");
    EXPECT(fb_lss_get_output(NULL, buf) == FB_ERR_ARG, "fb_lss_get_output");
    EXPECT(fb_lss_step(NULL, 1.0) == FB_ERR_ARG, "fb_lss_step");
    EXPECT(fb_lss_sync(NULL) == FB_ERR_ARG, "fb_lss_sync");
    EXPECT(fb_lss_stream(NULL, &p) == FB_ERR_ARG, "fb_lss_stream");
    EXPECT(fb_rhs(NULL, buf) == FB_ERR_ARG, "fb_rhs");
    EXPECT(fb_fstep(NULL, &i32) == FB_ERR_ARG, "fb_fstep");
    EXPECT(fb_fperiodic(NULL) == FB_ERR_ARG, "fb_fperiodic");
    EXPECT(fb_counters(NULL, &i64, &i64) == FB_ERR_ARG, "fb_counters");
    EXPECT(fb_measure_fp64_peak(0, NULL) == FB_ERR_ARG, "fb_measure_fp64_peak");
    EXPECT(fb_device_state(NULL, &p, &i64) == FB_ERR_ARG, "fb_device_state");
    EXPECT(fb_stream(NULL, &p) == FB_ERR_ARG, "fb_stream");
    printf("null: %d failures\n", failures);
    return failures ? 1 : 0;
}

static void* slurp(const char* path, size_t* n) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    void* b = malloc((size_t)sz);
    if (b && fread(b, 1, (size_t)sz, f) != (size_t)sz) { free(b); b = NULL; }
    fclose(f); *n = (size_t)sz;
    return b;
}

#define OK(call) do { int32_t rc_ = (call); if (rc_ != FB_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, fb_last_error()); return 1; } } while (0)

static int run(const char* tables) {
    enum { N = 1000 };
    size_t nb = 0; void* blob = slurp(tables, &nb);
    if (!blob) { fprintf(stderr, "cannot read %s\n", tables); return 1; }
    /* ---- Simulation(SimpleWorld(Cessna172Sv0(NED())); dt = 0.01, saveat = 0.5): trim, doublet, statistics, linearize */
    fb_config cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.variant = FB_C172S; cfg.kinematics = FB_NED; cfg.n_members = N; cfg.dt = 0.01; cfg.Dt = 0.01; cfg.t_end = 10.0; cfg.device = 0;
    cfg.log_every = 50; cfg.log_n_fields = 2; cfg.log_fields[0] = FB_Y_THETA; cfg.log_fields[1] = FB_Y_H_E; cfg.steps_per_launch = 100;
    fb_handle h = NULL;
    OK(fb_create(&cfg, &h));
    OK(fb_load_tables(h, blob, nb));
    static double EAS[N], h_e[N], ts[7 * N], cost[N], x[32 * N], ele[N], theta[N];
    for (int i = 0; i < N; i++) { EAS[i] = 32.0 + 22.0 * i / N; h_e[i] = 300.0 + 2.5 * i; }
    fb_trim_params_soa p; memset(&p, 0, sizeof(p));
    p.EAS = EAS; p.h_e = h_e;
    OK(fb_init_trim(h, &p, ts, cost));
    static int32_t st[N]; int ok = 0;
    OK(fb_status(h, st));
    for (int i = 0; i < N; i++) ok += st[i] == FB_ST_RUNNING;
    EXPECT(ok == N, "all %d trims converge (%d did)", N, ok);
    EXPECT(ts[0 * N + N / 2] > 0.0 && ts[0 * N + N / 2] < 0.3, "alpha of the trim state plausible: %g", ts[N / 2]);
    OK(fb_get_input(h, FB_U_ELEVATOR, ele));
    int64_t member[2 * N], step[2 * N]; int32_t field[2 * N]; double value[2 * N];
    for (int i = 0; i < N; i++) {
        member[i] = i; step[i] = 100; field[i] = FB_U_ELEVATOR; value[i] = ele[i] + 0.1;        /* elevator step at t = 1 s (README.md:166) */
        member[N + i] = i; step[N + i] = 300; field[N + i] = FB_U_ELEVATOR; value[N + i] = ele[i];
    }
    OK(fb_schedule_inputs(h, 2 * N, member, step, field, value));
    OK(fb_step(h, 5.0));
    OK(fb_sync(h));
    double t = 0.0; int64_t ns = 0; int32_t nx = 0, nf = 0;
    OK(fb_time(h, &t, &ns)); OK(fb_nx(h, &nx));
    EXPECT(fabs(t - 5.0) < 1e-12 && ns == 500 && nx == 24, "t = %g, steps = %lld, nx = %d", t, (long long)ns, nx);
    OK(fb_get_state(h, x));
    for (int i = 0; i < nx * N; i++) if (!isfinite(x[i])) { EXPECT(0, "state not finite at %d", i); break; }
    OK(fb_get_output(h, FB_Y_THETA, theta));
    fb_moments m, ms[11];
    OK(fb_stats(h, FB_Y_THETA, &m));
    double s = 0.0, mn = 1e300, mx = -1e300;
    for (int i = 0; i < N; i++) { s += theta[i]; mn = theta[i] < mn ? theta[i] : mn; mx = theta[i] > mx ? theta[i] : mx; }
    EXPECT(m.count == N && fabs(m.sum - s) < 1e-9 && m.min == mn && m.max == mx, "fb_stats vs host: %lld %g %g", (long long)m.count, m.sum, s);
    OK(fb_log_info(h, &ns, &nf));
    EXPECT(ns == 11 && nf == 2, "log: %lld samples of %d fields", (long long)ns, nf);
    OK(fb_log_stats(h, FB_Y_THETA, 0, 11, ms));
    EXPECT(ms[10].count == N && fabs(ms[10].sum - s) < 1e-9 && ms[10].min == mn && ms[10].max == mx, "moments of the last log sample = moments of the output now");
    EXPECT(ms[3].max > ms[0].max + 0.01, "the elevator step shows in the pitch-angle moments: %g -> %g", ms[0].max, ms[3].max);
    /* ---- linearize(aircraft, trim) + LinearizedSS (README example 3) on a fresh trim */
    OK(fb_init_trim(h, &p, NULL, NULL));
    int32_t dims[3];
    OK(fb_linearize_dims(h, dims));
    EXPECT(dims[0] == 16 && dims[1] == 4 && dims[2] == 33, "linearize dims %d %d %d", dims[0], dims[1], dims[2]);
    double* A = malloc(sizeof(double) * 16 * 16 * N); double* x0 = malloc(sizeof(double) * 16 * N); double* xd0 = malloc(sizeof(double) * 16 * N);
    OK(fb_linearize(h, A, NULL, NULL, NULL, x0, NULL, NULL, xd0));
    double worst = 0.0;
    for (int i = 0; i < 16 * N; i++) worst = fabs(xd0[i]) > worst ? fabs(xd0[i]) : worst;
    EXPECT(worst < 1e-3, "xdot0 of a trimmed ensemble is ~0 but for the fuel burn (~7e-5 /s): %g", worst);
    EXPECT(isfinite(A[0]) && A[(1 * 16 + 1) * N] != 0.0, "Jacobian entries written");
    fb_lss_handle l = NULL;
    OK(fb_lss_create(0, N, 16, 4, 33, 0.01, &l));
    OK(fb_lss_from_linearize(l, h));
    OK(fb_lss_step(l, 1.0));
    OK(fb_lss_sync(l));
    static double xl[16 * N];
    OK(fb_lss_get_state(l, xl));
    double dev = 0.0;
    for (int i = 0; i < 16 * N; i++) dev = fabs(xl[i] - x0[i]) > dev ? fabs(xl[i] - x0[i]) : dev;
    EXPECT(dev < 1e-4, "a linear model at its operating point stays there: %g", dev);
    OK(fb_lss_destroy(l));
    int64_t launches = 0, rhs = 0;
    OK(fb_counters(h, &launches, &rhs));
    EXPECT(launches >= 5 && rhs > 0, "kernel launches counted: %lld", (long long)launches);
    OK(fb_destroy(h));
    free(A); free(x0); free(xd0); free(blob);
    printf("run: %d failures\n", failures);
    return failures ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "layout")) return layout();
    if (argc >= 2 && !strcmp(argv[1], "null")) return null_calls();
    if (argc >= 3 && !strcmp(argv[1], "run")) return run(argv[2]);
    fprintf(stderr, "usage: abi_driver layout | null | run <tables.bin>\n");
    return 2;
}
