// fb_tables.h — layout of the lookup-table pack staged into shared memory by every physics kernel
// (one bulk async copy per CTA), plus the device-global tables that stay in L2 (geoid, gain schedules).
// Table semantics: SURVEY.md App. C; data come from the blob written by tools/build_tables.py.
#pragma once
#include <stdint.h>

// offsets in doubles inside the shared pack. Every explicit-knot axis carries a companion array of reciprocal
// knot spacings (I_*), filled by the host when it builds the pack, so that cell weights are multiplications.
enum {
    // ---- aero (lib/FlightApps/src/c172/c172.jl:51-159), gridded, Flat
    TB_GE_X = 0,                    // 13
    TB_I_GE = TB_GE_X + 13,         // 12
    TB_CD_GE = TB_I_GE + 12,        // 13
    TB_CL_GE = TB_CD_GE + 13,       // 13
    TB_DF_X = TB_CL_GE + 13,        // 4
    TB_I_DF = TB_DF_X + 4,          // 3
    TB_CD_DF = TB_I_DF + 3,         // 4
    TB_CL_DF = TB_CD_DF + 4,        // 4
    TB_CM_DF = TB_CL_DF + 4,        // 4
    TB_CD_A_X = TB_CM_DF + 4,       // 26
    TB_I_CD_A = TB_CD_A_X + 26,     // 25
    TB_CD_A_DF = TB_I_CD_A + 25,    // 26*4  [alpha][flap]
    TB_CL_A_X = TB_CD_A_DF + 104,   // 17
    TB_I_CL_A = TB_CL_A_X + 17,     // 16
    TB_CL_A = TB_I_CL_A + 16,       // 17*2  [alpha][stall]
    TB_CY_B_DF = TB_CL_A + 34,      // 3*2
    TB_CY_P = TB_CY_B_DF + 6,       // 2*2
    TB_CY_R = TB_CY_P + 4,          // 2*2
    TB_CL_R = TB_CY_R + 4,          // 2*2  (C_l.r)
    // ---- engine (lib/FlightPhysics/src/piston.jl:70-195)
    TB_D_WOT = TB_CL_R + 4,         // 2*9  scaled, Line
    TB_MU_WOT = TB_D_WOT + 18,      // 2*9  scaled, Line
    TB_N_STD = TB_MU_WOT + 18,      // 13
    TB_I_N_STD = TB_N_STD + 13,     // 12
    TB_PI_STD = TB_I_N_STD + 12,    // 13*3 gridded, Flat (mu knots 0, 0.568, 1)
    TB_N_WOT = TB_PI_STD + 39,      // 5
    TB_I_N_WOT = TB_N_WOT + 5,      // 4
    TB_PI_WOT = TB_I_N_WOT + 4,     // 5*3  gridded (delta knots 0, 0.441, 1), Flat / Line above in delta
    TB_F_X = TB_PI_WOT + 15,        // 11
    TB_I_F = TB_F_X + 11,           // 10
    TB_PI_RATIO = TB_I_F + 10,      // 11
    TB_SFC_RATIO = TB_PI_RATIO + 11,  // 11
    TB_N_SFC = TB_SFC_RATIO + 11,   // 5
    TB_I_N_SFC = TB_N_SFC + 5,      // 4
    TB_PI_SFC = TB_I_N_SFC + 4,     // 8
    TB_I_PI_SFC = TB_PI_SFC + 8,    // 7
    TB_SFC_POW = TB_I_PI_SFC + 7,   // 5*8  gridded, Line
    // ---- propeller (lib/FlightPhysics/src/propellers.jl:235-291): 21x21 over J, Mt in [0,1.5], scaled, Flat;
    //      interleaved [iJ][iMt][4] = C_Fx, C_Mx, C_Fz_a, C_Mz_a (C_P and eta_p are not on the path)
    TB_PROP = TB_SFC_POW + 40,
    // ---- cell finders of the nine explicit-knot axes (built by the host, fb_load_tables): the knot span of axis a is cut
    //      into FB_BK uniform buckets, none holding more than one interior knot; record b = {the knot inside bucket b
    //      (or +huge), cell index at the bucket's left edge (int32 in the low word of the second double)}, 16-byte aligned
    TB_BK_IH = TB_PROP + 21 * 21 * 4,                 // FB_AX_COUNT reciprocal bucket widths (+1 pad)
    TB_BK = (TB_BK_IH + 10 + 1) / 2 * 2,              // [FB_AX_COUNT][FB_BK][2]
    TB_COUNT = TB_BK + 9 * 32 * 2
};
#define FB_BK 32
enum { FB_AX_GE = 0, FB_AX_DF, FB_AX_CD_A, FB_AX_CL_A, FB_AX_N_STD, FB_AX_N_WOT, FB_AX_F, FB_AX_N_SFC, FB_AX_PI_SFC, FB_AX_COUNT };
#define FB_TB_BYTES ((TB_COUNT * 8 + 15) / 16 * 16)

#define FB_GEOID_NLAT 721
#define FB_GEOID_NLON 1441

// gain schedules (lib/FlightApps/src/c172/c172x/control/data/*.h5): [7 EAS][4 h][P] row-major per schedule (P = gs_P, the record stride), the ten
// schedules concatenated in GS order right behind the table pack in ONE device buffer, so that the integrator of the
// fly-by-wire variants stages pack + schedules into shared memory with a single bulk copy.
enum { GS_TE2TE = 0, GS_TV2TE, GS_VH2TE, GS_AR2AR, GS_PHIBETA2AR, GS_Q2E, GS_C2THETA, GS_V2T, GS_P2PHI, GS_CHI2PHI, GS_COUNT };
__host__ __device__ constexpr int gs_NV(int s) { return s < GS_Q2E ? (s == GS_VH2TE ? 39 : 36) : 4; }  // values per design point
// Record stride in doubles, odd: members of a warp sit in different cells, i.e. read the same value
// of different records; with a 36-double (or 4-double) stride every record starts on one of four shared-memory bank positions.
__host__ __device__ constexpr int gs_P(int s) { return gs_NV(s) | 1; }
__host__ __device__ constexpr int gs_off(int s) { int o = 0; for (int i = 0; i < s; i++) o += 28 * gs_P(i); return o; }
#define FB_GS_BYTES ((gs_off(GS_COUNT) * 8 + 15) / 16 * 16)

// Tail of the geoid buffer: what only the non-default components (FB_C172XA2) read. It sits behind the geoid table so that
// the right-hand side finds it from the pointer it already has (no extra parameter at any call site of the tuned kernels).
struct FbTablesTail {
    const double* prop3;      // [21 J][21 Mt][prop3_n][4] coefficients of the VariablePitch propeller (null: FixedPitch)
    double prop3_lo, prop3_hi, prop3_is;   // blade-pitch offset range (rad) and 1 / knot spacing
    int prop3_n, pad;
    double act2_wn2, act2_2zw;            // Actuator2: w_n^2 and 2 zeta w_n (c172x.jl:100)
};
#define FB_GEOID_TAIL_OFF (((size_t)FB_GEOID_NLAT * FB_GEOID_NLON * 4 + 15) / 16 * 16)     // bytes
#define FB_GEOID_BYTES (FB_GEOID_TAIL_OFF + sizeof(FbTablesTail))
#ifdef __CUDACC__
__device__ __forceinline__ const FbTablesTail* tables_tail(const float* geoid) {
    return reinterpret_cast<const FbTablesTail*>(reinterpret_cast<const char*>(geoid) + FB_GEOID_TAIL_OFF);
}
#endif

struct FbDeviceTables {
    const double* pack;   // TB_COUNT doubles, 16-byte aligned (global source of the shared pack), followed by
    const double* gains;  // the gain schedules (= pack + FB_TB_BYTES / 8)
    const float* geoid;   // [1441 lon][721 lat], followed by an FbTablesTail at byte FB_GEOID_TAIL_OFF
};
