// flightbatch_ext.cu — kernels of FB_C172XA2, the Cessna172X assembled from the non-default components the reference's Model
// tree admits: FlyByWireActuation of seven Actuator2 units (c172x.jl:68-106) and, optionally, a
// Propeller(VariablePitch) (propellers.jl:341-400). Its own translation unit: the integrator kernels of the BASELINE
// configurations are tuned to the register (flightbatch.cu, flightbatch_complete.cu), and ptxas' allocation of a kernel
// depends on what is compiled beside it. Complete integrator kernel only (no airborne layout); the C ABI reaches these
// kernels through fb_ext_ops().
#include "fb_host.cuh"

const FbOpsTable* fb_ext_ops(int kin) {
    static const FbOpsTable t[3] = {ops_table<FB_C172XA2, FB_WA>(), ops_table<FB_C172XA2, FB_ECEF>(), ops_table<FB_C172XA2, FB_NED>()};
    return &t[kin == FB_WA ? 0 : (kin == FB_ECEF ? 1 : 2)];
}
