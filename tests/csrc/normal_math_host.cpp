// normal_math_host.cpp — host build of the product's table-driven Box–Muller transform (TEST INFRASTRUCTURE ONLY).
//
// Compiles museinference.jl_b200/csrc/muse_normal_math.cuh and the generated tables with g++ so that tests can check
// the transform's arithmetic on the CPU (against long-double references and against oracle/philox.py) before it
// ever runs on a GPU: the header uses explicit fma() throughout, so host and device results agree bit for bit.
#include <cstdint>
#define __device__
#include "../../museinference.jl_b200/csrc/muse_draw_tables.cuh"
#undef __device__
#include "../../museinference.jl_b200/csrc/muse_normal_math.cuh"

extern "C" {

// t = −2 ln u₁ and (cos, sin)(2πu₂) for n 53-bit integers each
void muse_host_neg2log(const uint64_t* v, int n, double* out) {
    for (int i = 0; i < n; ++i) out[i] = muse::neg2_log_u(v[i], muse::kLogTab);
}
void muse_host_sincos(const uint64_t* v, int n, double* cs, double* sn) {
    for (int i = 0; i < n; ++i) muse::sincos_2pi_u(v[i], muse::kTrigTab, &cs[i], &sn[i]);
}
// normals of n Philox blocks (r0..r3 per block)
void muse_host_box_muller(const uint32_t* r, int n, double* out) {
    for (int i = 0; i < n; ++i)
        muse::box_muller_tab(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3], muse::kLogTab, muse::kTrigTab, &out[2 * i], &out[2 * i + 1]);
}

}
