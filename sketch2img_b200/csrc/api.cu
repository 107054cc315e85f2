// extern "C" surface of libs2i (declarations: include/s2i.h).
#include "../../include/s2i.h"
#include "common.cuh"
#include "gemm_tc.cuh"

extern "C" {

const char* s2i_last_error(void) { return s2i::last_error(); }
long long s2i_launch_count(void) { return s2i::g_launches; }

int s2i_gemm(const s2i_gemm_desc* d, void* cuda_stream) {
    if (!d) return s2i::set_error(S2I_ERR_ARG, "s2i_gemm: null descriptor");
    return s2i::gemm_launch(s2i::GemmDesc(*d), static_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
