// tcgen05 (5th-generation tensor core) GEMM path -- placeholder until the TMA/TMEM kernel lands.
#pragma once
#include "gemm_simt.cuh"

namespace mpmae {
inline bool tc_gemm_supported(int /*mode*/, const GemmArgs & /*a*/) { return false; }
template <int MODE>
inline cudaError_t launch_gemm_rows_tc(const GemmArgs &, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace mpmae
