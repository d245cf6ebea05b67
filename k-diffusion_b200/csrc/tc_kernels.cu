// tc_kernels.cu -- placeholder until the tcgen05 kernels land (next commit): everything reports
// "unsupported" so the engine takes the SIMT path.
#include "tc_kernels.cuh"

namespace kdb {
bool tc_gemm_supported(int64_t, int, int, const GemmEpi&) { return false; }
int launch_gemm_tc(const bf16*, const bf16*, bf16*, int64_t, int, int, const GemmEpi&, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return KDB_ERR_UNSUPPORTED;
}
bool tc_gemm_geglu_supported(int64_t, int, int) { return false; }
int launch_gemm_tc_geglu(const bf16*, const bf16*, bf16*, int64_t, int, int, cudaStream_t) {
  set_error("tcgen05 GEMM not built");
  return KDB_ERR_UNSUPPORTED;
}
bool tc_attention_supported(int, int, int, int, int, int) { return false; }
int launch_attention_tc(const bf16*, bf16*, int, int, int, int, int, int, int, int, cudaStream_t) {
  set_error("tcgen05 attention not built");
  return KDB_ERR_UNSUPPORTED;
}
}  // namespace kdb

extern "C" int kdb_gemm_bf16(const void*, const void*, void*, int, int, int, void*) {
  kdb::set_error("tcgen05 GEMM not built");
  return KDB_ERR_UNSUPPORTED;
}
