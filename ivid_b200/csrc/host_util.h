// Host-side helpers shared by the C-ABI implementation: error handling, TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace ivid {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// status codes of the C ABI (include/ivid_b200.h)
enum : int {
  kOk = 0,
  kErrInvalidArgument = 1,   // maps to Python AssertionError / ValueError
  kErrNotImplemented = 2,    // maps to Python NotImplementedError
  kErrCuda = 3,              // maps to RuntimeError
  kErrState = 4,
};

#define IVID_CHECK_CUDA(expr)                                                                              \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess)                                                                                 \
      throw ::ivid::Error(::ivid::kErrCuda, std::string(#expr) + " failed: " + cudaGetErrorString(_e) +   \
                                                " (" __FILE__ ":" + std::to_string(__LINE__) + ")");       \
  } while (0)

#define IVID_REQUIRE(cond, msg)                                                      \
  do {                                                                               \
    if (!(cond)) throw ::ivid::Error(::ivid::kErrInvalidArgument, std::string(msg)); \
  } while (0)

// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda).
CUtensorMap make_tensor_map(CUtensorMapDataType dtype, int rank, void* base, const uint64_t* dims,
                            const uint64_t* strides_bytes /* rank-1 entries, dim0 is dense */, const uint32_t* box,
                            CUtensorMapSwizzle swizzle);

// NHWC fp16 activation [N][H][W][C] viewed as (C, W, H, N); box = (64, TW, TH, TN), 128B swizzle.
CUtensorMap make_act_map(const void* base, int N, int H, int W, int C, int TW, int TH, int TN);
// fp16 weight matrix [rows][K] (K contiguous); box = (64, box_rows), 128B swizzle.
CUtensorMap make_weight_map(const void* base, int rows, int K, int box_rows);

int sm_count();

}  // namespace ivid
