// sweep_common.cuh -- arithmetic shared by the sweep kernels (bit-parity contract).
#ifndef VHP_SWEEP_COMMON_CUH
#define VHP_SWEEP_COMMON_CUH

namespace {

__device__ __forceinline__ double lerp_rn(double a, double b, double c) {
  return __dsub_rn(a, __dmul_rn(c, __dsub_rn(a, b)));
}

// Correctly rounded i/k for integers 0 <= i < k <= 16384 from r = RN(1/k):
// one Newton/Markstein correction.  Verified exhaustively against IEEE division
// (tests/test_gpu_sweep.py::test_ratio_exact_exhaustive and the host-side check
// described in DESIGN.md).
__device__ __forceinline__ double ratio_rn(double fi, double fk, double r) {
  const double q0 = __dmul_rn(fi, r);
  const double rem = __fma_rn(-q0, fk, fi);
  return __fma_rn(rem, r, q0);
}

template <typename OutT> __device__ __forceinline__ OutT to_out(double v);
template <> __device__ __forceinline__ float to_out<float>(double v) { return __double2float_rn(v); }
template <> __device__ __forceinline__ double to_out<double>(double v) { return v; }

} // namespace
#endif
