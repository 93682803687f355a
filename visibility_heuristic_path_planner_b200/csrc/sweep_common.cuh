// sweep_common.cuh -- arithmetic shared by the sweep kernels (bit-parity contract).
#ifndef VHP_SWEEP_COMMON_CUH
#define VHP_SWEEP_COMMON_CUH

namespace {

__device__ __forceinline__ double lerp_rn(double a, double b, double c) {
  return __dsub_rn(a, __dmul_rn(c, __dsub_rn(a, b)));
}

template <typename OutT> __device__ __forceinline__ OutT to_out(double v);
template <> __device__ __forceinline__ float to_out<float>(double v) { return __double2float_rn(v); }
template <> __device__ __forceinline__ double to_out<double>(double v) { return v; }

} // namespace
#endif
