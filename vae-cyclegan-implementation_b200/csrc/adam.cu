// adam.cu -- one-launch multi-tensor Adam, sm_100a.
//
// The reference builds torch.optim.Adam(params, lr, betas=(0.5, 0.999)) (Networks.py:312,894,
// 1032-1033,1212-1213,1372,1498,1669-1676,1928-1935); torch/optim/adam.py:457-547 then issues several
// foreach kernels per parameter group.  Here the two generators' (or the two discriminators')
// parameters are described by a table of <=64K-element chunks and updated by ONE kernel:
//   m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// HBM traffic: 16 B read + 12 B written per parameter (fp32 p, g, m, v) = 28 B/param.
#include "common.cuh"

namespace {

__device__ __forceinline__ float lerp_like_torch(float a, float b, float w) {
  // ATen lerp: weight < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
  const float d = b - a;
  return w < 0.5f ? a + w * d : b - d * (1.f - w);
}

// Device-side step counter and bias corrections: state = {step, lr/bc1, sqrt(bc2), unused}.  Keeping them on
// the device makes the optimiser step CUDA-graph replayable (no host scalar is baked into the launch).
__global__ void adam_tick_kernel(float* __restrict__ state, float lr, float beta1, float beta2) {
  const double step = static_cast<double>(state[0]) + 1.0;
  state[0] = static_cast<float>(step);
  state[1] = static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(beta1), step)));
  state[2] = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), step)));
}

__global__ void __launch_bounds__(256)
adam_multi_kernel(const vcg_adam_chunk* __restrict__ chunks, const float* __restrict__ state, float beta1, float beta2,
                  float eps, float grad_scale) {
  const float lr_over_bc1 = state[1], bc2_sqrt = state[2];
  const vcg_adam_chunk ck = chunks[blockIdx.x];
  const int n = ck.numel;
  const float w1 = 1.f - beta1, w2 = 1.f - beta2;
  const bool vec = ((reinterpret_cast<uintptr_t>(ck.p) | reinterpret_cast<uintptr_t>(ck.g) |
                     reinterpret_cast<uintptr_t>(ck.m) | reinterpret_cast<uintptr_t>(ck.v)) & 15) == 0;
  const int n4 = vec ? (n >> 2) : 0;
  for (int i = threadIdx.x; i < n4; i += 256) {
    float4 p = reinterpret_cast<float4*>(ck.p)[i];
    const float4 g4 = reinterpret_cast<const float4*>(ck.g)[i];
    float4 m = reinterpret_cast<float4*>(ck.m)[i];
    float4 v = reinterpret_cast<float4*>(ck.v)[i];
    float* pp = &p.x; const float* gp = &g4.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float g = gp[j] * grad_scale;
      mp[j] = lerp_like_torch(mp[j], g, w1);
      vp[j] = vp[j] * beta2 + w2 * g * g;
      const float denom = sqrtf(vp[j]) / bc2_sqrt + eps;
      pp[j] = pp[j] - lr_over_bc1 * (mp[j] / denom);
    }
    reinterpret_cast<float4*>(ck.p)[i] = p;
    reinterpret_cast<float4*>(ck.m)[i] = m;
    reinterpret_cast<float4*>(ck.v)[i] = v;
  }
  for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) {
    const float g = ck.g[i] * grad_scale;
    const float m = lerp_like_torch(ck.m[i], g, w1);
    const float v = ck.v[i] * beta2 + w2 * g * g;
    ck.m[i] = m; ck.v[i] = v;
    ck.p[i] = ck.p[i] - lr_over_bc1 * (m / (sqrtf(v) / bc2_sqrt + eps));
  }
}

}  // namespace

extern "C" int vcg_adam_multi(const vcg_adam_chunk* chunks_dev, int32_t nchunks, float* state_dev, float lr, float beta1,
                              float beta2, float eps, float grad_scale, void* stream) {
  if (nchunks <= 0) return VCG_OK;
  VCG_REQUIRE(state_dev, VCG_E_INVALID, "adam: device state is NULL");
  adam_tick_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(state_dev, lr, beta1, beta2);
  VCG_CHECK_LAUNCH("adam_tick_kernel");
  adam_multi_kernel<<<nchunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(chunks_dev, state_dev, beta1, beta2, eps, grad_scale);
  VCG_CHECK_LAUNCH("adam_multi_kernel");
  return VCG_OK;
}
