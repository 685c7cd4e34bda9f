// adam.cu -- one-launch multi-tensor Adam, sm_100a.
//
// The reference builds torch.optim.Adam(params, lr, betas=(0.5, 0.999)) (Networks.py:312,894,
// 1032-1033,1212-1213,1372,1498,1669-1676,1928-1935); torch/optim/adam.py:457-547 then issues several
// foreach kernels per parameter group.  Here the two generators' (or the two discriminators')
// parameters are described by a table of <=64K-element chunks and updated by ONE kernel:
//   m = lerp(m, g, 1-b1);  v = b2*v + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// HBM traffic: 16 B read + 12 B written per parameter (fp32 p, g, m, v) = 28 B/param.
//
// Variants selected by `flags` (the data-parallel / overlapped step of optim.py):
//   VCG_ADAM_TICK       advance the device-side step counter (once per optimiser step; the further buckets of the same
//                       step pass 0);
//   VCG_ADAM_GRAD_BF16  g points to bfloat16 gradients (the all-reduced wire buffer): 26 B/param;
//   VCG_ADAM_ZERO_GRAD  write zeros over the fp32 gradient after reading it, so that zero_grad() of the next step is
//                       free (32 B/param instead of 28 + a 4 B/param fill launch).
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

template <bool kBf16, bool kZero>
__global__ void __launch_bounds__(256)
adam_multi_kernel(const vcg_adam_chunk* __restrict__ chunks, const float* __restrict__ state, float beta1, float beta2,
                  float eps, float grad_scale) {
  const float lr_over_bc1 = state[1], bc2_sqrt = state[2];
  const vcg_adam_chunk ck = chunks[blockIdx.x];
  const int n = ck.numel;
  const float w1 = 1.f - beta1, w2 = 1.f - beta2;
  const __nv_bfloat16* gb = reinterpret_cast<const __nv_bfloat16*>(ck.g);
  float* gz = const_cast<float*>(ck.g);
  const bool vec = ((reinterpret_cast<uintptr_t>(ck.p) | reinterpret_cast<uintptr_t>(ck.m) | reinterpret_cast<uintptr_t>(ck.v)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(ck.g) & (kBf16 ? 7 : 15)) == 0;
  const int n4 = vec ? (n >> 2) : 0;
  for (int i = threadIdx.x; i < n4; i += 256) {
    float4 p = reinterpret_cast<float4*>(ck.p)[i];
    float4 g4;
    if (kBf16) {
      const uint2 r = reinterpret_cast<const uint2*>(gb)[i];
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
      g4 = make_float4(a.x, a.y, b.x, b.y);
    } else {
      g4 = reinterpret_cast<const float4*>(ck.g)[i];
    }
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
    // (after the arithmetic: a store to the address of a load that is still in flight stalls the thread until the
    //  load returns -- zeroing right behind the load made this kernel 5x slower)
    if (!kBf16 && kZero) reinterpret_cast<float4*>(gz)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) {
    const float g = (kBf16 ? __bfloat162float(gb[i]) : ck.g[i]) * grad_scale;
    const float m = lerp_like_torch(ck.m[i], g, w1);
    const float v = ck.v[i] * beta2 + w2 * g * g;
    ck.m[i] = m; ck.v[i] = v;
    ck.p[i] = ck.p[i] - lr_over_bc1 * (m / (sqrtf(v) / bc2_sqrt + eps));
    if (!kBf16 && kZero) gz[i] = 0.f;
  }
}


// dst (bf16) = src (fp32), optionally zeroing src: the gradient wire buffer of the data-parallel all-reduce
__global__ void __launch_bounds__(256) cast_bf16_kernel(float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n4,
                                                         long long n, int zero_src) {
  const long long stride = static_cast<long long>(gridDim.x) * 256;
  auto pack = [](const float4& v) {
    uint2 r;
    *reinterpret_cast<__nv_bfloat162*>(&r.x) = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(&r.y) = __floats2bfloat162_rn(v.z, v.w);
    return r;
  };
  // two independent 16-byte loads in flight per thread; the zero stores come last (a store to the address of a load
  // that is still in flight would stall the thread until the load returns)
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += 2 * stride) {
    const long long j = i + stride;
    const bool two = j < n4;
    const float4 a = reinterpret_cast<const float4*>(src)[i];
    const float4 b = two ? reinterpret_cast<const float4*>(src)[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<uint2*>(dst)[i] = pack(a);
    if (two) reinterpret_cast<uint2*>(dst)[j] = pack(b);
    if (zero_src) {
      reinterpret_cast<float4*>(src)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (two) reinterpret_cast<float4*>(src)[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  for (long long i = 4 * n4 + static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n; i += stride) {
    dst[i] = __float2bfloat16_rn(src[i]);
    if (zero_src) src[i] = 0.f;
  }
}

}  // namespace

extern "C" int vcg_adam_multi(const vcg_adam_chunk* chunks_dev, int32_t nchunks, float* state_dev, float lr, float beta1,
                              float beta2, float eps, float grad_scale, int32_t flags, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VCG_REQUIRE(state_dev, VCG_E_INVALID, "adam: device state is NULL");
  VCG_REQUIRE(!((flags & VCG_ADAM_GRAD_BF16) && (flags & VCG_ADAM_ZERO_GRAD)), VCG_E_INVALID,
              "adam: ZERO_GRAD applies to fp32 gradients (the bf16 wire buffer is overwritten by the next cast)");
  if (flags & VCG_ADAM_TICK) {
    adam_tick_kernel<<<1, 1, 0, stream>>>(state_dev, lr, beta1, beta2);
    VCG_CHECK_LAUNCH("adam_tick_kernel");
  }
  if (nchunks <= 0) return VCG_OK;
  if (flags & VCG_ADAM_GRAD_BF16) adam_multi_kernel<true, false><<<nchunks, 256, 0, stream>>>(chunks_dev, state_dev, beta1, beta2, eps, grad_scale);
  else if (flags & VCG_ADAM_ZERO_GRAD) adam_multi_kernel<false, true><<<nchunks, 256, 0, stream>>>(chunks_dev, state_dev, beta1, beta2, eps, grad_scale);
  else adam_multi_kernel<false, false><<<nchunks, 256, 0, stream>>>(chunks_dev, state_dev, beta1, beta2, eps, grad_scale);
  VCG_CHECK_LAUNCH("adam_multi_kernel");
  return VCG_OK;
}

extern "C" int vcg_cast_bf16(float* src, void* dst, int64_t n, int32_t zero_src, void* stream_) {
  if (n <= 0) return VCG_OK;
  VCG_REQUIRE(src && dst, VCG_E_INVALID, "cast_bf16: null argument");
  const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) | (reinterpret_cast<uintptr_t>(dst) & 7)) == 0;
  const long long n4 = vec ? n / 4 : 0;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = 16LL * vcg_num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cast_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      src, static_cast<__nv_bfloat16*>(dst), n4, n, zero_src);
  VCG_CHECK_LAUNCH("cast_bf16_kernel");
  return VCG_OK;
}
