// losses.cu -- fused, vectorised, warp-shuffle-reduced loss kernels and the VAE bottleneck, sm_100a.
//
//  * vcg_l1_fwd_bwd        : nn.L1Loss value + sign gradient in one pass       (Losses.py:21-24,36-39,63-65)
//  * vcg_mse_const_fwd_bwd : LSGAN mse_loss(d, zeros/ones) value + gradient   (Losses.py:80-81,99-100)
//  * vcg_kl_fwd_bwd        : KL(mu, clamp(logvar)) value + gradient            (Losses.py:115-121)
//  * vcg_reparam_fwd/bwd   : clamp, exp, z = mu + eps*std and its adjoint      (Networks.py:219-227)
//  * vcg_dhead_fwd/bwd     : spectral-normalised 512->1 16x16 conv = unit-vector dot product
//                            (Networks.py:248,267-269; torch/nn/utils/spectral_norm.py:92-114)
// All are HBM-bound: 16-byte loads, grid-stride over a grid sized to the SM count, one atomic per block.
#include "common.cuh"

namespace {

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sh[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.f;
  if (warp == 0) v = warp_sum(v);
  return v;   // valid in thread 0
}

__device__ __forceinline__ float sgn(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

__global__ void __launch_bounds__(256)
l1_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float scale, float* __restrict__ out,
          float* __restrict__ ga) {
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    acc += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
    if (ga) reinterpret_cast<float4*>(ga)[i] = make_float4(scale * sgn(d0), scale * sgn(d1), scale * sgn(d2), scale * sgn(d3));
  }
  for (long long i = (n4 << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = a[i] - b[i];
    acc += fabsf(d);
    if (ga) ga[i] = scale * sgn(d);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

__global__ void __launch_bounds__(256)
mse_const_kernel(const float* __restrict__ d, long long n, float target, float scale, float* __restrict__ out,
                 float* __restrict__ gd) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(gd)) & 15) == 0;
  const long long n4 = vec ? (n >> 2) : 0;
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 x = reinterpret_cast<const float4*>(d)[i];
    const float e0 = x.x - target, e1 = x.y - target, e2 = x.z - target, e3 = x.w - target;
    acc += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
    if (gd) reinterpret_cast<float4*>(gd)[i] = make_float4(2.f * scale * e0, 2.f * scale * e1, 2.f * scale * e2, 2.f * scale * e3);
  }
  for (long long i = (n4 << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float e = d[i] - target;
    acc += e * e;
    if (gd) gd[i] = scale * 2.f * e;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

__global__ void __launch_bounds__(256)
kl_kernel(const float* __restrict__ mu, const float* __restrict__ lv, long long n, float scale, float* __restrict__ out,
          float* __restrict__ gmu, float* __restrict__ glv) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(mu) | reinterpret_cast<uintptr_t>(lv) | reinterpret_cast<uintptr_t>(gmu) |
                     reinterpret_cast<uintptr_t>(glv)) & 15) == 0;
  const long long n4 = vec ? (n >> 2) : 0;
  float acc = 0.f;
  auto one = [&](float m, float l, float& gm, float& gl) {
    const float lc = fminf(fmaxf(l, -10.f), 10.f);
    const float e = expf(lc);
    acc += 1.f + lc - m * m - e;
    gm = scale * m;                                                  // d(-0.5*mean(...))/dmu = mu/N
    gl = (l >= -10.f && l <= 10.f) ? -0.5f * scale * (1.f - e) : 0.f;
  };
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 m = reinterpret_cast<const float4*>(mu)[i], l = reinterpret_cast<const float4*>(lv)[i];
    float4 gm, gl;
    one(m.x, l.x, gm.x, gl.x); one(m.y, l.y, gm.y, gl.y); one(m.z, l.z, gm.z, gl.z); one(m.w, l.w, gm.w, gl.w);
    if (gmu) { reinterpret_cast<float4*>(gmu)[i] = gm; reinterpret_cast<float4*>(glv)[i] = gl; }
  }
  for (long long i = (n4 << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float gm, gl;
    one(mu[i], lv[i], gm, gl);
    if (gmu) { gmu[i] = gm; glv[i] = gl; }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}

// ------------------------------------------------------------------ reparameterisation
template <typename T>
__global__ void __launch_bounds__(256)
reparam_fwd_kernel(const float* __restrict__ mu, int mu_pitch, const float* __restrict__ lv, int lv_pitch,
                   const float* __restrict__ eps, int n, int hw, int c, T* __restrict__ z, float* __restrict__ mu_out,
                   float* __restrict__ lv_out, float* __restrict__ kl_sum) {
  const long long total = static_cast<long long>(n) * hw * c;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int ch = static_cast<int>(i % c);
    const long long pix = i / c;                       // n*hw + p
    const int ni = static_cast<int>(pix / hw), pp = static_cast<int>(pix - static_cast<long long>(ni) * hw);
    const float m = mu[pix * mu_pitch + ch];
    const float l = lv[pix * lv_pitch + ch];
    const float lc = fminf(fmaxf(l, -10.f), 10.f);
    const long long nchw = (static_cast<long long>(ni) * c + ch) * hw + pp;
    const float e = eps[nchw];
    Elem<T>::st(z + i, m + e * expf(0.5f * lc));
    if (mu_out) { mu_out[nchw] = m; lv_out[nchw] = lc; }
    acc += 1.f + lc - m * m - expf(lc);
  }
  if (kl_sum) {
    acc = block_sum(acc);
    if (threadIdx.x == 0) atomicAdd(kl_sum, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
reparam_bwd_kernel(const float* __restrict__ mu, int mu_pitch, const float* __restrict__ lv, int lv_pitch,
                   const float* __restrict__ eps, const T* __restrict__ dz, int dz_pitch,
                   const float* __restrict__ gmu_ext, const float* __restrict__ glv_ext, float kl_scale, int n, int hw,
                   int c, T* __restrict__ dmu, int dmu_pitch, T* __restrict__ dlv, int dlv_pitch) {
  const long long total = static_cast<long long>(n) * hw * c;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int ch = static_cast<int>(i % c);
    const long long pix = i / c;
    const int ni = static_cast<int>(pix / hw), pp = static_cast<int>(pix - static_cast<long long>(ni) * hw);
    const float m = mu[pix * mu_pitch + ch];
    const float l = lv[pix * lv_pitch + ch];
    const float lc = fminf(fmaxf(l, -10.f), 10.f);
    const long long nchw = (static_cast<long long>(ni) * c + ch) * hw + pp;
    const float g = Elem<T>::ld(dz + pix * dz_pitch + ch);
    float gm = g + kl_scale * m;
    float gl = g * eps[nchw] * 0.5f * expf(0.5f * lc) - 0.5f * kl_scale * (1.f - expf(lc));
    if (gmu_ext) gm += gmu_ext[nchw];
    if (glv_ext) gl += glv_ext[nchw];
    if (!(l >= -10.f && l <= 10.f)) gl = 0.f;          // clamp passes gradient on the closed interval
    Elem<T>::st(dmu + pix * dmu_pitch + ch, gm);
    Elem<T>::st(dlv + pix * dlv_pitch + ch, gl);
  }
}

// ------------------------------------------------------------------ discriminator head
// grid (chunks, n+1): row y<n accumulates <x[y], w> into score[y]; row y==n accumulates |w|^2
template <typename T>
__global__ void __launch_bounds__(256)
dhead_dot_kernel(const T* __restrict__ x, const float* __restrict__ w, int n, int k, float* __restrict__ score,
                 float* __restrict__ wnorm2) {
  const int row = blockIdx.y;
  const int stride = gridDim.x * blockDim.x * 8;
  float acc = 0.f;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 8; i < k; i += stride) {
    float wv[8];
    ld8<float>(w + i, wv);
    if (row < n) {
      float xv[8];
      ld8<T>(x + static_cast<size_t>(row) * k + i, xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(xv[j], wv[j], acc);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(wv[j], wv[j], acc);
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(row < n ? score + row : wnorm2, acc);
}

__global__ void dhead_finish_kernel(float* score, const float* bias, const float* wnorm2, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) score[i] = score[i] * rsqrtf(*wnorm2) + bias[0];
}

// dx[n,:] = gs[n] * w/|w| ; G[k] = sum_n gs[n] x[n,k] -> scratch[k]; scratch[k_total] += <G, w>
template <typename T>
__global__ void __launch_bounds__(256)
dhead_bwd1_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ wnorm2,
                  const float* __restrict__ gs, int n, int k, T* __restrict__ dx, float* __restrict__ scratch) {
  const float inv = rsqrtf(*wnorm2);
  const int stride = gridDim.x * blockDim.x * 8;
  float gw = 0.f;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 8; i < k; i += stride) {
    float wv[8], G[8] = {};
    ld8<float>(w + i, wv);
    for (int r = 0; r < n; ++r) {
      const float g = gs[r];
      float xv[8], o[8];
      ld8<T>(x + static_cast<size_t>(r) * k + i, xv);
#pragma unroll
      for (int j = 0; j < 8; ++j) { G[j] = fmaf(g, xv[j], G[j]); o[j] = g * wv[j] * inv; }
      if (dx) st8<T>(dx + static_cast<size_t>(r) * k + i, o);
    }
    st8<float>(scratch + i, G);
#pragma unroll
    for (int j = 0; j < 8; ++j) gw = fmaf(G[j], wv[j], gw);
  }
  gw = block_sum(gw);
  if (threadIdx.x == 0) atomicAdd(scratch + k, gw);
}

__global__ void __launch_bounds__(256)
dhead_bwd2_kernel(const float* __restrict__ w, const float* __restrict__ wnorm2, const float* __restrict__ scratch,
                  const float* __restrict__ gs, int n, int k, float* __restrict__ dw, float* __restrict__ dbias, int dw_c) {
  const float n2 = *wnorm2, inv = rsqrtf(n2), coef = scratch[k] / n2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // dw_c > 0: dw is the OIHW gradient of the [1, dw_c, kh, kw] filter ((h,w,c) index i -> c * (k / dw_c) + pixel)
  if (i < k) dw[dw_c > 0 ? (i % dw_c) * (k / dw_c) + i / dw_c : i] += (scratch[i] - coef * w[i]) * inv;
  if (i == 0 && dbias) {
    float s = 0.f;
    for (int r = 0; r < n; ++r) s += gs[r];
    dbias[0] += s;
  }
}

// Spectral-norm power iteration of a 1 x k weight matrix (torch/nn/utils/spectral_norm.py:92-114 as run by every training
// forward of Networks.py:248), the (h, w, c)-ordered copy of the filter that vcg_dhead_* read, and
// aux = {sigma = u . (W v), |W|}.  With one output row the iteration converges in one step: v = W u / |W u|, u = +-1.
// Two small launches: (a) one block per 32-channel tile transposes it through shared memory (both sides coalesced) and
// accumulates sum w^2 (and sum w*v_stale for the eval-mode sigma); (b) writes v and finishes u / aux.
//   scratch (3 doubles, zeroed by the caller): {sum w^2, sum w * v_old, u_old}
__global__ void __launch_bounds__(256)
dhead_prepare_a_kernel(const float* __restrict__ w, int c, int hw, const float* __restrict__ u, const float* __restrict__ v,
                       float* __restrict__ w_hwc, double* __restrict__ scratch) {
  extern __shared__ float tile[];                          // [32][hw + 1]
  const int c0 = blockIdx.x * 32, t = threadIdx.x, pitch = hw + 1;
  const int nch = min(32, c - c0);
  double ww = 0.0, wv = 0.0;
  for (int i = t; i < nch * hw; i += 256) {
    const int ch = i / hw, px = i - ch * hw;
    const float x = w[static_cast<size_t>(c0 + ch) * hw + px];
    tile[ch * pitch + px] = x;
    ww += static_cast<double>(x) * x;
    wv += static_cast<double>(x) * v[static_cast<size_t>(c0 + ch) * hw + px];
  }
  __syncthreads();
  if (w_hwc)
    for (int i = t; i < nch * hw; i += 256) {
      const int px = i / nch, ch = i - px * nch;
      w_hwc[static_cast<size_t>(px) * c + c0 + ch] = tile[ch * pitch + px];
    }
  ww = warp_sum_d(ww);
  wv = warp_sum_d(wv);
  if ((t & 31) == 0) { atomicAdd(scratch, ww); atomicAdd(scratch + 1, wv); }
  if (blockIdx.x == 0 && t == 0) scratch[2] = static_cast<double>(u[0]);
}

__global__ void __launch_bounds__(256)
dhead_prepare_b_kernel(const float* __restrict__ w, int k, float* __restrict__ u, float* __restrict__ v,
                       float* __restrict__ aux, const double* __restrict__ scratch, int do_iter) {
  const double ww = scratch[0];
  const float wnorm = static_cast<float>(sqrt(ww)), u0 = static_cast<float>(scratch[2]);
  const float denom = fmaxf(fabsf(u0) * wnorm, 1e-12f);                     // |W^T u| = |u| |W|
  if (do_iter)
    for (int i = blockIdx.x * 256 + threadIdx.x; i < k; i += gridDim.x * 256) v[i] = w[i] * u0 / denom;   // v = normalize(W^T u)
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float un = u0, s;
    if (do_iter) {
      s = static_cast<float>(static_cast<double>(u0) * ww / denom);          // W v with the v written above
      un = s / fmaxf(fabsf(s), 1e-12f);                                       // u = normalize(W v)
      u[0] = un;
    } else {
      s = static_cast<float>(scratch[1]);                                     // W v_stale (eval mode)
    }
    if (aux) { aux[0] = un * s; aux[1] = wnorm; }
  }
}

int grid_for(long long n, int per_thread) {
  long long b = (n / per_thread + 255) / 256;
  const long long cap = 4LL * vcg_num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

extern "C" int vcg_l1_fwd_bwd(const float* a, const float* b, int64_t numel, float scale, float* out_sum, float* grad_a,
                              void* stream) {
  VCG_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(grad_a)) & 15) == 0,
              VCG_E_INVALID, "l1: pointers must be 16-byte aligned");
  l1_kernel<<<grid_for(numel, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, numel, scale, out_sum, grad_a);
  VCG_CHECK_LAUNCH("l1_kernel");
  return VCG_OK;
}

extern "C" int vcg_mse_const_fwd_bwd(const float* d, int64_t numel, float target, float scale, float* out_sum,
                                     float* grad_d, void* stream) {
  mse_const_kernel<<<grid_for(numel, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(d, numel, target, scale, out_sum, grad_d);
  VCG_CHECK_LAUNCH("mse_const_kernel");
  return VCG_OK;
}

extern "C" int vcg_kl_fwd_bwd(const float* mu, const float* lv, int64_t numel, float scale, float* out_sum, float* gmu,
                              float* glv, void* stream) {
  VCG_REQUIRE((gmu == nullptr) == (glv == nullptr), VCG_E_INVALID, "kl: gmu and glv must both be given or both NULL");
  kl_kernel<<<grid_for(numel, 1), 256, 0, static_cast<cudaStream_t>(stream)>>>(mu, lv, numel, scale, out_sum, gmu, glv);
  VCG_CHECK_LAUNCH("kl_kernel");
  return VCG_OK;
}

extern "C" int vcg_reparam_fwd(int32_t dtype, const float* mu, int32_t mu_pitch, const float* lv, int32_t lv_pitch,
                               const float* eps, int32_t n, int32_t hw, int32_t c, void* z, float* mu_out, float* lv_out,
                               float* kl_sum, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VCG_REQUIRE((mu_out == nullptr) == (lv_out == nullptr), VCG_E_INVALID, "reparam_fwd: mu_out/lv_out together");
  const int grid = grid_for(static_cast<long long>(n) * hw * c, 1);
  if (dtype == VCG_F32)
    reparam_fwd_kernel<float><<<grid, 256, 0, stream>>>(mu, mu_pitch, lv, lv_pitch, eps, n, hw, c, static_cast<float*>(z),
                                                        mu_out, lv_out, kl_sum);
  else
    reparam_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(mu, mu_pitch, lv, lv_pitch, eps, n, hw, c,
                                                                static_cast<__nv_bfloat16*>(z), mu_out, lv_out, kl_sum);
  VCG_CHECK_LAUNCH("reparam_fwd_kernel");
  return VCG_OK;
}

extern "C" int vcg_reparam_bwd(int32_t dtype, const float* mu, int32_t mu_pitch, const float* lv, int32_t lv_pitch,
                               const float* eps, const void* dz, int32_t dz_pitch, const float* gmu_ext,
                               const float* glv_ext, float kl_scale, int32_t n, int32_t hw, int32_t c, void* dmu,
                               int32_t dmu_pitch, void* dlv, int32_t dlv_pitch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int grid = grid_for(static_cast<long long>(n) * hw * c, 1);
  if (dtype == VCG_F32)
    reparam_bwd_kernel<float><<<grid, 256, 0, stream>>>(mu, mu_pitch, lv, lv_pitch, eps, static_cast<const float*>(dz), dz_pitch,
                                                        gmu_ext, glv_ext, kl_scale, n, hw, c, static_cast<float*>(dmu),
                                                        dmu_pitch, static_cast<float*>(dlv), dlv_pitch);
  else
    reparam_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        mu, mu_pitch, lv, lv_pitch, eps, static_cast<const __nv_bfloat16*>(dz), dz_pitch, gmu_ext, glv_ext, kl_scale, n, hw, c,
        static_cast<__nv_bfloat16*>(dmu), dmu_pitch, static_cast<__nv_bfloat16*>(dlv), dlv_pitch);
  VCG_CHECK_LAUNCH("reparam_bwd_kernel");
  return VCG_OK;
}

extern "C" int vcg_dhead_prepare(const float* w_oihw, int32_t c, int32_t hw, float* u, float* v, float* w_hwc, float* aux,
                                 double* scratch, int32_t do_iter, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VCG_REQUIRE(w_oihw && u && v && scratch && c > 0 && hw > 0, VCG_E_INVALID, "dhead_prepare: null / empty argument");
  VCG_REQUIRE(static_cast<size_t>(32) * (hw + 1) * sizeof(float) <= 48 * 1024, VCG_E_UNSUPPORTED, "dhead_prepare: %d taps", hw);
  cudaError_t e = cudaMemsetAsync(scratch, 0, 3 * sizeof(double), stream);
  VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "dhead_prepare: memset: %s", cudaGetErrorString(e));
  dhead_prepare_a_kernel<<<(c + 31) / 32, 256, static_cast<size_t>(32) * (hw + 1) * sizeof(float), stream>>>(w_oihw, c, hw, u, v, w_hwc, scratch);
  VCG_CHECK_LAUNCH("dhead_prepare_a_kernel");
  const int k = c * hw;
  dhead_prepare_b_kernel<<<do_iter ? (k + 2047) / 2048 : 1, 256, 0, stream>>>(w_oihw, k, u, v, aux, scratch, do_iter);
  VCG_CHECK_LAUNCH("dhead_prepare_b_kernel");
  return VCG_OK;
}

extern "C" int vcg_dhead_fwd(int32_t dtype, const void* x, const float* w_khwc, const float* bias, int32_t n, int32_t k,
                             float* score, float* wnorm2, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VCG_REQUIRE(k % 8 == 0, VCG_E_UNSUPPORTED, "dhead: k=%d must be a multiple of 8", k);
  // scores / wnorm2 are tiny: clear them with memset nodes (graph-capturable)
  cudaError_t e = cudaMemsetAsync(score, 0, static_cast<size_t>(n) * sizeof(float), stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(wnorm2, 0, sizeof(float), stream);
  VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "dhead: memset: %s", cudaGetErrorString(e));
  int chunks = k / (256 * 8 * 4);
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, n + 1);
  if (dtype == VCG_F32) dhead_dot_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), w_khwc, n, k, score, wnorm2);
  else dhead_dot_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), w_khwc, n, k, score, wnorm2);
  VCG_CHECK_LAUNCH("dhead_dot_kernel");
  dhead_finish_kernel<<<(n + 255) / 256, 256, 0, stream>>>(score, bias, wnorm2, n);
  VCG_CHECK_LAUNCH("dhead_finish_kernel");
  return VCG_OK;
}

extern "C" int vcg_dhead_bwd(int32_t dtype, const void* x, const float* w_khwc, const float* wnorm2, const float* gscore,
                             int32_t n, int32_t k, void* dx, float* dw, float* dbias, float* scratch, int32_t dw_c,
                             void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VCG_REQUIRE(k % 8 == 0, VCG_E_UNSUPPORTED, "dhead: k=%d must be a multiple of 8", k);
  cudaError_t e = cudaMemsetAsync(scratch + k, 0, sizeof(float), stream);
  VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "dhead: memset: %s", cudaGetErrorString(e));
  int blocks = k / (256 * 8);
  if (blocks < 1) blocks = 1;
  if (dtype == VCG_F32)
    dhead_bwd1_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(x), w_khwc, wnorm2, gscore, n, k,
                                                         static_cast<float*>(dx), scratch);
  else
    dhead_bwd1_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), w_khwc, wnorm2, gscore,
                                                                 n, k, static_cast<__nv_bfloat16*>(dx), scratch);
  VCG_CHECK_LAUNCH("dhead_bwd1_kernel");
  if (dw) {
    VCG_REQUIRE(dw_c >= 0 && (dw_c == 0 || k % dw_c == 0), VCG_E_INVALID, "dhead_bwd: dw_c=%d does not divide k=%d", dw_c, k);
    dhead_bwd2_kernel<<<(k + 255) / 256, 256, 0, stream>>>(w_khwc, wnorm2, scratch, gscore, n, k, dw, dbias, dw_c);
    VCG_CHECK_LAUNCH("dhead_bwd2_kernel");
  }
  return VCG_OK;
}
