// wgrad_thin.cu -- weight gradient of convolutions with very few output channels (cout <= 4: the
// 64->3 7x7 output convolution of the decoder, Networks.py:192), sm_100a.
//
//   dw[co][kh][kw*c + ci] += sum_{n,h,w} dy[n,h,w,co] * x[n, h+kh, w+kw, ci]
//
// With M = 3 this is not GEMM-shaped for tcgen05 (minimum M is 64): 97% of the MMA rows would be padding.
// Instead every thread owns ~13 (tap, ci) pairs x 4 cout accumulators in registers and a persistent CTA
// streams 16x16 output tiles through shared memory (the 22x22x64 input patch is read once from HBM/L2 and
// reused by all 49 taps).  Bound: FP32 FMA pipe + shared-memory reads; HBM traffic = x once + dy once.
#include "common.cuh"

namespace {

constexpr int TILE = 16, CO = 4, MAXP = 16;

struct ThinArgs {
  int n, hp, wp, c, kh, kw, kwc_pad, ho, wo, cout, dy_halo, dy_c;
  int tiles_w, tiles_h, ntiles, npairs;
};

template <typename T>
__global__ void __launch_bounds__(256)
wgrad_thin_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, ThinArgs p) {
  extern __shared__ __align__(16) uint8_t smem_thin[];
  const int pw = TILE + p.kw - 1, ph = TILE + p.kh - 1;
  T* xs = reinterpret_cast<T*>(smem_thin);                                   // [ph][pw][c]
  float* dys = reinterpret_cast<float*>(smem_thin + ((static_cast<size_t>(ph) * pw * p.c * sizeof(T) + 15) / 16) * 16);  // [256][CO]
  const int tid = threadIdx.x;
  float acc[MAXP][CO];
  int off[MAXP];
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int pr = tid + i * 256;
    if (pr < p.npairs) {
      const int tap = pr / p.c, ci = pr - tap * p.c;
      const int khi = tap / p.kw, kwi = tap - khi * p.kw;
      off[i] = (khi * pw + kwi) * p.c + ci;
    } else off[i] = -1;
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[i][j] = 0.f;
  }
  const int wpd = p.wo + 2 * p.dy_halo, hpd = p.ho + 2 * p.dy_halo;
  const int vec_per_row = pw * p.c / 8;
  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int img = tile / (p.tiles_w * p.tiles_h), rem = tile - img * p.tiles_w * p.tiles_h;
    const int h0 = (rem / p.tiles_w) * TILE, w0 = (rem % p.tiles_w) * TILE;
    __syncthreads();
    // input patch rows are contiguous runs of pw*c elements in the padded NHWC input
    for (int v = tid; v < ph * vec_per_row; v += 256) {
      const int r = v / vec_per_row, q = v - r * vec_per_row;
      const T* src = x + ((static_cast<size_t>(img) * p.hp + h0 + r) * p.wp + w0) * p.c + q * 8;
      *reinterpret_cast<uint4*>(xs + (static_cast<size_t>(r) * pw) * p.c + q * 8) = *reinterpret_cast<const uint4*>(src);
      if (sizeof(T) == 4)
        *reinterpret_cast<uint4*>(reinterpret_cast<float*>(xs) + (static_cast<size_t>(r) * pw) * p.c + q * 8 + 4) =
            *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(src) + 4);
    }
    {
      const int py = tid / TILE, px = tid - py * TILE;
      const T* src = dy + ((static_cast<size_t>(img) * hpd + h0 + py + p.dy_halo) * wpd + w0 + px + p.dy_halo) * p.dy_c;
#pragma unroll
      for (int j = 0; j < CO; ++j) dys[tid * CO + j] = j < p.cout ? Elem<T>::ld(src + j) : 0.f;
    }
    __syncthreads();
    for (int pix = 0; pix < TILE * TILE; ++pix) {
      const int py = pix / TILE, px = pix - py * TILE;
      const float4 d = *reinterpret_cast<const float4*>(dys + pix * CO);
      const T* base = xs + (py * pw + px) * p.c;
#pragma unroll
      for (int i = 0; i < MAXP; ++i) {
        if (off[i] >= 0) {
          const float xv = Elem<T>::ld(base + off[i]);
          acc[i][0] = fmaf(xv, d.x, acc[i][0]);
          acc[i][1] = fmaf(xv, d.y, acc[i][1]);
          acc[i][2] = fmaf(xv, d.z, acc[i][2]);
          acc[i][3] = fmaf(xv, d.w, acc[i][3]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int pr = tid + i * 256;
    if (pr < p.npairs) {
      const int tap = pr / p.c, ci = pr - tap * p.c;
      const int khi = tap / p.kw, kwi = tap - khi * p.kw;
#pragma unroll
      for (int j = 0; j < CO; ++j)
        if (j < p.cout) atomicAdd(dw + (static_cast<size_t>(j) * p.kh + khi) * p.kwc_pad + kwi * p.c + ci, acc[i][j]);
    }
  }
}

}  // namespace

bool vcg_wgrad_thin_supported(const vcg_conv_desc* d) {
  const int ho = d->hp - d->kh + 1, wo = d->wp - d->kw + 1;
  return d->cout <= CO && d->kh * d->kw * d->c <= 256 * MAXP && ho % TILE == 0 && wo % TILE == 0 && d->c % 8 == 0;
}

int vcg_conv_wgrad_thin(const vcg_conv_desc* d, const void* x, const void* dy, int dy_halo, int dy_c, float* dw,
                        cudaStream_t stream) {
  ThinArgs a{};
  a.n = d->n; a.hp = d->hp; a.wp = d->wp; a.c = d->c; a.kh = d->kh; a.kw = d->kw; a.kwc_pad = d->kwc_pad;
  a.ho = d->hp - d->kh + 1; a.wo = d->wp - d->kw + 1; a.cout = d->cout; a.dy_halo = dy_halo; a.dy_c = dy_c;
  a.tiles_w = a.wo / TILE; a.tiles_h = a.ho / TILE; a.ntiles = d->n * a.tiles_w * a.tiles_h;
  a.npairs = d->kh * d->kw * d->c;
  const size_t es = d->dtype == VCG_F32 ? 4 : 2;
  const size_t xs_bytes = ((static_cast<size_t>(TILE + d->kh - 1) * (TILE + d->kw - 1) * d->c * es + 15) / 16) * 16;
  const size_t smem = xs_bytes + TILE * TILE * CO * sizeof(float);
  VCG_REQUIRE(smem <= 227 * 1024, VCG_E_UNSUPPORTED, "wgrad_thin: patch does not fit shared memory");
  int grid = 3 * vcg_num_sms();
  if (grid > a.ntiles) grid = a.ntiles;
  cudaError_t e;
  if (d->dtype == VCG_F32) {
    e = cudaFuncSetAttribute(wgrad_thin_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "wgrad_thin: %s", cudaGetErrorString(e));
    wgrad_thin_kernel<float><<<grid, 256, smem, stream>>>(static_cast<const float*>(x), static_cast<const float*>(dy), dw, a);
  } else {
    e = cudaFuncSetAttribute(wgrad_thin_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "wgrad_thin: %s", cudaGetErrorString(e));
    wgrad_thin_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(static_cast<const __nv_bfloat16*>(x),
                                                                static_cast<const __nv_bfloat16*>(dy), dw, a);
  }
  VCG_CHECK_LAUNCH("wgrad_thin_kernel");
  return VCG_OK;
}
