// wgrad_thin.cu -- weight gradient of convolutions with very few output channels (cout <= 8: the
// 64->3 7x7 output convolution of the decoder, Networks.py:192), sm_100a, bf16 mode.
//
//   dw[co][kh][kw*C + ci] += sum_{n,h,w} dy[n,h,w,co] * x[n, h+kh, w+kw, ci]
//
// With M = cout = 3 this is no tcgen05 shape (minimum M is 64: >95% of the MMA rows would be padding, and
// the A operand -- the 49-tap window of a 64-channel input -- would be re-read 49x through L2).  Instead a
// persistent CTA streams 16x16 output tiles: the (16+6)x(16+6)x64 input patch is staged ONCE in shared
// memory and reused by all 49 taps; the products run on the warp-level tensor-core path
// (mma.sync m16n8k16: M = cout padded to 16, N = 8 input channels, K = 16 pixels of one tile row), with the
// B fragments fetched by ldmatrix.trans straight from the NHWC patch.  Warp w owns input-channel block w
// for all taps, so its 49 x (8 ci x cout) accumulators stay in registers across tiles; one red.add per
// weight per CTA at the end.  HBM traffic = x once + dy once.
#include "common.cuh"

namespace {

constexpr int TILE = 16;

struct ThinArgs {
  int n, hp, wp, ho, wo, cout, dy_halo, dy_c, kwc_pad;
  int tiles_w, tiles_h, ntiles;
};

__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16_16816(float& d0, float& d1, float& d2, float& d3, uint32_t a0, uint32_t a1,
                                               uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d0), "+f"(d1), "+f"(d2), "+f"(d3)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// KH x KW taps, C input channels (C/8 == number of warps)
template <int KH, int KW, int C>
__global__ void __launch_bounds__(C / 8 * 32, 1)
wgrad_thin_mma_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw,
                      ThinArgs p) {
  constexpr int PW = TILE + KW - 1, PH = TILE + KH - 1, CP = C + 8;   // +8 elements: conflict-free ldmatrix rows
  constexpr int NT = C / 8 * 32;
  extern __shared__ __align__(16) uint8_t smem_thin[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem_thin);            // [PH][PW][CP]
  __nv_bfloat16* dyT = xs + PH * PW * CP;                                      // [8][256]: dy transposed, co-major
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  float acc[KH * KW][2];
#pragma unroll
  for (int i = 0; i < KH * KW; ++i) acc[i][0] = acc[i][1] = 0.f;
  float z0 = 0.f, z1 = 0.f;                      // rows 8..15 of the MMA tile: always zero, shared by all taps
  const int wpd = p.wo + 2 * p.dy_halo, hpd = p.ho + 2 * p.dy_halo;
  const uint32_t xs_s = smem_u32(xs);
  // ldmatrix row address of this lane: pixel (lane & 15) of the k-step, channel block `warp`
  const uint32_t lane_off = static_cast<uint32_t>(((lane & 15) * CP + warp * 8) * 2);

  for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const int img = tile / (p.tiles_w * p.tiles_h), rem = tile - img * p.tiles_w * p.tiles_h;
    const int h0 = (rem / p.tiles_w) * TILE, w0 = (rem % p.tiles_w) * TILE;
    __syncthreads();
    for (int v = tid; v < PH * PW * (C / 8); v += NT) {
      const int q = v % (C / 8), pix = v / (C / 8);
      const int r = pix / PW, cx = pix - r * PW;
      const __nv_bfloat16* src = x + ((static_cast<size_t>(img) * p.hp + h0 + r) * p.wp + w0 + cx) * C + q * 8;
      *reinterpret_cast<uint4*>(xs + (r * PW + cx) * CP + q * 8) = *reinterpret_cast<const uint4*>(src);
    }
    for (int v = tid; v < TILE * TILE; v += NT) {
      const int py = v / TILE, px = v - py * TILE;
      const __nv_bfloat16* src = dy + ((static_cast<size_t>(img) * hpd + h0 + py + p.dy_halo) * wpd + w0 + px + p.dy_halo) * p.dy_c;
#pragma unroll
      for (int j = 0; j < 8; ++j) dyT[j * 256 + v] = j < p.cout ? src[j] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
#pragma unroll 1
    for (int py = 0; py < TILE; ++py) {
      // A = dy^T (16 co x 16 px of tile row py): a0 = (co g, px 2t..2t+1), a2 = (co g, px 2t+8..2t+9); rows 8..15 = 0
      const uint32_t a0 = *reinterpret_cast<const uint32_t*>(dyT + g * 256 + py * TILE + 2 * t);
      const uint32_t a2 = *reinterpret_cast<const uint32_t*>(dyT + g * 256 + py * TILE + 2 * t + 8);
#pragma unroll
      for (int kh = 0; kh < KH; ++kh) {
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) {
          uint32_t b0, b1;
          ldmatrix_x2_trans(b0, b1, xs_s + static_cast<uint32_t>((((py + kh) * PW + kw) * CP) * 2) + lane_off);
          mma_bf16_16816(acc[kh * KW + kw][0], acc[kh * KW + kw][1], z0, z1, a0, 0u, a2, 0u, b0, b1);
        }
      }
    }
  }
  // D[co = g][ci = warp*8 + 2t, +1] for every tap
  if (g < p.cout) {
#pragma unroll
    for (int kh = 0; kh < KH; ++kh)
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
        float* dst = dw + (static_cast<size_t>(g) * KH + kh) * p.kwc_pad + kw * C + warp * 8 + 2 * t;
        atomicAdd(dst, acc[kh * KW + kw][0]);
        atomicAdd(dst + 1, acc[kh * KW + kw][1]);
      }
  }
}

}  // namespace

bool vcg_wgrad_thin_supported(const vcg_conv_desc* d) {
  const int ho = d->hp - d->kh + 1, wo = d->wp - d->kw + 1;
  return d->dtype == VCG_BF16 && d->cout <= 8 && d->kh == 7 && d->kw == 7 && d->c == 64 && ho % TILE == 0 && wo % TILE == 0;
}

int vcg_conv_wgrad_thin(const vcg_conv_desc* d, const void* x, const void* dy, int dy_halo, int dy_c, float* dw,
                        cudaStream_t stream) {
  VCG_REQUIRE(vcg_wgrad_thin_supported(d), VCG_E_UNSUPPORTED, "wgrad_thin: unsupported shape");
  ThinArgs a{};
  a.n = d->n; a.hp = d->hp; a.wp = d->wp; a.ho = d->hp - d->kh + 1; a.wo = d->wp - d->kw + 1; a.cout = d->cout;
  a.dy_halo = dy_halo; a.dy_c = dy_c; a.kwc_pad = d->kwc_pad;
  a.tiles_w = a.wo / TILE; a.tiles_h = a.ho / TILE; a.ntiles = d->n * a.tiles_w * a.tiles_h;
  constexpr int KH = 7, KW = 7, C = 64;
  const size_t smem = static_cast<size_t>(TILE + KH - 1) * (TILE + KW - 1) * (C + 8) * 2 + 8 * 256 * 2;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_thin_mma_kernel<KH, KW, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "wgrad_thin: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int grid = vcg_gemm_sms();      // 246 registers x 256 threads: one persistent CTA per SM
  if (grid > a.ntiles) grid = a.ntiles;
  wgrad_thin_mma_kernel<KH, KW, C><<<grid, C / 8 * 32, smem, stream>>>(static_cast<const __nv_bfloat16*>(x),
                                                                     static_cast<const __nv_bfloat16*>(dy), dw, a);
  VCG_CHECK_LAUNCH("wgrad_thin_mma_kernel");
  return VCG_OK;
}
