// wgrad_tc.cu -- weight-gradient pass of every convolution on tcgen05 tensor cores, sm_100a.
//
//   dW[co, kh, j] += sum_{n,h,w} dY[n,h,w,co] * X[n, h+kh, w*c + j]        (j indexes (kw, cin))
//
// is a GEMM whose reduction dimension is the pixel index, so BOTH operands are "MN-major": the
// NHWC channel axis is contiguous and becomes M (cout, from dY) and N (a run of the packed filter
// row, from the saved forward input X).  A 64-pixel box of 64 channels lands in shared memory via
// TMA as 64 rows of 128 B, which is exactly the MN-major SWIZZLE_128B UMMA layout (K = pixel rows,
// 8-row groups 1024 B apart, 64-channel atoms one sub-box apart).  D (128 x BN fp32) lives in
// TMEM; the reduction over pixels is split across CTAs and combined with red.global.add.v4.f32.
//
// Replaces the weight half of autograd's convolution_backward for Networks.py:60,87,101,104,122,
// 136,145.  warp roles as in conv_tc.cu.
#include "common.cuh"

namespace {

struct WgradTcArgs {
  int n_img, ho, wo;
  int tw, th, tiles_w, tiles_h;
  int halo, window, kw, cchunks;
  int kh, nsub, n_tiles_per_row, m_tiles;
  int splits, kb_total, kb_per_split, num_items;
  int cout, row_len, kwc_pad, stages;
  uint32_t idesc;
  float* dw;
};

constexpr int kSub = 8192;  // one 64-pixel x 64-channel sub-box
constexpr int kThreads = 256;

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
                const WgradTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int S = p.stages;
  const int bn = p.nsub * 64;
  const uint32_t stage_bytes = (2 + p.nsub) * kSub;
  const uint32_t bar0 = base + S * stage_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S * stage_bytes + (2 * S + 4) * 8);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 + a); };

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * bn) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmDy); tma_prefetch_desc(&tmX); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(smem_u32(tmem_slot), tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  // item -> (split, m_tile, khi, jc0)
  // (m_tile, kh, n-tile) fastest, K-split slowest: CTAs that run at the same time reduce the SAME pixel range
  // for different filter tiles, so dY / X are read from HBM once and re-served from L2
  const int base_items = p.m_tiles * p.kh * p.n_tiles_per_row;
  auto decode = [&](int item, int& split, int& m_tile, int& khi, int& jc0) {
    split = item / base_items;
    int rest = item - split * base_items;
    m_tile = rest % p.m_tiles;
    const int nt = rest / p.m_tiles;
    khi = nt / p.n_tiles_per_row;
    jc0 = (nt - khi * p.n_tiles_per_row) * p.nsub;
  };

  // producer / MMA warps: warp-uniform loops, one elected lane issues (see conv_tc.cu)
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int split, m_tile, khi, jc0;
      decode(item, split, m_tile, khi, jc0);
      const int kb_begin = split * p.kb_per_split;
      const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int img = kb / tiles_per_img, rem = kb - img * tiles_per_img;
        const int h0 = (rem / p.tiles_w) * p.th, w0 = (rem % p.tiles_w) * p.tw;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = base + stage * stage_bytes, sb = sa + 2 * kSub;
        if (elect_one_sync()) {
          mbar_expect_tx(full_bar(stage), stage_bytes);
          // rank-5 maps put the 64-channel block index last, so ONE box load lands all sub-boxes of an operand
          // back to back in the MN-major layout ([block][pixel][64 ch]): 2 TMA instructions per stage, not 6
          tma_load_5d(sa, &tmDy, full_bar(stage), 0, w0 + p.halo, h0 + p.halo, img, m_tile * 2);
          if (p.window) {
            for (int s = 0; s < p.nsub; ++s)
              tma_load_4d(sb + s * kSub, &tmX, full_bar(stage), (jc0 + s) * 64, w0, h0 + khi, img);
          } else {
            tma_load_5d(sb, &tmX, full_bar(stage), 0, w0, h0 + khi, img, jc0);
          }
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int split, m_tile, khi, jc0;
      decode(item, split, m_tile, khi, jc0);
      const int kb_begin = split * p.kb_per_split;
      const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * bn);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = base + stage * stage_bytes, sb = sa + 2 * kSub;
        const uint64_t ad = umma_desc_sw128(sa, kSub, 1024), bd = umma_desc_sw128(sb, kSub, 1024);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)      // K step = 16 pixel rows = 2048 B (descriptor address field is in 16-byte units)
            umma_bf16(d_tmem, ad + 128 * k, bd + 128 * k, p.idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
      if (elect_one_sync()) umma_commit(tfull_bar(as));
      __syncwarp();
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int as = 0; uint32_t aphase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      int split, m_tile, khi, jc0;
      decode(item, split, m_tile, khi, jc0);
      const int co = m_tile * 128 + row;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * bn);
      float* dst = p.dw + static_cast<size_t>(co) * p.row_len + static_cast<size_t>(khi) * p.kwc_pad + jc0 * 64;
      for (int c0 = 0; c0 < bn; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        if (co < p.cout) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            red_add_v4(dst + c0 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                       __uint_as_float(r[j + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace

int vcg_conv_wgrad_tc(const vcg_conv_desc* d, const void* x, const void* dy, int dy_halo, int dy_c, float* dw,
                      cudaStream_t stream) {
  const int ho = d->hp - d->kh + 1, wo = d->wp - d->kw + 1;
  VCG_REQUIRE(d->c % 8 == 0 && d->kwc_pad % 64 == 0 && dy_c % 8 == 0, VCG_E_UNSUPPORTED,
              "wgrad_tc: unsupported channel geometry c=%d kwc_pad=%d dy_c=%d", d->c, d->kwc_pad, dy_c);
  const bool window = (d->c % 64) != 0;
  WgradTcArgs a{};
  a.n_img = d->n; a.ho = ho; a.wo = wo;
  a.tw = wo < 64 ? wo : 64;
  a.th = 64 / a.tw;
  VCG_REQUIRE(a.tw * a.th == 64 && wo % a.tw == 0 && ho % a.th == 0, VCG_E_UNSUPPORTED,
              "wgrad_tc: output %dx%d is not tileable by 64-pixel boxes", ho, wo);
  a.tiles_w = wo / a.tw; a.tiles_h = ho / a.th;
  a.halo = dy_halo; a.window = window ? 1 : 0; a.kw = d->kw;
  a.cchunks = window ? d->kwc_pad / 64 : d->c / 64;
  a.kh = d->kh;
  const int chunks_per_row = d->kwc_pad / 64;
  int nsub = 4;
  while (chunks_per_row % nsub) --nsub;
  a.nsub = nsub;
  a.n_tiles_per_row = chunks_per_row / nsub;
  a.m_tiles = (d->cout + 127) / 128;
  a.kb_total = d->n * a.tiles_w * a.tiles_h;
  const int sms = vcg_gemm_sms();
  const int base_items = a.m_tiles * d->kh * a.n_tiles_per_row;
  // K splits: as many as keep the item count at or below a whole number of waves (a single extra item would
  // add a full wave: 297 items on 148 SMs take 3 rounds, 288 take 2)
  int splits = (2 * sms) / base_items;
  if (splits > a.kb_total / 4) splits = a.kb_total / 4;
  if (splits < 1) splits = 1;
  a.kb_per_split = (a.kb_total + splits - 1) / splits;
  a.splits = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
  a.num_items = base_items * a.splits;
  a.cout = d->cout; a.row_len = d->kh * d->kwc_pad; a.kwc_pad = d->kwc_pad;
  a.idesc = umma_idesc_bf16(128, 64 * nsub, 1, 1);
  a.dw = dw;
  const int stage_bytes = (2 + nsub) * kSub;
  int stages = (227 * 1024 - 2048) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages > a.kb_per_split) stages = a.kb_per_split;
  if (stages < 2) stages = 2;
  a.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 2048;

  CUtensorMap tmDy, tmX;
  const uint64_t es = 2;
  {
    // dY as (64 ch, w, h, n, ch/64): the channel-block index is the LAST (slowest) box dimension
    const uint64_t wpd = wo + 2 * dy_halo, hpd = ho + 2 * dy_halo;
    uint64_t dims[5] = {static_cast<uint64_t>(dy_c < 64 ? dy_c : 64), wpd, hpd, static_cast<uint64_t>(d->n),
                        static_cast<uint64_t>((dy_c + 63) / 64)};
    uint64_t str[4] = {dy_c * es, wpd * dy_c * es, hpd * wpd * dy_c * es, 128};
    uint32_t box[5] = {64, static_cast<uint32_t>(a.tw), static_cast<uint32_t>(a.th), 1, 2};
    int rc = vcg_encode_tmap(&tmDy, dy, 5, dims, str, box, "wgrad_tc dY");
    if (rc) return rc;
  }
  if (window) {
    const uint64_t pix = d->c * es, row = d->wp * pix, img = d->hp * row;
    uint64_t dims[4] = {static_cast<uint64_t>(d->kw) * d->c, static_cast<uint64_t>(wo), static_cast<uint64_t>(d->hp),
                        static_cast<uint64_t>(d->n)};
    uint64_t str[3] = {pix, row, img};
    uint32_t box[4] = {64, static_cast<uint32_t>(a.tw), static_cast<uint32_t>(a.th), 1};
    int rc = vcg_encode_tmap(&tmX, x, 4, dims, str, box, "wgrad_tc X (window)");
    if (rc) return rc;
  } else {
    // X as (64 elems, w positions, h, n, chunk of the (kw, c) run): chunk jc starts jc*128 B into the pixel's window
    const uint64_t pix = d->c * es, row = d->wp * pix, img = d->hp * row;
    uint64_t dims[5] = {64, static_cast<uint64_t>(wo), static_cast<uint64_t>(d->hp), static_cast<uint64_t>(d->n),
                        static_cast<uint64_t>(d->kwc_pad / 64)};
    uint64_t str[4] = {pix, row, img, 128};
    uint32_t box[5] = {64, static_cast<uint32_t>(a.tw), static_cast<uint32_t>(a.th), 1, static_cast<uint32_t>(nsub)};
    int rc = vcg_encode_tmap(&tmX, x, 5, dims, str, box, "wgrad_tc X");
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int grid = a.num_items < sms ? a.num_items : sms;
  wgrad_tc_kernel<<<grid, kThreads, smem, stream>>>(tmDy, tmX, a);
  VCG_CHECK_LAUNCH("wgrad_tc_kernel");
  return VCG_OK;
}
