// wgrad_tc2.cu -- CTA-pair (tcgen05 cta_group::2) variant of the weight-gradient GEMM of wgrad_tc.cu for layers with a
// multiple of 256 output channels and 64-channel-multiple inputs, sm_100a.
//
//   dW[co, kh, j] += sum_{n,h,w} dY[n,h,w,co] * X[n, h+kh, w*c + j]
//
// Both operands are MN-major (the reduction runs over pixels; the NHWC channel axis is M for dY and N for X).  A single
// CTA reads 4 KB of dY and 8 KB of X from shared memory per K=16 step of a 128 x 256 tile, which paces the MMA below
// the tensor pipe's rate.  A CTA pair computes a 256 x 256 tile: each CTA stages its own 128 output channels of dY and
// only HALF of the X columns (128 of 256), so a step reads 4 + 4 KB per CTA and a stage is 32 KB instead of 48 KB.
// Protocol and warp roles as in conv_tc2.cu (leader issues the MMAs, loads complete on the leader's barrier, commits
// are multicast to both CTAs, each CTA reduces its own 128 TMEM lanes into dW with red.global.add.v4.f32).
#include <stdlib.h>

#include "common.cuh"

namespace {

struct Wgrad2Args {
  int tw, th, tiles_w, tiles_h;
  int halo, kh, n_tiles_per_row, m_tiles;       // m tile = 256 output channels
  int splits, kb_total, kb_per_split, num_items;
  int cout, row_len, kwc_pad, stages;
  uint32_t idesc;
  float* dw;
};

constexpr int kSub = 8192;            // one 64-pixel x 64-channel sub-box
constexpr int kStage = 4 * kSub;      // dY: 2 sub-boxes (128 couts), X: 2 sub-boxes (this CTA's 128 of 256 columns)
constexpr int kThreads = 256;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
wgrad_tc2_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX, const Wgrad2Args p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int S = p.stages;
  const uint32_t bar0 = base + S * kStage;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S * kStage + (2 * S + 4) * 8);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 + a); };

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int rank = static_cast<int>(uniform_u32(cluster_ctarank()));
  const bool leader = rank == 0;
  constexpr uint32_t tmem_cols = 512;           // two 256-column accumulators

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmDy); tma_prefetch_desc(&tmX); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }     // 4 warps x 2 CTAs
    fence_mbar_init();
  }
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int npairs = gridDim.x / 2, pair = blockIdx.x / 2;

  // item -> (split, m_tile, khi, jc0): (m_tile, kh, n-tile) fastest, K-split slowest, so pairs that run at the same
  // time reduce the SAME pixel range for different filter tiles (dY / X re-served from L2)
  const int base_items = p.m_tiles * p.kh * p.n_tiles_per_row;
  auto decode = [&](int item, int& split, int& m_tile, int& khi, int& jc0) {
    split = item / base_items;
    const int rest = item - split * base_items;
    m_tile = rest % p.m_tiles;
    const int nt = rest / p.m_tiles;
    khi = nt / p.n_tiles_per_row;
    jc0 = (nt - khi * p.n_tiles_per_row) * 4;
  };

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int item = pair; item < p.num_items; item += npairs) {
      int split, m_tile, khi, jc0;
      decode(item, split, m_tile, khi, jc0);
      const int kb_begin = split * p.kb_per_split;
      const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        const int img = kb / tiles_per_img, rem = kb - img * tiles_per_img;
        const int h0 = (rem / p.tiles_w) * p.th, w0 = (rem % p.tiles_w) * p.tw;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = base + stage * kStage, sb = sa + 2 * kSub;
        const uint32_t lbar = full_bar(stage) & kPeerMask;
        if (elect_one_sync()) {
          if (leader) mbar_expect_tx(full_bar(stage), 2u * kStage);
          // this CTA's 128 output channels of dY and its 128 of the 256 filter-row columns of X
          tma2_load_5d(sa, &tmDy, lbar, 0, w0 + p.halo, h0 + p.halo, img, m_tile * 4 + rank * 2);
          tma2_load_5d(sb, &tmX, lbar, 0, w0, h0 + khi, img, jc0 + rank * 2);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
      for (int item = pair; item < p.num_items; item += npairs) {
        int split, m_tile, khi, jc0;
        decode(item, split, m_tile, khi, jc0);
        const int kb_begin = split * p.kb_per_split;
        const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * 256);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * kStage, sb = sa + 2 * kSub;
          const uint64_t ad = umma_desc_sw128(sa, kSub, 1024), bd = umma_desc_sw128(sb, kSub, 1024);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)      // K step = 16 pixel rows = 2048 B
              umma2_bf16(d_tmem, ad + 128 * k, bd + 128 * k, p.idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            umma2_commit_both(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if (elect_one_sync()) umma2_commit_both(tfull_bar(as));
        __syncwarp();
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // Epilogue: dW += accumulator (red.global.add.v4.f32).  A thread owns one accumulator ROW (output channel), so a
    // straight drain makes every red instruction hit 32 different filter rows (32 separate 16-byte atomics).  Each
    // 32 x 32 block goes through a warp-private shared-memory tile instead, so that one instruction covers four rows
    // x 128 contiguous bytes (4 cache lines instead of 32 sectors): at a per-GPU batch of 8 the reduction length is 32
    // k-blocks and this epilogue, not the MMAs, paced the kernel.
    const int quad = warp & 3;
    float* tile = reinterpret_cast<float*>(smem + S * kStage + 2048) + (warp - 4) * (32 * 36);
    const int rr = lane >> 3, cc = (lane & 7) * 4;
    int as = 0; uint32_t aphase = 0;
    for (int item = pair; item < p.num_items; item += npairs) {
      int split, m_tile, khi, jc0;
      decode(item, split, m_tile, khi, jc0);
      const int co0 = m_tile * 256 + rank * 128 + quad * 32;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * 256);
      float* dst0 = p.dw + static_cast<size_t>(co0) * p.row_len + static_cast<size_t>(khi) * p.kwc_pad + jc0 * 64;
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        __syncwarp();                                   // the previous block's reads of the tile are done
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(tile + lane * 36 + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = it * 4 + rr;
          if (co0 + row < p.cout) {
            const float4 v = *reinterpret_cast<const float4*>(tile + row * 36 + cc);
            red_add_v4(dst0 + static_cast<size_t>(row) * p.row_len + c0 + cc, v.x, v.y, v.z, v.w);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

}  // namespace

bool vcg_wgrad2_supported(const vcg_conv_desc* d, int dy_c) {
  if (d->c % 64 != 0 || d->kwc_pad != d->kw * d->c) return false;
  if (d->cout % 256 != 0 || dy_c != d->cout) return false;
  if ((d->kwc_pad / 64) % 4 != 0) return false;
  const int ho = d->hp - d->kh + 1, wo = d->wp - d->kw + 1;
  const int tw = wo < 64 ? wo : 64;
  if (tw <= 0 || 64 % tw != 0) return false;
  const int th = 64 / tw;
  return wo % tw == 0 && ho % th == 0;
}

int vcg_conv_wgrad_tc2(const vcg_conv_desc* d, const void* x, const void* dy, int dy_halo, int dy_c, float* dw,
                       cudaStream_t stream) {
  const int ho = d->hp - d->kh + 1, wo = d->wp - d->kw + 1;
  Wgrad2Args a{};
  a.tw = wo < 64 ? wo : 64;
  a.th = 64 / a.tw;
  a.tiles_w = wo / a.tw; a.tiles_h = ho / a.th;
  a.halo = dy_halo; a.kh = d->kh;
  a.n_tiles_per_row = (d->kwc_pad / 64) / 4;
  a.m_tiles = d->cout / 256;
  a.kb_total = d->n * a.tiles_w * a.tiles_h;
  const int sms = vcg_gemm_sms();
  const int npairs_max = sms / 2;
  const int base_items = a.m_tiles * d->kh * a.n_tiles_per_row;
  // K splits: as many as keep the item count at or below a whole number of rounds of the CTA pairs
  int splits = (2 * npairs_max) / base_items;
  if (splits > a.kb_total / 4) splits = a.kb_total / 4;
  if (splits < 1) splits = 1;
  a.kb_per_split = (a.kb_total + splits - 1) / splits;
  a.splits = (a.kb_total + a.kb_per_split - 1) / a.kb_per_split;
  a.num_items = base_items * a.splits;
  a.cout = d->cout; a.row_len = d->kh * d->kwc_pad; a.kwc_pad = d->kwc_pad;
  a.idesc = umma_idesc_bf16(256, 256, 1, 1);
  a.dw = dw;
  constexpr int kEpiBytes = 4 * 32 * 36 * 4;            // warp-private transpose tiles of the epilogue
  int stages = (227 * 1024 - 2048 - kEpiBytes) / kStage;
  if (stages > 6) stages = 6;
  if (stages > a.kb_per_split) stages = a.kb_per_split;
  if (stages < 2) stages = 2;
  a.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * kStage + 2048 + kEpiBytes;

  CUtensorMap tmDy, tmX;
  const uint64_t es = 2;
  {
    // dY as (64 ch, w, h, n, ch/64): the channel-block index is the LAST (slowest) box dimension
    const uint64_t wpd = wo + 2 * dy_halo, hpd = ho + 2 * dy_halo;
    uint64_t dims[5] = {64, wpd, hpd, static_cast<uint64_t>(d->n), static_cast<uint64_t>(dy_c / 64)};
    uint64_t str[4] = {dy_c * es, wpd * dy_c * es, hpd * wpd * dy_c * es, 128};
    uint32_t box[5] = {64, static_cast<uint32_t>(a.tw), static_cast<uint32_t>(a.th), 1, 2};
    int rc = vcg_encode_tmap(&tmDy, dy, 5, dims, str, box, "wgrad_tc2 dY");
    if (rc) return rc;
  }
  {
    // X as (64 elems, w positions, h, n, chunk of the (kw, c) run): chunk jc starts jc*128 B into the pixel's window
    const uint64_t pix = d->c * es, row = d->wp * pix, img = d->hp * row;
    uint64_t dims[5] = {64, static_cast<uint64_t>(wo), static_cast<uint64_t>(d->hp), static_cast<uint64_t>(d->n),
                        static_cast<uint64_t>(d->kwc_pad / 64)};
    uint64_t str[4] = {pix, row, img, 128};
    uint32_t box[5] = {64, static_cast<uint32_t>(a.tw), static_cast<uint32_t>(a.th), 1, 2};
    int rc = vcg_encode_tmap(&tmX, x, 5, dims, str, box, "wgrad_tc2 X");
    if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "wgrad_tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int grid = npairs_max * 2;
  if (grid > 2 * a.num_items) grid = 2 * a.num_items;
  wgrad_tc2_kernel<<<grid, kThreads, smem, stream>>>(tmDy, tmX, a);
  VCG_CHECK_LAUNCH("wgrad_tc2_kernel");
  return VCG_OK;
}
