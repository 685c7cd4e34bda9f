// conv_tc2.cu -- CTA-pair (tcgen05 cta_group::2) variant of the implicit-GEMM convolution of conv_tc.cu for the wide
// layers (64-channel-multiple inputs, BN = 128 / 256), sm_100a.
//
// With both operands in shared memory a single-CTA tcgen05.mma re-reads A (4 KB) and the whole B tile (BN x 32 B) per
// K=16 step; below ~N = 256 that operand traffic, not the tensor pipe, paces the MMA.  A CTA pair
// computes a 256 x BN tile: each CTA stages its own 128 rows of A and only HALF of the filter tile (BN/2 rows); the
// pair's tensor cores share the two B halves, so the per-CTA shared-memory traffic drops from 4 KB + 32*BN to
// 4 KB + 16*BN per step and a stage shrinks from 16 KB + 128*BN to 16 KB + 64*BN bytes (more pipeline stages).
//
// Protocol (the DeepGEMM / CUTLASS 2-SM pattern):
//   * cluster of 2 CTAs; TMEM allocated with cta_group::2 by warp 2 of both CTAs;
//   * both producers issue cta_group::2 TMA loads into their OWN shared memory that complete on the LEADER's full
//     barrier (mbarrier address with the peer bit cleared); the leader expects the bytes of all loads of the pair;
//   * only the leader issues tcgen05.mma.cta_group::2 (M = 256); tcgen05.commit multicasts the arrival to the
//     empty / accumulator-full barriers of both CTAs;
//   * each CTA's epilogue warps drain their own 128 TMEM lanes; "accumulator empty" is collected on the leader's
//     barrier (remote mbarrier.arrive from the second CTA).
//
// Work items.  Every pair walks a list of items; an item is (kind, pair tile, BN).  Three schedules:
//   * plain: contiguous chunks of 256 x BN tiles (m fastest);
//   * tail split (BN = 256): whole rounds of tiles, then the left-over tiles cut into two BN = 128 halves so the last
//     round costs about half a tile (1024-channel layers at 16 x 16: 256 tiles on 74 pairs = 3.46 rounds);
//   * interior + ring (data gradient of a 3 x 3 reflect-padded layer on a small map).  The data gradient is the full
//     correlation of the zero-haloed dY with the flipped filter: (H, W) = (ho + 2, wo + 2) outputs per image.  Tiling
//     that in flat input-pitch order computes 384 rows per 16 x 16 image for 324 outputs, and two thirds of the taps
//     of the 68 border outputs multiply zero rows.  Here the ho x wo interior is tiled exactly (kind 0: 2 tiles per
//     16 x 16 image, all 9 taps) and the border ring is computed by four cheap tile kinds that batch images in the
//     TMA box and run only the 3 taps that can be non-zero: top / bottom row (box 64ch x W x 1 x tn images, taps
//     kh = 2 / kh = 0) and left / right column (box 64ch x 1 x ho x tn, taps kw = 2 / kw = 0).
// Epilogue: bias + activation + InstanceNorm sum / sum-of-squares + bf16 NHWC store (direct, or staged through shared
// memory at BN = 128; see conv_tc.cu).
#include <stdlib.h>

#include "common.cuh"

namespace {

struct Conv2Args {
  int n_img, ho, wo, wp;                        // ho, wo: output rows / columns of the convolution
  int tw, th, tiles_w, tiles_h;
  int flat, kh, kw, cchunks;
  int bn, cout, out_c, act, stats, epi2;
  int num_m_tiles, num_pair_tiles, stages;      // pair tile = two consecutive m tiles x one n tile
  int tail_r, full_per_pair;                    // tail split
  int ring;                                     // 1: interior + ring schedule (num_pair_tiles = interior pair tiles)
  int tn_tb, tn_lr, p_tb, p_lr, ring_pairs;     // images per ring box; pair tiles per ring side (x n tiles = ring_pairs)
  uint32_t idesc, idesc_half, a_tx_bytes, tb_tx_bytes, lr_tx_bytes;
  const float* bias;
  float* stats_acc;
  void* out;
};

// one work item as seen by this CTA of the pair
struct Item {
  int kind;                 // 0: convolution tile; 1..4: ring top / bottom / left / right
  int bn, n0;               // output columns of this item
  int img, h0, w0;          // kind 0: image and tile origin; ring: first image of the box (h0 = w0 = 0)
  int ok;                   // this CTA's m tile exists (odd counts: the last pair has one tile)
  int kh_lo, kh_hi, kw_lo, kw_hi;   // taps that contribute
  int ch, cw;               // input-coordinate origin of the A box for tap (0, 0)
  uint32_t a_bytes;
};

constexpr int kAStageBytes = 16384;
constexpr int kThreads = 384;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;     // shared::cluster address of the even (leader) CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
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
// arrive on the barrier at the same shared-memory offset in BOTH CTAs when all previously issued MMAs are done
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(static_cast<uint16_t>(3)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmTB, const __grid_constant__ CUtensorMap tmLR, const Conv2Args p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int S = p.stages;
  const uint32_t stage_bytes = kAStageBytes + static_cast<uint32_t>(p.bn / 2) * 128u;   // A tile + this CTA's half of B
  const uint32_t bar0 = base + S * stage_bytes;            // full[S], empty[S], tfull[2], tempty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S * stage_bytes + (2 * S + 4) * 8);
  const uint32_t bias0 = bar0 + 1024u;
  const uint32_t epi0 = bias0 + 8192u;                     // staged epilogue: 128 x BN bf16 tile, row table, slab
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 + a); };

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int rank = static_cast<int>(uniform_u32(cluster_ctarank()));
  const bool leader = rank == 0;
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * p.bn) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    if (p.ring) { tma_prefetch_desc(&tmTB); tma_prefetch_desc(&tmLR); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }      // full: the leader's expect-tx arrive
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 16); }   // tempty: 8 warps x 2 CTAs
    fence_mbar_init();
  }
  cluster_sync_all();                                       // barrier inits visible to the peer before any remote use
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int m_pairs = (p.num_m_tiles + 1) / 2;
  const int npairs = gridDim.x / 2, pair = blockIdx.x / 2;
  const int per_pair = (p.num_pair_tiles + npairs - 1) / npairs;
  const int tile_begin = pair * per_pair;
  const int tile_end = min(p.num_pair_tiles, tile_begin + per_pair);
  const int ring_full = p.num_pair_tiles / npairs, ring_left = p.num_pair_tiles % npairs;
  const int ring_sides = 2 * p.p_tb + 2 * p.p_lr;           // ring pair tiles per n tile

  // convolution tile t (m pair, n tile; m fastest) as seen by this CTA
  auto conv_item = [&](int t, int bn_cur, int nsub, Item& it) {
    const int mp = t % m_pairs, n_tile = t / m_pairs;
    const int m_tile = 2 * mp + rank;
    it.kind = 0; it.bn = bn_cur; it.n0 = n_tile * p.bn + nsub * 128;
    it.ok = m_tile < p.num_m_tiles;
    it.img = m_tile / tiles_per_img;
    const int rem = m_tile - it.img * tiles_per_img;
    it.h0 = (rem / p.tiles_w) * p.th; it.w0 = (rem % p.tiles_w) * p.tw;
    it.kh_lo = 0; it.kh_hi = p.kh; it.kw_lo = 0; it.kw_hi = p.kw;
    it.a_bytes = p.a_tx_bytes;
    if (p.ring) { it.ch = it.h0 + 1; it.cw = it.w0 + 1; }              // interior of the (ho + 2) x (wo + 2) gradient
    else if (p.flat) { it.ch = 0; it.cw = it.w0; }
    else { it.ch = it.h0; it.cw = it.w0; }
  };
  // ring pair tile j: (n tile, side, image group)
  auto ring_item = [&](int j, Item& it) {
    const int n_tile = j / ring_sides;
    const int r = j - n_tile * ring_sides;
    int kind, g;
    if (r < p.p_tb) { kind = 1; g = r; }
    else if (r < 2 * p.p_tb) { kind = 2; g = r - p.p_tb; }
    else if (r < 2 * p.p_tb + p.p_lr) { kind = 3; g = r - 2 * p.p_tb; }
    else { kind = 4; g = r - 2 * p.p_tb - p.p_lr; }
    const int tn = kind <= 2 ? p.tn_tb : p.tn_lr;
    it.kind = kind; it.bn = p.bn; it.n0 = n_tile * p.bn;
    it.img = (2 * g + rank) * tn; it.h0 = 0; it.w0 = 0;
    it.ok = it.img < p.n_img;
    it.kh_lo = 0; it.kh_hi = p.kh; it.kw_lo = 0; it.kw_hi = p.kw;
    it.ch = 0; it.cw = 0;
    // gradient row 0 only sees dY row 0 (tap kh = 2), row H-1 only dY's last row (tap kh = 0); same for the columns
    if (kind == 1) { it.kh_lo = p.kh - 1; }
    else if (kind == 2) { it.kh_hi = 1; it.ch = p.ho + 1; }
    else if (kind == 3) { it.kw_lo = p.kw - 1; it.ch = 1; }
    else { it.kw_hi = 1; it.ch = 1; it.cw = p.wo + 1; }
    it.a_bytes = kind <= 2 ? p.tb_tx_bytes : p.lr_tx_bytes;
  };
  // i-th work item of this pair
  auto work = [&](int i, Item& it) -> bool {
    if (p.ring) {
      if (i < ring_full) { conv_item(pair * ring_full + i, p.bn, 0, it); return true; }
      int k = i - ring_full;
      if (pair < ring_left) { if (k == 0) { conv_item(npairs * ring_full + pair, p.bn, 0, it); return true; } --k; }
      // ring tiles (3 of 9 taps: a third of a tile) first fill the pairs without a left-over convolution tile, three
      // each, then go round-robin over all pairs
      const int spread = npairs - ring_left;
      const int lim1 = ring_left > 0 ? min(p.ring_pairs, 3 * spread) : 0;
      if (ring_left > 0 && pair >= ring_left) {
        const int first = pair - ring_left;
        const int n1 = first < lim1 ? (lim1 - first + spread - 1) / spread : 0;
        if (k < n1) { ring_item(first + k * spread, it); return true; }
        k -= n1;
      }
      const int j = lim1 + pair + k * npairs;
      if (j >= p.ring_pairs) return false;
      ring_item(j, it);
      return true;
    }
    if (!p.tail_r) { const int t = tile_begin + i; if (t >= tile_end) return false; conv_item(t, p.bn, 0, it); return true; }
    if (i < p.full_per_pair) { conv_item(pair * p.full_per_pair + i, p.bn, 0, it); return true; }
    if (i == p.full_per_pair && pair < 2 * p.tail_r) { conv_item(npairs * p.full_per_pair + (pair >> 1), 128, pair & 1, it); return true; }
    return false;
  };
  // accumulator row -> output pixel
  auto rowmap = [&](const Item& it, int row, size_t& pix) -> bool {
    int img = it.img, h, w; bool valid;
    if (it.kind == 0) {
      if (p.flat && !p.ring) { const int f = it.w0 + row; h = f / p.wp; w = f - h * p.wp; valid = (h < p.ho) && (w < p.wo); }
      else { const int hh = row / p.tw; h = it.h0 + hh; w = it.w0 + (row - hh * p.tw);
             valid = (row < p.tw * p.th) && (h < p.ho) && (w < p.wo); }
      valid = valid && it.ok;
      if (p.ring) { pix = (static_cast<size_t>(img) * (p.ho + 2) + h + 1) * (p.wo + 2) + w + 1; return valid; }
      pix = (static_cast<size_t>(img) * p.ho + h) * p.wo + w;
      return valid;
    }
    const int H = p.ho + 2, W = p.wo + 2;
    int nl;
    if (it.kind <= 2) { nl = row / W; w = row - nl * W; h = it.kind == 1 ? 0 : H - 1; valid = nl < p.tn_tb; }
    else { nl = row / p.ho; h = 1 + row - nl * p.ho; w = it.kind == 3 ? 0 : W - 1; valid = nl < p.tn_lr; }
    img += nl;
    pix = (static_cast<size_t>(img) * H + h) * W + w;
    return valid && img < p.n_img;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0; uint32_t phase = 0;
    for (int i = 0;; ++i) {
      Item it;
      if (!work(i, it)) break;
      const uint32_t bh_cur = static_cast<uint32_t>(it.bn / 2) * 128u;
      const CUtensorMap* mapA = it.kind == 0 ? &tmA : (it.kind <= 2 ? &tmTB : &tmLR);
      const int nrow = it.n0 + rank * (it.bn / 2);
      for (int khi = it.kh_lo; khi < it.kh_hi; ++khi)
        for (int kwi = it.kw_lo; kwi < it.kw_hi; ++kwi)
          for (int q = 0; q < p.cchunks; ++q) {
            const int kb = (khi * p.kw + kwi) * p.cchunks + q;          // k-block of the packed filter
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = base + stage * stage_bytes, sb = sa + kAStageBytes;
            const uint32_t lbar = full_bar(stage) & kPeerMask;           // the LEADER's full barrier
            if (elect_one_sync()) {
              // the leader expects the bytes of all loads of the pair; the second CTA's loads only complete_tx there
              if (leader) mbar_expect_tx(full_bar(stage), 2u * (it.a_bytes + bh_cur));
              if (p.flat && !p.ring) tma2_load_4d(sa, mapA, lbar, q * 64, it.cw + khi * p.wp + kwi, 0, it.img);
              else                   tma2_load_4d(sa, mapA, lbar, q * 64, it.cw + kwi, it.ch + khi, it.img);
              // filter rows in 64-row boxes: this CTA's half of the n-tile is one (BN = 128) or two (BN = 256) of them
              tma2_load_2d(sb, &tmB, lbar, kb * 64, nrow);
              if (it.bn == 256) tma2_load_2d(sb + 8192, &tmB, lbar, kb * 64, nrow + 64);
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
      for (int i = 0;; ++i) {
        Item it;
        if (!work(i, it)) break;
        const uint32_t idesc = it.bn == p.bn ? p.idesc : p.idesc_half;
        const int nkb = (it.kh_hi - it.kh_lo) * (it.kw_hi - it.kw_lo) * p.cchunks;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.bn);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * stage_bytes, sb = sa + kAStageBytes;
          const uint64_t ad = umma_desc_sw128(sa, 16, 1024), bd = umma_desc_sw128(sb, 16, 1024);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma2_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
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
  } else if (warp >= 4 && p.epi2) {
    // ===================== staged epilogue (BN = 128; see conv_tc.cu): TMEM -> bias/act -> bf16 tile in shared memory
    // (16-byte chunks XOR-swizzled by row), then coalesced stores and column sums over the stored values
    const int et = static_cast<int>(threadIdx.x) - 128;             // 0..255
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const int row_bytes = p.bn * 2, cpr = row_bytes >> 4;
    uint8_t* stg = smem + (epi0 - base);
    long long* rowinfo = reinterpret_cast<long long*>(stg + 128 * row_bytes);
    float* red = reinterpret_cast<float*>(rowinfo + 128);
    auto swz = [&](int k, int r) { return (k & ~7) | ((k ^ r) & 7); };
    const int pairs = p.bn >> 1, groups = 256 / pairs, rpg = 128 / groups;
    int as = 0; uint32_t aphase = 0;
    float* sb = reinterpret_cast<float*>(smem + (bias0 - base)) + (warp - 4) * 256;
    int cur_n0 = -1;
    float run = 0.f;
    int run_img = -1, run_n0 = 0;
    auto flush_stats = [&]() {
      if (run_img >= 0 && et < 2 * p.bn) {
        const int ch = run_n0 + (et % p.bn);
        if (ch < p.cout) atomicAdd(p.stats_acc + (static_cast<size_t>(run_img) * p.cout + ch) * 2 + et / p.bn, run);
      }
      run = 0.f;
    };
    for (int i = 0;; ++i) {
      Item it;
      if (!work(i, it)) break;
      const int n0 = it.n0;
      if (p.stats && it.ok && (it.img != run_img || n0 != run_n0)) { flush_stats(); run_img = it.img; run_n0 = n0; }
      size_t pix;
      const bool valid = rowmap(it, row, pix);
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * p.bn);
      uint8_t* srow = stg + row * row_bytes;
      if (n0 != cur_n0) { load_bias_tile(sb, p.bias, n0, p.bn, p.cout, lane); cur_n0 = n0; }
#pragma unroll 1
      for (int c0 = half * 32; c0 < p.bn; c0 += 64) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        bias_act32(v, sb + c0, p.act, valid ? p.cout - (n0 + c0) : 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 pk;
          __nv_bfloat162* hp2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) hp2[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
          *reinterpret_cast<uint4*>(srow + swz(c0 / 8 + j, row) * 16) = pk;
        }
      }
      if (half == 0) rowinfo[row] = valid ? static_cast<long long>(pix) : -1;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int idx = et; idx < 128 * cpr; idx += 256) {
        const int r2 = idx / cpr, k = idx - r2 * cpr;
        const long long pi = rowinfo[r2];
        const int col0 = n0 + k * 8;
        if (pi >= 0 && col0 < p.cout) {
          const uint4 val = *reinterpret_cast<const uint4*>(stg + r2 * row_bytes + swz(k, r2) * 16);
          *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out) + (static_cast<size_t>(pi) * p.out_c + col0) * 2) = val;
        }
      }
      if (p.stats) {
        const int cp = et % pairs, g = et / pairs;
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
        for (int r2 = g * rpg; r2 < (g + 1) * rpg; ++r2) {
          const uint32_t wv = *reinterpret_cast<const uint32_t*>(stg + r2 * row_bytes + swz(cp >> 2, r2) * 16 + (cp & 3) * 4);
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv));
          s1a += f.x; s1b += f.y; s2a = fmaf(f.x, f.x, s2a); s2b = fmaf(f.y, f.y, s2b);
        }
        red[(g * 2 + 0) * p.bn + 2 * cp] = s1a; red[(g * 2 + 0) * p.bn + 2 * cp + 1] = s1b;
        red[(g * 2 + 1) * p.bn + 2 * cp] = s2a; red[(g * 2 + 1) * p.bn + 2 * cp + 1] = s2b;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (p.stats && et < 2 * p.bn) {
        const int which = et / p.bn, ch = et % p.bn;
        float tsum = 0.f;
        for (int g = 0; g < groups; ++g) tsum += red[(g * 2 + which) * p.bn + ch];
        run += tsum;
      }
    }
    if (p.stats) flush_stats();
  } else if (warp >= 4) {
    // ===================== direct epilogue (both CTAs, own 128 TMEM lanes) =====================
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    int as = 0; uint32_t aphase = 0;
    float run1[8], run2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) run1[i] = run2[i] = 0.f;
    int run_img = -1, run_n0 = 0;
    float* sb = reinterpret_cast<float*>(smem + (bias0 - base)) + (warp - 4) * 256;
    int cur_n0 = -1;
    auto flush_stats = [&]() {
      if (run_img >= 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int col = run_n0 + i * 32 + lane;
          if ((i & 1) == half && i * 32 < p.bn && col < p.cout) {
            float* dst = p.stats_acc + (static_cast<size_t>(run_img) * p.cout + col) * 2;
            atomicAdd(dst, run1[i]);
            atomicAdd(dst + 1, run2[i]);
          }
          run1[i] = run2[i] = 0.f;
        }
      }
    };
    for (int i = 0;; ++i) {
      Item it;
      if (!work(i, it)) break;
      const int n0 = it.n0;
      if (p.stats && it.ok && (it.img != run_img || n0 != run_n0)) { flush_stats(); run_img = it.img; run_n0 = n0; }
      size_t pix;
      const bool valid = rowmap(it, row, pix);
      if (n0 != cur_n0) { load_bias_tile(sb, p.bias, n0, it.bn, p.cout, lane); cur_n0 = n0; }
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * p.bn);
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const int c0 = ci * 32;
        if (c0 >= it.bn) break;
        if ((ci & 1) != half) continue;
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const int col0 = n0 + c0;
        bias_act32(v, sb + c0, p.act, p.cout - col0);
        if (valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            if (col < p.cout) {
              float tt[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) tt[j] = v[g * 8 + j];
              st8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_c + col, tt);
            }
          }
        }
        if (p.stats && col0 < p.cout) {
          float s1[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { const float x = valid ? v[j] : 0.f; s1[j] = x; s2[j] = x * x; }
          transposed_warp_sum32(s1, lane);
          transposed_warp_sum32(s2, lane);
          run1[ci] += s1[0];
          run2[ci] += s2[0];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(as));      // both CTAs report to the leader's barrier
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (p.stats) flush_stats();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

void pick_box2(int wo, int ho, int* tw, int* th) {
  if (wo >= 128) { *tw = 128; *th = 1; return; }
  *tw = wo;
  int t = 128 / wo;
  if (t > ho) t = ho;
  if (t < 1) t = 1;
  *th = t;
}

// relative cost of a launch in "rounds of BN = 256 tiles": whole rounds + the tail (split into BN = 128 halves when
// they fit); a BN = 128 tile costs ~0.69 of a BN = 256 one (its MMAs are bound by shared-memory operand reads)
double rounds_cost(int tiles256, int max_pairs, bool bn256) {
  if (bn256) {
    const int np = tiles256 < max_pairs ? tiles256 : max_pairs;
    const int full = tiles256 / np, rem = tiles256 % np;
    return full + (rem == 0 ? 0.0 : (2 * rem <= np ? 0.69 : 1.0));
  }
  const int t = 2 * tiles256, np = t < max_pairs ? t : max_pairs;
  return 0.69 * ((t + np - 1) / np);
}

}  // namespace

bool vcg_conv2_supported(const vcg_conv_desc* d, int out_f32) {
  if (out_f32) return false;
  if (d->c % 64 != 0 || d->kwc_pad != d->kw * d->c) return false;
  // 128-channel-multiple outputs; 64 output channels only for data gradients (no statistics: the thin-N epilogue of
  // conv_tc.cu is the better forward path at N = 64)
  const bool n64 = d->cout_pad == 64 && d->flat != 0 && !d->stats;
  if ((d->cout_pad % 128 != 0 && !n64) || d->cout != d->cout_pad || d->out_c != d->cout) return false;
  const int kblocks = d->kh * (d->kwc_pad / 64);
  return kblocks >= 8;
}

int vcg_conv_fwd_tc2(const vcg_conv_desc* d, const void* x, const void* w, const float* bias, void* y, float* stats,
                     cudaStream_t stream) {
  int ho = d->hp - d->kh + 1, wo = d->wp - d->kw + 1;
  const int sms = vcg_gemm_sms();
  Conv2Args a{};
  // interior + ring data gradient: flat == 2 promises a zero halo of (kh - 1, kw - 1) around dY; 3 x 3 filters on maps
  // whose interior tiles exactly (16 x 16, 32 x 32, 64 x 64), no statistics
  const int hi = ho - 2, wi = wo - 2;                       // interior = the forward layer's input map
  a.ring = (d->flat == 2 && d->kh == 3 && d->kw == 3 && !(d->stats && stats) && (wi == 16 || wi == 32 || wi == 64) &&
            hi >= 8 && hi <= 128 && 128 % hi == 0 && hi % (128 / wi) == 0) ? 1 : 0;
  a.n_img = d->n; a.wp = d->wp;
  a.flat = d->flat ? 1 : 0; a.kh = d->kh; a.kw = d->kw;
  a.cchunks = d->c / 64;
  const int kblocks = d->kh * (d->kwc_pad / 64);
  if (a.ring) {
    ho = hi; wo = wi;                                       // the kernel's (ho, wo) is the interior
    a.tw = wi; a.th = 128 / wi; a.tiles_w = 1; a.tiles_h = hi / a.th;
    a.a_tx_bytes = 128 * 128;
    a.tn_tb = 128 / (wi + 2); a.tn_lr = 128 / hi;
    a.tb_tx_bytes = static_cast<uint32_t>(a.tn_tb * (wi + 2)) * 128u;
    a.lr_tx_bytes = static_cast<uint32_t>(a.tn_lr * hi) * 128u;
    a.p_tb = ((d->n + a.tn_tb - 1) / a.tn_tb + 1) / 2;
    a.p_lr = ((d->n + a.tn_lr - 1) / a.tn_lr + 1) / 2;
  } else if (a.flat) {
    a.tw = 128; a.th = 1; a.tiles_h = 1;
    a.tiles_w = (ho * d->wp + 127) / 128;
    a.a_tx_bytes = 128 * 128;
  } else {
    pick_box2(wo, ho, &a.tw, &a.th);
    a.tiles_w = (wo + a.tw - 1) / a.tw;
    a.tiles_h = (ho + a.th - 1) / a.th;
    a.a_tx_bytes = static_cast<uint32_t>(a.tw * a.th) * 128u;
  }
  a.ho = ho; a.wo = wo;
  a.num_m_tiles = d->n * a.tiles_w * a.tiles_h;
  const int m_pairs = (a.num_m_tiles + 1) / 2;
  int bn = d->cout_pad % 256 == 0 ? 256 : (d->cout_pad == 64 ? 64 : 128);
  if (bn == 256) {
    const int t256 = m_pairs * (d->cout_pad / 256);
    if (rounds_cost(t256, sms / 2, false) < rounds_cost(t256, sms / 2, true)) bn = 128;
  }
  a.bn = bn;
  const int ntn = d->cout_pad / bn;
  a.num_pair_tiles = m_pairs * ntn;
  a.ring_pairs = a.ring ? (2 * a.p_tb + 2 * a.p_lr) * ntn : 0;
  a.cout = d->cout; a.out_c = d->out_c; a.act = d->act; a.stats = (d->stats && stats) ? 1 : 0;
  a.bias = bias; a.stats_acc = stats; a.out = y;
  a.idesc = umma_idesc_bf16(256, bn, 0, 0);
  a.idesc_half = umma_idesc_bf16(256, 128, 0, 0);
  a.epi2 = bn == 128 ? 1 : 0;
  const int epi_bytes = a.epi2 ? 128 * bn * 2 + 1024 + 4096 : 0;
  const int stage_bytes = kAStageBytes + (bn / 2) * 128;
  int stages = (227 * 1024 - 3072 - 8192 - epi_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages > kblocks) stages = kblocks;
  a.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 3072 + 8192 + epi_bytes;

  CUtensorMap tmA, tmB;
  const uint64_t es = 2;
  const uint64_t pix_stride = static_cast<uint64_t>(d->c) * es, row_stride = d->wp * pix_stride, img_stride = d->hp * row_stride;
  uint64_t dims[4], strides[3];
  uint32_t box[4];
  dims[0] = d->c;
  if (a.flat && !a.ring) {
    dims[1] = static_cast<uint64_t>(d->hp) * d->wp; dims[2] = 1; dims[3] = d->n;
    strides[0] = pix_stride; strides[1] = img_stride; strides[2] = img_stride;
    box[0] = 64; box[1] = 128; box[2] = 1; box[3] = 1;
  } else {
    dims[1] = d->wp; dims[2] = d->hp; dims[3] = d->n;
    strides[0] = pix_stride; strides[1] = row_stride; strides[2] = img_stride;
    box[0] = 64; box[1] = a.tw; box[2] = a.th; box[3] = 1;
  }
  int rc = vcg_encode_tmap(&tmA, x, 4, dims, strides, box, "conv_tc2 A");
  if (rc) return rc;
  CUtensorMap tmTB = tmA, tmLR = tmA;
  if (a.ring) {
    uint32_t btb[4] = {64, static_cast<uint32_t>(wi + 2), 1, static_cast<uint32_t>(a.tn_tb)};
    rc = vcg_encode_tmap(&tmTB, x, 4, dims, strides, btb, "conv_tc2 A (ring rows)");
    if (rc) return rc;
    uint32_t blr[4] = {64, 1, static_cast<uint32_t>(hi), static_cast<uint32_t>(a.tn_lr)};
    rc = vcg_encode_tmap(&tmLR, x, 4, dims, strides, blr, "conv_tc2 A (ring columns)");
    if (rc) return rc;
  }
  const uint64_t ktot = static_cast<uint64_t>(d->kh) * d->kwc_pad;
  uint64_t bdims[2] = {ktot, static_cast<uint64_t>(d->cout_pad)};
  uint64_t bstr[1] = {ktot * es};
  uint32_t bbox[2] = {64, bn == 64 ? 32u : 64u};            // filter rows per box: this CTA's half at BN <= 128, two boxes at BN = 256
  rc = vcg_encode_tmap(&tmB, w, 2, bdims, bstr, bbox, "conv_tc2 B");
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "conv_tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int grid = (sms / 2) * 2;
  if (grid > 2 * a.num_pair_tiles) grid = 2 * a.num_pair_tiles;
  if (!a.ring) {
    // wave quantisation: run the whole rounds as BN = 256 tiles and cut the left-over tiles into two BN = 128 halves
    const int npairs = grid / 2, full = a.num_pair_tiles / npairs, rem = a.num_pair_tiles % npairs;
    if (bn == 256 && full >= 1 && rem > 0 && 2 * rem <= npairs) { a.tail_r = rem; a.full_per_pair = full; }
  }
  conv_tc2_kernel<<<grid, kThreads, smem, stream>>>(tmA, tmB, tmTB, tmLR, a);
  VCG_CHECK_LAUNCH("conv_tc2_kernel");
  return VCG_OK;
}
