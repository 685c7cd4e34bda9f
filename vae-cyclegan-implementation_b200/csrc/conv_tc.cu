// conv_tc.cu -- implicit-GEMM convolution on 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// One persistent, warp-specialised kernel serves the forward pass and the data-gradient pass of
// every convolution of the reference (Networks.py:60,87,101,104,122,136,145 executed through
// torch/nn/modules/conv.py:534-550 and autograd's convolution_backward):
//
//   D[pixel, cout] = sum_{kh} sum_{j < kw*c} X[n, h+kh, (w*c + j)] * W[cout, kh, j]
//
//   A operand  = X, NHWC bf16 with a materialised halo: a [tw x th] box of output pixels at tap
//                (kh,kw) and channel chunk q is ONE 4-D TMA box load shifted by (kw,kh) -- 128 rows of
//                64 bf16 land in shared memory in exactly the K-major SWIZZLE_128B UMMA layout.
//                "window" maps (dim0 = kw*c with an overlapping pixel stride) serve thin inputs
//                (c = 8/16/32); "flat" maps tile the output in input-pitch order (data gradient).
//   B operand  = packed filter [cout_pad][kh][kwc_pad] bf16, 2-D TMA boxes of 64 x BN.
//   D          = fp32 in TMEM, double buffered (2 x BN columns) so the epilogue of tile i overlaps
//                the main loop of tile i+1.
//   epilogue   = tcgen05.ld -> +bias -> ReLU/LeakyReLU -> InstanceNorm sum/sumsq (warp-shuffle
//                transposed reduction + one atomic per column per warp) -> bf16/fp32 NHWC store.
//
// warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..11 = epilogue (TMEM lane
// quadrant = warp & 3; two warps per quadrant split the 32-column chunks).
#include <stdlib.h>

#include "common.cuh"

namespace {

struct ConvTcArgs {
  int n_img, ho, wo, wp;
  int tw, th, tiles_w, tiles_h;
  int flat, window;
  int kw, cchunks, kblocks;
  int bn, cout, out_c, act, stats, out_f32;
  int num_m_tiles, num_tiles, stages;
  int rowwin;               // 1: 8-channel input: the A operand is the raw pixel row in shared memory read through an
                            //    un-swizzled descriptor whose row pitch (16 B) equals the pixel pitch, so the
                            //    (kw, c) window of every pixel overlaps its neighbours' IN SHARED MEMORY: one 2 KB
                            //    TMA load per (tile, kh) instead of a 16 KB overlapping-stride box
  int nacc;                 // TMEM accumulator stages (2: measured no gain from 4 or 8 on short-K tiles)
  int stats_smem;           // 1: column sums of the direct epilogue through a warp-private shared-memory transpose
                            //    (8 STS.128 + 32 LDS + 64 FMA per 32x32 block instead of 62 shuffles + 124 selects +
                            //    62 adds); used when the main loop is short, i.e. when the epilogue paces the kernel
  int epi2;                 // 1: staged epilogue (TMEM -> shared memory tile -> coalesced stores + statistics)
  int epi3;                 // 1: BN <= 64 epilogue: one 32-column chunk per warp, InstanceNorm partial sums kept per lane in
                            //    registers across tiles (cross-lane reduction only when the image changes)
  int tstore;               // 1 (with epi3): the 128-pixel x 64-channel tile leaves through shared memory and one TMA store
  int exp;                  // timing experiments (VCG_EXP_EPI): 1 no epilogue work, 2 no stores, 4 no statistics
  int ashift;               // experiment: A rows loaded once per (kh, chunk), taps read through row-shifted descriptors
  int bres;                 // 1: the whole filter (kblocks x BN x 64) is loaded once per CTA and stays in shared memory
  uint32_t idesc, a_tx_bytes;
  const float* bias;
  float* stats_acc;
  void* out;
};

constexpr int kAStageBytes = 16384;  // 128 rows x 128 B
constexpr int kRowWinPix = 136;      // rowwin mode: 128 output pixels + 8 window pixels
constexpr int kRowWinStage = 2304;   // 136 pixels x 16 B, rounded up to a multiple of 128 B
constexpr int kThreads = 384;   // warps 0-2: TMA / MMA / TMEM alloc, 3: idle, 4-11: epilogue (2 per TMEM lane quadrant)

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const ConvTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int S = p.stages;
  const uint32_t b_bytes = static_cast<uint32_t>(p.bn) * 128u;
  // resident-filter mode (short K, one n-tile: the thin image-side layers, which are L2->SM bound): stages hold
  // only the A tile and the filter k-blocks live behind them; otherwise every stage carries its own B tile
  const uint32_t a_stage = p.ashift ? 17408u : static_cast<uint32_t>(kAStageBytes);
  const uint32_t stage_bytes = p.rowwin ? kRowWinStage : (p.bres ? a_stage : a_stage + b_bytes);
  const uint32_t bres0 = (base + S * stage_bytes + 1023u) & ~1023u;      // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bres_bytes = p.bres ? static_cast<uint32_t>(p.kblocks) * b_bytes : 0u;
  const int NA = p.nacc;
  const uint32_t bar0 = bres0 + bres_bytes;               // full[S], empty[S], tfull[NA], tempty[NA], bready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bar0 - base) + (2 * S + 2 * NA + 1) * 8);
  const uint32_t bready_bar = bar0 + 8u * (2 * S + 2 * NA);
  const uint32_t bias0 = bar0 + 1024u;                    // 8 epilogue warps x 256 floats: bias of the current n-tile
  const uint32_t epi0 = bias0 + 8192u;                    // staged-epilogue area (tile, row table, reduction slab)
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + NA + a); };

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(NA * p.bn)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < NA; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    mbar_init(bready_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  // contiguous chunk of tiles per CTA (m fastest): consecutive tiles share the image (statistics stay in
  // registers until the image changes) and neighbouring input rows / the same filter tile (L2 reuse)
  const int tiles_per_cta = (p.num_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_begin = blockIdx.x * tiles_per_cta;
  const int tile_end = min(p.num_tiles, tile_begin + tiles_per_cta);

  // Producer and MMA warps run their loops with all 32 lanes (warp-uniform control flow and values, so the
  // compiler keeps descriptors / coordinates in uniform registers) and let ONE elected lane issue the TMA / MMA
  // instructions: issuing from inside an `if (lane == 0)` region costs an R2UR waterfall loop per instruction.
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (p.bres && tile_begin < tile_end) {
      if (elect_one_sync()) {
        mbar_expect_tx(bready_bar, bres_bytes);
        for (int kb = 0; kb < p.kblocks; ++kb) tma_load_2d(bres0 + kb * b_bytes, &tmB, bready_bar, kb * 64, 0);
      }
      __syncwarp();
    }
    int stage = 0; uint32_t phase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int m_tile = tile % p.num_m_tiles, n_tile = tile / p.num_m_tiles;
      const int img = m_tile / tiles_per_img, rem = m_tile % tiles_per_img;
      const int h0 = (rem / p.tiles_w) * p.th, w0 = (rem % p.tiles_w) * p.tw;
      const int n0 = n_tile * p.bn;
      int khi = 0, kwi = 0, q = 0;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = base + stage * stage_bytes, sb = sa + a_stage;
        const int kws = p.ashift ? 0 : kwi;
        if (elect_one_sync()) {
          mbar_expect_tx(full_bar(stage), p.bres ? p.a_tx_bytes : p.a_tx_bytes + b_bytes);
          if (p.rowwin) {
            if (p.flat) tma_load_4d(sa, &tmA, full_bar(stage), 0, w0 + khi * p.wp, 0, img);
            else        tma_load_4d(sa, &tmA, full_bar(stage), 0, w0, h0 + khi, img);
          } else if (p.flat) tma_load_4d(sa, &tmA, full_bar(stage), q * 64, w0 + khi * p.wp + kws, 0, img);
          else               tma_load_4d(sa, &tmA, full_bar(stage), q * 64, w0 + kws, h0 + khi, img);
          if (!p.bres) tma_load_2d(sb, &tmB, full_bar(stage), kb * 64, n0);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
        // next (kh, kw, channel chunk): window maps fold kw into the chunk index
        if (++q == p.cchunks) { q = 0; if (p.window || ++kwi == p.kw) { kwi = 0; ++khi; } }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
    if (p.bres && tile_begin < tile_end) mbar_wait(bready_bar, 0);
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      mbar_wait(tempty_bar(as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.bn);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = base + stage * stage_bytes;
        const uint32_t sb = p.bres ? bres0 + kb * b_bytes : sa + a_stage;
        // rowwin: row pitch 16 B (one pixel), 8-row groups 128 B apart, second K chunk = next pixel (+16 B)
        uint64_t ad = p.rowwin ? umma_desc_linear(sa, 16, 128) : umma_desc_sw128(sa, 16, 1024);
        if (p.ashift) {
          const int kwi = (kb / p.cchunks) % p.kw;
          ad = umma_desc_sw128(sa + kwi * 128, 16, 1024);
          if (p.ashift == 1) ad |= static_cast<uint64_t>(kwi & 7) << 49;
        }
        const uint64_t bd = umma_desc_sw128(sb, 16, 1024);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
      if (elect_one_sync()) umma_commit(tfull_bar(as));
      __syncwarp();
      if (++as == NA) { as = 0; aphase ^= 1u; }
    }
  } else if (warp >= 4 && p.epi2) {
    // ===================== staged epilogue (BN <= 128) =====================
    // phase A: TMEM -> registers -> bias/activation -> output dtype -> 128 x BN tile in shared memory (16-byte chunks
    //          XOR-swizzled by row so that both phases are bank-conflict free); the accumulator is released here.
    // phase B: the tile leaves as fully coalesced 16-byte stores (a pixel's BN channels are contiguous in NHWC) and
    //          the InstanceNorm sum / sum-of-squares are column sums over the STORED values, read back from the tile
    //          (no warp-shuffle transposes); per-(image, channel) running sums stay in registers until the image
    //          changes.  Direct per-thread stores touched 32 different lines per instruction and the shuffle
    //          reductions cost ~250 instructions per 32 columns: tiles with short K were bound by this epilogue.
    const int et = static_cast<int>(threadIdx.x) - 128;             // 0..255
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const int esz = p.out_f32 ? 4 : 2;
    const int row_bytes = p.bn * esz, cpr = row_bytes >> 4;          // 16-byte chunks per tile row
    uint8_t* stg = smem + (epi0 - base);
    int* rowinfo = reinterpret_cast<int*>(stg + 128 * row_bytes);
    float* red = reinterpret_cast<float*>(rowinfo + 128);
    auto swz = [&](int k, int r) { return (k & ~7) | ((k ^ r) & 7); };
    const int pairs = p.bn >> 1, groups = 256 / pairs, rpg = 128 / groups;
    int as = 0; uint32_t aphase = 0;
    float* sb = reinterpret_cast<float*>(smem + (bias0 - base)) + (warp - 4) * 256;
    int cur_n0 = -1;
    float run = 0.f;                // thread et < 2*BN: running sum (et < BN) / sum of squares of channel n0 + et % BN
    int run_img = -1, run_n0 = 0;
    auto flush_stats = [&]() {
      if (run_img >= 0 && et < 2 * p.bn) {
        const int ch = run_n0 + (et % p.bn);
        if (ch < p.cout) atomicAdd(p.stats_acc + (static_cast<size_t>(run_img) * p.cout + ch) * 2 + et / p.bn, run);
      }
      run = 0.f;
    };
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int m_tile = tile % p.num_m_tiles, n_tile = tile / p.num_m_tiles;
      const int img = m_tile / tiles_per_img, rem = m_tile % tiles_per_img;
      const int h0 = (rem / p.tiles_w) * p.th, w0 = (rem % p.tiles_w) * p.tw;
      const int n0 = n_tile * p.bn;
      if (p.stats && (img != run_img || n0 != run_n0)) { flush_stats(); run_img = img; run_n0 = n0; }
      int h, w; bool valid;
      if (p.flat) { const int f = w0 + row; h = f / p.wp; w = f - h * p.wp; valid = (h < p.ho) && (w < p.wo); }
      else { const int hh = row / p.tw; h = h0 + hh; w = w0 + (row - hh * p.tw);
             valid = (row < p.tw * p.th) && (h < p.ho) && (w < p.wo); }
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
        if (p.out_f32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(srow + swz(c0 / 4 + j, row) * 16) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            __nv_bfloat162* hp2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) hp2[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
            *reinterpret_cast<uint4*>(srow + swz(c0 / 8 + j, row) * 16) = pk;
          }
        }
      }
      if (half == 0) rowinfo[row] = valid ? static_cast<int>((static_cast<size_t>(img) * p.ho + h) * p.wo + w) : -1;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));          // accumulator drained
      if (++as == NA) { as = 0; aphase ^= 1u; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // ---- phase B: coalesced stores
      const int cpe = 16 / esz;                              // channels per 16-byte chunk
      for (int idx = et; idx < 128 * cpr; idx += 256) {
        const int r2 = idx / cpr, k = idx - r2 * cpr;
        const int pi = rowinfo[r2];
        const int col0 = n0 + k * cpe;
        if (pi >= 0 && (col0 & ~7) < p.cout) {
          const uint4 val = *reinterpret_cast<const uint4*>(stg + r2 * row_bytes + swz(k, r2) * 16);
          *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out) + (static_cast<size_t>(pi) * p.out_c + col0) * esz) = val;
        }
      }
      // ---- statistics over the stored (rounded) values: thread = (channel pair, group of rows)
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
      asm volatile("bar.sync 1, 256;" ::: "memory");          // tile + row table free again; slab complete
      if (p.stats && et < 2 * p.bn) {
        const int which = et / p.bn, ch = et % p.bn;
        float t = 0.f;
        for (int g = 0; g < groups; ++g) t += red[(g * 2 + which) * p.bn + ch];
        run += t;
      }
    }
    if (p.stats) flush_stats();
  } else if (warp >= 4 && p.epi3) {
    // ===================== thin-N epilogue (BN <= 64, one n-tile) =====================
    // Short-K tiles with 64 output channels are paced by this code, not by the MMAs (measured: 64->64 1x1 @256^2 took
    // 0.42 ms with the generic epilogue, 0.14 ms without any).  Two changes: (1) statistics: every lane adds its row's
    // 32 values into per-lane partial sums that live in registers across ALL tiles of an image; the 32x32 transpose
    // reduction runs once per image instead of once per tile; (2) stores: the bf16 tile is written to shared memory
    // in the 128-byte-swizzle layout and leaves with ONE TMA tensor store (full 128-byte lines) instead of 4 STG.128
    // per lane that each touch 32 different lines; three tile buffers, one named barrier per tile.
    const int et = static_cast<int>(threadIdx.x) - 128;
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const int c0 = half * 32;
    const bool has_chunk = c0 < p.bn;
    float a1[32], a2[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) a1[j] = a2[j] = 0.f;
    int run_img = -1;
    float* sb = reinterpret_cast<float*>(smem + (bias0 - base)) + (warp - 4) * 256;
    load_bias_tile(sb, p.bias, 0, p.bn, p.cout, lane);
    auto flush_stats = [&]() {
      if (run_img >= 0 && has_chunk) {
        transposed_warp_sum32(a1, lane);
        transposed_warp_sum32(a2, lane);
        const int col = c0 + lane;
        if (col < p.cout) {
          float* dst = p.stats_acc + (static_cast<size_t>(run_img) * p.cout + col) * 2;
          atomicAdd(dst, a1[0]);
          atomicAdd(dst + 1, a2[0]);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) a1[j] = a2[j] = 0.f;
    };
    int as = 0; uint32_t aphase = 0; int buf = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int m_tile = tile % p.num_m_tiles;
      const int img = m_tile / tiles_per_img, rem = m_tile % tiles_per_img;
      const int h0 = (rem / p.tiles_w) * p.th, w0 = (rem % p.tiles_w) * p.tw;
      if (p.stats && img != run_img) { flush_stats(); run_img = img; }
      int h, w; bool valid;
      if (p.flat) { const int f = w0 + row; h = f / p.wp; w = f - h * p.wp; valid = (h < p.ho) && (w < p.wo); }
      else { const int hh = row / p.tw; h = h0 + hh; w = w0 + (row - hh * p.tw);
             valid = (row < p.tw * p.th) && (h < p.ho) && (w < p.wo); }
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      uint32_t r[32];
      if (has_chunk) {
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * p.bn + c0), r);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));          // accumulator is in registers: the next tile's MMAs may start
      if (++as == NA) { as = 0; aphase ^= 1u; }
      if (has_chunk && !(p.exp & 1)) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        bias_act32(v, sb + c0, p.act, valid ? p.cout - c0 : 0);
        if (p.stats && !(p.exp & 4)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { a1[j] += v[j]; a2[j] = fmaf(v[j], v[j], a2[j]); }
        }
        if (p.tstore) {
          uint8_t* srow = smem + (epi0 - base) + buf * 16384 + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            __nv_bfloat162* hp2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) hp2[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
            *reinterpret_cast<uint4*>(srow + (((half * 4 + j) ^ (row & 7)) << 4)) = pk;
          }
        } else if (valid && !(p.exp & 2)) {
          const size_t pix = (static_cast<size_t>(img) * p.ho + h) * p.wo + w;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = c0 + g * 8;
            if (col < p.cout) {
              float t[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) t[j] = v[g * 8 + j];
              st8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_c + col, t);
            }
          }
        }
      }
      if (p.tstore) {
        fence_proxy_async_smem();                           // generic-proxy tile writes -> visible to the TMA engine
        // the store of two tiles ago has finished READING its buffer, which is the one the next tile writes
        if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0 && !(p.exp & 2)) {
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(epi0 + buf * 16384), "r"(0), "r"(w0), "r"(h0), "r"(img) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (++buf == 3) buf = 0;
      }
    }
    if (p.stats) flush_stats();
    if (p.tstore && et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // two epilogue warps per TMEM lane quadrant: warp (4+q) takes the even 32-column chunks, warp (8+q) the odd
    // ones, so a short-K tile (epilogue-bound) drains in half the time
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    int as = 0; uint32_t aphase = 0;
    // per-lane running InstanceNorm sums: lane l of this warp owns column (chunk*32 + l) of the current
    // (image, n_tile); flushed with ONE atomic pair per column when either changes (not once per tile)
    float run1[8], run2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) run1[i] = run2[i] = 0.f;
    int run_img = -1, run_n0 = 0;
    float* sb = reinterpret_cast<float*>(smem + (bias0 - base)) + (warp - 4) * 256;
    float* tr = reinterpret_cast<float*>(smem + (epi0 - base)) + (warp - 4) * (32 * 36);     // stats_smem: 32 x 36 floats per warp
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
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int m_tile = tile % p.num_m_tiles, n_tile = tile / p.num_m_tiles;
      const int img = m_tile / tiles_per_img, rem = m_tile % tiles_per_img;
      const int h0 = (rem / p.tiles_w) * p.th, w0 = (rem % p.tiles_w) * p.tw;
      const int n0 = n_tile * p.bn;
      if (p.stats && (img != run_img || n0 != run_n0)) { flush_stats(); run_img = img; run_n0 = n0; }
      int h, w; bool valid;
      if (p.flat) { const int f = w0 + row; h = f / p.wp; w = f - h * p.wp; valid = (h < p.ho) && (w < p.wo); }
      else { const int hh = row / p.tw; h = h0 + hh; w = w0 + (row - hh * p.tw);
             valid = (row < p.tw * p.th) && (h < p.ho) && (w < p.wo); }
      const size_t pix = (static_cast<size_t>(img) * p.ho + h) * p.wo + w;
      if (n0 != cur_n0) { load_bias_tile(sb, p.bias, n0, p.bn, p.cout, lane); cur_n0 = n0; }
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * p.bn);
      if (!(p.exp & 1))
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        const int c0 = ci * 32;
        if (c0 >= p.bn) break;
        if ((ci & 1) != half) continue;
        float v[32];
        if (p.bn - c0 >= 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        } else {  // bn == 16 (mod 32)
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = __uint_as_float(r[j]); v[j + 16] = 0.f; }
        }
        const int col0 = n0 + c0;
        bias_act32(v, sb + c0, p.act, p.cout - col0);
        if (valid && !(p.exp & 2)) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            if (col < p.cout) {
              float t[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) t[j] = v[g * 8 + j];
              if (p.out_f32) st8<float>(reinterpret_cast<float*>(p.out) + pix * p.out_c + col, t);
              else st8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_c + col, t);
            }
          }
        }
        if (p.stats && col0 < p.cout && !(p.exp & 4)) {
          if (p.stats_smem) {
            // lane r parks its 32 values in row r (pitch 36 floats: conflict-free 16-byte stores), then lane j sums column j
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(tr + lane * 36 + 4 * q) =
                  valid ? make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            float a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) { const float x = tr[r * 36 + lane]; a1 += x; a2 = fmaf(x, x, a2); }
            run1[ci] += a1;
            run2[ci] += a2;
          } else {
            float s1[32], s2[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float x = valid ? v[j] : 0.f; s1[j] = x; s2[j] = x * x; }
            transposed_warp_sum32(s1, lane);
            transposed_warp_sum32(s2, lane);
            run1[ci] += s1[0];
            run2[ci] += s2[0];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
      if (++as == NA) { as = 0; aphase ^= 1u; }
    }
    if (p.stats) flush_stats();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace

// --------------------------------------------------------------------------------- host side
static void pick_box(int wo, int ho, int* tw, int* th) {
  // largest tw*th <= 128 box: full rows when they fit, otherwise the best divisor-free cover
  if (wo >= 128) { *tw = 128; *th = 1; return; }
  *tw = wo;
  int t = 128 / wo;
  if (t > ho) t = ho;
  if (t < 1) t = 1;
  *th = t;
}

bool vcg_conv_fold_supported(const vcg_conv_desc* d, bool has_stats);
int vcg_conv_fwd_tc_fold(const vcg_conv_desc* d, const void* x, const void* w, const float* bias, void* y, int out_f32,
                         cudaStream_t stream);

bool vcg_conv2_supported(const vcg_conv_desc* d, int out_f32);
int vcg_conv_fwd_tc2(const vcg_conv_desc* d, const void* x, const void* w, const float* bias, void* y, float* stats,
                     cudaStream_t stream);

int vcg_conv_fwd_tc(const vcg_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                    float* stats, int out_f32, cudaStream_t stream) {
  const int ho = d->hp - d->kh + 1, wo = d->wp - d->kw + 1;
  // thin outputs on wide maps (64->3 7x7, the data gradients of 3->64 7x7 and 32->64 3x3): horizontal taps folded
  // into N, input rows streamed once through a ring of TMEM accumulators (conv_tc_fold.cu)
  if (vcg_conv_fold_supported(d, d->stats && stats)) return vcg_conv_fwd_tc_fold(d, x, w, bias, y, out_f32, stream);
  if (vcg_conv2_supported(d, out_f32)) return vcg_conv_fwd_tc2(d, x, w, bias, y, stats, stream);
  VCG_REQUIRE(d->c % 8 == 0 && d->kwc_pad % 64 == 0 && d->cout_pad % 16 == 0 && d->out_c % 8 == 0,
              VCG_E_UNSUPPORTED, "conv_tc: unsupported channel geometry c=%d kwc_pad=%d cout_pad=%d cout=%d",
              d->c, d->kwc_pad, d->cout_pad, d->cout);
  VCG_REQUIRE(ho > 0 && wo > 0, VCG_E_INVALID, "conv_tc: empty output");
  const bool window = (d->c % 64) != 0;
  VCG_REQUIRE(window || d->kwc_pad == d->kw * d->c, VCG_E_INVALID, "conv_tc: kwc_pad mismatch");
  VCG_REQUIRE(!window || d->kwc_pad == ((d->kw * d->c + 63) / 64) * 64, VCG_E_INVALID, "conv_tc: kwc_pad mismatch (window)");

  ConvTcArgs a{};
  a.n_img = d->n; a.ho = ho; a.wo = wo; a.wp = d->wp;
  a.flat = d->flat ? 1 : 0; a.window = window ? 1 : 0;
  a.kw = d->kw;
  a.cchunks = window ? d->kwc_pad / 64 : d->c / 64;
  a.kblocks = d->kh * (d->kwc_pad / 64);
  if (a.flat) {
    a.tw = 128; a.th = 1; a.tiles_h = 1;
    a.tiles_w = (ho * d->wp + 127) / 128;   // rows >= ho are skipped by the epilogue anyway
    a.a_tx_bytes = 128 * 128;
  } else {
    pick_box(wo, ho, &a.tw, &a.th);
    a.tiles_w = (wo + a.tw - 1) / a.tw;
    a.tiles_h = (ho + a.th - 1) / a.th;
    a.a_tx_bytes = static_cast<uint32_t>(a.tw * a.th) * 128u;
  }
  a.num_m_tiles = d->n * a.tiles_w * a.tiles_h;
  const int sms = vcg_gemm_sms();
  int bn = d->cout_pad < 256 ? d->cout_pad : 256;
  if (bn == 256 && static_cast<long long>(a.num_m_tiles) * (d->cout_pad / 256) < sms) bn = 128;
  if (bn > 64 && (bn % 64) != 0) bn = 64;
  VCG_REQUIRE(bn % 16 == 0 && bn >= 16, VCG_E_UNSUPPORTED, "conv_tc: bn=%d", bn);
  a.bn = bn;
  const int ntn = (d->cout_pad + bn - 1) / bn;
  a.num_tiles = a.num_m_tiles * ntn;
  a.cout = d->cout; a.out_c = d->out_c; a.act = d->act; a.stats = (d->stats && stats) ? 1 : 0;
  a.out_f32 = out_f32;
  a.bias = bias; a.stats_acc = stats; a.out = y;
  a.idesc = umma_idesc_bf16(128, bn, 0, 0);
  // resident filter: one n-tile, many tiles per CTA, and the whole filter fits beside >= 4 A stages
  const size_t filt_bytes = static_cast<size_t>(a.kblocks) * bn * 128;
  a.bres = (ntn == 1 && a.num_tiles >= 4 * sms && filt_bytes + 4 * kAStageBytes + 3072 + 8192 + 40960 <= 227 * 1024) ? 1 : 0;
  // staged epilogue: needs a 128 x BN output tile (+ row table + reduction slab) in shared memory
  const int esz = out_f32 ? 4 : 2;
  // measured (tools/bench_conv.py, B=64): BN=128 layers gain (256->128 @128^2 forward 1020 -> 1178 TFLOP/s); BN=64 tiles
  // are bound by the MMA's shared-memory operand reads and the extra tile traffic costs 10 %, so they keep the
  // direct epilogue
  a.epi2 = (bn == 128 && !out_f32 && static_cast<long long>(d->n) * ho * wo < (1LL << 31)) ? 1 : 0;
  a.nacc = 2;
  a.stats_smem = (a.stats && !a.epi2 && a.kblocks <= 24) ? 1 : 0;   // measured: +13 % at 18 k-blocks, -3 % at 36
  a.epi3 = (!a.epi2 && !out_f32 && ntn == 1 && (bn == 64 || bn == 32)) ? 1 : 0;
  a.tstore = (a.epi3 && !a.flat && bn == 64 && d->cout == 64 && d->out_c == 64 && a.tw == 128 && a.th == 1) ? 1 : 0;
  if (a.epi3) a.stats_smem = 0;
  const size_t epi_bytes = a.tstore ? 3 * 16384 : a.epi2 ? static_cast<size_t>(128) * bn * esz + 512 + 4096 : (a.stats_smem ? 8 * 32 * 36 * 4 : 0);
  a.rowwin = (a.bres && window && d->c == 8 && d->kwc_pad == 64 && a.tw == 128 && a.th == 1) ? 1 : 0;
  if (a.rowwin) a.a_tx_bytes = kRowWinPix * 16;
  a.exp = 0;        // timing experiments of round 1 (epilogue-only / row-shifted A descriptors), see DESIGN.md section 6
  a.ashift = 0;
  const int a_stage = a.ashift ? 17408 : kAStageBytes;
  const int stage_bytes = a.rowwin ? kRowWinStage : (a.bres ? a_stage : a_stage + bn * 128);
  int stages = static_cast<int>((227 * 1024 - 3072 - 8192 - (a.bres ? filt_bytes : 0) - epi_bytes) / stage_bytes);
  if (stages > (a.rowwin ? 16 : 8)) stages = a.rowwin ? 16 : 8;
  if (stages > a.kblocks && !a.bres) stages = a.kblocks;
  if (stages < 2) stages = 2;
  a.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + (a.bres ? filt_bytes : 0) + 3072 + 8192 + epi_bytes;

  // ---- tensor maps
  CUtensorMap tmA, tmB;
  const uint64_t es = 2;
  const uint64_t pix_stride = static_cast<uint64_t>(d->c) * es;
  const uint64_t row_stride = static_cast<uint64_t>(d->wp) * pix_stride;
  const uint64_t img_stride = static_cast<uint64_t>(d->hp) * row_stride;
  uint64_t dims[4], strides[3];
  uint32_t box[4];
  dims[0] = window ? static_cast<uint64_t>(d->kw) * d->c : static_cast<uint64_t>(d->c);
  if (a.flat) {
    dims[1] = static_cast<uint64_t>(d->hp) * d->wp - (window ? (d->kw - 1) : 0);
    dims[2] = 1; dims[3] = d->n;
    strides[0] = pix_stride; strides[1] = img_stride; strides[2] = img_stride;
    box[0] = 64; box[1] = a.ashift ? 136 : 128; box[2] = 1; box[3] = 1;
  } else {
    dims[1] = window ? static_cast<uint64_t>(wo) : static_cast<uint64_t>(d->wp);
    dims[2] = d->hp; dims[3] = d->n;
    strides[0] = pix_stride; strides[1] = row_stride; strides[2] = img_stride;
    box[0] = 64; box[1] = a.ashift ? 136 : a.tw; box[2] = a.th; box[3] = 1;
  }
  int rc;
  if (a.rowwin) {
    // plain (un-swizzled) map over the 8-channel pixels: box = 136 consecutive pixels of one row (flat: of the image)
    dims[0] = 8; box[0] = 8; box[1] = kRowWinPix;
    dims[1] = a.flat ? static_cast<uint64_t>(d->hp) * d->wp : static_cast<uint64_t>(d->wp);
    rc = vcg_encode_tmap_linear(&tmA, x, 4, dims, strides, box, "conv_tc A (row window)");
  } else {
    rc = vcg_encode_tmap(&tmA, x, 4, dims, strides, box, "conv_tc A");
  }
  if (rc) return rc;
  const uint64_t ktot = static_cast<uint64_t>(d->kh) * d->kwc_pad;
  uint64_t bdims[2] = {ktot, static_cast<uint64_t>(d->cout_pad)};
  uint64_t bstr[1] = {ktot * es};
  uint32_t bbox[2] = {64, static_cast<uint32_t>(bn)};
  rc = vcg_encode_tmap(&tmB, w, 2, bdims, bstr, bbox, "conv_tc B");
  if (rc) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap tmO = tmA;                                   // placeholder unless the TMA-store epilogue is on
  if (a.tstore) {
    uint64_t od[4] = {static_cast<uint64_t>(d->out_c), static_cast<uint64_t>(wo), static_cast<uint64_t>(ho), static_cast<uint64_t>(d->n)};
    uint64_t os[3] = {od[0] * es, od[0] * es * wo, od[0] * es * wo * ho};
    uint32_t ob[4] = {64, 128, 1, 1};
    rc = vcg_encode_tmap(&tmO, y, 4, od, os, ob, "conv_tc out");
    if (rc) return rc;
  }
  const int grid = a.num_tiles < sms ? a.num_tiles : sms;
  conv_tc_kernel<<<grid, kThreads, smem, stream>>>(tmA, tmB, tmO, a);
  VCG_CHECK_LAUNCH("conv_tc_kernel");
  return VCG_OK;
}
