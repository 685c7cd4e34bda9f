// conv_tc_fold.cu -- tcgen05 convolution for THIN outputs (<= 32 channels) on wide maps, sm_100a:
// the 64->3 7x7 output convolution (Networks.py:192), the data gradient of the 3->64 7x7 input
// convolution (Networks.py:158) and the data gradient of the last up-sampling conv (Networks.py:191).
//
// With N = cout the implicit GEMM of conv_tc.cu is bound by shared-memory operand reads: every MMA re-reads a
// 128 x 64 A tile (4 KB per K=16 step) to produce only 16 columns, once per tap (49x for 7x7).  Here the
// HORIZONTAL taps are folded into N instead:
//
//     P_h[w', (kw, co)] = sum_{kh, c} X[h + kh, w', c] * W[co, kh, kw, c]          (GEMM, N = kw * co)
//     Y[h, w, co]       = sum_{kw} P_h[w + kw, (kw, co)]                            (epilogue shift-add)
//
// so an input row strip of 128 pixels x 64 channels is loaded ONCE (no kw shift) and every MMA produces kw*co
// columns: 7x fewer MMAs and 7x fewer operand bytes for the 7x7 layers.  Input rows stream through a ring of
// TMEM accumulators (one per output row in flight): input row rho feeds the accumulators of output rows
// rho-kh+1 .. rho; an output row is complete after its last tap row and is drained by the epilogue warps while
// the MMAs of the following rows continue.  The filter ((kh*cchunks) x N x 64 bf16) stays resident in shared
// memory.  A tile of 128 input columns yields 128-kw+1 output columns.
//
// Several output rows per MMA: input row rho feeds output row r with tap kh = rho - r, i.e. CONSECUTIVE ring slots take
// consecutive taps in descending order.  The resident filter is therefore stored tap-descending (block kh-1-khi), so
// the filter blocks of a run of g consecutive slots are contiguous in shared memory and one MMA with N = g * bn
// (<= 256) feeds all g accumulators: the 128 x 64 A tile is read once per run instead of once per slot (an MMA with
// both operands in shared memory costs max(N/2, (4096 + 32 N) / 70) cycles: 88 at N = 64, 175 at N = 256).  Only the
// first touch of a row (tap 0, first channel chunk: accumulate = 0) and the ring wrap need their own MMA.
//
// warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = epilogue (TMEM lane quadrant = warp & 3).
#include <stdlib.h>

#include "common.cuh"

namespace {

struct FoldConvArgs {
  int n_img, ho, wo;
  int tiles_w, segs_h, seg_rows, tile_w_out;
  int kh, kw, cchunks, kwc_pad, c;
  int co8, bn, nslots;
  int cout, out_c, act, out_f32;
  int num_items, stages;
  uint32_t idesc, b_tx_bytes;
  const float* bias;
  void* out;
};

constexpr int kAStage = 16384;
constexpr int kThreads = 256;

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int KW, int NB4>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_fold_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FoldConvArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int S = p.stages;
  const uint32_t b_chunk = static_cast<uint32_t>(p.bn) * 128u;                 // one (kh, q) filter chunk: N rows x 128 B
  const uint32_t nb_chunks = static_cast<uint32_t>(p.kh * p.cchunks);
  const uint32_t bres = base + S * kAStage;                                    // resident filter
  const uint32_t stg_bytes = static_cast<uint32_t>(KW * NB4) * 128u * 16u;   // one staging buffer: [kw*nb4][128] float4
  const uint32_t stg0 = bres + nb_chunks * b_chunk;
  const uint32_t bar0 = stg0 + 2 * stg_bytes;       // full[S], empty[S], tfull[nslots], tempty[nslots], bready
  const int NS = p.nslots;
  uint8_t* bar_ptr = smem + (bar0 - base);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + (2 * S + 2 * NS + 1) * 8);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + NS + a); };
  const uint32_t bready_bar = bar0 + 8u * (2 * S + 2 * NS);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(NS * p.bn)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < NS; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    mbar_init(bready_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(smem_u32(tmem_slot), tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  const int items_per_img = p.tiles_w * p.segs_h;
  const int per_cta = (p.num_items + gridDim.x - 1) / gridDim.x;
  const int item_begin = blockIdx.x * per_cta;
  const int item_end = min(p.num_items, item_begin + per_cta);
  auto decode = [&](int item, int& img, int& h0, int& rows, int& w0) {
    img = item / items_per_img;
    const int rem = item - img * items_per_img;
    const int seg = rem / p.tiles_w;
    h0 = seg * p.seg_rows;
    rows = min(p.seg_rows, p.ho - h0);
    w0 = (rem - seg * p.tiles_w) * p.tile_w_out;
  };

  if (warp == 0) {
    if (lane == 0 && item_begin < item_end) {
      // ---- resident filter: for every (kh, q) the kw blocks of co8 rows, one 8-row-aligned TMA box each
      mbar_expect_tx(bready_bar, p.b_tx_bytes);
      for (int khi = 0; khi < p.kh; ++khi)
        for (int q = 0; q < p.cchunks; ++q)
          for (int kwi = 0; kwi < p.kw; ++kwi)
            tma_load_2d(bres + static_cast<uint32_t>(q * p.kh + (p.kh - 1 - khi)) * b_chunk + static_cast<uint32_t>(kwi * p.co8) * 128u,
                        &tmB, bready_bar, khi * p.kwc_pad + kwi * p.c + q * 64, 0);
      int stage = 0; uint32_t phase = 0;
      for (int item = item_begin; item < item_end; ++item) {
        int img, h0, rows, w0;
        decode(item, img, h0, rows, w0);
        for (int rho = 0; rho < rows + p.kh - 1; ++rho)
          for (int q = 0; q < p.cchunks; ++q) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), kAStage);
            tma_load_4d(base + stage * kAStage, &tmA, full_bar(stage), q * 64, w0, h0 + rho, img);
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp runs the (uniform) loop, one elected lane issues
    if (item_begin < item_end) {
      mbar_wait(bready_bar, 0);
      int stage = 0; uint32_t phase = 0;
      int slot0 = 0; uint32_t sphase0 = 0;      // ring slot / phase of the current item's output row 0
      const int maxg = 256 / p.bn;              // ring slots one MMA may span
      for (int item = item_begin; item < item_end; ++item) {
        int img, h0, rows, w0;
        decode(item, img, h0, rows, w0);
        int slot_lo = slot0; uint32_t sphase_lo = sphase0;     // ring position of output row r_lo
        for (int rho = 0; rho < rows + p.kh - 1; ++rho) {
          const int r_lo = rho - (p.kh - 1) > 0 ? rho - (p.kh - 1) : 0;
          const int r_hi = rho < rows - 1 ? rho : rows - 1;
          for (int q = 0; q < p.cchunks; ++q) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(base + stage * kAStage, 16, 1024);
            int slot = slot_lo; uint32_t sphase = sphase_lo;
            int khi = rho - r_lo;
            int r = r_lo;
            while (r <= r_hi) {
              // first touch of the newest output row (tap 0, first chunk): its accumulator must have been drained, and
              // the MMA overwrites instead of accumulating -> always a run of its own
              const bool first = (khi == 0) && (q == 0);
              int g = 1;
              if (first) {
                mbar_wait(tempty_bar(slot), sphase ^ 1u);
                tc_fence_after();
              } else {
                while (g < maxg && r + g <= r_hi && slot + g < NS && !((khi - g == 0) && (q == 0))) ++g;
              }
              const uint64_t bdesc = umma_desc_sw128(bres + static_cast<uint32_t>(q * p.kh + (p.kh - 1 - khi)) * b_chunk, 16, 1024);
              const uint32_t d = tmem_base + static_cast<uint32_t>(slot * p.bn);
              const uint32_t idesc = umma_idesc_bf16(128, g * p.bn, 0, 0);
              if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
                // the oldest row of the run (largest tap) completes with its last tap row and chunk
                if (khi == p.kh - 1 && q == p.cchunks - 1) umma_commit(tfull_bar(slot));
              }
              __syncwarp();
              r += g; khi -= g; slot += g;
              if (slot >= NS) { slot -= NS; sphase ^= 1u; }
            }
            if (elect_one_sync()) umma_commit(empty_bar(stage));
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1u; }
          }
          if (rho >= p.kh - 1) { if (++slot_lo == NS) { slot_lo = 0; sphase_lo ^= 1u; } }   // r_lo advances with rho
        }
        slot0 += rows % NS; if (slot0 >= NS) { slot0 -= NS; sphase0 ^= 1u; }
        sphase0 ^= static_cast<uint32_t>(rows / NS) & 1u;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> staging smem -> shift-add over kw -> bias/act -> store
    // One TMEM column group of 8 per tap (co8 == 8).  Staging granularity = 4 channels (one float4 per lane:
    // consecutive lanes hit consecutive 16-byte words, no bank conflicts); only the NB4 = ceil(cout/4) blocks that
    // hold real channels are staged.  KW and NB4 are compile-time so that all TMEM loads are issued back to back
    // and the shift-add is fully unrolled: the per-row latency of these 4 warps paces the whole kernel.
    const int quad = warp & 3;
    const int j = quad * 32 + lane;                     // TMEM lane = input column w0 + j; also the output column handled
    float4* stg = reinterpret_cast<float4*>(smem + (stg0 - base));
    const uint32_t stg_vecs = stg_bytes / 16;
    float bias_r[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) bias_r[e] = (p.bias && e < p.cout) ? __ldg(p.bias + e) : 0.f;
    int slot = 0; uint32_t sphase = 0, buf = 0;
    for (int item = item_begin; item < item_end; ++item) {
      int img, h0, rows, w0;
      decode(item, img, h0, rows, w0);
      const int w = w0 + j;
      const bool valid = (j < p.tile_w_out) && (w < p.wo);
      size_t pix = (static_cast<size_t>(img) * p.ho + h0) * p.wo + w;
      for (int r = 0; r < rows; ++r, pix += p.wo) {
        float4* sbuf = stg + buf * stg_vecs;
        buf ^= 1u;
        mbar_wait(tfull_bar(slot), sphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(slot * p.bn);
        uint32_t rr[KW][8];
#pragma unroll
        for (int t = 0; t < KW; ++t) tmem_ld8(taddr + static_cast<uint32_t>(t * 8), rr[t]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(slot));   // accumulator drained: the MMA warp may start a new row in it
#pragma unroll
        for (int t = 0; t < KW; ++t) {
          sbuf[(t * NB4) * 128 + j] =
              make_float4(__uint_as_float(rr[t][0]), __uint_as_float(rr[t][1]), __uint_as_float(rr[t][2]), __uint_as_float(rr[t][3]));
          if (NB4 == 2)
            sbuf[(t * NB4 + 1) * 128 + j] =
                make_float4(__uint_as_float(rr[t][4]), __uint_as_float(rr[t][5]), __uint_as_float(rr[t][6]), __uint_as_float(rr[t][7]));
        }
        epi_bar_sync();                                  // staging complete (and everyone is done with the other buffer)
        if (valid) {
          float acc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = bias_r[e];
#pragma unroll
          for (int t = 0; t < KW; ++t) {
            const float4 v = sbuf[(t * NB4) * 128 + j + t];
            acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            if (NB4 == 2) {
              const float4 u = sbuf[(t * NB4 + 1) * 128 + j + t];
              acc[4] += u.x; acc[5] += u.y; acc[6] += u.z; acc[7] += u.w;
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = e < p.cout ? act_apply(acc[e], p.act) : 0.f;
          if (p.out_f32) st8<float>(reinterpret_cast<float*>(p.out) + pix * p.out_c, acc);
          else st8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_c, acc);
          // physical channel groups beyond the logical output channels are zero
          for (int col = 8; col < p.out_c; col += 8) {
            float z[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) z[e] = 0.f;
            if (p.out_f32) st8<float>(reinterpret_cast<float*>(p.out) + pix * p.out_c + col, z);
            else st8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_c + col, z);
          }
        }
        if (++slot == NS) { slot = 0; sphase ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

struct FoldGeom { int co8, bn, nslots, stages; size_t smem; };

bool fold_geometry(const vcg_conv_desc* d, FoldGeom* g) {
  const int wo = d->wp - d->kw + 1;
  if (d->c % 64 != 0 || d->kwc_pad != d->kw * d->c || d->cout > 8 || wo < 64 || !(d->kw == 7 || d->kw == 3 || d->kw == 2)) return false;
  g->co8 = (d->cout + 7) / 8 * 8;
  if (g->co8 > d->cout_pad) return false;
  g->bn = (d->kw * g->co8 + 15) / 16 * 16;
  if (g->bn > 256) return false;
  g->nslots = 512 / g->bn;
  if (g->nslots > 8) g->nslots = 8;
  if (g->nslots < d->kh + 1) return false;
  const size_t bres = static_cast<size_t>(d->kh) * (d->c / 64) * g->bn * 128;
  const size_t stg = 2 * static_cast<size_t>(d->kw) * ((d->cout + 3) / 4) * 128 * 16;
  const size_t fixed = bres + stg + 2048;
  if (fixed + 3 * kAStage > 227 * 1024) return false;
  int stages = static_cast<int>((227 * 1024 - fixed) / kAStage);
  if (stages > 8) stages = 8;
  g->stages = stages;
  g->smem = fixed + static_cast<size_t>(stages) * kAStage;
  return true;
}

}  // namespace

bool vcg_conv_fold_supported(const vcg_conv_desc* d, bool has_stats) {
  FoldGeom g;
  return !has_stats && fold_geometry(d, &g);
}

int vcg_conv_fwd_tc_fold(const vcg_conv_desc* d, const void* x, const void* w, const float* bias, void* y, int out_f32,
                         cudaStream_t stream) {
  FoldGeom g;
  VCG_REQUIRE(fold_geometry(d, &g), VCG_E_UNSUPPORTED, "conv_tc_fold: unsupported geometry");
  const int ho = d->hp - d->kh + 1, wo = d->wp - d->kw + 1;
  const int sms = vcg_gemm_sms();
  FoldConvArgs a{};
  a.n_img = d->n; a.ho = ho; a.wo = wo;
  a.tile_w_out = 128 - d->kw + 1;
  a.tiles_w = (wo + a.tile_w_out - 1) / a.tile_w_out;
  // output rows per work item: long segments re-read fewer halo rows, short ones balance the persistent CTAs
  int seg = 64;
  while (seg > 8 && static_cast<long long>(d->n) * a.tiles_w * ((ho + seg - 1) / seg) < 4LL * sms) seg >>= 1;
  if (seg > ho) seg = ho;
  a.seg_rows = seg;
  a.segs_h = (ho + seg - 1) / seg;
  a.kh = d->kh; a.kw = d->kw; a.cchunks = d->c / 64; a.kwc_pad = d->kwc_pad; a.c = d->c;
  a.co8 = g.co8; a.bn = g.bn; a.nslots = g.nslots;
  a.cout = d->cout; a.out_c = d->out_c; a.act = d->act; a.out_f32 = out_f32;
  a.num_items = d->n * a.tiles_w * a.segs_h;
  a.stages = g.stages;
  a.idesc = umma_idesc_bf16(128, g.bn, 0, 0);
  a.b_tx_bytes = static_cast<uint32_t>(d->kh * a.cchunks * d->kw * g.co8) * 128u;
  a.bias = bias; a.out = y;

  CUtensorMap tmA, tmB;
  const uint64_t es = 2, pix = d->c * es, row = d->wp * pix, img = d->hp * row;
  uint64_t dims[4] = {static_cast<uint64_t>(d->c), static_cast<uint64_t>(d->wp), static_cast<uint64_t>(d->hp),
                      static_cast<uint64_t>(d->n)};
  uint64_t strides[3] = {pix, row, img};
  uint32_t box[4] = {64, 128, 1, 1};
  int rc = vcg_encode_tmap(&tmA, x, 4, dims, strides, box, "conv_tc_fold A");
  if (rc) return rc;
  const uint64_t ktot = static_cast<uint64_t>(d->kh) * d->kwc_pad;
  uint64_t bdims[2] = {ktot, static_cast<uint64_t>(d->cout_pad)};
  uint64_t bstr[1] = {ktot * es};
  uint32_t bbox[2] = {64, static_cast<uint32_t>(g.co8)};
  rc = vcg_encode_tmap(&tmB, w, 2, bdims, bstr, bbox, "conv_tc_fold B");
  if (rc) return rc;
  const int grid = a.num_items < sms ? a.num_items : sms;
#define VCG_FOLD_LAUNCH(KW, NB4)                                                                                         \
  do {                                                                                                                   \
    static bool attr_set = false;                                                                                        \
    if (!attr_set) {                                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(conv_tc_fold_kernel<KW, NB4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "conv_tc_fold: cudaFuncSetAttribute: %s", cudaGetErrorString(e));         \
      attr_set = true;                                                                                                   \
    }                                                                                                                    \
    conv_tc_fold_kernel<KW, NB4><<<grid, kThreads, g.smem, stream>>>(tmA, tmB, a);                                        \
  } while (0)
  const int nb4 = (d->cout + 3) / 4;
  if (d->kw == 7 && nb4 == 1) VCG_FOLD_LAUNCH(7, 1);
  else if (d->kw == 7) VCG_FOLD_LAUNCH(7, 2);
  else if (d->kw == 3 && nb4 == 1) VCG_FOLD_LAUNCH(3, 1);
  else if (d->kw == 3) VCG_FOLD_LAUNCH(3, 2);
  else if (nb4 == 1) VCG_FOLD_LAUNCH(2, 1);
  else VCG_FOLD_LAUNCH(2, 2);
#undef VCG_FOLD_LAUNCH
  VCG_CHECK_LAUNCH("conv_tc_fold_kernel");
  return VCG_OK;
}
