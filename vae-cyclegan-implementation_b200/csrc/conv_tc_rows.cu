// conv_tc_rows.cu -- tcgen05 implicit-GEMM convolution for THIN outputs (cout_pad <= 32) on wide maps:
// the 64->3 7x7 output convolution (Networks.py:192) and the data gradient of the 3->64 7x7 input
// convolution (Networks.py:158).
//
// With N = 16 the generic kernel (conv_tc.cu) is bound by L2->SM traffic, not by the tensor pipe: every
// 128-pixel output tile re-reads its A operand once per tap (49 x 16 KB) and the whole filter.  Here a CTA
// computes a block of R = 16 output rows x 128 pixels with R accumulators in TMEM (R x 16 columns, double
// buffered = 512 columns): an input row strip shifted by kw is loaded ONCE (one TMA box, one pipeline
// stage) and feeds the accumulators of all output rows r = rho - kh it contributes to, and the complete
// filter (cout_pad x K bf16, ~100 KB) stays resident in shared memory.  A traffic drops from
// R*kh*kw to (R+kh-1)*kw boxes per block (5.7x for 7x7), B traffic to one load per CTA.
#include "common.cuh"

namespace {

struct RowsArgs {
  int n_img, ho, wo, tiles_w, blocks_h, R;
  int kh, kw, cchunks, nb_chunks;
  int bn, cout, out_c, act, out_f32;
  int num_tiles, stages;
  uint32_t idesc;
  const float* bias;
  void* out;
};

constexpr int kAStage = 16384;
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads, 1)
conv_tc_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const RowsArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int S = p.stages;
  const uint32_t b_bytes = static_cast<uint32_t>(p.bn) * 128u;
  const uint32_t bres = base + S * kAStage;                                  // resident filter chunks
  const uint32_t bar0 = bres + p.nb_chunks * b_bytes;                        // full[S], empty[S], tfull[2], tempty[2], bready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S * kAStage + p.nb_chunks * b_bytes + (2 * S + 5) * 8);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 + a); };
  const uint32_t bready_bar = bar0 + 8u * (2 * S + 4);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t acc_cols = static_cast<uint32_t>(p.R * p.bn);
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * acc_cols) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    mbar_init(bready_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(smem_u32(tmem_slot), tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  const int blocks_per_img = p.tiles_w * p.blocks_h;
  const int per_cta = (p.num_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_begin = blockIdx.x * per_cta;
  const int tile_end = min(p.num_tiles, tile_begin + per_cta);
  const int n_rho = p.R + p.kh - 1;

  if (warp == 0) {
    if (lane == 0 && tile_begin < tile_end) {
      // ---- resident filter: all K chunks once
      mbar_expect_tx(bready_bar, p.nb_chunks * b_bytes);
      for (int c = 0; c < p.nb_chunks; ++c) tma_load_2d(bres + c * b_bytes, &tmB, bready_bar, c * 64, 0);
      int stage = 0; uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int img = tile / blocks_per_img, rem = tile - img * blocks_per_img;
        const int h0 = (rem / p.tiles_w) * p.R, w0 = (rem % p.tiles_w) * 128;
        for (int rho = 0; rho < n_rho; ++rho)
          for (int kwi = 0; kwi < p.kw; ++kwi)
            for (int q = 0; q < p.cchunks; ++q) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              mbar_expect_tx(full_bar(stage), kAStage);
              tma_load_4d(base + stage * kAStage, &tmA, full_bar(stage), q * 64, w0 + kwi, h0 + rho, img);
              if (++stage == S) { stage = 0; phase ^= 1u; }
            }
      }
    }
  } else if (warp == 1) {
    // warp-uniform loop, one elected lane issues (see conv_tc.cu)
    if (tile_begin < tile_end) {
      mbar_wait(bready_bar, 0);
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + static_cast<uint32_t>(as) * acc_cols;
        for (int rho = 0; rho < n_rho; ++rho)
          for (int kwi = 0; kwi < p.kw; ++kwi)
            for (int q = 0; q < p.cchunks; ++q) {
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint64_t ad = umma_desc_sw128(base + stage * kAStage, 16, 1024);
              const int r_lo = rho - (p.kh - 1) > 0 ? rho - (p.kh - 1) : 0;
              const int r_hi = rho < p.R - 1 ? rho : p.R - 1;
              if (elect_one_sync()) {
                for (int r = r_lo; r <= r_hi; ++r) {
                  const int khi = rho - r;
                  const uint64_t bd = umma_desc_sw128(bres + static_cast<uint32_t>((khi * p.kw + kwi) * p.cchunks + q) * b_bytes, 16, 1024);
                  const bool first = (khi == 0) && (kwi == 0) && (q == 0);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16(d0 + static_cast<uint32_t>(r * p.bn), ad + 2 * k, bd + 2 * k, p.idesc, (first && k == 0) ? 0u : 1u);
                }
                umma_commit(empty_bar(stage));
              }
              __syncwarp();
              if (++stage == S) { stage = 0; phase ^= 1u; }
            }
        if (elect_one_sync()) umma_commit(tfull_bar(as));
        __syncwarp();
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int as = 0; uint32_t aphase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int img = tile / blocks_per_img, rem = tile - img * blocks_per_img;
      const int h0 = (rem / p.tiles_w) * p.R, w0 = (rem % p.tiles_w) * 128;
      const int w = w0 + row;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      for (int r = 0; r < p.R; ++r) {
        const int h = h0 + r;
        if (h >= p.ho) break;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as) * acc_cols +
                               static_cast<uint32_t>(r * p.bn);
        const size_t pix = (static_cast<size_t>(img) * p.ho + h) * p.wo + w;
        for (int c0 = 0; c0 < p.bn; c0 += 16) {
          uint32_t rr[16];
          tmem_ld16(taddr + c0, rr);
          tmem_ld_wait();
          if (w < p.wo) {
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8) {
              const int col = c0 + g8 * 8;
              if (col < p.cout) {
                float t[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float x = __uint_as_float(rr[g8 * 8 + j]);
                  if (col + j < p.cout) { if (p.bias) x += __ldg(p.bias + col + j); x = act_apply(x, p.act); }
                  else x = 0.f;
                  t[j] = x;
                }
                if (p.out_f32) st8<float>(reinterpret_cast<float*>(p.out) + pix * p.out_c + col, t);
                else st8<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.out_c + col, t);
              }
            }
          }
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

bool vcg_conv_rows_supported(const vcg_conv_desc* d, bool has_stats) {
  const int wo = d->wp - d->kw + 1;
  if (has_stats || d->c % 64 != 0 || d->cout_pad > 32 || wo < 128) return false;
  const size_t bres = static_cast<size_t>(d->kh) * d->kw * (d->c / 64) * d->cout_pad * 128;
  return bres + 4 * kAStage + 2048 <= 227 * 1024;
}

int vcg_conv_fwd_tc_rows(const vcg_conv_desc* d, const void* x, const void* w, const float* bias, void* y, int out_f32,
                         cudaStream_t stream) {
  const int ho = d->hp - d->kh + 1, wo = d->wp - d->kw + 1;
  RowsArgs a{};
  a.n_img = d->n; a.ho = ho; a.wo = wo;
  a.bn = d->cout_pad;
  a.R = 256 / a.bn;
  if (a.R > 16) a.R = 16;
  if (a.R > ho) a.R = ho;
  a.tiles_w = (wo + 127) / 128;
  a.blocks_h = (ho + a.R - 1) / a.R;
  a.kh = d->kh; a.kw = d->kw; a.cchunks = d->c / 64; a.nb_chunks = d->kh * d->kw * a.cchunks;
  a.cout = d->cout; a.out_c = d->out_c; a.act = d->act; a.out_f32 = out_f32;
  a.num_tiles = d->n * a.tiles_w * a.blocks_h;
  a.idesc = umma_idesc_bf16(128, a.bn, 0, 0);
  a.bias = bias; a.out = y;
  const size_t bres = static_cast<size_t>(a.nb_chunks) * a.bn * 128;
  int stages = static_cast<int>((227 * 1024 - 2048 - bres) / kAStage);
  if (stages > 7) stages = 7;
  VCG_REQUIRE(stages >= 2, VCG_E_UNSUPPORTED, "conv_tc_rows: filter does not fit shared memory");
  a.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * kAStage + bres + 2048;

  CUtensorMap tmA, tmB;
  const uint64_t es = 2, pix = d->c * es, row = d->wp * pix, img = d->hp * row;
  uint64_t dims[4] = {static_cast<uint64_t>(d->c), static_cast<uint64_t>(d->wp), static_cast<uint64_t>(d->hp),
                      static_cast<uint64_t>(d->n)};
  uint64_t strides[3] = {pix, row, img};
  uint32_t box[4] = {64, 128, 1, 1};
  int rc = vcg_encode_tmap(&tmA, x, 4, dims, strides, box, "conv_tc_rows A");
  if (rc) return rc;
  const uint64_t ktot = static_cast<uint64_t>(d->kh) * d->kwc_pad;
  uint64_t bdims[2] = {ktot, static_cast<uint64_t>(d->cout_pad)};
  uint64_t bstr[1] = {ktot * es};
  uint32_t bbox[2] = {64, static_cast<uint32_t>(a.bn)};
  rc = vcg_encode_tmap(&tmB, w, 2, bdims, bstr, bbox, "conv_tc_rows B");
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "conv_tc_rows: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int sms = vcg_num_sms();
  const int grid = a.num_tiles < sms ? a.num_tiles : sms;
  conv_tc_rows_kernel<<<grid, kThreads, smem, stream>>>(tmA, tmB, a);
  VCG_CHECK_LAUNCH("conv_tc_rows_kernel");
  return VCG_OK;
}
