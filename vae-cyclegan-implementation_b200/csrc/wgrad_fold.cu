// wgrad_fold.cu -- weight gradient of the two 7x7 image-side convolutions on tcgen05 tensor cores, sm_100a:
//   (A) 64 -> 3 output convolution of the decoder (Networks.py:192): thick tensor T = saved input x (64 ch),
//       thin tensor U = dY (3 -> 8 ch)
//   (B) 3 -> 64 input convolution of the encoder (Networks.py:158): T = dY (64 ch), U = saved input x (3 -> 8 ch)
//
//   dW[co, kh, kw, ci] = sum_{n,h,w} dY[n,h,w,co] * X[n, h+kh, w+kw, ci]
//
// One side of the product has only 8 physical channels, so a plain GEMM (M = cout, N = a run of the filter row,
// K = pixels) would be >90 % padding.  Instead the 7 HORIZONTAL taps are folded into N through an
// overlapping-stride ("window") TMA map of the thin tensor -- 64 consecutive elements from pixel p are the 8
// channels of pixels p..p+7 -- and two consecutive rows of the thick tensor are stacked into M = 2 x 64:
//
//   D_d[(r, c64), (j, c8)] += sum_{p in strip} T[tau + r, p, c64] * U[tau + d, p + j, c8],   d = 0..7
//
// Accumulator d pairs thick row tau+r with thin row tau+d, i.e. the vertical tap is a function of (d, r) and the
// horizontal tap a function of j.  Eight accumulators (8 x 64 = 512 TMEM columns) stay resident for the CTA's whole
// pixel range; thin rows are loaded once and reused by four consecutive row pairs (ring of stages holding one T
// pair + the two NEW thin rows each); the result is added to the packed fp32 gradient with red.global.add once
// per CTA.  Both operands are MN-major (pixels are K), exactly as TMA lands NHWC boxes.
//
// warp roles: 0 = TMA producer, 1 = MMA issuer (warp-uniform, elected lane), 2 = TMEM allocator, 4..7 = epilogue.
#include "common.cuh"

namespace {

struct WfArgs {
  int n_img, pairs, strips;        // row pairs per image, 64-pixel strips per row
  int t_row0, t_col0;              // thick-tensor buffer offset of logical (row 0, pixel 0)
  int last_ksteps;                 // K=16 steps that hold real pixels in the last strip (1..4)
  int mode;                        // 0: (A) T = x, U = dY;  1: (B) T = dY, U = x
  int cout, kwc_pad, row_len;      // dw geometry: [rows][kh][kwc_pad], row_len = kh * kwc_pad
  int units, stages;
  uint32_t idesc;
  float* dw;
};

constexpr int kStage = 32768;      // T pair box (2 x 8 KB) + two thin rows (2 x 8 KB)
constexpr int kSub = 8192;
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads, 1)
wgrad_fold_kernel(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmU, const WfArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int S = p.stages;
  const uint32_t bar0 = base + S * kStage;                 // full[S], empty[S], tfull
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S * kStage + (2 * S + 1) * 8);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * S);

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmT); tma_prefetch_desc(&tmU); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);

  // contiguous range of (image, strip, pair) units, pair fastest: consecutive pairs of one (image, strip) form a
  // "run" that shares thin rows through the stage ring; every run starts with 3 warm-up stages (thin rows only)
  const int per_cta = (p.units + gridDim.x - 1) / gridDim.x;
  const int u_begin = blockIdx.x * per_cta;
  const int u_end = min(p.units, u_begin + per_cta);

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0; uint32_t phase = 0;
    int u = u_begin;
    while (u < u_end) {
      const int col = u / p.pairs, t0 = u - col * p.pairs;            // col = image * strips + strip
      const int img = col / p.strips, strip = col - img * p.strips;
      const int t1 = min(p.pairs, t0 + (u_end - u));                  // run = pairs [t0, t1)
      const int w0 = strip * 64;
      for (int t = t0 - 3; t < t1; ++t) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = base + stage * kStage;
        if (elect_one_sync()) {
          const bool warm = t < t0;
          mbar_expect_tx(full_bar(stage), warm ? 2 * kSub : kStage);
          if (!warm) tma_load_4d(sa, &tmT, full_bar(stage), 0, w0 + p.t_col0, 2 * t + p.t_row0, img);
          // the two thin rows first needed by pair t: tau + 6 and tau + 7 (tau = 2t)
          tma_load_4d(sa + 2 * kSub, &tmU, full_bar(stage), 0, w0, 2 * t + 6, img);
          tma_load_4d(sa + 3 * kSub, &tmU, full_bar(stage), 0, w0, 2 * t + 7, img);
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
      u += t1 - t0;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0; uint32_t phase = 0;
    bool started = false;
    int u = u_begin;
    while (u < u_end) {
      const int col = u / p.pairs, t0 = u - col * p.pairs;
      const int strip = col % p.strips;
      const int t1 = min(p.pairs, t0 + (u_end - u));
      const int nk = (strip == p.strips - 1) ? p.last_ksteps : 4;
      for (int t = t0 - 3; t < t1; ++t) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (t >= t0) {
          const uint64_t ad = umma_desc_sw128(base + stage * kStage, kSub, 1024);
          if (elect_one_sync()) {
#pragma unroll
            for (int d = 0; d < 8; ++d) {
              // thin row tau + d lives in the stage of pair t - 3 + (d >> 1), sub-row d & 1
              int sd = stage - 3 + (d >> 1);
              if (sd < 0) sd += S;
              const uint64_t bd = umma_desc_sw128(base + sd * kStage + (2 + (d & 1)) * kSub, kSub, 1024);
              for (int k = 0; k < nk; ++k)          // K step = 16 pixel rows = 2048 B
                umma_bf16(tmem_base + d * 64, ad + 128 * k, bd + 128 * k, p.idesc, (started || k > 0) ? 1u : 0u);
            }
            // the oldest of the four stages this pair read is free once these MMAs are done
            int so = stage - 3;
            if (so < 0) so += S;
            umma_commit(empty_bar(so));
            if (t == t1 - 1) {                    // end of the run: nobody will read the last three stages again
              for (int b = 2; b >= 0; --b) { int sr = stage - b; if (sr < 0) sr += S; umma_commit(empty_bar(sr)); }
            }
          }
          __syncwarp();
          started = true;
        }
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
      u += t1 - t0;
    }
    if (elect_one_sync()) umma_commit(tfull_bar);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue (once per CTA): TMEM -> red.global.add into the packed gradient
    if (u_begin < u_end) {
      const int quad = warp & 3;
      const int r = quad >> 1;                         // which row of the pair this TMEM lane belongs to
      const int c64 = (quad & 1) * 32 + lane;          // thick-tensor channel of this lane
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      for (int d = 0; d < 8; ++d) {
        const int kh = p.mode == 0 ? 6 - d + r : d - r;
        if (kh < 0 || kh > 6) continue;               // (warp-uniform: r depends on the warp only)
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(d * 64);
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          if (p.mode == 1) {
            // (B) rows = (r, co), columns = kw * 8 + ci: one contiguous filter row segment
            float* dst = p.dw + (static_cast<size_t>(c64) * 7 + kh) * p.kwc_pad + c0;
            if (c64 < p.cout) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                             "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
            }
          } else {
            // (A) rows = (r, ci), columns = j * 8 + co with kw = 6 - j
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = c0 / 8 + jj;
              if (j > 6) continue;
#pragma unroll
              for (int co = 0; co < 8; ++co)
                if (co < p.cout)
                  atomicAdd(p.dw + (static_cast<size_t>(co) * 7 + kh) * p.kwc_pad + (6 - j) * 64 + c64, __uint_as_float(v[jj * 8 + co]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// which of the two supported shapes (or -1)
int fold_mode(const vcg_conv_desc* d, int dy_c) {
  if (d->dtype != VCG_BF16 || d->kh != 7 || d->kw != 7) return -1;
  if (d->c == 64 && d->cout <= 8 && dy_c == 8 && d->kwc_pad == 7 * 64) return 0;
  if (d->c == 8 && d->cout == 64 && dy_c == 64 && d->kwc_pad == 64) return 1;
  return -1;
}

}  // namespace

bool vcg_wgrad_fold_supported(const vcg_conv_desc* d, int dy_halo, int dy_c) {
  const int wo = d->wp - d->kw + 1;
  return fold_mode(d, dy_c) >= 0 && dy_halo == 6 && wo >= 64;
}

int vcg_conv_wgrad_fold(const vcg_conv_desc* d, const void* x, const void* dy, int dy_halo, int dy_c, float* dw,
                        cudaStream_t stream) {
  const int mode = fold_mode(d, dy_c);
  VCG_REQUIRE(mode >= 0 && dy_halo == 6, VCG_E_UNSUPPORTED, "wgrad_fold: unsupported shape");
  const int ho = d->hp - 6, wo = d->wp - 6;
  const int hpd = ho + 12, wpd = wo + 12;
  WfArgs a{};
  a.n_img = d->n; a.mode = mode;
  // logical extent of the thick tensor the pairs / strips walk over
  const int t_rows = mode == 0 ? d->hp : ho, t_cols = mode == 0 ? d->wp : wo;
  a.pairs = (t_rows + 1) / 2;
  a.strips = (t_cols + 63) / 64;
  a.t_row0 = mode == 0 ? 0 : 6;
  a.t_col0 = mode == 0 ? 0 : 6;
  const int last = t_cols - (a.strips - 1) * 64;
  a.last_ksteps = (last + 15) / 16;
  a.cout = d->cout; a.kwc_pad = d->kwc_pad; a.row_len = 7 * d->kwc_pad;
  a.units = d->n * a.strips * a.pairs;
  a.stages = 6;
  a.idesc = umma_idesc_bf16(128, 64, 1, 1);
  a.dw = dw;

  CUtensorMap tmT, tmU;
  const uint64_t es = 2;
  {
    const void* tbase = mode == 0 ? x : dy;
    const uint64_t tw = mode == 0 ? d->wp : wpd, th = mode == 0 ? d->hp : hpd;
    uint64_t dims[4] = {64, tw, th, static_cast<uint64_t>(d->n)};
    uint64_t str[3] = {64 * es, tw * 64 * es, th * tw * 64 * es};
    uint32_t box[4] = {64, 64, 2, 1};
    int rc = vcg_encode_tmap(&tmT, tbase, 4, dims, str, box, "wgrad_fold T");
    if (rc) return rc;
  }
  {
    // window map of the thin tensor: 56 = 7 pixels x 8 channels contiguous elements from every pixel position
    const void* ubase = mode == 0 ? dy : x;
    const uint64_t uw = mode == 0 ? wpd : d->wp, uh = mode == 0 ? hpd : d->hp;
    uint64_t dims[4] = {56, uw - 6, uh, static_cast<uint64_t>(d->n)};
    uint64_t str[3] = {8 * es, uw * 8 * es, uh * uw * 8 * es};
    uint32_t box[4] = {64, 64, 1, 1};
    int rc = vcg_encode_tmap(&tmU, ubase, 4, dims, str, box, "wgrad_fold U (window)");
    if (rc) return rc;
  }
  const size_t smem = static_cast<size_t>(a.stages) * kStage + 2048;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "wgrad_fold: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int sms = vcg_gemm_sms();
  int grid = sms;
  if (grid > a.units / 8) grid = a.units / 8 > 0 ? a.units / 8 : 1;     // keep the 3-stage warm-up per run amortised
  wgrad_fold_kernel<<<grid, kThreads, smem, stream>>>(tmT, tmU, a);
  VCG_CHECK_LAUNCH("wgrad_fold_kernel");
  return VCG_OK;
}
