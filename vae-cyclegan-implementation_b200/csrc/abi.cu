// abi.cu -- C-ABI glue: error reporting, TMA descriptor encoding, convolution dispatch.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

std::atomic<long long> g_vcg_launches{0};
static thread_local char g_err[512] = "";

void vcg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int vcg_version(void) { return VCG_ABI_VERSION; }
extern "C" const char* vcg_last_error(void) { return g_err; }
extern "C" long long vcg_launch_count(void) { return g_vcg_launches.load(); }

static std::atomic<int> g_sm_budget{0};
extern "C" int vcg_set_sm_budget(int32_t sms) {
  VCG_REQUIRE(sms >= 0, VCG_E_INVALID, "set_sm_budget: negative budget %d", sms);
  g_sm_budget.store(sms);
  return VCG_OK;
}
static std::atomic<int> g_l2_prefetch{1};   // measured best on B200 (tools/bench_xform.py: 1 > 2 >> 4, 8)
extern "C" int vcg_set_l2_prefetch(int32_t chunks) {
  VCG_REQUIRE(chunks >= 0 && chunks <= 16, VCG_E_INVALID, "set_l2_prefetch: %d chunks (0..16)", chunks);
  g_l2_prefetch.store(chunks);
  return VCG_OK;
}
int vcg_l2_prefetch() { return g_l2_prefetch.load(std::memory_order_relaxed); }

int vcg_gemm_sms() {
  const int phys = vcg_num_sms(), b = g_sm_budget.load(std::memory_order_relaxed);
  // pairs of CTAs (cta_group::2 kernels) need an even count
  return (b > 0 && b < phys) ? (b < 2 ? 2 : b) : phys;
}

// ---------------------------------------------------------------- cuTensorMapEncodeTiled
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    else (void)cudaGetLastError();
  }
  return fn;
}

static int encode_tmap_impl(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                            const uint32_t* box, const char* what, CUtensorMapSwizzle swz);

int vcg_encode_tmap(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const char* what) {
  return encode_tmap_impl(map, base, rank, dims, strides_bytes, box, what, CU_TENSOR_MAP_SWIZZLE_128B);
}
// same without shared-memory swizzle: the box lands as plain contiguous rows (inner box = 16 bytes)
int vcg_encode_tmap_linear(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, const char* what) {
  return encode_tmap_impl(map, base, rank, dims, strides_bytes, box, what, CU_TENSOR_MAP_SWIZZLE_NONE);
}

static int encode_tmap_impl(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                            const uint32_t* box, const char* what, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  VCG_REQUIRE(fn, VCG_E_DRIVER, "%s: cuTensorMapEncodeTiled is not available from the driver", what);
  VCG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, VCG_E_INVALID, "%s: base pointer not 16-byte aligned", what);
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gd, gs,
                  bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    int o = 0;
    for (int i = 0; i < rank; ++i) o += snprintf(buf + o, sizeof(buf) - o, "%llu/%u ", (unsigned long long)dims[i], box[i]);
    for (int i = 0; i + 1 < rank; ++i) o += snprintf(buf + o, sizeof(buf) - o, "s%llu ", (unsigned long long)strides_bytes[i]);
    vcg_set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d) dims/box: %s", what, static_cast<int>(r), buf);
    return VCG_E_DRIVER;
  }
  return VCG_OK;
}

// test hook: encode only, so that descriptor legality (e.g. overlapping "window" strides) can be probed
extern "C" int vcg_probe_tmap(const void* base, int32_t rank, const uint64_t* dims, const uint64_t* strides_bytes,
                              const uint32_t* box) {
  CUtensorMap m;
  return vcg_encode_tmap(&m, base, rank, dims, strides_bytes, box, "probe");
}

// ---------------------------------------------------------------- convolution dispatch
int vcg_conv_fwd_tc(const vcg_conv_desc*, const void*, const void*, const float*, void*, float*, int, cudaStream_t);
int vcg_conv_wgrad_tc(const vcg_conv_desc*, const void*, const void*, int, int, float*, cudaStream_t);
bool vcg_wgrad2_supported(const vcg_conv_desc*, int);
int vcg_conv_wgrad_tc2(const vcg_conv_desc*, const void*, const void*, int, int, float*, cudaStream_t);
int vcg_conv_fwd_simt(const vcg_conv_desc*, int, const void*, const void*, const float*, void*, int, cudaStream_t);
int vcg_conv_wgrad_simt(const vcg_conv_desc*, int, const void*, const void*, int, int, float*, cudaStream_t);
int vcg_conv_wgrad_thin(const vcg_conv_desc*, const void*, const void*, int, int, float*, cudaStream_t);
bool vcg_wgrad_thin_supported(const vcg_conv_desc*);
int vcg_conv_wgrad_fold(const vcg_conv_desc*, const void*, const void*, int, int, float*, cudaStream_t);
bool vcg_wgrad_fold_supported(const vcg_conv_desc*, int, int);


extern "C" int vcg_conv_fwd(const vcg_conv_desc* d, const void* x, const void* w, const float* bias, void* y,
                            float* stats, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VCG_REQUIRE(d && x && w && y, VCG_E_INVALID, "conv_fwd: null argument");
  VCG_REQUIRE(d->dtype == VCG_F32 || d->dtype == VCG_BF16, VCG_E_UNSUPPORTED, "conv_fwd: dtype %d", d->dtype);
  if (d->dtype == VCG_F32) {
    VCG_REQUIRE(!(d->stats && stats), VCG_E_UNSUPPORTED, "conv_fwd: the SIMT path has no fused statistics; use vcg_in_stats");
    return vcg_conv_fwd_simt(d, d->dtype, x, w, bias, y, d->out_f32, stream);
  }
  VCG_REQUIRE(d->act <= VCG_ACT_LEAKY, VCG_E_UNSUPPORTED,
              "conv_fwd: activation %d is not fused in the tensor-core epilogues (Tanh / Sigmoid run in vcg_xform_fwd)", d->act);
  return vcg_conv_fwd_tc(d, x, w, bias, y, stats, d->out_f32, stream);
}

extern "C" int vcg_conv_wgrad(const vcg_conv_desc* d, const void* x, const void* dy, int32_t dy_halo, int32_t dy_c,
                              float* dw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VCG_REQUIRE(d && x && dy && dw, VCG_E_INVALID, "conv_wgrad: null argument");
  VCG_REQUIRE(d->dtype == VCG_F32 || d->dtype == VCG_BF16, VCG_E_UNSUPPORTED, "conv_wgrad: dtype %d", d->dtype);
  // degenerate M (cout < 16, i.e. the 64->3 output convolution) and feature maps that cannot be cut
  // into 64-pixel TMA boxes (inputs smaller than 256x256) stay on the SIMT kernel
  const int ho = d->hp - d->kh + 1, wo = d->wp - d->kw + 1;
  const bool tileable = wo >= 64 ? (wo % 64 == 0) : (64 % wo == 0 && ho % (64 / wo) == 0);
  // the two 7x7 image-side layers (64->3 and 3->64): horizontal taps folded into N, row pairs stacked into M
  if (d->dtype == VCG_BF16 && vcg_wgrad_fold_supported(d, dy_halo, dy_c))
    return vcg_conv_wgrad_fold(d, x, dy, dy_halo, dy_c, dw, stream);
  // bf16 mode, cout <= 4 (the 64->3 output convolution) on maps the fold kernel does not take: HMMA kernel
  if (d->dtype == VCG_BF16 && vcg_wgrad_thin_supported(d)) return vcg_conv_wgrad_thin(d, x, dy, dy_halo, dy_c, dw, stream);
  if (d->dtype == VCG_F32) return vcg_conv_wgrad_simt(d, d->dtype, x, dy, dy_halo, dy_c, dw, stream);
  if (d->cout < 16 || !tileable) {
    // no tensor-core kernel takes this shape.  The SIMT kernel is ~50x slower, so it runs only when the caller asked
    // for it (VCG_WGRAD_ALLOW_SIMT in desc->flat: maps below 8x8, i.e. inputs smaller than the 256x256 the networks
    // are built for); anything else is an error, not a silent cliff
    VCG_REQUIRE(d->flat & VCG_WGRAD_ALLOW_SIMT, VCG_E_UNSUPPORTED,
                "conv_wgrad: no tensor-core kernel for cout=%d on a %dx%d map (set VCG_WGRAD_ALLOW_SIMT to run the SIMT kernel)",
                d->cout, ho, wo);
    return vcg_conv_wgrad_simt(d, d->dtype, x, dy, dy_halo, dy_c, dw, stream);
  }
  // 256-channel-multiple outputs: CTA-pair kernel (256 x 256 tiles, half of the X tile per CTA)
  if (vcg_wgrad2_supported(d, dy_c)) return vcg_conv_wgrad_tc2(d, x, dy, dy_halo, dy_c, dw, stream);
  return vcg_conv_wgrad_tc(d, x, dy, dy_halo, dy_c, dw, stream);
}
