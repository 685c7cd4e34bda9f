// common.cuh -- shared device helpers (PTX wrappers for mbarrier / TMA / tcgen05) and host-side
// error plumbing for libvcg_b200.so.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/vcg.h"

// ------------------------------------------------------------------ host-side error plumbing
void vcg_set_error(const char* fmt, ...);
extern std::atomic<long long> g_vcg_launches;
#define VCG_COUNT_LAUNCH() (g_vcg_launches.fetch_add(1, std::memory_order_relaxed))
#define VCG_CHECK_LAUNCH(name)                                                     \
  do {                                                                             \
    cudaError_t e__ = cudaPeekAtLastError();                                       \
    VCG_COUNT_LAUNCH();                                                            \
    if (e__ != cudaSuccess) {                                                      \
      vcg_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));       \
      (void)cudaGetLastError();                                                    \
      return VCG_E_CUDA;                                                           \
    }                                                                              \
  } while (0)
#define VCG_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      vcg_set_error(__VA_ARGS__);     \
      return code;                    \
    }                                 \
  } while (0)

static inline int vcg_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// SMs a persistent tensor-core kernel may occupy: the physical count, or the budget set through vcg_set_sm_budget()
// (two independent network passes running on two streams each take half of the machine; abi.cu)
int vcg_gemm_sms();
// read-ahead distance (loop iterations) of the streaming transform kernels, vcg_set_l2_prefetch() (abi.cu)
int vcg_l2_prefetch();

// HBM -> L2 read-ahead of a contiguous byte range (16-byte aligned start; the size is rounded down to 16 bytes)
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, long long bytes) {
  const unsigned b = static_cast<unsigned>(bytes) & ~15u;
  if (b) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(b) : "memory");
}

// ------------------------------------------------------------------ element traits
template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 8-element vector load/store as floats (32 B for fp32, 16 B for bf16); p must be aligned
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}

// the activations of every shipped network: ReLU / LeakyReLU(0.2) / identity (hot loops, conv epilogues)
__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == VCG_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == VCG_ACT_LEAKY) return v > 0.f ? v : 0.2f * v;
  return v;
}
// ... plus CaSb's Tanh / Sigmoid choices (Networks.py:63-71; no shipped network uses them).  Kept out of the conv
// epilogues: inlining tanhf / expf into bias_act32 took conv_tc2_kernel from 128 to 162 registers.  They run in the
// transform pass (vcg_xform_fwd) and in the SIMT parity kernels.
__device__ __forceinline__ float act_apply_any(float v, int act) {
  if (act == VCG_ACT_TANH) return tanhf(v);
  if (act == VCG_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  return act_apply(v, act);
}
// derivative expressed through the activation OUTPUT o (ReLU/LeakyReLU keep the sign; tanh' = 1 - o^2, sigmoid' = o(1-o))
__device__ __forceinline__ float act_grad(float o, int act) {
  if (act == VCG_ACT_RELU) return o > 0.f ? 1.f : 0.f;
  if (act == VCG_ACT_LEAKY) return o > 0.f ? 1.f : 0.2f;
  if (act == VCG_ACT_TANH) return 1.f - o * o;
  if (act == VCG_ACT_SIGMOID) return o * (1.f - o);
  return 1.f;
}
// derivative expressed through the activation INPUT z (where the saved tensor is the pre-activation value)
__device__ __forceinline__ float act_grad_in(float z, int act) {
  if (act == VCG_ACT_TANH || act == VCG_ACT_SIGMOID) return act_grad(act_apply_any(z, act), act);
  return act_grad(z, act);
}
// branch-free form for inner loops: slope = act_slope(act) once per kernel, then one compare + select per element
__device__ __forceinline__ float act_slope(int act) { return act == VCG_ACT_RELU ? 0.f : (act == VCG_ACT_LEAKY ? 0.2f : 1.f); }
__device__ __forceinline__ float act_grad_s(float o, float slope) { return o > 0.f ? 1.f : slope; }
// PyTorch 'reflect' index (edge not repeated) for t in [-p, L-1+p], p < L
__device__ __forceinline__ int reflect_idx(int t, int L) {
  t = t < 0 ? -t : t;
  return t >= L ? 2 * (L - 1) - t : t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ sm_100a PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA tiled loads (global -> shared, completion on an mbarrier); SASS: UTMALDG
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// warp index as a value the compiler knows to be warp-uniform (role branches become uniform branches and the
// code inside can use the uniform datapath: no per-instruction R2UR waterfall around UTCHMMA / UTMALDG)
__device__ __forceinline__ int warp_idx_uniform() { return __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0); }
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
// one elected lane of a fully active warp
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n}"
      : "=r"(pred));
  return pred != 0;
}

// ---- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; SASS: UTCHMMA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t = lane t of the quadrant)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// UMMA shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4  [16,30) LBO>>4  [32,46) SBO>>4  [46,48) version=1  [61,64) layout (2 = SWIZZLE_128B)
// K-major : rows of 64 bf16 (128 B), 8-row atoms 1024 B apart (SBO); LBO unused.
// MN-major: 64-element MN atoms (128 B) x 8 K rows = 1024 B; SBO = stride between 8-row K groups,
//           LBO = stride between 64-element MN atoms.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major operand WITHOUT swizzle ("interleaved" canonical layout, cute ((8,m),(T,2)):((1T,SBO),(1,LBO))): rows of a
// core matrix are 16 bytes apart, 8-row groups SBO apart, the two 16-byte K chunks of one K=16 step LBO apart.
__device__ __forceinline__ uint64_t umma_desc_linear(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor for kind::f16: bf16 x bf16 -> fp32 (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ inline uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ |
         (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// ------------------------------------------------------------------ epilogue helpers shared by the conv kernels

// Sum over the 32 lanes of a warp of v[j] for each j: afterwards lane l holds column l in v[0].
__device__ __forceinline__ void transposed_warp_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16, n = 32; s >= 1; s >>= 1, n >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = upper ? v[i + n / 2] : v[i];
      const float send = upper ? v[i] : v[i + n / 2];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

// bias + activation on 32 accumulator columns.  The bias of the current n-tile sits in a warp-private shared-memory
// copy (8 broadcast LDS.128 instead of 32 global loads), the activation switch is hoisted out of the element loop
// and columns >= ncols (beyond the logical output channels) are zeroed: ~3 instructions per element instead of ~10
// (short-K tiles are bound by the instruction count of this epilogue).
__device__ __forceinline__ void bias_act32(float (&v)[32], const float* __restrict__ sb, int act, int ncols) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(sb + 4 * q);
    v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
  }
  if (act == VCG_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (act == VCG_ACT_LEAKY) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
  }
  if (ncols < 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) if (j >= ncols) v[j] = 0.f;
  }
}
// warp-private copy of bias[n0 .. n0+bn) (zeros where there is no bias / beyond cout)
__device__ __forceinline__ void load_bias_tile(float* sb, const float* __restrict__ bias, int n0, int bn, int cout, int lane) {
  __syncwarp();
  for (int i = lane; i < bn; i += 32) sb[i] = (bias && n0 + i < cout) ? __ldg(bias + n0 + i) : 0.f;
  __syncwarp();
}


// ------------------------------------------------------------------ host: TMA descriptor encode
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
int vcg_encode_tmap(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes /*rank-1*/, const uint32_t* box, const char* what);
int vcg_encode_tmap_linear(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes /*rank-1*/, const uint32_t* box, const char* what);
