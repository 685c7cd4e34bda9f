// xform.cu -- the memory-bound passes between convolutions (HBM-roofline kernels), sm_100a.
//
//  * vcg_in_stats / vcg_in_finalize : InstanceNorm statistics (nn.InstanceNorm2d, Networks.py:61,88,
//    102,105,123; executed by ATen native_batch_norm in the reference)
//  * vcg_xform_fwd : normalise -> ReLU/LeakyReLU -> +residual -> PixelShuffle / PixelUnshuffle /
//    space-to-depth addressing -> reflect halo -> store (Networks.py:76-81, 91-96, 108-116, 126-131)
//  * vcg_xform_bwd_gather / _norm : the exact adjoint: fold the reflect halo, invert the shuffle, apply the
//    activation derivative and the InstanceNorm backward (two per-(n,c) reductions), produce the
//    zero-haloed dY the data-/weight-gradient GEMMs consume, and the bias gradient.
//  * vcg_pack_nchw / vcg_unpack_nchw : NCHW fp32 API tensors <-> NHWC kernel tensors.
//
// All kernels move 8 channels (16 B bf16 / 32 B fp32) per thread with consecutive threads on
// consecutive channel groups, so every warp access is a contiguous 512 B / 1 KB run.
#include <stdlib.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------ statistics
template <typename T>
__global__ void __launch_bounds__(256)
in_stats_kernel(const T* __restrict__ y, int hw, int c, int c_pitch, int pix_per_block, double* __restrict__ acc) {
  // grid: (pixel chunks, n, channel-group chunks of 32 groups)
  // thread layout: cgl = min(32, c/8) channel-group lanes x 256/cgl pixel lanes (all 256 threads busy for thin c)
  const int cgb = min(32, c / 8 - blockIdx.z * 32);
  const int cgl = c / 8 < 32 ? c / 8 : 32;
  const int lanes = 256 / cgl;
  const int cg = threadIdx.x % cgl, pl = threadIdx.x / cgl;
  __shared__ double sacc[32 * 16];
  for (int i = threadIdx.x; i < 32 * 16; i += 256) sacc[i] = 0.0;
  __syncthreads();
  if (cg < cgb && pl < lanes) {
    const int ch = (blockIdx.z * 32 + cg) * 8;
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(hw, p0 + pix_per_block);
    double s1[8] = {}, s2[8] = {};
    const T* base = y + (static_cast<size_t>(blockIdx.y) * hw) * c_pitch + ch;
    for (int p = p0 + pl; p < p1; p += lanes) {
      float v[8];
      ld8<T>(base + static_cast<size_t>(p) * c_pitch, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] += v[j]; s2[j] += static_cast<double>(v[j]) * v[j]; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(&sacc[cg * 16 + j], s1[j]); atomicAdd(&sacc[cg * 16 + 8 + j], s2[j]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cgb * 16; i += 256) {
    const int g = i / 16, r = i % 16;
    const int ch = (blockIdx.z * 32 + g) * 8 + (r & 7);
    atomicAdd(acc + (static_cast<size_t>(blockIdx.y) * c + ch) * 2 + (r >> 3), sacc[i]);
  }
}

__global__ void in_finalize_d_kernel(const double* __restrict__ acc, int nc, int hw, float* __restrict__ mr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  const double mean = acc[2 * i] / hw;
  double var = acc[2 * i + 1] / hw - mean * mean;
  if (var < 0) var = 0;
  mr[2 * i] = static_cast<float>(mean);
  mr[2 * i + 1] = static_cast<float>(1.0 / sqrt(var + 1e-5));
}

__global__ void in_finalize_f_kernel(const float* __restrict__ acc, int nc, int hw, float* __restrict__ mr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  const float inv = 1.f / hw;
  const float mean = acc[2 * i] * inv;
  float var = acc[2 * i + 1] * inv - mean * mean;
  if (var < 0.f) var = 0.f;
  mr[2 * i] = mean;
  mr[2 * i + 1] = rsqrtf(var + 1e-5f);
}

// ------------------------------------------------------------------ forward transform
struct XfArgs {
  int n, h, w, c, src_c, norm, act, mode, pad, dst_c, hd, wd, cd;  // cd = logical dst channels
  int res_hp, res_wp, res_c, res_off, stats_hw;
};

// raw 8-element vectors: loads are issued into packed registers and unpacked only when consumed, so several
// independent 16-byte loads per thread can be in flight without a float register per element
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> { uint4 r; };
template <> struct Raw8<float> { float4 a, b; };
__device__ __forceinline__ void ldraw(const __nv_bfloat16* p, Raw8<__nv_bfloat16>& v) { v.r = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void ldraw(const float* p, Raw8<float>& v) {
  v.a = *reinterpret_cast<const float4*>(p); v.b = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void unpack(const Raw8<__nv_bfloat16>& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v.r);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void unpack(const Raw8<float>& v, float (&f)[8]) {
  f[0] = v.a.x; f[1] = v.a.y; f[2] = v.a.z; f[3] = v.a.w; f[4] = v.b.x; f[5] = v.b.y; f[6] = v.b.z; f[7] = v.b.w;
}

constexpr int kXfRows = 4;   // rows per unrolled iteration: 4 independent 16-byte loads in flight per thread

// per-channel statistics of the pending InstanceNorm for 8 consecutive channels: sc = rstd, sf = mean
// (applied as (v - mean) * rstd, the reference's operation order, in both precisions).
// stats_hw > 0: the buffer holds the raw {sum, sum of squares} pairs accumulated by the convolution epilogue and
// mean / rstd are derived here (same arithmetic as in_finalize_f_kernel), which saves one launch per layer.
__device__ __forceinline__ void finalize_pair(float& a, float& b, int stats_hw) {
  if (stats_hw > 0) {
    const float inv = 1.f / stats_hw;
    const float mean = a * inv;
    float var = b * inv - mean * mean;
    if (var < 0.f) var = 0.f;
    a = mean;
    b = rsqrtf(var + 1e-5f);
  }
}
__device__ __forceinline__ void load_scale_shift(const float* __restrict__ m, float (&sc)[8], float (&sf)[8], int stats_hw) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 t = *reinterpret_cast<const float4*>(m + 4 * q);   // {mean, rstd, mean, rstd} (or raw sums)
    finalize_pair(t.x, t.y, stats_hw);
    finalize_pair(t.z, t.w, stats_hw);
    sc[2 * q] = t.y; sf[2 * q] = t.x;
    sc[2 * q + 1] = t.w; sf[2 * q + 1] = t.z;
  }
}

// zero this block's share of the halo ring of an NHWC buffer [n, h+2*halo, w+2*halo, c_pitch] (8-channel group ch)
template <typename T>
__device__ __forceinline__ void clear_halo_share(T* __restrict__ buf, int n, int h, int w, int halo, int c_pitch, int ch,
                                                 int first, int step) {
  const int hp = h + 2 * halo, wp = w + 2 * halo;
  const int ring = hp * wp - h * w, top = halo * wp;
  float z[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) z[j] = 0.f;
  for (int r = first; r < ring; r += step) {
    int a, b;
    if (r < top) { a = r / wp; b = r - a * wp; }
    else if (r < 2 * top) { const int q = r - top; a = halo + h + q / wp; b = q % wp; }
    else { const int q = r - 2 * top; const int row = q / (2 * halo), k = q - row * 2 * halo;
           a = halo + row; b = k < halo ? k : w + k; }
    st8<T>(buf + ((static_cast<size_t>(n) * hp + a) * wp + b) * c_pitch + ch, z);
  }
}

template <typename T> __device__ __forceinline__ void st2(T* p, float a, float b);
template <> __device__ __forceinline__ void st2<float>(float* p, float a, float b) {
  *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
template <> __device__ __forceinline__ void st2<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// k-th padded coordinate t in [0, L+2p) whose reflect source is i (k=0 direct, 1 low mirror, 2 high mirror); -1 = none
__device__ __forceinline__ int mirror_k(int i, int L, int p, int k) {
  if (k == 0) return i + p;
  if (k == 1) return (i >= 1 && i <= p) ? p - i : -1;
  return (i >= L - 1 - p && i <= L - 2) ? p + 2 * (L - 1) - i : -1;
}

// Destination-driven (PLAIN / UNSHUFFLE / PAD_S2D).  grid = (x chunks of a destination row, row chunks, image);
// a thread owns one (destination column b, 8-channel group g) and walks down the rows of its chunk: the column
// mapping, the reflect index and the InstanceNorm scale/shift are computed once, every row costs one 16-byte load
// (+ one for the residual) and one 16-byte store, kXfRows rows in flight.
template <typename T, int MODE>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 3 : 2)
xform_fwd_kernel(const T* __restrict__ src, const float* __restrict__ mr, const T* __restrict__ res,
                 T* __restrict__ dst, const XfArgs p, int rows_per_block) {
  const int groups = p.dst_c / 8;
  const int xi = blockIdx.x * 256 + threadIdx.x;
  if (xi >= p.wd * groups) return;
  const int b = xi / groups, g = xi - b * groups, cd0 = g * 8;
  const int n = blockIdx.z;
  const int a0 = blockIdx.y * rows_per_block, a1 = min(p.hd, a0 + rows_per_block);
  const size_t orow = static_cast<size_t>(p.wd) * p.dst_c;
  T* out = dst + (static_cast<size_t>(n) * p.hd * p.wd + b) * p.dst_c + cd0;
  if (cd0 >= p.cd) {     // padding channel group of the destination
    float z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = 0.f;
    for (int a = a0; a < a1; ++a) st8<T>(out + a * orow, z);
    return;
  }
  int sub = 0, sc0 = cd0, sw;
  if (MODE == VCG_MODE_PLAIN) sw = reflect_idx(b - p.pad, p.w);
  else if (MODE == VCG_MODE_UNSHUFFLE) {
    sub = cd0 / p.c; sc0 = cd0 - sub * p.c;
    sw = 2 * reflect_idx(b - p.pad, p.w / 2) + (sub & 1);
  } else {
    sub = cd0 / p.c; sc0 = cd0 - sub * p.c;
    sw = reflect_idx(2 * b + (sub & 1) - p.pad, p.w);
  }
  auto src_row = [&](int a) {
    if (MODE == VCG_MODE_PLAIN) return reflect_idx(a - p.pad, p.h);
    if (MODE == VCG_MODE_UNSHUFFLE) return 2 * reflect_idx(a - p.pad, p.h / 2) + (sub >> 1);
    return reflect_idx(2 * a + (sub >> 1) - p.pad, p.h);
  };
  float sc[8], sf[8];
  if (p.norm) load_scale_shift(mr + (static_cast<size_t>(n) * p.c + sc0) * 2, sc, sf, p.stats_hw);
  const size_t srow = static_cast<size_t>(p.w) * p.src_c;
  const T* sbase = src + (static_cast<size_t>(n) * p.h * p.w + sw) * p.src_c + sc0;
  const size_t rrow = static_cast<size_t>(p.res_wp) * p.res_c;
  const T* rbase = res ? res + ((static_cast<size_t>(n) * p.res_hp + p.res_off) * p.res_wp + sw + p.res_off) * p.res_c + sc0
                       : nullptr;
  for (int a = a0; a < a1; a += kXfRows) {
    Raw8<T> vr[kXfRows], rr[kXfRows];
    int sr[kXfRows];
#pragma unroll
    for (int u = 0; u < kXfRows; ++u) {
      sr[u] = src_row(min(a + u, a1 - 1));
      ldraw(sbase + sr[u] * srow, vr[u]);
    }
    if (rbase) {
#pragma unroll
      for (int u = 0; u < kXfRows; ++u) ldraw(rbase + sr[u] * rrow, rr[u]);
    }
#pragma unroll
    for (int u = 0; u < kXfRows; ++u) {
      if (a + u >= a1) break;
      float v[8];
      unpack(vr[u], v);
      if (p.norm) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (v[j] - sf[j]) * sc[j];
      }
      if (p.act == VCG_ACT_RELU || p.act == VCG_ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_apply(v[j], p.act);
      } else if (p.act) {          // Tanh / Sigmoid (CaSb's other choices): cold path
#pragma unroll 1
        for (int j = 0; j < 8; ++j) v[j] = act_apply_any(v[j], p.act);
      }
      if (rbase) {
        float r[8];
        unpack(rr[u], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += r[j];
      }
      st8<T>(out + (a + u) * orow, v);
    }
  }
}

// PixelShuffle is source-driven: a thread reads 8 consecutive SOURCE channels (= 2 destination channels x 4
// sub-pixels) of one source pixel with one 16-byte load and scatters four channel pairs (plus their reflect
// mirrors) -- consecutive threads write consecutive pairs, so the stores coalesce too.
template <typename T>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 3 : 2)
xform_fwd_shuffle_kernel(const T* __restrict__ src, const float* __restrict__ mr, const T* __restrict__ res,
                         T* __restrict__ dst, const XfArgs p, int rows_per_block) {
  const int sgroups = p.c / 8;
  const int xi = blockIdx.x * 256 + threadIdx.x;
  if (xi >= p.w * sgroups) return;
  const int sw = xi / sgroups, sg = xi - sw * sgroups, sc0 = sg * 8;
  const int n = blockIdx.z;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(p.h, r0 + rows_per_block);
  const int H2 = 2 * p.h, W2 = 2 * p.w;
  float sc[8], sf[8];
  if (p.norm) load_scale_shift(mr + (static_cast<size_t>(n) * p.c + sc0) * 2, sc, sf, p.stats_hw);
  int cols[2][3];
#pragma unroll
  for (int jj = 0; jj < 2; ++jj)
#pragma unroll
    for (int k = 0; k < 3; ++k) cols[jj][k] = mirror_k(2 * sw + jj, W2, p.pad, k);
  const bool col_plain = (cols[0][1] < 0) & (cols[0][2] < 0) & (cols[1][1] < 0) & (cols[1][2] < 0);
  const size_t srow = static_cast<size_t>(p.w) * p.src_c;
  const T* sbase = src + (static_cast<size_t>(n) * p.h * p.w + sw) * p.src_c + sc0;
  const size_t rrow = static_cast<size_t>(p.res_wp) * p.res_c;
  const T* rbase = res ? res + ((static_cast<size_t>(n) * p.res_hp + p.res_off) * p.res_wp + sw + p.res_off) * p.res_c + sc0
                       : nullptr;
  T* obase = dst + static_cast<size_t>(n) * p.hd * p.wd * p.dst_c + sg * 2;
  for (int sh = r0; sh < r1; sh += kXfRows) {
    Raw8<T> vr[kXfRows], rr[kXfRows];
#pragma unroll
    for (int u = 0; u < kXfRows; ++u) ldraw(sbase + min(sh + u, r1 - 1) * srow, vr[u]);
    if (rbase) {
#pragma unroll
      for (int u = 0; u < kXfRows; ++u) ldraw(rbase + min(sh + u, r1 - 1) * rrow, rr[u]);
    }
#pragma unroll
    for (int u = 0; u < kXfRows; ++u) {
      if (sh + u >= r1) break;
      float v[8];
      unpack(vr[u], v);
      if (p.norm) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (v[j] - sf[j]) * sc[j];
      }
      if (p.act == VCG_ACT_RELU || p.act == VCG_ACT_LEAKY) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act_apply(v[j], p.act);
      } else if (p.act) {          // Tanh / Sigmoid (CaSb's other choices): cold path
#pragma unroll 1
        for (int j = 0; j < 8; ++j) v[j] = act_apply_any(v[j], p.act);
      }
      if (rbase) {
        float r[8];
        unpack(rr[u], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += r[j];
      }
      // interior source pixels (almost all) have no reflect mirrors: four plain stores instead of 36 predicated ones
      // (the general path made this kernel issue-bound: 2.6-2.9 TB/s against 4.9-5.4 for the other modes)
      const int hh = 2 * (sh + u);
      const bool row_plain = (mirror_k(hh, H2, p.pad, 1) < 0) & (mirror_k(hh, H2, p.pad, 2) < 0) &
                             (mirror_k(hh + 1, H2, p.pad, 1) < 0) & (mirror_k(hh + 1, H2, p.pad, 2) < 0);
      if (col_plain & row_plain) {
        T* o = obase + (static_cast<size_t>(hh + p.pad) * p.wd + 2 * sw + p.pad) * p.dst_c;
        const size_t orow_stride = static_cast<size_t>(p.wd) * p.dst_c;
        st2<T>(o, v[0], v[4]);
        st2<T>(o + p.dst_c, v[1], v[5]);
        st2<T>(o + orow_stride, v[2], v[6]);
        st2<T>(o + orow_stride + p.dst_c, v[3], v[7]);
        continue;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int th = mirror_k(hh + i, H2, p.pad, kh);
          if (th < 0) continue;
          T* orow = obase + static_cast<size_t>(th) * p.wd * p.dst_c;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const int tw = cols[jj][kw];
              if (tw >= 0) st2<T>(orow + static_cast<size_t>(tw) * p.dst_c, v[i * 2 + jj], v[4 + i * 2 + jj]);
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ backward transform
struct GSrc { const void* dxp; int mode, pad, c_pitch, folded; };
struct XbArgs {
  int n, h, w, c, y_c, norm, act, pre_act, dy_halo, dy_c, nsrc, stats_hw, clear_halo;
  int l2pf;                    // L2 read-ahead distance in loop iterations (xform_bwd_norm_kernel), 0 = off
  GSrc s[3];
};

// f(th, tw) for every padded position that reflects onto (i, j); interior pixels (the vast majority) take one call
template <typename F>
__device__ __forceinline__ void for_mirrors(int i, int Lh, int j, int Lw, int pad, bool folded, F&& f) {
  if (pad == 0 || folded || ((i > pad) & (i < Lh - 1 - pad) & (j > pad) & (j < Lw - 1 - pad))) { f(i + pad, j + pad); return; }
#pragma unroll 1
  for (int kh = 0; kh < 3; ++kh) {
    const int th = mirror_k(i, Lh, pad, kh);
    if (th < 0) continue;
#pragma unroll 1
    for (int kw = 0; kw < 3; ++kw) {
      const int tw = mirror_k(j, Lw, pad, kw);
      if (tw >= 0) f(th, tw);
    }
  }
}

template <typename T> __device__ __forceinline__ void ld2(const T* p, float& a, float& b);
template <> __device__ __forceinline__ void ld2<float>(const float* p, float& a, float& b) {
  const float2 v = *reinterpret_cast<const float2*>(p); a = v.x; b = v.y;
}
template <> __device__ __forceinline__ void ld2<__nv_bfloat16>(const __nv_bfloat16* p, float& a, float& b) {
  const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p)); a = v.x; b = v.y;
}

template <typename T>
__device__ __forceinline__ void gather_src(const GSrc& s, const XbArgs& p, int n, int h, int w, int ch, float (&g)[8]) {
  const T* dxp = static_cast<const T*>(s.dxp);
  const int pad = s.pad, pitch = s.c_pitch;
  if (s.mode == VCG_MODE_PLAIN) {
    const int wp = p.w + 2 * pad;
    const T* base = dxp + static_cast<size_t>(n) * (p.h + 2 * pad) * wp * pitch + ch;
    for_mirrors(h, p.h, w, p.w, pad, s.folded != 0, [&](int th, int tw) {
      float v[8];
      ld8<T>(base + (static_cast<size_t>(th) * wp + tw) * pitch, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) g[q] += v[q];
    });
  } else if (s.mode == VCG_MODE_SHUFFLE) {
    // source channels ch..ch+7 = destination channels ch/4, ch/4+1 at the four sub-pixels (PixelShuffle inverse)
    const int H2 = 2 * p.h, W2 = 2 * p.w, wp = W2 + 2 * pad;
    const T* base = dxp + static_cast<size_t>(n) * (H2 + 2 * pad) * wp * pitch + (ch >> 2);
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      float a0 = 0.f, a1 = 0.f;
      for_mirrors(2 * h + (sub >> 1), H2, 2 * w + (sub & 1), W2, pad, s.folded != 0, [&](int th, int tw) {
        float x0, x1;
        ld2<T>(base + (static_cast<size_t>(th) * wp + tw) * pitch, x0, x1);
        a0 += x0; a1 += x1;
      });
      g[sub] += a0; g[sub + 4] += a1;
    }
  } else if (s.mode == VCG_MODE_UNSHUFFLE) {
    const int Hh = p.h / 2, Wh = p.w / 2, wp = Wh + 2 * pad;
    const int sub = (h & 1) * 2 + (w & 1);
    const T* base = dxp + static_cast<size_t>(n) * (Hh + 2 * pad) * wp * pitch + sub * p.c + ch;
    for_mirrors(h >> 1, Hh, w >> 1, Wh, pad, s.folded != 0, [&](int th, int tw) {
      float v[8];
      ld8<T>(base + (static_cast<size_t>(th) * wp + tw) * pitch, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) g[q] += v[q];
    });
  } else {  // PAD_S2D: mirrors live in the padded full-resolution domain, then map to (pixel/2, sub-pixel channel block)
    const int hp = (p.h + 2 * pad) / 2, wp = (p.w + 2 * pad) / 2;
    const T* base = dxp + static_cast<size_t>(n) * hp * wp * pitch + ch;
    for_mirrors(h, p.h, w, p.w, pad, s.folded != 0, [&](int th, int tw) {
      const int sub = (th & 1) * 2 + (tw & 1);
      float v[8];
      ld8<T>(base + (static_cast<size_t>(th >> 1) * wp + (tw >> 1)) * pitch + sub * p.c, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) g[q] += v[q];
    });
  }
}

// Generic fallback (channel-group counts that are not a power of two, e.g. latent_dim 96).
// grid: (pixel chunks, n, channel-group chunks of 32); block 256 = 32 channel groups x 8 pixel lanes
template <typename T, bool PHASE2>
__global__ void __launch_bounds__(256, PHASE2 ? 2 : 3)
xform_bwd_generic_kernel(const __grid_constant__ XbArgs p, const T* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gsums_in,
                 T* __restrict__ dy, float* __restrict__ gsums, float* __restrict__ dbias, int pix_per_block) {
  const int cgb = min(32, p.c / 8 - blockIdx.z * 32);
  const int cgl = p.c / 8 < 32 ? p.c / 8 : 32;          // channel-group lanes; the rest of the block strides pixels
  const int lanes = 256 / cgl;
  const int cg = threadIdx.x % cgl, pl = threadIdx.x / cgl;
  __shared__ float sacc[32 * 24];
  for (int i = threadIdx.x; i < 32 * 24; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  if (cg < cgb && pl < lanes) {
    const int ch = (blockIdx.z * 32 + cg) * 8;
    const int hw = p.h * p.w;
    const int p0 = blockIdx.x * pix_per_block, p1 = min(hw, p0 + pix_per_block);
    const int wpd = p.w + 2 * p.dy_halo, hpd = p.h + 2 * p.dy_halo;
    float s1[8] = {}, s2[8] = {};      // norm phase 1: sum g, sum g*zhat; otherwise s1 = bias gradient
    float mean[8], rstd[8], m1[8], m2[8];
    if (p.norm) {
      const float* m = mr + (static_cast<size_t>(n) * p.c + ch) * 2;
#pragma unroll
      for (int j = 0; j < 8; ++j) { mean[j] = m[2 * j]; rstd[j] = m[2 * j + 1]; finalize_pair(mean[j], rstd[j], p.stats_hw); }
      if (PHASE2) {
        const float* gs = gsums_in + (static_cast<size_t>(n) * p.c + ch) * 2;
        const float inv = 1.f / hw;
#pragma unroll
        for (int j = 0; j < 8; ++j) { m1[j] = gs[2 * j] * inv; m2[j] = gs[2 * j + 1] * inv; }
      }
    }
    if (!PHASE2 && p.clear_halo && p.dy_halo > 0) clear_halo_share<T>(dy, n, p.h, p.w, p.dy_halo, p.dy_c, ch, blockIdx.x * lanes + pl, gridDim.x * lanes);
    const bool need_y = p.norm || p.act || p.pre_act;
    auto dy_ptr = [&](int pp) {
      const int h = pp / p.w, w = pp - h * p.w;
      return dy + ((static_cast<size_t>(n) * hpd + h + p.dy_halo) * wpd + w + p.dy_halo) * p.dy_c + ch;
    };
    // loads of one pixel (saved output y, and either the gathered consumer gradients or phase-1's g)
    auto load_px = [&](int pp, float (&g)[8], float (&yv)[8]) {
      if (need_y) ld8<T>(y + (static_cast<size_t>(n) * hw + pp) * p.y_c + ch, yv);
      if (!PHASE2) {
        const int h = pp / p.w, w = pp - h * p.w;
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
        for (int k = 0; k < p.nsrc; ++k) gather_src<T>(p.s[k], p, n, h, w, ch, g);
      } else {
        ld8<T>(dy_ptr(pp), g);
      }
    };
    auto finish_px = [&](int pp, float (&g)[8], float (&yv)[8]) {
      if (!PHASE2) {
        if (p.norm) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float z = (yv[j] - mean[j]) * rstd[j];
            g[j] *= act_grad_in(z, p.act);
            s1[j] += g[j]; s2[j] += g[j] * z;
          }
        } else {
          // no norm: at most one activation (fused in the conv epilogue or applied after): y or act(y) share sign
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            g[j] *= act_grad_in(yv[j], p.act) * act_grad(yv[j], p.pre_act);      // y: input of act, output of pre_act
            s1[j] += g[j];
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = (yv[j] - mean[j]) * rstd[j];
          g[j] = rstd[j] * (g[j] - m1[j] - z * m2[j]) * act_grad(yv[j], p.pre_act);
          s1[j] += g[j];
        }
      }
      st8<T>(dy_ptr(pp), g);
    };
    if (PHASE2) {
      // elementwise phase: two pixels per iteration, all loads issued before the first store (more bytes in flight)
      for (int pp = p0 + pl; pp < p1; pp += 2 * lanes) {
        float gA[8], yA[8], gB[8], yB[8];
        const bool hasB = pp + lanes < p1;
        load_px(pp, gA, yA);
        if (hasB) load_px(pp + lanes, gB, yB);
        finish_px(pp, gA, yA);
        if (hasB) finish_px(pp + lanes, gB, yB);
      }
    } else {
      for (int pp = p0 + pl; pp < p1; pp += lanes) {
        float g[8], yv[8];
        load_px(pp, g, yv);
        finish_px(pp, g, yv);
      }
    }
    if (!PHASE2 && p.norm) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(&sacc[cg * 24 + j], s1[j]); atomicAdd(&sacc[cg * 24 + 8 + j], s2[j]); }
    } else if (dbias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sacc[cg * 24 + 16 + j], s1[j]);
    }
  }
  __syncthreads();
  if (!PHASE2 && p.norm) {
    for (int i = threadIdx.x; i < cgb * 16; i += 256) {
      const int g = i / 16, r = i % 16;
      const int ch = (blockIdx.z * 32 + g) * 8 + (r & 7);
      atomicAdd(gsums + (static_cast<size_t>(n) * p.c + ch) * 2 + (r >> 3), sacc[g * 24 + r]);
    }
  } else if (dbias) {
    for (int i = threadIdx.x; i < cgb * 8; i += 256) {
      const int g = i / 8, r = i % 8;
      atomicAdd(dbias + (blockIdx.z * 32 + g) * 8 + r, sacc[g * 24 + 16 + r]);
    }
  }
}

// ---- fast path (power-of-two channel-group counts: every layer of the reference networks) ----------------
// Same grid and thread mapping as the generic kernel, but
//  (a) the reflect halo of every consumer gradient has been folded into its interior beforehand (vcg_fold_halo),
//      so each consumer contributes exactly ONE 16-byte load per pixel at an address that is linear in
//      (h>>s, w>>s, h&1, w&1) -- no mode switch and no mirror loop in the hot loop;
//  (b) kXbU pixels per thread per iteration with all their loads issued before the first use;
//  (c) the per-(n,c) sums are reduced with warp shuffles and one shared-memory slab row per warp.
constexpr int kXbUmax = 4;

__device__ __forceinline__ void ldraw_pairs(const __nv_bfloat16* p0, const __nv_bfloat16* p1, const __nv_bfloat16* p2,
                                            const __nv_bfloat16* p3, Raw8<__nv_bfloat16>& v) {
  v.r.x = *reinterpret_cast<const uint32_t*>(p0); v.r.y = *reinterpret_cast<const uint32_t*>(p1);
  v.r.z = *reinterpret_cast<const uint32_t*>(p2); v.r.w = *reinterpret_cast<const uint32_t*>(p3);
}
__device__ __forceinline__ void ldraw_pairs(const float* p0, const float* p1, const float* p2, const float* p3,
                                            Raw8<float>& v) {
  const float2 a = *reinterpret_cast<const float2*>(p0), b = *reinterpret_cast<const float2*>(p1);
  const float2 c = *reinterpret_cast<const float2*>(p2), d = *reinterpret_cast<const float2*>(p3);
  v.a = make_float4(a.x, a.y, b.x, b.y); v.b = make_float4(c.x, c.y, d.x, d.y);
}

// direct-term addressing of one consumer, precomputed on the host (element offsets within one image):
//   H = h + hoff, W = w + woff;  off = (H>>sh)*rs + (W>>sh)*cs + (H&msk)*ph + (W&msk)*pw
// PixelShuffle consumers (shuffle=1) read four channel pairs at off, off+pitch, off+ph, off+ph+pitch.
struct FSrc {
  const void* base;            // first direct element of image 0 (pad offsets applied)
  long long img_stride;
  int hoff, woff, sh, msk, rs, cs, ph, pw, pitch, shuffle;
};
struct XgArgs {
  int n, h, w, c, y_c, norm, act, pre_act, dy_halo, dy_c, nsrc, stats_hw, clear_halo;
  int l2pf;                    // L2 read-ahead distance in loop iterations (vcg_set_l2_prefetch), 0 = off
  FSrc s[3];
};

// f(th, tw) for the mirrored padded positions only (the direct one is excluded)
template <typename F>
__device__ __forceinline__ void for_mirrors_only(int i, int Lh, int j, int Lw, int pad, F&& f) {
#pragma unroll 1
  for (int kh = 0; kh < 3; ++kh) {
    const int th = mirror_k(i, Lh, pad, kh);
    if (th < 0) continue;
#pragma unroll 1
    for (int kw = (kh == 0 ? 1 : 0); kw < 3; ++kw) {
      const int tw = mirror_k(j, Lw, pad, kw);
      if (tw >= 0) f(th, tw);
    }
  }
}

// ---- halo fold: dxp[direct(h,w)] += sum of dxp at the padded positions that reflect onto (h,w), in place.
// Only pixels within `pad` of the border have mirrors; the kernel enumerates exactly those:
// border rows x all columns, then the remaining rows x border columns.
struct FoldArgs {
  int n, h, w, c, mode, pad, pitch;
  int ra0, rlen0, rb0, rlen1;      // border row ranges [ra0, ra0+rlen0), [rb0, rb0+rlen1) in activation coordinates
  int ca0, clen0, cb0, clen1;      // border column ranges
  int total;                       // work items (pixels) per image
};

__device__ __forceinline__ int nth_outside(int r, int a0, int len0, int b0, int len1) {
  // r-th index not inside [a0,a0+len0) or [b0,b0+len1)  (a0 < b0, ranges disjoint)
  if (r < a0) return r;
  r -= a0;
  const int gap = b0 - (a0 + len0);
  if (r < gap) return a0 + len0 + r;
  return b0 + len1 + (r - gap);
}

template <typename T>
__global__ void __launch_bounds__(256)
fold_halo_kernel(T* __restrict__ dxp, const FoldArgs p) {
  const int groups = p.c / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= static_cast<long long>(p.total) * groups) return;
  const int cg = static_cast<int>(idx % groups);
  int item = static_cast<int>(idx / groups);
  const int n = blockIdx.y, ch = cg * 8;
  const int nbh = p.rlen0 + p.rlen1, nbw = p.clen0 + p.clen1;
  int h, w;
  if (item < nbh * p.w) {
    const int r = item / p.w;
    w = item - r * p.w;
    h = r < p.rlen0 ? p.ra0 + r : p.rb0 + (r - p.rlen0);
  } else {
    item -= nbh * p.w;
    const int r = item / nbw, ci = item - r * nbw;
    w = ci < p.clen0 ? p.ca0 + ci : p.cb0 + (ci - p.clen0);
    h = nth_outside(r, p.ra0, p.rlen0, p.rb0, p.rlen1);
  }
  const int pad = p.pad, pitch = p.pitch;
  float m[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) m[q] = 0.f;
  if (p.mode == VCG_MODE_PLAIN) {
    const int wp = p.w + 2 * pad;
    T* base = dxp + static_cast<size_t>(n) * (p.h + 2 * pad) * wp * pitch + ch;
    for_mirrors_only(h, p.h, w, p.w, pad, [&](int th, int tw) {
      float v[8];
      ld8<T>(base + (static_cast<size_t>(th) * wp + tw) * pitch, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) m[q] += v[q];
    });
    T* d = base + (static_cast<size_t>(h + pad) * wp + w + pad) * pitch;
    float v[8];
    ld8<T>(d, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] += m[q];
    st8<T>(d, v);
  } else if (p.mode == VCG_MODE_SHUFFLE) {
    const int H2 = 2 * p.h, W2 = 2 * p.w, wp = W2 + 2 * pad;
    T* base = dxp + static_cast<size_t>(n) * (H2 + 2 * pad) * wp * pitch + (ch >> 2);
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      float a0 = 0.f, a1 = 0.f;
      const int A = 2 * h + (sub >> 1), B = 2 * w + (sub & 1);
      for_mirrors_only(A, H2, B, W2, pad, [&](int th, int tw) {
        float x0, x1;
        ld2<T>(base + (static_cast<size_t>(th) * wp + tw) * pitch, x0, x1);
        a0 += x0; a1 += x1;
      });
      T* d = base + (static_cast<size_t>(A + pad) * wp + B + pad) * pitch;
      float x0, x1;
      ld2<T>(d, x0, x1);
      st2<T>(d, x0 + a0, x1 + a1);
    }
  } else if (p.mode == VCG_MODE_UNSHUFFLE) {
    const int Hh = p.h / 2, Wh = p.w / 2, wp = Wh + 2 * pad;
    const int sub = (h & 1) * 2 + (w & 1);
    T* base = dxp + static_cast<size_t>(n) * (Hh + 2 * pad) * wp * pitch + sub * p.c + ch;
    for_mirrors_only(h >> 1, Hh, w >> 1, Wh, pad, [&](int th, int tw) {
      float v[8];
      ld8<T>(base + (static_cast<size_t>(th) * wp + tw) * pitch, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) m[q] += v[q];
    });
    T* d = base + (static_cast<size_t>((h >> 1) + pad) * wp + (w >> 1) + pad) * pitch;
    float v[8];
    ld8<T>(d, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] += m[q];
    st8<T>(d, v);
  } else {
    const int hp = (p.h + 2 * pad) / 2, wp = (p.w + 2 * pad) / 2;
    T* base = dxp + static_cast<size_t>(n) * hp * wp * pitch + ch;
    for_mirrors_only(h, p.h, w, p.w, pad, [&](int th, int tw) {
      const int sub = (th & 1) * 2 + (tw & 1);
      float v[8];
      ld8<T>(base + (static_cast<size_t>(th >> 1) * wp + (tw >> 1)) * pitch + sub * p.c, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) m[q] += v[q];
    });
    const int th = h + pad, tw = w + pad;
    T* d = base + (static_cast<size_t>(th >> 1) * wp + (tw >> 1)) * pitch + ((th & 1) * 2 + (tw & 1)) * p.c;
    float v[8];
    ld8<T>(d, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] += m[q];
    st8<T>(d, v);
  }
}

// Block-wide sum of r[0..NV) over the threads that share a channel group (cg = threadIdx.x % cgl, cgl | 32):
// shuffle across the pixel lanes of each warp, one slab row per warp, then emit(cg, j, total) once per block.
template <int NV, typename F>
__device__ __forceinline__ void reduce_groups(float (&r)[NV], int cgl, float* sred /* [8][NV][32] */, F&& emit) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int off = 16; off >= cgl; off >>= 1) {
#pragma unroll
    for (int j = 0; j < NV; ++j) r[j] += __shfl_xor_sync(0xffffffffu, r[j], off);
  }
  if (lane < cgl) {
#pragma unroll
    for (int j = 0; j < NV; ++j) sred[(warp * NV + j) * 32 + lane] = r[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cgl * NV; i += 256) {
    const int l = i % cgl, j = i / cgl;
    float t = 0.f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) t += sred[(wp * NV + j) * 32 + l];
    emit(l, j, t);
  }
}

template <typename T, int kXbU, int MINB>
__global__ void __launch_bounds__(256, MINB)
xform_bwd_gather_kernel(const __grid_constant__ XgArgs p, const T* __restrict__ y, const float* __restrict__ mr,
                        T* __restrict__ dy, float* __restrict__ gsums, float* __restrict__ dbias, int pix_per_block) {
  __shared__ float sred[8 * 16 * 32];
  const int cgl = p.c / 8 < 32 ? p.c / 8 : 32;     // power of two; p.c/8 is a multiple of it
  const int lanes = 256 / cgl;
  const int cg = threadIdx.x % cgl, pl = threadIdx.x / cgl;
  const int n = blockIdx.y;
  const int ch = (blockIdx.z * 32 + cg) * 8;
  const int hw = p.h * p.w;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(hw, p0 + pix_per_block);
  const int wpd = p.w + 2 * p.dy_halo, hpd = p.h + 2 * p.dy_halo;
  float s[16];      // norm: [0,8) = sum g, [8,16) = sum g*zhat; otherwise [0,8) = bias gradient
#pragma unroll
  for (int j = 0; j < 16; ++j) s[j] = 0.f;
  float sc[8], sf[8];
  if (p.norm) load_scale_shift(mr + (static_cast<size_t>(n) * p.c + ch) * 2, sc, sf, p.stats_hw);
  if (p.clear_halo && p.dy_halo > 0) clear_halo_share<T>(dy, n, p.h, p.w, p.dy_halo, p.dy_c, ch, blockIdx.x * lanes + pl, gridDim.x * lanes);
  const bool need_y = p.norm || p.act || p.pre_act;
  const float slope_act = act_slope(p.act), slope_pre = act_slope(p.pre_act);
  // L2 read-ahead.  The demand loads below are register-limited (8 x 16 B in flight per thread, and none while a warp
  // computes and stores), so HBM -> L2 is kept streaming by bulk prefetches of the contiguous byte ranges the block
  // touches p.l2pf iterations from now: lane 0 of warp 0 for y, of warps 1.. for the (plain-layout) sources.  Blocks
  // that cover every channel of their pixels only (c <= 256: the large planes, where the kernel is HBM-bound).
  const int step = kXbU * lanes;
  const bool pf_on = p.l2pf > 0 && gridDim.z == 1 && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) <= p.nsrc;
  auto prefetch = [&](int q0) {
    if (q0 >= p1) return;
    const int q1 = min(q0 + step, p1);
    const int wi = threadIdx.x >> 5;
    if (wi == 0) {
      if (need_y) l2_prefetch_bulk(y + (static_cast<size_t>(n) * hw + q0) * p.y_c, static_cast<long long>(q1 - q0) * p.y_c * sizeof(T));
      return;
    }
    const FSrc& fs = p.s[wi - 1];
    if (fs.shuffle || fs.sh) return;
    const T* b = static_cast<const T*>(fs.base) + static_cast<size_t>(n) * fs.img_stride;
    const int h0 = q0 / p.w, w0 = q0 - h0 * p.w, h1 = (q1 - 1) / p.w, w1 = (q1 - 1) - h1 * p.w;
    const T* a0 = b + ((h0 + fs.hoff) * fs.rs + (w0 + fs.woff) * fs.cs);
    const T* a1 = b + ((h1 + fs.hoff) * fs.rs + (w1 + fs.woff) * fs.cs) + p.c;
    l2_prefetch_bulk(a0, (a1 - a0) * static_cast<long long>(sizeof(T)));
  };
  int q_ahead = p0;
  if (pf_on)
    for (; q_ahead < p0 + p.l2pf * step; q_ahead += step) prefetch(q_ahead);
  for (int pp = p0 + pl; pp < p1; pp += kXbU * lanes) {
    if (pf_on) { prefetch(q_ahead); q_ahead += step; }
    int hh[kXbU], ww[kXbU];
    float g[kXbU][8];
    Raw8<T> yr[kXbU];
#pragma unroll
    for (int u = 0; u < kXbU; ++u) {
      const int q = min(pp + u * lanes, p1 - 1);
      hh[u] = q / p.w; ww[u] = q - hh[u] * p.w;
      if (need_y) ldraw(y + (static_cast<size_t>(n) * hw + q) * p.y_c + ch, yr[u]);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[u][j] = 0.f;
    }
#pragma unroll 1
    for (int k = 0; k < p.nsrc; ++k) {
      const FSrc& fs = p.s[k];
      const T* b = static_cast<const T*>(fs.base) + static_cast<size_t>(n) * fs.img_stride;
      Raw8<T> raw[kXbU];
      if (!fs.shuffle) {
        b += ch;
#pragma unroll
        for (int u = 0; u < kXbU; ++u) {
          const int H = hh[u] + fs.hoff, W = ww[u] + fs.woff;
          ldraw(b + ((H >> fs.sh) * fs.rs + (W >> fs.sh) * fs.cs + (H & fs.msk) * fs.ph + (W & fs.msk) * fs.pw), raw[u]);
        }
#pragma unroll
        for (int u = 0; u < kXbU; ++u) {
          float f[8];
          unpack(raw[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[u][j] += f[j];
        }
      } else {
        b += ch >> 2;
#pragma unroll
        for (int u = 0; u < kXbU; ++u) {
          const T* b0 = b + (hh[u] * fs.rs + ww[u] * fs.cs);
          ldraw_pairs(b0, b0 + fs.pitch, b0 + fs.ph, b0 + fs.ph + fs.pitch, raw[u]);
        }
#pragma unroll
        for (int u = 0; u < kXbU; ++u) {
          float f[8];
          unpack(raw[u], f);
#pragma unroll
          for (int sub = 0; sub < 4; ++sub) { g[u][sub] += f[2 * sub]; g[u][sub + 4] += f[2 * sub + 1]; }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kXbU; ++u) {
      if (pp + u * lanes >= p1) break;
      float yv[8];
      if (need_y) unpack(yr[u], yv);
      if (p.norm) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = (yv[j] - sf[j]) * sc[j];
          g[u][j] *= act_grad_s(z, slope_act);
          s[j] += g[u][j]; s[8 + j] = fmaf(g[u][j], z, s[8 + j]);
        }
      } else if (need_y) {
        // no norm: at most one activation (fused in the conv epilogue or applied after): y or act(y) share sign
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          g[u][j] *= act_grad_s(yv[j], slope_act) * act_grad_s(yv[j], slope_pre);
          s[j] += g[u][j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += g[u][j];
      }
      st8<T>(dy + ((static_cast<size_t>(n) * hpd + hh[u] + p.dy_halo) * wpd + ww[u] + p.dy_halo) * p.dy_c + ch, g[u]);
    }
  }
  if (p.norm) {
    reduce_groups<16>(s, cgl, sred, [&](int l, int j, float t) {
      atomicAdd(gsums + (static_cast<size_t>(n) * p.c + (blockIdx.z * 32 + l) * 8 + (j & 7)) * 2 + (j >> 3), t);
    });
  } else if (dbias) {
    float b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = s[j];
    reduce_groups<8>(b, cgl, sred, [&](int l, int j, float t) { atomicAdd(dbias + (blockIdx.z * 32 + l) * 8 + j, t); });
  }
}

// phase 2 of the InstanceNorm backward: dY = rstd * (g - mean(g) - zhat * mean(g*zhat)) * pre_act'(y), in place
template <typename T>
__global__ void __launch_bounds__(256, 2)
xform_bwd_norm_kernel(const __grid_constant__ XbArgs p, const T* __restrict__ y, const float* __restrict__ mr,
                      const float* __restrict__ gsums_in, T* __restrict__ dy, float* __restrict__ dbias, int pix_per_block) {
  constexpr int kXbU = kXbUmax;
  __shared__ float sred[8 * 8 * 32];
  const int cgl = p.c / 8 < 32 ? p.c / 8 : 32;
  const int lanes = 256 / cgl;
  const int cg = threadIdx.x % cgl, pl = threadIdx.x / cgl;
  const int n = blockIdx.y;
  const int ch = (blockIdx.z * 32 + cg) * 8;
  const int hw = p.h * p.w;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(hw, p0 + pix_per_block);
  const int wpd = p.w + 2 * p.dy_halo, hpd = p.h + 2 * p.dy_halo;
  float mean[8], rstd[8], m1[8], m2[8], s[8];
  const float slope_pre = act_slope(p.pre_act);
  Raw8<T> gr[kXbU], yr[kXbU];
  T* dp[kXbU];
  auto issue = [&](int pp) {
#pragma unroll
    for (int u = 0; u < kXbU; ++u) {
      const int q = min(pp + u * lanes, p1 - 1);
      const int h = q / p.w, w = q - h * p.w;
      dp[u] = dy + ((static_cast<size_t>(n) * hpd + h + p.dy_halo) * wpd + w + p.dy_halo) * p.dy_c + ch;
      ldraw(dp[u], gr[u]);
      ldraw(y + (static_cast<size_t>(n) * hw + q) * p.y_c + ch, yr[u]);
    }
  };
  // L2 read-ahead (see xform_bwd_gather_kernel): lane 0 of warp 0 for y, of warp 1 for the gradient buffer
  const int step = kXbU * lanes;
  const bool pf_on = p.l2pf > 0 && gridDim.z == 1 && (threadIdx.x & 31) == 0 && threadIdx.x < 64;
  auto prefetch = [&](int q0) {
    if (q0 >= p1) return;
    const int q1 = min(q0 + step, p1);
    if (threadIdx.x == 0) {
      l2_prefetch_bulk(y + (static_cast<size_t>(n) * hw + q0) * p.y_c, static_cast<long long>(q1 - q0) * p.y_c * sizeof(T));
    } else {
      const int h0 = q0 / p.w, w0 = q0 - h0 * p.w, h1 = (q1 - 1) / p.w, w1 = (q1 - 1) - h1 * p.w;
      const T* a0 = dy + ((static_cast<size_t>(n) * hpd + h0 + p.dy_halo) * wpd + w0 + p.dy_halo) * p.dy_c;
      const T* a1 = dy + ((static_cast<size_t>(n) * hpd + h1 + p.dy_halo) * wpd + w1 + p.dy_halo) * p.dy_c + p.c;
      l2_prefetch_bulk(a0, (a1 - a0) * static_cast<long long>(sizeof(T)));
    }
  };
  int q_ahead = p0 + step;       // the first chunk's demand loads go out right below
  if (pf_on)
    for (; q_ahead < p0 + (1 + p.l2pf) * step; q_ahead += step) prefetch(q_ahead);
  // the first batch of data loads goes out BEFORE the per-channel statistics are fetched: small planes run only a
  // few iterations per block, and the dependent statistics -> data round trips were a quarter of the kernel time
  int pp = p0 + pl;
  if (pp < p1) issue(pp);
  {
    load_scale_shift(mr + (static_cast<size_t>(n) * p.c + ch) * 2, rstd, mean, p.stats_hw);
    const float* gs = gsums_in + (static_cast<size_t>(n) * p.c + ch) * 2;
    const float inv = 1.f / hw;
#pragma unroll
    for (int q = 0; q < 4; ++q) {                           // 16-byte loads (scalar ones cost 8x the L2 sectors)
      const float4 t = *reinterpret_cast<const float4*>(gs + 4 * q);
      m1[2 * q] = t.x * inv; m2[2 * q] = t.y * inv; m1[2 * q + 1] = t.z * inv; m2[2 * q + 1] = t.w * inv;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
  }
  while (pp < p1) {
#pragma unroll
    for (int u = 0; u < kXbU; ++u) {
      if (pp + u * lanes >= p1) break;
      float g[8], yv[8];
      unpack(gr[u], g);
      unpack(yr[u], yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = (yv[j] - mean[j]) * rstd[j];
        g[j] = rstd[j] * (g[j] - m1[j] - z * m2[j]) * act_grad_s(yv[j], slope_pre);
        s[j] += g[j];
      }
      st8<T>(dp[u], g);
    }
    pp += kXbU * lanes;
    if (pf_on) { prefetch(q_ahead); q_ahead += step; }
    if (pp < p1) issue(pp);
  }
  if (dbias)
    reduce_groups<8>(s, cgl, sred, [&](int l, int j, float t) { atomicAdd(dbias + (blockIdx.z * 32 + l) * 8 + j, t); });
}

// ------------------------------------------------------------------ NCHW <-> NHWC
template <typename T>
__global__ void pack_nchw_kernel(const float* __restrict__ src, int n, int c, int h, int w, T* __restrict__ dst,
                                 int dst_c, int halo, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;   // one thread per (n, a, b, 8-channel group) of the destination
  const int groups = dst_c / 8, hd = h + 2 * halo, wd = w + 2 * halo;
  const int g = static_cast<int>(idx % groups);
  long long t = idx / groups;
  const int b = static_cast<int>(t % wd); t /= wd;
  const int a = static_cast<int>(t % hd);
  const int ni = static_cast<int>(t / hd);
  float v[8];
  const int hh = a - halo, ww = b - halo;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = g * 8 + j;
    v[j] = (ch < c && hh >= 0 && hh < h && ww >= 0 && ww < w)
               ? src[((static_cast<size_t>(ni) * c + ch) * h + hh) * w + ww] : 0.f;
  }
  st8<T>(dst + idx * 8, v);
}

template <typename T>
__global__ void unpack_nchw_kernel(const T* __restrict__ src, int src_c, int n, int c, int h, int w,
                                   float* __restrict__ dst, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;   // one thread per destination element, w fastest (coalesced writes)
  const int ww = static_cast<int>(idx % w);
  long long t = idx / w;
  const int hh = static_cast<int>(t % h); t /= h;
  const int ch = static_cast<int>(t % c);
  const int ni = static_cast<int>(t / c);
  dst[idx] = Elem<T>::ld(src + ((static_cast<size_t>(ni) * h + hh) * w + ww) * src_c + ch);
}

// zero only the halo ring of an NHWC buffer [n, h+2*halo, w+2*halo, c] (the interior is fully overwritten by
// the backward transform, so clearing the whole dY buffer would double its write traffic)
__global__ void zero_halo_kernel(uint4* __restrict__ buf, int n, int h, int w, int halo, int vec_per_px) {
  const int hp = h + 2 * halo, wp = w + 2 * halo;
  const int ring = hp * wp - h * w;                       // halo pixels per image
  const long long total = static_cast<long long>(n) * ring * vec_per_px;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec_per_px);
    long long t = i / vec_per_px;
    const int r = static_cast<int>(t % ring);
    const int img = static_cast<int>(t / ring);
    int a, b;
    const int top = halo * wp;
    if (r < top) { a = r / wp; b = r - a * wp; }
    else if (r < 2 * top) { const int q = r - top; a = halo + h + q / wp; b = q % wp; }
    else { const int q = r - 2 * top; const int row = q / (2 * halo), k = q - row * 2 * halo;
           a = halo + row; b = k < halo ? k : w + k; }
    buf[((static_cast<size_t>(img) * hp + a) * wp + b) * vec_per_px + v] = make_uint4(0u, 0u, 0u, 0u);
  }
}

__global__ void zero_kernel(float4* p, size_t n16, char* tail, size_t ntail) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t k = i; k < n16; k += stride) p[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < ntail) tail[i] = 0;
}

}  // namespace

// ===================================================================================== host API
static int pix_chunk(int hw, int n, int zc) {
  // aim for ~4 blocks per SM, at least 64 pixels per block
  long long want = 4LL * vcg_num_sms();
  long long per = (static_cast<long long>(hw) * n * zc + want - 1) / want;
  if (per < 64) per = 64;
  if (per > hw) per = hw;
  return static_cast<int>(per);
}

// fast backward kernels: ~8 blocks per SM in flight over the whole launch, whole unrolled iterations per thread
// pixels per block of the backward fast paths.  Every block pays a fixed prologue (per-channel statistics) and
// epilogue (bias-gradient reduction + atomics), so small planes want FEWER, longer blocks; large planes want
// several waves for load balance.  mult = blocks per SM aimed for (measured on B200, tools/bench_xform.py:
// InstanceNorm backward 1024ch 16x16: 55 us at 8, 37 us at 2; 64ch 256x256: 291 us at 8, 376 us at 2).
static int pix_chunk_fast(int hw, int n, int zc, int cg_total, int mult) {
  const int cgl = cg_total < 32 ? cg_total : 32;
  const int step = kXbUmax * (256 / cgl);                  // pixels one block covers per unrolled iteration
  long long want = static_cast<long long>(mult) * vcg_num_sms();
  long long per = (static_cast<long long>(hw) * n * zc + want - 1) / want;
  if (per < 2 * step) per = 2 * step;
  per = (per + step - 1) / step * step;
  if (per > hw) per = hw;
  return static_cast<int>(per);
}

extern "C" int vcg_in_finalize(const float* sums, int32_t nc, int32_t hw, float* mean_rstd, void* stream) {
  in_finalize_f_kernel<<<(nc + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(sums, nc, hw, mean_rstd);
  VCG_CHECK_LAUNCH("in_finalize_f_kernel");
  return VCG_OK;
}

// fp64 accumulation needs n*c*2 doubles of scratch; to keep the ABI allocation-free it lives in the
// tail of the caller's buffer: mean_rstd must have room for n*c*2 floats + n*c*2 doubles.
extern "C" int vcg_in_stats(int32_t dtype, const void* y, int32_t n, int32_t hw, int32_t c, int32_t c_pitch,
                            float* mean_rstd, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VCG_REQUIRE(c % 8 == 0 && c_pitch % 8 == 0, VCG_E_UNSUPPORTED, "in_stats: c=%d pitch=%d", c, c_pitch);
  double* acc = reinterpret_cast<double*>(mean_rstd + static_cast<size_t>(n) * c * 2);
  cudaError_t e = cudaMemsetAsync(acc, 0, static_cast<size_t>(n) * c * 2 * sizeof(double), stream);
  VCG_REQUIRE(e == cudaSuccess, VCG_E_CUDA, "in_stats: memset: %s", cudaGetErrorString(e));
  const int zc = (c / 8 + 31) / 32;
  const int ppb = pix_chunk(hw, n, zc);
  dim3 grid((hw + ppb - 1) / ppb, n, zc);
  if (dtype == VCG_F32)
    in_stats_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(y), hw, c, c_pitch, ppb, acc);
  else
    in_stats_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(y), hw, c, c_pitch, ppb, acc);
  VCG_CHECK_LAUNCH("in_stats_kernel");
  in_finalize_d_kernel<<<(n * c + 255) / 256, 256, 0, stream>>>(acc, n * c, hw, mean_rstd);
  VCG_CHECK_LAUNCH("in_finalize_d_kernel");
  return VCG_OK;
}

extern "C" int vcg_xform_fwd(const vcg_xform_desc* d, const void* src, const float* mean_rstd, const void* residual,
                             void* dst, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VCG_REQUIRE(d->c % 8 == 0 && d->src_c % 8 == 0 && d->dst_c % 8 == 0, VCG_E_UNSUPPORTED,
              "xform_fwd: channels must be multiples of 8 (c=%d src_c=%d dst_c=%d)", d->c, d->src_c, d->dst_c);
  VCG_REQUIRE(!d->norm || mean_rstd, VCG_E_INVALID, "xform_fwd: norm without statistics");
  XfArgs a{};
  a.n = d->n; a.h = d->h; a.w = d->w; a.c = d->c; a.src_c = d->src_c; a.norm = d->norm; a.act = d->act;
  a.mode = d->mode; a.pad = d->pad; a.dst_c = d->dst_c;
  a.res_hp = d->res_hp; a.res_wp = d->res_wp; a.res_c = d->res_c; a.res_off = d->res_off; a.stats_hw = d->stats_hw;
  switch (d->mode) {
    case VCG_MODE_PLAIN: a.hd = d->h + 2 * d->pad; a.wd = d->w + 2 * d->pad; a.cd = d->c; break;
    case VCG_MODE_SHUFFLE:
      VCG_REQUIRE(d->c % 32 == 0, VCG_E_UNSUPPORTED, "xform_fwd: shuffle needs c%%32==0");
      a.hd = 2 * d->h + 2 * d->pad; a.wd = 2 * d->w + 2 * d->pad; a.cd = d->c / 4; break;
    case VCG_MODE_UNSHUFFLE:
      VCG_REQUIRE(d->h % 2 == 0 && d->w % 2 == 0, VCG_E_UNSUPPORTED, "xform_fwd: unshuffle needs even dims");
      a.hd = d->h / 2 + 2 * d->pad; a.wd = d->w / 2 + 2 * d->pad; a.cd = d->c * 4; break;
    case VCG_MODE_PAD_S2D:
      VCG_REQUIRE((d->h + 2 * d->pad) % 2 == 0 && (d->w + 2 * d->pad) % 2 == 0, VCG_E_UNSUPPORTED,
                  "xform_fwd: s2d needs even padded dims");
      a.hd = (d->h + 2 * d->pad) / 2; a.wd = (d->w + 2 * d->pad) / 2; a.cd = d->c * 4; break;
    default: VCG_REQUIRE(false, VCG_E_INVALID, "xform_fwd: bad mode %d", d->mode);
  }
  VCG_REQUIRE(a.cd <= d->dst_c, VCG_E_INVALID, "xform_fwd: dst_c=%d < %d", d->dst_c, a.cd);
  VCG_REQUIRE(d->n <= 65535, VCG_E_UNSUPPORTED, "xform_fwd: n=%d", d->n);
  VCG_REQUIRE(!d->norm || (reinterpret_cast<uintptr_t>(mean_rstd) & 15) == 0, VCG_E_INVALID, "xform_fwd: statistics must be 16-byte aligned");
  const bool shuffle = d->mode == VCG_MODE_SHUFFLE;
  VCG_REQUIRE(!shuffle || a.cd == d->dst_c, VCG_E_UNSUPPORTED, "xform_fwd: shuffle needs dst_c == c/4");
  // grid: x = 256-thread chunks of one row of (column, channel-group) pairs; y = row chunks; z = image.
  // rows per block: 8 (two unrolled batches) when that still leaves >= 4 waves of blocks, else 4
  const int row_len = shuffle ? d->w * (d->c / 8) : a.wd * (d->dst_c / 8);
  const int nrows = shuffle ? d->h : a.hd;
  const int xb = (row_len + 255) / 256;
  int rpb = 2 * kXfRows;
  if (static_cast<long long>(xb) * ((nrows + rpb - 1) / rpb) * d->n < 32LL * vcg_num_sms()) rpb = kXfRows;
  const dim3 blocks(xb, (nrows + rpb - 1) / rpb, d->n);
#define VCG_XF_LAUNCH(T, KERN)                                                                                   \
  KERN<<<blocks, 256, 0, stream>>>(static_cast<const T*>(src), mean_rstd, static_cast<const T*>(residual),       \
                                   static_cast<T*>(dst), a, rpb)
#define VCG_XF_MODES(T)                                                                     \
  switch (d->mode) {                                                                        \
    case VCG_MODE_PLAIN: VCG_XF_LAUNCH(T, (xform_fwd_kernel<T, VCG_MODE_PLAIN>)); break;     \
    case VCG_MODE_UNSHUFFLE: VCG_XF_LAUNCH(T, (xform_fwd_kernel<T, VCG_MODE_UNSHUFFLE>)); break; \
    case VCG_MODE_PAD_S2D: VCG_XF_LAUNCH(T, (xform_fwd_kernel<T, VCG_MODE_PAD_S2D>)); break; \
    default: VCG_XF_LAUNCH(T, xform_fwd_shuffle_kernel<T>); break;                          \
  }
  if (d->dtype == VCG_F32) { VCG_XF_MODES(float) } else { VCG_XF_MODES(__nv_bfloat16) }
#undef VCG_XF_MODES
#undef VCG_XF_LAUNCH
  VCG_CHECK_LAUNCH("xform_fwd_kernel");
  return VCG_OK;
}

static int fill_xb(const vcg_xbwd_desc* d, const vcg_gsrc* srcs, XbArgs& a) {
  VCG_REQUIRE(d->c % 8 == 0 && d->y_c % 8 == 0 && d->dy_c % 8 == 0, VCG_E_UNSUPPORTED, "xform_bwd: channel multiples of 8");
  VCG_REQUIRE(d->nsrc >= 0 && d->nsrc <= 3, VCG_E_INVALID, "xform_bwd: nsrc=%d", d->nsrc);
  a.n = d->n; a.h = d->h; a.w = d->w; a.c = d->c; a.y_c = d->y_c; a.norm = d->norm; a.act = d->act;
  a.pre_act = d->pre_act; a.dy_halo = d->dy_halo; a.dy_c = d->dy_c; a.nsrc = d->nsrc;
  a.stats_hw = d->stats_hw; a.clear_halo = d->clear_halo;
  a.l2pf = vcg_l2_prefetch();
  for (int k = 0; k < d->nsrc && srcs; ++k) {
    a.s[k].dxp = srcs[k].dxp; a.s[k].mode = srcs[k].mode; a.s[k].pad = srcs[k].pad; a.s[k].c_pitch = srcs[k].c_pitch;
    a.s[k].folded = srcs[k].folded;
    if (srcs[k].mode == VCG_MODE_SHUFFLE) VCG_REQUIRE(d->c % 32 == 0, VCG_E_UNSUPPORTED, "xform_bwd: shuffle needs c%%32==0");
  }
  return VCG_OK;
}

extern "C" int vcg_xform_bwd_gather(const vcg_xbwd_desc* d, const vcg_gsrc* srcs, const void* y, const float* mean_rstd,
                                    void* dy, float* gsums, float* dbias, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  XbArgs a{};
  int rc = fill_xb(d, srcs, a);
  if (rc) return rc;
  VCG_REQUIRE(!d->norm || (mean_rstd && gsums), VCG_E_INVALID, "xform_bwd_gather: norm needs statistics and gsums");
  const int zc = (d->c / 8 + 31) / 32, hw = d->h * d->w;
  const int cg_total = d->c / 8;
  bool fast = (cg_total & (cg_total - 1)) == 0 && (!d->norm || (reinterpret_cast<uintptr_t>(mean_rstd) & 15) == 0);
  // the fast kernel knows the piecewise-linear activations only (slope select); tanh / sigmoid take the generic one
  fast = fast && d->act <= VCG_ACT_LEAKY && d->pre_act <= VCG_ACT_LEAKY;
  for (int k = 0; k < d->nsrc; ++k) fast = fast && (srcs[k].pad == 0 || srcs[k].folded);
  if (fast) {
    XgArgs g{};
    g.n = d->n; g.h = d->h; g.w = d->w; g.c = d->c; g.y_c = d->y_c; g.norm = d->norm; g.act = d->act;
    g.pre_act = d->pre_act; g.dy_halo = d->dy_halo; g.dy_c = d->dy_c; g.nsrc = d->nsrc;
    g.stats_hw = d->stats_hw; g.clear_halo = d->clear_halo;
    g.l2pf = vcg_l2_prefetch();
    const size_t es = d->dtype == VCG_F32 ? 4 : 2;
    for (int k = 0; k < d->nsrc; ++k) {
      FSrc& f = g.s[k];
      const int pad = srcs[k].pad, pitch = srcs[k].c_pitch;
      long long first = 0;       // element offset of the direct position of activation pixel (0,0), channel 0
      f.pitch = pitch;
      switch (srcs[k].mode) {
        case VCG_MODE_PLAIN: {
          const int wp = d->w + 2 * pad;
          first = (static_cast<long long>(pad) * wp + pad) * pitch;
          f.rs = wp * pitch; f.cs = pitch; f.img_stride = static_cast<long long>(d->h + 2 * pad) * wp * pitch;
          break;
        }
        case VCG_MODE_UNSHUFFLE: {
          const int wp = d->w / 2 + 2 * pad;
          first = (static_cast<long long>(pad) * wp + pad) * pitch;
          f.sh = 1; f.msk = 1; f.rs = wp * pitch; f.cs = pitch; f.ph = 2 * d->c; f.pw = d->c;
          f.img_stride = static_cast<long long>(d->h / 2 + 2 * pad) * wp * pitch;
          break;
        }
        case VCG_MODE_PAD_S2D: {
          const int hp = (d->h + 2 * pad) / 2, wp = (d->w + 2 * pad) / 2;
          f.hoff = pad; f.woff = pad; f.sh = 1; f.msk = 1; f.rs = wp * pitch; f.cs = pitch; f.ph = 2 * d->c; f.pw = d->c;
          f.img_stride = static_cast<long long>(hp) * wp * pitch;
          break;
        }
        default: {
          const int wp = 2 * d->w + 2 * pad;
          first = (static_cast<long long>(pad) * wp + pad) * pitch;
          f.shuffle = 1; f.rs = 2 * wp * pitch; f.cs = 2 * pitch; f.ph = wp * pitch;
          f.img_stride = static_cast<long long>(2 * d->h + 2 * pad) * wp * pitch;
          break;
        }
      }
      VCG_REQUIRE(f.img_stride < (1LL << 31), VCG_E_UNSUPPORTED, "xform_bwd_gather: image too large for 32-bit offsets");
      f.base = static_cast<const char*>(srcs[k].dxp) + first * es;
    }
    const int ppb = pix_chunk_fast(hw, d->n, zc, cg_total, hw <= 1024 ? 4 : 8);
    dim3 grid((hw + ppb - 1) / ppb, d->n, zc);
    // 4 pixels per thread per iteration at 2 blocks/SM measured best on B200 (4.2 TB/s; 2 px x 3 blocks: 3.7)
    if (d->dtype == VCG_F32)
      xform_bwd_gather_kernel<float, 2, 2><<<grid, 256, 0, stream>>>(g, static_cast<const float*>(y), mean_rstd,
                                                                     static_cast<float*>(dy), gsums, dbias, ppb);
    else
      xform_bwd_gather_kernel<__nv_bfloat16, 4, 2><<<grid, 256, 0, stream>>>(g, static_cast<const __nv_bfloat16*>(y), mean_rstd,
                                                                             static_cast<__nv_bfloat16*>(dy), gsums, dbias, ppb);
    VCG_CHECK_LAUNCH("xform_bwd_gather_kernel");
    return VCG_OK;
  }
  const int ppb = pix_chunk(hw, d->n, zc);
  dim3 grid((hw + ppb - 1) / ppb, d->n, zc);
  if (d->dtype == VCG_F32)
    xform_bwd_generic_kernel<float, false><<<grid, 256, 0, stream>>>(a, static_cast<const float*>(y), mean_rstd, nullptr,
                                                                     static_cast<float*>(dy), gsums, dbias, ppb);
  else
    xform_bwd_generic_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>(a, static_cast<const __nv_bfloat16*>(y), mean_rstd,
                                                                             nullptr, static_cast<__nv_bfloat16*>(dy), gsums, dbias, ppb);
  VCG_CHECK_LAUNCH("xform_bwd_kernel<gather>");
  return VCG_OK;
}

// border index ranges (activation coordinates) of one axis: indices whose reflect mirrors exist
static void border_ranges(int mode, int L_act, int pad, int* a0, int* len0, int* b0, int* len1) {
  int lo0, lo1, hi0, hi1;      // inclusive
  if (mode == VCG_MODE_UNSHUFFLE) {
    const int Lh = L_act / 2;
    lo0 = 2; lo1 = 2 * pad + 1; hi0 = 2 * (Lh - 1 - pad); hi1 = 2 * (Lh - 2) + 1;
  } else if (mode == VCG_MODE_SHUFFLE) {
    const int L2 = 2 * L_act;
    lo0 = 0; lo1 = pad >> 1; hi0 = (L2 - 1 - pad) >> 1; hi1 = (L2 - 2) >> 1;
  } else {
    lo0 = 1; lo1 = pad; hi0 = L_act - 1 - pad; hi1 = L_act - 2;
  }
  if (lo0 < 0) lo0 = 0;
  if (hi1 > L_act - 1) hi1 = L_act - 1;
  if (hi0 < 0) hi0 = 0;
  if (lo1 > L_act - 1) lo1 = L_act - 1;
  if (lo1 + 1 >= hi0) {          // ranges touch or overlap: one range
    *a0 = lo0 < hi0 ? lo0 : hi0; *len0 = (hi1 > lo1 ? hi1 : lo1) - *a0 + 1; *b0 = *a0 + *len0; *len1 = 0;
  } else {
    *a0 = lo0; *len0 = lo1 - lo0 + 1; *b0 = hi0; *len1 = hi1 - hi0 + 1;
  }
}

extern "C" int vcg_fold_halo(int32_t dtype, void* dxp, int32_t n, int32_t h, int32_t w, int32_t c, int32_t mode,
                             int32_t pad, int32_t c_pitch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (pad <= 0) return VCG_OK;
  VCG_REQUIRE(c % 8 == 0 && c_pitch % 8 == 0 && n <= 65535, VCG_E_UNSUPPORTED, "fold_halo: c=%d pitch=%d n=%d", c, c_pitch, n);
  VCG_REQUIRE(mode != VCG_MODE_SHUFFLE || c % 32 == 0, VCG_E_UNSUPPORTED, "fold_halo: shuffle needs c%%32==0");
  const int dom_h = mode == VCG_MODE_UNSHUFFLE ? h / 2 : (mode == VCG_MODE_SHUFFLE ? 2 * h : h);
  const int dom_w = mode == VCG_MODE_UNSHUFFLE ? w / 2 : (mode == VCG_MODE_SHUFFLE ? 2 * w : w);
  VCG_REQUIRE(pad < dom_h && pad < dom_w, VCG_E_INVALID, "fold_halo: pad %d >= extent", pad);
  FoldArgs a{};
  a.n = n; a.h = h; a.w = w; a.c = c; a.mode = mode; a.pad = pad; a.pitch = c_pitch;
  border_ranges(mode, h, pad, &a.ra0, &a.rlen0, &a.rb0, &a.rlen1);
  border_ranges(mode, w, pad, &a.ca0, &a.clen0, &a.cb0, &a.clen1);
  const int nbh = a.rlen0 + a.rlen1, nbw = a.clen0 + a.clen1;
  a.total = nbh * w + (h - nbh) * nbw;
  const long long threads = static_cast<long long>(a.total) * (c / 8);
  dim3 grid(static_cast<unsigned>((threads + 255) / 256), n);
  if (dtype == VCG_F32) fold_halo_kernel<float><<<grid, 256, 0, stream>>>(static_cast<float*>(dxp), a);
  else fold_halo_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(dxp), a);
  VCG_CHECK_LAUNCH("fold_halo_kernel");
  return VCG_OK;
}

extern "C" int vcg_xform_bwd_norm(const vcg_xbwd_desc* d, const void* y, const float* mean_rstd, const float* gsums,
                                  void* dy, float* dbias, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  XbArgs a{};
  int rc = fill_xb(d, nullptr, a);
  if (rc) return rc;
  VCG_REQUIRE(d->norm && mean_rstd && gsums, VCG_E_INVALID, "xform_bwd_norm: needs norm statistics");
  VCG_REQUIRE(d->pre_act <= VCG_ACT_LEAKY, VCG_E_UNSUPPORTED, "xform_bwd_norm: pre-norm activation %d (ReLU / LeakyReLU only)", d->pre_act);
  const int zc = (d->c / 8 + 31) / 32, hw = d->h * d->w;
  const int cg_total = d->c / 8;
  if ((cg_total & (cg_total - 1)) == 0 && (reinterpret_cast<uintptr_t>(mean_rstd) & 15) == 0) {
    const int ppb = pix_chunk_fast(hw, d->n, zc, cg_total, hw >= 16384 ? 8 : (hw >= 4096 ? 4 : 2));
    dim3 grid((hw + ppb - 1) / ppb, d->n, zc);
    if (d->dtype == VCG_F32)
      xform_bwd_norm_kernel<float><<<grid, 256, 0, stream>>>(a, static_cast<const float*>(y), mean_rstd, gsums,
                                                             static_cast<float*>(dy), dbias, ppb);
    else
      xform_bwd_norm_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(a, static_cast<const __nv_bfloat16*>(y), mean_rstd, gsums,
                                                                     static_cast<__nv_bfloat16*>(dy), dbias, ppb);
    VCG_CHECK_LAUNCH("xform_bwd_norm_kernel");
    return VCG_OK;
  }
  const int ppb = pix_chunk(hw, d->n, zc);
  dim3 grid((hw + ppb - 1) / ppb, d->n, zc);
  if (d->dtype == VCG_F32)
    xform_bwd_generic_kernel<float, true><<<grid, 256, 0, stream>>>(a, static_cast<const float*>(y), mean_rstd, gsums,
                                                                    static_cast<float*>(dy), nullptr, dbias, ppb);
  else
    xform_bwd_generic_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>(a, static_cast<const __nv_bfloat16*>(y), mean_rstd,
                                                                            gsums, static_cast<__nv_bfloat16*>(dy), nullptr, dbias, ppb);
  VCG_CHECK_LAUNCH("xform_bwd_kernel<norm>");
  return VCG_OK;
}

extern "C" int vcg_pack_nchw(int32_t dtype, const float* src, int32_t n, int32_t c, int32_t h, int32_t w, void* dst,
                             int32_t dst_c, int32_t dst_halo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VCG_REQUIRE(dst_c % 8 == 0 && dst_c >= c, VCG_E_INVALID, "pack_nchw: dst_c=%d c=%d", dst_c, c);
  const long long total = static_cast<long long>(n) * (h + 2 * dst_halo) * (w + 2 * dst_halo) * (dst_c / 8);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (dtype == VCG_F32)
    pack_nchw_kernel<float><<<blocks, 256, 0, stream>>>(src, n, c, h, w, static_cast<float*>(dst), dst_c, dst_halo, total);
  else
    pack_nchw_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(src, n, c, h, w, static_cast<__nv_bfloat16*>(dst), dst_c,
                                                                dst_halo, total);
  VCG_CHECK_LAUNCH("pack_nchw_kernel");
  return VCG_OK;
}

extern "C" int vcg_unpack_nchw(int32_t dtype, const void* src, int32_t src_c, int32_t n, int32_t c, int32_t h, int32_t w,
                               float* dst, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long total = static_cast<long long>(n) * c * h * w;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (dtype == VCG_F32)
    unpack_nchw_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(src), src_c, n, c, h, w, dst, total);
  else
    unpack_nchw_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), src_c, n, c, h, w,
                                                                  dst, total);
  VCG_CHECK_LAUNCH("unpack_nchw_kernel");
  return VCG_OK;
}

extern "C" int vcg_zero_halo(int32_t dtype, void* buf, int32_t n, int32_t h, int32_t w, int32_t c, int32_t halo, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (halo <= 0) return VCG_OK;
  const int es = dtype == VCG_F32 ? 4 : 2;
  VCG_REQUIRE((c * es) % 16 == 0, VCG_E_UNSUPPORTED, "zero_halo: pixel size must be a multiple of 16 bytes");
  const int vec = c * es / 16;
  const long long total = static_cast<long long>(n) * ((h + 2 * halo) * (w + 2 * halo) - h * w) * vec;
  long long blocks = (total + 255) / 256;
  const long long cap = 8LL * vcg_num_sms();
  if (blocks > cap) blocks = cap;
  zero_halo_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<uint4*>(buf), n, h, w, halo, vec);
  VCG_CHECK_LAUNCH("zero_halo_kernel");
  return VCG_OK;
}

extern "C" int vcg_zero(void* p, size_t bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!bytes) return VCG_OK;
  VCG_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, VCG_E_INVALID, "vcg_zero: pointer must be 16-byte aligned");
  const size_t n16 = bytes / 16, ntail = bytes % 16;
  size_t blocks = (n16 + 255) / 256;
  const size_t cap = static_cast<size_t>(vcg_num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  zero_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<float4*>(p), n16,
                                                                 static_cast<char*>(p) + n16 * 16, ntail);
  VCG_CHECK_LAUNCH("zero_kernel");
  return VCG_OK;
}
