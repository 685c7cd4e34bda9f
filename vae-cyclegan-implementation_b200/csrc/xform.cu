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
  int res_hp, res_wp, res_c, res_off;
};

template <typename T>
__global__ void __launch_bounds__(256)
xform_fwd_kernel(const T* __restrict__ src, const float* __restrict__ mr, const T* __restrict__ res,
                 T* __restrict__ dst, XfArgs p, long long total) {
  // grid.y = destination row (n*hd + a), grid.x covers the (pixel b, 8-channel group g) pairs of that row:
  // one 32-bit division per thread instead of four 64-bit ones (the kernel is a byte mover)
  const int groups = p.dst_c / 8;
  const int xi = blockIdx.x * 256 + threadIdx.x;
  if (xi >= p.wd * groups) return;
  const int b = xi / groups, g = xi - b * groups;
  for (int rowi = blockIdx.y; rowi < p.n * p.hd; rowi += gridDim.y) {
  const int n = rowi / p.hd, a = rowi - n * p.hd;
  const int cd0 = g * 8;
  float v[8];
  T* out = dst + ((static_cast<size_t>(n) * p.hd + a) * p.wd + b) * p.dst_c + cd0;
  if (cd0 >= p.cd) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    st8<T>(out, v);
    continue;
  }
  int sh, sw, sc0, sstride = 1;
  if (p.mode == VCG_MODE_PLAIN) {
    sh = reflect_idx(a - p.pad, p.h); sw = reflect_idx(b - p.pad, p.w); sc0 = cd0;
  } else if (p.mode == VCG_MODE_SHUFFLE) {
    const int A = reflect_idx(a - p.pad, 2 * p.h), B = reflect_idx(b - p.pad, 2 * p.w);
    sh = A >> 1; sw = B >> 1; sc0 = cd0 * 4 + (A & 1) * 2 + (B & 1); sstride = 4;
  } else if (p.mode == VCG_MODE_UNSHUFFLE) {
    const int A = reflect_idx(a - p.pad, p.h / 2), B = reflect_idx(b - p.pad, p.w / 2);
    const int sub = cd0 / p.c;
    sh = 2 * A + (sub >> 1); sw = 2 * B + (sub & 1); sc0 = cd0 - sub * p.c;
  } else {  // PAD_S2D
    const int sub = cd0 / p.c;
    sh = reflect_idx(2 * a + (sub >> 1) - p.pad, p.h); sw = reflect_idx(2 * b + (sub & 1) - p.pad, p.w);
    sc0 = cd0 - sub * p.c;
  }
  const size_t spix = (static_cast<size_t>(n) * p.h + sh) * p.w + sw;
  const T* sp = src + spix * p.src_c + sc0;
  if (sstride == 1) ld8<T>(sp, v);
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = Elem<T>::ld(sp + j * 4);
  }
  if (p.norm) {
    const float* m = mr + (static_cast<size_t>(n) * p.c + sc0) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (v[j] - m[2 * j * sstride]) * m[2 * j * sstride + 1];
  }
  if (p.act) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act_apply(v[j], p.act);
  }
  if (res) {
    const T* rp = res + ((static_cast<size_t>(n) * p.res_hp + sh + p.res_off) * p.res_wp + sw + p.res_off) * p.res_c + sc0;
    if (sstride == 1) { float r[8]; ld8<T>(rp, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j]; }
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += Elem<T>::ld(rp + j * 4);
    }
  }
  st8<T>(out, v);
  }
}

// ------------------------------------------------------------------ backward transform
struct GSrc { const void* dxp; int mode, pad, c_pitch; };
struct XbArgs {
  int n, h, w, c, y_c, norm, act, pre_act, dy_halo, dy_c, nsrc;
  GSrc s[3];
};

// k-th padded coordinate t in [0, L+2p) whose reflect source is i (k=0 direct, 1 low mirror, 2 high mirror); -1 = none
__device__ __forceinline__ int mirror_k(int i, int L, int p, int k) {
  if (k == 0) return i + p;
  if (k == 1) return (i >= 1 && i <= p) ? p - i : -1;
  return (i >= L - 1 - p && i <= L - 2) ? p + 2 * (L - 1) - i : -1;
}

// f(th, tw) for every padded position that reflects onto (i, j); interior pixels (the vast majority) take one call
template <typename F>
__device__ __forceinline__ void for_mirrors(int i, int Lh, int j, int Lw, int pad, F&& f) {
  if (pad == 0 || ((i > pad) & (i < Lh - 1 - pad) & (j > pad) & (j < Lw - 1 - pad))) { f(i + pad, j + pad); return; }
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
    for_mirrors(h, p.h, w, p.w, pad, [&](int th, int tw) {
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
      for_mirrors(2 * h + (sub >> 1), H2, 2 * w + (sub & 1), W2, pad, [&](int th, int tw) {
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
    for_mirrors(h >> 1, Hh, w >> 1, Wh, pad, [&](int th, int tw) {
      float v[8];
      ld8<T>(base + (static_cast<size_t>(th) * wp + tw) * pitch, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) g[q] += v[q];
    });
  } else {  // PAD_S2D: mirrors live in the padded full-resolution domain, then map to (pixel/2, sub-pixel channel block)
    const int hp = (p.h + 2 * pad) / 2, wp = (p.w + 2 * pad) / 2;
    const T* base = dxp + static_cast<size_t>(n) * hp * wp * pitch + ch;
    for_mirrors(h, p.h, w, p.w, pad, [&](int th, int tw) {
      const int sub = (th & 1) * 2 + (tw & 1);
      float v[8];
      ld8<T>(base + (static_cast<size_t>(th >> 1) * wp + (tw >> 1)) * pitch + sub * p.c, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) g[q] += v[q];
    });
  }
}

// grid: (pixel chunks, n, channel-group chunks of 32); block 256 = 32 channel groups x 8 pixel lanes
template <typename T, bool PHASE2>
__global__ void __launch_bounds__(256, PHASE2 ? 2 : 3)
xform_bwd_kernel(const __grid_constant__ XbArgs p, const T* __restrict__ y, const float* __restrict__ mr, const float* __restrict__ gsums_in,
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
      for (int j = 0; j < 8; ++j) { mean[j] = m[2 * j]; rstd[j] = m[2 * j + 1]; }
      if (PHASE2) {
        const float* gs = gsums_in + (static_cast<size_t>(n) * p.c + ch) * 2;
        const float inv = 1.f / hw;
#pragma unroll
        for (int j = 0; j < 8; ++j) { m1[j] = gs[2 * j] * inv; m2[j] = gs[2 * j + 1] * inv; }
      }
    }
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
            g[j] *= act_grad(z, p.act);
            s1[j] += g[j]; s2[j] += g[j] * z;
          }
        } else {
          // no norm: at most one activation (fused in the conv epilogue or applied after): y or act(y) share sign
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            g[j] *= act_grad(yv[j], p.act) * act_grad(yv[j], p.pre_act);
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
  a.res_hp = d->res_hp; a.res_wp = d->res_wp; a.res_c = d->res_c; a.res_off = d->res_off;
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
  const long long total = static_cast<long long>(d->n) * a.hd * a.wd * (d->dst_c / 8);
  const long long rows = static_cast<long long>(d->n) * a.hd;
  const dim3 blocks((a.wd * (d->dst_c / 8) + 255) / 256, static_cast<unsigned>(rows < 65535 ? rows : 65535));
  if (d->dtype == VCG_F32)
    xform_fwd_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(src), mean_rstd,
                                                        static_cast<const float*>(residual), static_cast<float*>(dst), a, total);
  else
    xform_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), mean_rstd,
                                                                static_cast<const __nv_bfloat16*>(residual),
                                                                static_cast<__nv_bfloat16*>(dst), a, total);
  VCG_CHECK_LAUNCH("xform_fwd_kernel");
  return VCG_OK;
}

static int fill_xb(const vcg_xbwd_desc* d, const vcg_gsrc* srcs, XbArgs& a) {
  VCG_REQUIRE(d->c % 8 == 0 && d->y_c % 8 == 0 && d->dy_c % 8 == 0, VCG_E_UNSUPPORTED, "xform_bwd: channel multiples of 8");
  VCG_REQUIRE(d->nsrc >= 0 && d->nsrc <= 3, VCG_E_INVALID, "xform_bwd: nsrc=%d", d->nsrc);
  a.n = d->n; a.h = d->h; a.w = d->w; a.c = d->c; a.y_c = d->y_c; a.norm = d->norm; a.act = d->act;
  a.pre_act = d->pre_act; a.dy_halo = d->dy_halo; a.dy_c = d->dy_c; a.nsrc = d->nsrc;
  for (int k = 0; k < d->nsrc && srcs; ++k) {
    a.s[k].dxp = srcs[k].dxp; a.s[k].mode = srcs[k].mode; a.s[k].pad = srcs[k].pad; a.s[k].c_pitch = srcs[k].c_pitch;
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
  const int ppb = pix_chunk(hw, d->n, zc);
  dim3 grid((hw + ppb - 1) / ppb, d->n, zc);
  if (d->dtype == VCG_F32)
    xform_bwd_kernel<float, false><<<grid, 256, 0, stream>>>(a, static_cast<const float*>(y), mean_rstd, nullptr,
                                                             static_cast<float*>(dy), gsums, dbias, ppb);
  else
    xform_bwd_kernel<__nv_bfloat16, false><<<grid, 256, 0, stream>>>(a, static_cast<const __nv_bfloat16*>(y), mean_rstd,
                                                                     nullptr, static_cast<__nv_bfloat16*>(dy), gsums, dbias, ppb);
  VCG_CHECK_LAUNCH("xform_bwd_kernel<gather>");
  return VCG_OK;
}

extern "C" int vcg_xform_bwd_norm(const vcg_xbwd_desc* d, const void* y, const float* mean_rstd, const float* gsums,
                                  void* dy, float* dbias, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  XbArgs a{};
  int rc = fill_xb(d, nullptr, a);
  if (rc) return rc;
  VCG_REQUIRE(d->norm && mean_rstd && gsums, VCG_E_INVALID, "xform_bwd_norm: needs norm statistics");
  const int zc = (d->c / 8 + 31) / 32, hw = d->h * d->w;
  const int ppb = pix_chunk(hw, d->n, zc);
  dim3 grid((hw + ppb - 1) / ppb, d->n, zc);
  if (d->dtype == VCG_F32)
    xform_bwd_kernel<float, true><<<grid, 256, 0, stream>>>(a, static_cast<const float*>(y), mean_rstd, gsums,
                                                            static_cast<float*>(dy), nullptr, dbias, ppb);
  else
    xform_bwd_kernel<__nv_bfloat16, true><<<grid, 256, 0, stream>>>(a, static_cast<const __nv_bfloat16*>(y), mean_rstd,
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
