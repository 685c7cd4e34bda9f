// wpack.cu -- filter layout conversion at the state_dict boundary, sm_100a.
//
// Master weights stay fp32 OIHW exactly as the reference's state_dict holds them
// (e.g. G.encoder.model.1.conv.weight (128,256,3,3), Networks.py:87); the kernels consume
//   forward       : [cout_pad][kh][kwc_pad]   K index = (kw, physical cin)          (K-major B operand)
//   data-gradient : [cin_pad ][kh][kwc_pad]   K index = (flipped kw, physical cout), taps flipped
// where "physical cin" already contains the PixelUnshuffle channel order used by vcg_xform_fwd
// ((i,j,c) instead of the reference's c*4+i*2+j, Networks.py:86) or the space-to-depth order of the
// stride-2 discriminator convolutions (Networks.py:244-247).  vcg_wunpack_grad is the inverse map for
// the fp32 weight gradient produced by the wgrad GEMM.
#include "common.cuh"

namespace {

struct WpArgs {
  int co, ci, kh, kw, wmap, c_phys, co_phys, rows_pad, pkh, pkw, kwc_pad, tflip;
};

// physical input channel -> (ci, dkh, dkw); returns false when the channel is padding
__device__ __forceinline__ bool map_phys(const WpArgs& p, int ph, int& ci, int& dkh, int& dkw) {
  dkh = dkw = 0;
  if (p.wmap == VCG_WMAP_PLAIN) { ci = ph; return ph < p.ci; }
  if (p.wmap == VCG_WMAP_UNSHUFFLE) {
    const int c0 = p.ci / 4;
    if (ph >= p.ci) return false;
    const int sub = ph / c0, c = ph - sub * c0;
    ci = c * 4 + sub;
    return true;
  }
  const int cp = p.c_phys / 4;   // S2D: physical channels per sub-pixel
  const int sub = ph / cp, c = ph - sub * cp;
  ci = c; dkh = sub >> 1; dkw = sub & 1;
  return sub < 4 && c < p.ci;
}

template <typename T>
__global__ void wpack_kernel(const float* __restrict__ w, T* __restrict__ out, WpArgs p, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int q = static_cast<int>(idx % p.kwc_pad);
  long long t = idx / p.kwc_pad;
  const int khp = static_cast<int>(t % p.pkh);
  const int r = static_cast<int>(t / p.pkh);
  const int s = (p.wmap == VCG_WMAP_S2D) ? 2 : 1;
  const int kext = p.tflip ? p.co_phys : p.c_phys;
  const int kwp = q / kext, e = q - kwp * kext;
  float v = 0.f;
  if (kwp < p.pkw) {
    int co, ph, kh_e = khp, kw_e = kwp;
    if (p.tflip) { ph = r; co = e; kh_e = p.pkh - 1 - khp; kw_e = p.pkw - 1 - kwp; }
    else { co = r; ph = e; }
    int ci, dkh, dkw;
    if (co < p.co && ph < p.c_phys && map_phys(p, ph, ci, dkh, dkw))
      v = w[((static_cast<size_t>(co) * p.ci + ci) * p.kh + kh_e * s + dkh) * p.kw + kw_e * s + dkw];
  }
  Elem<T>::st(out + idx, v);
}

__global__ void wunpack_kernel(const float* __restrict__ dwp, float* __restrict__ g, WpArgs p, int accumulate,
                               long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;   // one thread per OIHW element
  const int kw = static_cast<int>(idx % p.kw);
  long long t = idx / p.kw;
  const int kh = static_cast<int>(t % p.kh); t /= p.kh;
  const int ci = static_cast<int>(t % p.ci);
  const int co = static_cast<int>(t / p.ci);
  int ph, khp = kh, kwp = kw;
  if (p.wmap == VCG_WMAP_PLAIN) ph = ci;
  else if (p.wmap == VCG_WMAP_UNSHUFFLE) { const int c0 = p.ci / 4; ph = (ci & 3) * c0 + (ci >> 2); }
  else { khp = kh >> 1; kwp = kw >> 1; ph = ((kh & 1) * 2 + (kw & 1)) * (p.c_phys / 4) + ci; }
  const float v = dwp[(static_cast<size_t>(co) * p.pkh + khp) * p.kwc_pad + kwp * p.c_phys + ph];
  g[idx] = accumulate ? g[idx] + v : v;
}

// ---- multi-tensor variants: ONE launch converts every filter of a network ------------------------------
// A block owns a (CO_T output channels) x (32 input channels) x (all taps) tile of one filter and moves it through
// shared memory, so that the OIHW side is accessed as contiguous runs of 32*taps floats and the packed side as
// runs of 32 channels: both sides coalesced (the per-element kernels above read with a stride of kh*kw floats).
struct WJob {
  WpArgs a;
  float* oihw;          // pack: fp32 master weight (read); unpack: gradient (written)
  void* packed;         // forward layout (pack: dtype T; unpack: fp32 dw accumulator, re-zeroed after reading)
  void* packed_t;       // pack only: data-gradient layout (may be null)
  int t_kwc_pad;        // row length per tap row of packed_t
  int accumulate;       // unpack: grad += instead of grad =
  int co_t, tiles_ci, tile0, ntiles;
  int vec, reserved;
};

constexpr int kWTileFloats = 10240;   // 40 KB: 32 x 32 x 9(+1) fp32 for 3x3 filters

__device__ __forceinline__ void map_tap(const WpArgs& p, int ci, int kh, int kw, int& ph, int& khp, int& kwp) {
  khp = kh; kwp = kw;
  if (p.wmap == VCG_WMAP_PLAIN) ph = ci;
  else if (p.wmap == VCG_WMAP_UNSHUFFLE) { const int c0 = p.ci / 4; ph = (ci & 3) * c0 + (ci >> 2); }
  else { khp = kh >> 1; kwp = kw >> 1; ph = ((kh & 1) * 2 + (kw & 1)) * (p.c_phys / 4) + ci; }
}

// the job record of this block, staged through shared memory (one global read per block, no per-use reloads)
__device__ __forceinline__ WJob find_job(const WJob* __restrict__ jobs, int njobs, int tile) {
  __shared__ WJob s_job;
  if (threadIdx.x < 32) {
    // 32 lanes test 32 jobs at a time: the last job whose first tile is <= tile
    int found = 0;
    for (int j0 = 0; j0 < njobs; j0 += 32) {
      const int j = j0 + threadIdx.x;
      const bool ok = j < njobs && jobs[j].tile0 <= tile;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (m) found = j0 + 31 - __clz(m);
    }
    if (threadIdx.x == 0) s_job = jobs[found];
  }
  __syncthreads();
  return s_job;
}

// per-tap offset tables (shared memory): packed offsets are separable into an output-channel term, a tap term and an
// input-channel term, so the hot loops need no division:
//   forward layout  : off = co * (pkh*kwc_pad)      + t_fwd[tap] + ph_ci(ci)
//   data-grad layout: off = ph_ci(ci) * (pkh*t_kwc) + t_bwd[tap] + co
__device__ __forceinline__ int ph_ci(const WpArgs& p, int ci) {
  if (p.wmap == VCG_WMAP_UNSHUFFLE) { const int c0 = p.ci / 4; return (ci & 3) * c0 + (ci >> 2); }
  return ci;
}
__device__ __forceinline__ void build_tap_tables(const WpArgs& p, int t_kwc, int* t_fwd, int* t_bwd) {
  const int taps = p.kh * p.kw;
  for (int tap = threadIdx.x; tap < taps; tap += blockDim.x) {
    int ph, khp, kwp;
    map_tap(p, 0, tap / p.kw, tap % p.kw, ph, khp, kwp);       // ph = tap-dependent part of the physical channel
    if (p.wmap != VCG_WMAP_S2D) ph = 0;
    t_fwd[tap] = khp * p.kwc_pad + kwp * p.c_phys + ph;
    t_bwd[tap] = ph * p.pkh * t_kwc + (p.pkh - 1 - khp) * t_kwc + (p.pkw - 1 - kwp) * p.co_phys;
  }
}

// ---- 16-byte vector path (3x3 filters, 32 co x 32 ci tiles).  Shared tile s[co_l][ci_l * 9 + tap] with an ODD row
// pitch (289 floats) so that reading 8 output channels of one (ci, tap) is conflict free.  A lane owns 8 channels of a
// packed row: PLAIN cil = part*8 + i;  UNSHUFFLE cil = 4*i + part (those map to 8 consecutive physical channels).
constexpr int kVecPitch = 289;

__device__ __forceinline__ int vec_cil(const WpArgs& p, int part, int i) {
  return p.wmap == VCG_WMAP_UNSHUFFLE ? 4 * i + part : part * 8 + i;
}
__device__ __forceinline__ int vec_ph0(const WpArgs& p, int ci0, int part) {      // first physical channel of the lane's 8
  return p.wmap == VCG_WMAP_UNSHUFFLE ? part * (p.ci / 4) + (ci0 >> 2) : ci0 + part * 8;
}

template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) { st8<float>(p, v); }
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) { st8<__nv_bfloat16>(p, v); }

template <typename T>
__device__ __forceinline__ void wpack_tile_vec(const WJob& jb, const WpArgs& p, int co0, int ci0, float* s_tile) {
  const int tid = threadIdx.x;
  // phase 1: 32 OIHW runs of 288 floats -> s[col][k]  (float4 loads, 4 in flight)
  for (int e0 = tid; e0 < 32 * 72; e0 += 4 * 256) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * 256;
      if (e < 32 * 72) {
        const int col = e / 72, k4 = e - col * 72;
        v[u] = *reinterpret_cast<const float4*>(jb.oihw + (static_cast<size_t>(co0 + col) * p.ci + ci0) * 9 + 4 * k4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * 256;
      if (e < 32 * 72) {
        const int col = e / 72, k4 = e - col * 72;
        float* d = s_tile + col * kVecPitch + 4 * k4;
        d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
      }
    }
  }
  __syncthreads();
  const int part = tid & 3, rlane = tid >> 2;           // 64 row slots per pass, 4 lanes (x 8 channels) per row
  // phase 2a: forward layout rows (co, tap): 32 input channels
  {
    T* out = static_cast<T*>(jb.packed);
    const int ph0 = vec_ph0(p, ci0, part);
    for (int r = rlane; r < 32 * 9; r += 64) {
      const int col = r / 9, tap = r - col * 9;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = s_tile[col * kVecPitch + vec_cil(p, part, i) * 9 + tap];
      store8<T>(out + (static_cast<size_t>(co0 + col) * 3 + tap / 3) * p.kwc_pad + (tap % 3) * p.c_phys + ph0, v);
    }
  }
  // phase 2b: data-gradient layout rows (physical ci, flipped tap): 32 output channels
  if (jb.packed_t) {
    T* outT = static_cast<T*>(jb.packed_t);
    for (int r = rlane; r < 32 * 9; r += 64) {
      const int cil = r / 9, tap = r - cil * 9;
      const int ph = p.wmap == VCG_WMAP_UNSHUFFLE ? (cil & 3) * (p.ci / 4) + ((ci0 + cil) >> 2) : ci0 + cil;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = s_tile[(part * 8 + i) * kVecPitch + cil * 9 + tap];
      store8<T>(outT + (static_cast<size_t>(ph) * 3 + (2 - tap / 3)) * jb.t_kwc_pad + (2 - tap % 3) * p.co_phys + co0 + part * 8, v);
    }
  }
}

__device__ __forceinline__ void wunpack_tile_vec(const WJob& jb, const WpArgs& p, int co0, int ci0, float* s_tile) {
  const int tid = threadIdx.x;
  const int part = tid & 3, rlane = tid >> 2;
  // phase 1: packed fp32 rows (co, tap) x 8 channels per lane -> s[col][cil*9+tap]; accumulator re-zeroed
  {
    float* dwp = static_cast<float*>(jb.packed);
    const int ph0 = vec_ph0(p, ci0, part);
    for (int r0 = rlane; r0 < 32 * 9; r0 += 2 * 64) {
      float4 a[2][2];
      float* ptr[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int r = r0 + 64 * u;
        ptr[u] = nullptr;
        if (r < 32 * 9) {
          const int col = r / 9, tap = r - col * 9;
          ptr[u] = dwp + (static_cast<size_t>(co0 + col) * 3 + tap / 3) * p.kwc_pad + (tap % 3) * p.c_phys + ph0;
          a[u][0] = *reinterpret_cast<const float4*>(ptr[u]);
          a[u][1] = *reinterpret_cast<const float4*>(ptr[u] + 4);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (ptr[u]) {
          const int r = r0 + 64 * u;
          const int col = r / 9, tap = r - col * 9;
          const float v[8] = {a[u][0].x, a[u][0].y, a[u][0].z, a[u][0].w, a[u][1].x, a[u][1].y, a[u][1].z, a[u][1].w};
#pragma unroll
          for (int i = 0; i < 8; ++i) s_tile[col * kVecPitch + vec_cil(p, part, i) * 9 + tap] = v[i];
          // (the accumulator is re-zeroed AFTER the barrier: a store to the address of a load still in flight
          //  stalls the thread until the load returns -- measured 330 us of a 540 us launch)
        }
      }
    }
  }
  __syncthreads();
  // re-zero the accumulator rows this block read (stores only)
  {
    float* dwp = static_cast<float*>(jb.packed);
    const int ph0 = vec_ph0(p, ci0, part);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = rlane; r < 32 * 9; r += 64) {
      const int col = r / 9, tap = r - col * 9;
      float* q = dwp + (static_cast<size_t>(co0 + col) * 3 + tap / 3) * p.kwc_pad + (tap % 3) * p.c_phys + ph0;
      *reinterpret_cast<float4*>(q) = z;
      *reinterpret_cast<float4*>(q + 4) = z;
    }
  }
  // phase 2: OIHW runs of 288 floats (float4 read-modify-write, 4 in flight)
  const bool acc = jb.accumulate != 0;
  for (int e0 = tid; e0 < 32 * 72; e0 += 4 * 256) {
    float4 g[4];
    float* gp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * 256;
      gp[u] = nullptr;
      if (e < 32 * 72) {
        const int col = e / 72, k4 = e - col * 72;
        gp[u] = jb.oihw + (static_cast<size_t>(co0 + col) * p.ci + ci0) * 9 + 4 * k4;
        if (acc) g[u] = *reinterpret_cast<const float4*>(gp[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (gp[u]) {
        const int e = e0 + u * 256;
        const int col = e / 72, k4 = e - col * 72;
        const float* sp = s_tile + col * kVecPitch + 4 * k4;
        float4 o = make_float4(sp[0], sp[1], sp[2], sp[3]);
        if (acc) { o.x += g[u].x; o.y += g[u].y; o.z += g[u].z; o.w += g[u].w; }
        *reinterpret_cast<float4*>(gp[u]) = o;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
wpack_multi_kernel(const WJob* __restrict__ jobs, int njobs) {
  extern __shared__ float s_tile[];
  __shared__ int t_fwd[64], t_bwd[64];
  const WJob jb = find_job(jobs, njobs, blockIdx.x);
  const WpArgs p = jb.a;
  const int t = blockIdx.x - jb.tile0;
  const int co0 = (t / jb.tiles_ci) * jb.co_t, ci0 = (t % jb.tiles_ci) * 32;
  const int nco = min(jb.co_t, p.co - co0), nci = min(32, p.ci - ci0);
  const int taps = p.kh * p.kw, ts = taps | 1;
  if (jb.vec) { wpack_tile_vec<T>(jb, p, co0, ci0, s_tile); return; }
  build_tap_tables(p, jb.t_kwc_pad, t_fwd, t_bwd);
  // phase 1: OIHW runs (nci*taps contiguous floats per output channel) -> s[co_l][ci_l][tap]; 4 loads in flight
  const float* __restrict__ wsrc = jb.oihw;
  const int run = nci * taps, n1 = nco * run;
  for (int e0 = threadIdx.x; e0 < n1; e0 += 4 * 256) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * 256;
      if (e < n1) {
        const int col = e / run, k = e - col * run;
        v[u] = wsrc[(static_cast<size_t>(co0 + col) * p.ci + ci0) * taps + k];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * 256;
      if (e < n1) {
        const int col = e / run, k = e - col * run;
        const int cil = k / taps, tap = k - cil * taps;
        s_tile[(col * 32 + cil) * ts + tap] = v[u];
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // phase 2a: forward layout: one (output channel, tap) row of 32 input channels per warp iteration
  {
    T* out = static_cast<T*>(jb.packed);
    const size_t co_stride = static_cast<size_t>(p.pkh) * p.kwc_pad;
    const int pc = lane < nci ? ph_ci(p, ci0 + lane) : 0;
    int tap = warp % taps, col = warp / taps;           // rows r = warp, warp+8, ... ; r = col*taps + tap
    for (; col < nco; ) {
      if (lane < nci) Elem<T>::st(out + (co0 + col) * co_stride + t_fwd[tap] + pc, s_tile[(col * 32 + lane) * ts + tap]);
      tap += 8;
      while (tap >= taps) { tap -= taps; ++col; }
    }
  }
  // phase 2b: data-gradient layout (rows = physical cin, taps flipped): one (input channel, tap) row of co_t
  // output channels per warp iteration
  if (jb.packed_t) {
    T* outT = static_cast<T*>(jb.packed_t);
    const size_t ci_stride = static_cast<size_t>(p.pkh) * jb.t_kwc_pad;
    int tap = warp % taps, cil = warp / taps;
    for (; cil < nci; ) {
      if (lane < nco)
        Elem<T>::st(outT + ph_ci(p, ci0 + cil) * ci_stride + t_bwd[tap] + co0 + lane, s_tile[(lane * 32 + cil) * ts + tap]);
      tap += 8;
      while (tap >= taps) { tap -= taps; ++cil; }
    }
  }
}

__global__ void __launch_bounds__(256)
wunpack_multi_kernel(const WJob* __restrict__ jobs, int njobs) {
  extern __shared__ float s_tile[];
  __shared__ int t_fwd[64], t_bwd[64];
  const WJob jb = find_job(jobs, njobs, blockIdx.x);
  const WpArgs p = jb.a;
  const int t = blockIdx.x - jb.tile0;
  const int co0 = (t / jb.tiles_ci) * jb.co_t, ci0 = (t % jb.tiles_ci) * 32;
  const int nco = min(jb.co_t, p.co - co0), nci = min(32, p.ci - ci0);
  const int taps = p.kh * p.kw, ts = taps | 1;
  if (jb.vec) { wunpack_tile_vec(jb, p, co0, ci0, s_tile); return; }
  build_tap_tables(p, 0, t_fwd, t_bwd);
  __syncthreads();
  // phase 1: packed fp32 accumulator rows (32 input channels of one (output channel, tap)) -> s[co_l][ci_l][tap];
  // the accumulator is re-zeroed.  Four rows per warp iteration: the loads are issued before the stores.
  {
    float* __restrict__ dwp = static_cast<float*>(jb.packed);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t co_stride = static_cast<size_t>(p.pkh) * p.kwc_pad;
    const int pc = lane < nci ? ph_ci(p, ci0 + lane) : 0;
    const int nrows = nco * taps;
    for (int r0 = warp; r0 < nrows; r0 += 32) {
      float v[4];
      float* ptr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + 8 * u;
        ptr[u] = nullptr;
        if (r < nrows && lane < nci) {
          const int col = r / taps, tap = r - col * taps;
          ptr[u] = dwp + (co0 + col) * co_stride + t_fwd[tap] + pc;
          v[u] = *ptr[u];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ptr[u]) {
          const int r = r0 + 8 * u;
          const int col = r / taps, tap = r - col * taps;
          s_tile[(col * 32 + lane) * ts + tap] = v[u];
        }
      }
    }
    __syncthreads();
    for (int r = warp; r < nrows; r += 8) {        // re-zero after the loads have landed (see the vector path)
      if (lane < nci) {
        const int col = r / taps, tap = r - col * taps;
        dwp[(co0 + col) * co_stride + t_fwd[tap] + pc] = 0.f;
      }
    }
  }
  // phase 2: OIHW runs: the tile's rows are contiguous runs of nci*taps floats, CI*taps apart
  float* __restrict__ g = jb.oihw;
  const int run = nci * taps;
  const int n2 = nco * run;
  const bool acc = jb.accumulate != 0;
  for (int e0 = threadIdx.x; e0 < n2; e0 += 4 * 256) {
    float v[4];
    size_t off[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * 256;
      off[u] = static_cast<size_t>(-1);
      if (e < n2) {
        const int col = e / run, k = e - col * run;
        off[u] = (static_cast<size_t>(co0 + col) * p.ci + ci0) * taps + k;
        const int cil = k / taps, tap = k - cil * taps;
        v[u] = s_tile[(col * 32 + cil) * ts + tap];
        if (acc) v[u] += g[off[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (off[u] != static_cast<size_t>(-1)) g[off[u]] = v[u];
  }
}

WpArgs to_args(const vcg_wpack_desc* d) {
  WpArgs a;
  a.co = d->co; a.ci = d->ci; a.kh = d->kh; a.kw = d->kw; a.wmap = d->wmap; a.c_phys = d->c_phys;
  a.co_phys = d->co_phys; a.rows_pad = d->rows_pad; a.pkh = d->pkh; a.pkw = d->pkw; a.kwc_pad = d->kwc_pad;
  a.tflip = d->transpose_flip;
  return a;
}

}  // namespace

extern "C" int vcg_wpack(const vcg_wpack_desc* d, const float* w_oihw, void* packed, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int kext = d->transpose_flip ? d->co_phys : d->c_phys;
  VCG_REQUIRE(d->pkw * kext <= d->kwc_pad, VCG_E_INVALID, "wpack: kwc_pad=%d < %d", d->kwc_pad, d->pkw * kext);
  VCG_REQUIRE(d->wmap != VCG_WMAP_UNSHUFFLE || d->ci % 4 == 0, VCG_E_INVALID, "wpack: unshuffle needs ci%%4==0");
  VCG_REQUIRE(d->wmap != VCG_WMAP_S2D || (d->kh == 2 * d->pkh && d->kw == 2 * d->pkw && d->c_phys % 4 == 0),
              VCG_E_INVALID, "wpack: bad S2D geometry");
  const long long total = static_cast<long long>(d->rows_pad) * d->pkh * d->kwc_pad;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  const WpArgs a = to_args(d);
  if (d->dtype == VCG_F32) wpack_kernel<float><<<blocks, 256, 0, stream>>>(w_oihw, static_cast<float*>(packed), a, total);
  else wpack_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(w_oihw, static_cast<__nv_bfloat16*>(packed), a, total);
  VCG_CHECK_LAUNCH("wpack_kernel");
  return VCG_OK;
}

extern "C" int vcg_wunpack_grad(const vcg_wpack_desc* d, const float* dw_packed, float* grad_oihw, int32_t accumulate,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VCG_REQUIRE(!d->transpose_flip, VCG_E_INVALID, "wunpack_grad: expects the forward layout descriptor");
  const long long total = static_cast<long long>(d->co) * d->ci * d->kh * d->kw;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  wunpack_kernel<<<blocks, 256, 0, stream>>>(dw_packed, grad_oihw, to_args(d), accumulate, total);
  VCG_CHECK_LAUNCH("wunpack_kernel");
  return VCG_OK;
}

namespace {
struct VecJob { float* src; float* dst; int n, pad; };
// dst[i] += src[i]; src[i] = 0  for every job (bias-gradient accumulators -> .grad), one block per job
__global__ void vecflush_multi_kernel(const VecJob* __restrict__ jobs) {
  const VecJob jb = jobs[blockIdx.x];
  for (int i = threadIdx.x; i < jb.n; i += blockDim.x) { jb.dst[i] += jb.src[i]; jb.src[i] = 0.f; }
}
}  // namespace

static_assert(sizeof(vcg_vecjob) == sizeof(VecJob), "vcg_vecjob layout mismatch");
extern "C" int vcg_vecflush_multi(const vcg_vecjob* jobs_dev, int32_t njobs, void* stream_) {
  if (njobs <= 0) return VCG_OK;
  vecflush_multi_kernel<<<njobs, 128, 0, static_cast<cudaStream_t>(stream_)>>>(reinterpret_cast<const VecJob*>(jobs_dev));
  VCG_CHECK_LAUNCH("vecflush_multi_kernel");
  return VCG_OK;
}

// ---- multi-tensor entry points.  jobs: DEVICE array of njobs vcg_wjob records whose tile0 / ntiles fields were
// filled by vcg_wjob_plan (host); total_tiles = sum of ntiles.
extern "C" int vcg_wjob_plan(vcg_wjob* jobs_host, int32_t njobs, int32_t* total_tiles) {
  int tile0 = 0;
  for (int j = 0; j < njobs; ++j) {
    vcg_wjob& jb = jobs_host[j];
    const int taps = jb.kh * jb.kw;
    VCG_REQUIRE(taps > 0 && 32 * (taps | 1) <= kWTileFloats, VCG_E_UNSUPPORTED, "wjob_plan: %d taps", taps);
    int co_t = kWTileFloats / (32 * (taps | 1));
    if (co_t > 32) co_t = 32;
    jb.co_t = co_t;
    jb.tiles_ci = (jb.ci + 31) / 32;
    jb.tile0 = tile0;
    jb.ntiles = ((jb.co + co_t - 1) / co_t) * jb.tiles_ci;
    // 16-byte vector path: 3x3 filters, whole 32x32 tiles, channel runs of 8 on the packed side
    jb.vec = (taps == 9 && co_t == 32 && jb.ci % 32 == 0 && jb.co % 32 == 0 && jb.wmap != VCG_WMAP_S2D &&
              jb.c_phys == jb.ci && jb.co_phys % 8 == 0 && jb.kwc_pad % 8 == 0 && jb.t_kwc_pad % 8 == 0) ? 1 : 0;
    tile0 += jb.ntiles;
  }
  *total_tiles = tile0;
  return VCG_OK;
}

static_assert(sizeof(vcg_wjob) == sizeof(WJob), "vcg_wjob / WJob layout mismatch");

extern "C" int vcg_wpack_multi(int32_t dtype, const vcg_wjob* jobs_dev, int32_t njobs, int32_t total_tiles, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (njobs <= 0 || total_tiles <= 0) return VCG_OK;
  const WJob* jobs = reinterpret_cast<const WJob*>(jobs_dev);
  const size_t smem = kWTileFloats * sizeof(float);
  if (dtype == VCG_F32) wpack_multi_kernel<float><<<total_tiles, 256, smem, stream>>>(jobs, njobs);
  else wpack_multi_kernel<__nv_bfloat16><<<total_tiles, 256, smem, stream>>>(jobs, njobs);
  VCG_CHECK_LAUNCH("wpack_multi_kernel");
  return VCG_OK;
}

extern "C" int vcg_wunpack_multi(const vcg_wjob* jobs_dev, int32_t njobs, int32_t total_tiles, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (njobs <= 0 || total_tiles <= 0) return VCG_OK;
  wunpack_multi_kernel<<<total_tiles, 256, kWTileFloats * sizeof(float), stream>>>(reinterpret_cast<const WJob*>(jobs_dev), njobs);
  VCG_CHECK_LAUNCH("wunpack_multi_kernel");
  return VCG_OK;
}
