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
