// Device-side image path: the reference preprocesses every input image on the host — PIL bicubic
// antialiased resize to a stride multiple (data/transforms.py:15-115 -> Pillow ImagingResample), ToTensor + Normalize,
// then `patchify` (data/data_utils.py:43-50). Here the uint8 image is uploaded once and everything else runs on the GPU,
// BIT-EXACTLY: integer byte work like Pillow's 8-bit resampler must not differ by a single level.
//
// Pillow's algorithm (src/libImaging/Resample.c, 8 bits per channel): separable, horizontal pass then vertical pass, each
//   out = clip8( (2^21 + sum_k in[xmin + k] * kk[k]) >> 22 )
// with per-output-pixel windows [xmin, xmin + n) and filter taps kk = round(w * 2^22) of the bicubic kernel (a = -0.5)
// stretched by max(scale, 1) for antialiasing, normalised to sum 1 in double precision (taps are computed on the host by
// bagel_b200.transforms.pil_bicubic_coeffs — a few hundred integers per axis). The intermediate image is uint8 again.
// Both passes are pure streaming byte kernels (HBM-bound): one thread per output byte, taps through the read-only path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "host_util.h"

namespace bagel {

// generic axis pass over an interleaved [A_in, ...] image: output element (o, r, c) = along-index o, other-index r, channel c
__global__ void __launch_bounds__(256)
resample_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ kk,
                   const int* __restrict__ bounds, int ksize, int n_out, int n_other, int C, long long s_along,
                   long long s_other, long long d_along, long long d_other) {
  const long long total = (long long)n_out * n_other * C;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // fastest index = channel, then the index that is contiguous in memory for coalescing: the caller orders (o, r) so that
  // `inner` runs along the image row
  const int c = (int)(idx % C);
  const long long t = idx / C;
  int o, r;
  if (d_along < d_other) { o = (int)(t % n_out); r = (int)(t / n_out); }   // horizontal pass: o = x is the inner index
  else { r = (int)(t % n_other); o = (int)(t / n_other); }                 // vertical pass: r = x is the inner index
  const int xmin = bounds[2 * o], n = bounds[2 * o + 1];
  const int* k = kk + (long long)o * ksize;
  const uint8_t* p = src + (long long)xmin * s_along + (long long)r * s_other + c;
  int acc = 1 << 21;
  for (int i = 0; i < n; ++i) acc += (int)p[(long long)i * s_along] * __ldg(k + i);
  acc >>= 22;                                           // arithmetic shift, as Pillow's clip8 lookup index
  dst[(long long)o * d_along + (long long)r * d_other + c] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
}

// ToTensor + Normalize (+ patchify): value = ((u8 / 255) - mean[c]) / std[c] with the three fp32 roundings torch performs.
//   patch == 0: out is planar CHW fp32 [3, H, W]
//   patch  > 0: out[(py * (W/patch) + px) * ld + (r * patch + q) * 3 + c]   ("chpwq->hwpqc", data/data_utils.py:43-50)
__global__ void __launch_bounds__(256)
image_normalize_u8_kernel(const uint8_t* __restrict__ src, int H, int W, float m0, float m1, float m2, float s0, float s1,
                          float s2, float* __restrict__ out, long long ld, int patch) {
  const long long total = (long long)H * W * 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % 3);
  const long long px_i = idx / 3;
  const int x = (int)(px_i % W), y = (int)(px_i / W);
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
  const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[idx], 255.0f), mean), sd);
  if (patch == 0) {
    out[((long long)c * H + y) * W + x] = v;
  } else {
    const int py = y / patch, r = y - py * patch, pxx = x / patch, q = x - pxx * patch;
    out[((long long)py * (W / patch) + pxx) * ld + (r * patch + q) * 3 + c] = v;
  }
}

}  // namespace bagel

using namespace bagel;

extern "C" int bagel_image_resize_bicubic_u8(const uint8_t* src, int Hi, int Wi, uint8_t* dst, int Ho, int Wo, uint8_t* tmp,
                                             const int* kk_h, const int* bounds_h, int ksize_h, const int* kk_v,
                                             const int* bounds_v, int ksize_v, void* stream) {
  if (Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0) return set_error(BAGEL_ERR_SHAPE, "bagel_image_resize_bicubic_u8: bad sizes");
  if (!src || !dst) return set_error(BAGEL_ERR_ARG, "bagel_image_resize_bicubic_u8: null image pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int C = 3;
  const bool need_h = Wo != Wi, need_v = Ho != Hi;
  if (need_h && (!kk_h || !bounds_h)) return set_error(BAGEL_ERR_ARG, "bagel_image_resize_bicubic_u8: horizontal taps missing");
  if (need_v && (!kk_v || !bounds_v)) return set_error(BAGEL_ERR_ARG, "bagel_image_resize_bicubic_u8: vertical taps missing");
  if (need_h && need_v && !tmp) return set_error(BAGEL_ERR_ARG, "bagel_image_resize_bicubic_u8: tmp [Hi, Wo, 3] required");
  if (!need_h && !need_v) {
    BAGEL_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)Hi * Wi * C, cudaMemcpyDeviceToDevice, s));
    return 0;
  }
  const uint8_t* cur = src;
  if (need_h) {   // [Hi, Wi, 3] -> [Hi, Wo, 3]
    uint8_t* o = need_v ? tmp : dst;
    const long long total = (long long)Hi * Wo * C;
    resample_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(cur, o, kk_h, bounds_h, ksize_h, Wo, Hi, C, C,
                                                                       (long long)Wi * C, C, (long long)Wo * C);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cur = o;
  }
  if (need_v) {   // [Hi, Wo, 3] -> [Ho, Wo, 3]
    const long long total = (long long)Ho * Wo * C;
    resample_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(cur, dst, kk_v, bounds_v, ksize_v, Ho, Wo, C,
                                                                       (long long)Wo * C, C, (long long)Wo * C, C);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int bagel_image_normalize_u8(const uint8_t* src, int H, int W, float mean0, float mean1, float mean2, float std0,
                                        float std1, float std2, float* out, long long ld, int patch, void* stream) {
  if (H <= 0 || W <= 0) return set_error(BAGEL_ERR_SHAPE, "bagel_image_normalize_u8: bad sizes");
  if (patch < 0 || (patch > 0 && ((H % patch) || (W % patch) || ld < (long long)patch * patch * 3)))
    return set_error(BAGEL_ERR_SHAPE, "bagel_image_normalize_u8: H, W must be multiples of patch and ld >= patch*patch*3");
  const long long total = (long long)H * W * 3;
  image_normalize_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, H, W, mean0, mean1, mean2, std0, std1, std2, out, ld, patch);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  BAGEL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
