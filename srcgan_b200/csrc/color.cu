// RGB <-> CIE-LAB (D65, 2 degree observer) on NCHW fp32 planes, fused elementwise (HBM / SFU bound).
// Moves onto the device what the reference does on the CPU with scikit-image:
// rgb2lab at src/dataset.py:148-159 (normalised L/100, (a,b+128)/255) and lab2rgb at
// src/utils.py:22-26, src/testCasConstLAB.py:37-41.
#include "common.cuh"

namespace srcgan {

__device__ __forceinline__ float srgb_to_linear(float v) {
  return v > 0.04045f ? powf((v + 0.055f) / 1.055f, 2.4f) : v / 12.92f;
}
__device__ __forceinline__ float linear_to_srgb(float v) {
  return v > 0.0031308f ? 1.055f * powf(fmaxf(v, 0.f), 1.f / 2.4f) - 0.055f : 12.92f * v;
}
__device__ __forceinline__ float lab_f(float t) { return t > 0.008856f ? cbrtf(t) : 7.787f * t + 16.f / 116.f; }
__device__ __forceinline__ float lab_finv(float f) {
  return f > 0.2068966f ? f * f * f : (f - 16.f / 116.f) / 7.787f;
}

__global__ void rgb2lab_k(const float* __restrict__ rgb, float* __restrict__ lab, int64_t hw, int64_t total,
                          int normalised) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t n = i / hw, q = i - n * hw;
  const float* s = rgb + n * 3 * hw + q;
  float r = srgb_to_linear(__ldg(s)), g = srgb_to_linear(__ldg(s + hw)), b = srgb_to_linear(__ldg(s + 2 * hw));
  float X = (0.412453f * r + 0.357580f * g + 0.180423f * b) / 0.95047f;
  float Y = (0.212671f * r + 0.715160f * g + 0.072169f * b);
  float Z = (0.019334f * r + 0.119193f * g + 0.950227f * b) / 1.08883f;
  float fx = lab_f(X), fy = lab_f(Y), fz = lab_f(Z);
  float L = 116.f * fy - 16.f, A = 500.f * (fx - fy), B = 200.f * (fy - fz);
  if (normalised) { L = L / 100.f; A = (A + 128.f) / 255.f; B = (B + 128.f) / 255.f; }
  float* d = lab + n * 3 * hw + q;
  d[0] = L; d[hw] = A; d[2 * hw] = B;
}

__global__ void lab2rgb_k(const float* __restrict__ lab, float* __restrict__ rgb, int64_t hw, int64_t total,
                          int normalised) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t n = i / hw, q = i - n * hw;
  const float* s = lab + n * 3 * hw + q;
  float L = __ldg(s), A = __ldg(s + hw), B = __ldg(s + 2 * hw);
  if (normalised) { L = L * 100.f; A = A * 255.f - 128.f; B = B * 255.f - 128.f; }
  float fy = (L + 16.f) / 116.f, fx = A / 500.f + fy, fz = fmaxf(fy - B / 200.f, 0.f);
  float X = lab_finv(fx) * 0.95047f, Y = lab_finv(fy), Z = lab_finv(fz) * 1.08883f;
  // inverse of the 3x3 above (float64-inverted, rounded to fp32)
  float r = 3.2404813432f * X - 1.5371515163f * Y - 0.4985363262f * Z;
  float g = -0.9692549500f * X + 1.8759900015f * Y + 0.0415559266f * Z;
  float b = 0.0556466391f * X - 0.2040413384f * Y + 1.0573110696f * Z;
  float* d = rgb + n * 3 * hw + q;
  d[0] = fminf(fmaxf(linear_to_srgb(r), 0.f), 1.f);
  d[hw] = fminf(fmaxf(linear_to_srgb(g), 0.f), 1.f);
  d[2 * hw] = fminf(fmaxf(linear_to_srgb(b), 0.f), 1.f);
}

int rgb2lab(const float* rgb, float* lab, int n, int h, int w, int normalised, cudaStream_t st) {
  SRCGAN_REQUIRE(rgb && lab && n > 0, "rgb2lab: null pointer");
  int64_t hw = (int64_t)h * w, total = hw * n;
  rgb2lab_k<<<ceil_div(total, 256), 256, 0, st>>>(rgb, lab, hw, total, normalised);
  count_launch();
  return check_launch("rgb2lab");
}
int lab2rgb(const float* lab, float* rgb, int n, int h, int w, int normalised, cudaStream_t st) {
  SRCGAN_REQUIRE(rgb && lab && n > 0, "lab2rgb: null pointer");
  int64_t hw = (int64_t)h * w, total = hw * n;
  lab2rgb_k<<<ceil_div(total, 256), 256, 0, st>>>(lab, rgb, hw, total, normalised);
  count_launch();
  return check_launch("lab2rgb");
}

}  // namespace srcgan
