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

// ---- the reference's dataset glue, exact: uint8 HWC image <-> normalised LAB CHW float32 tensor -----------------------------
// Basic._arr2lab (src/dataset.py:148-159) is skimage rgb2lab in float64 on arr/255, then L/100, (a,b+128)/255, then .float();
// Basic._lab2img / utils.tensor2img (src/dataset.py:94-104, src/utils.py:22-26) denormalise IN float32, run skimage lab2rgb
// in float64, multiply by 255 and truncate to uint8.  The truncation makes the result sensitive to the last bits (the
// reference's own example tiles come back as k or k-1 for a source value k, half and half), so these two kernels compute in
// float64 like skimage does; tests/test_color.py reproduces the reference's example PNG tiles bit for bit with them.
struct Mat3d { double m[9]; };

__device__ __forceinline__ double lab_f64(double t) { return t > 0.008856 ? cbrt(t) : 7.787 * t + 16.0 / 116.0; }
__device__ __forceinline__ double lab_finv64(double f) { return f > 0.2068966 ? f * f * f : (f - 16.0 / 116.0) / 7.787; }

__global__ void rgb2lab_u8_k(const uint8_t* __restrict__ rgb, float* __restrict__ lab, int64_t hw, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t n = i / hw, q = i - n * hw;
  const uint8_t* s = rgb + i * 3;
  double c[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double v = (double)s[k] / 255.0;
    c[k] = v > 0.04045 ? pow((v + 0.055) / 1.055, 2.4) : v / 12.92;
  }
  double X = (0.412453 * c[0] + 0.357580 * c[1] + 0.180423 * c[2]) / 0.95047;
  double Y = (0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2]);
  double Z = (0.019334 * c[0] + 0.119193 * c[1] + 0.950227 * c[2]) / 1.08883;
  double fx = lab_f64(X), fy = lab_f64(Y), fz = lab_f64(Z);
  float* d = lab + n * 3 * hw + q;
  d[0] = (float)((116.0 * fy - 16.0) / 100.0);
  d[hw] = (float)((500.0 * (fx - fy) + 128.0) / 255.0);
  d[2 * hw] = (float)((200.0 * (fy - fz) + 128.0) / 255.0);
}

__global__ void lab2rgb_u8_k(const float* __restrict__ lab, uint8_t* __restrict__ rgb, int64_t hw, int64_t total, Mat3d inv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t n = i / hw, q = i - n * hw;
  const float* s = lab + n * 3 * hw + q;
  // float32 denormalisation, as the reference's in-place numpy arithmetic on the float32 array
  // (two roundings each, like numpy: an FMA contraction of x * 255 - 128 changes the last bit and, through the truncation
  //  below, one uint8 value in five)
  const double L = (double)__fmul_rn(__ldg(s), 100.f);
  const double A = (double)__fsub_rn(__fmul_rn(__ldg(s + hw), 255.f), 128.f);
  const double B = (double)__fsub_rn(__fmul_rn(__ldg(s + 2 * hw), 255.f), 128.f);
  double fy = (L + 16.0) / 116.0, fx = A / 500.0 + fy, fz = fmax(fy - B / 200.0, 0.0);
  double X = lab_finv64(fx) * 0.95047, Y = lab_finv64(fy), Z = lab_finv64(fz) * 1.08883;
  uint8_t* d = rgb + i * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double v = inv.m[3 * k] * X + inv.m[3 * k + 1] * Y + inv.m[3 * k + 2] * Z;
    v = v > 0.0031308 ? 1.055 * pow(v, 1.0 / 2.4) - 0.055 : 12.92 * v;
    v = fmin(fmax(v, 0.0), 1.0);
    d[k] = (uint8_t)(v * 255.0);       // truncation, as .astype("uint8")
  }
}

int rgb2lab_u8(const uint8_t* rgb, float* lab, int n, int h, int w, cudaStream_t st) {
  SRCGAN_REQUIRE(rgb && lab && n > 0 && h > 0 && w > 0, "rgb2lab_u8: null pointer or empty image");
  int64_t hw = (int64_t)h * w, total = hw * n;
  rgb2lab_u8_k<<<ceil_div(total, 256), 256, 0, st>>>(rgb, lab, hw, total);
  count_launch();
  return check_launch("rgb2lab_u8");
}
int lab2rgb_u8(const float* lab, uint8_t* rgb, int n, int h, int w, cudaStream_t st) {
  SRCGAN_REQUIRE(rgb && lab && n > 0 && h > 0 && w > 0, "lab2rgb_u8: null pointer or empty image");
  // rgb_from_xyz = inverse of the sRGB matrix above, in float64 (skimage: scipy.linalg.inv(xyz_from_rgb))
  const double a[9] = {0.412453, 0.357580, 0.180423, 0.212671, 0.715160, 0.072169, 0.019334, 0.119193, 0.950227};
  const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
  Mat3d inv;
  inv.m[0] = (a[4] * a[8] - a[5] * a[7]) / det; inv.m[1] = (a[2] * a[7] - a[1] * a[8]) / det; inv.m[2] = (a[1] * a[5] - a[2] * a[4]) / det;
  inv.m[3] = (a[5] * a[6] - a[3] * a[8]) / det; inv.m[4] = (a[0] * a[8] - a[2] * a[6]) / det; inv.m[5] = (a[2] * a[3] - a[0] * a[5]) / det;
  inv.m[6] = (a[3] * a[7] - a[4] * a[6]) / det; inv.m[7] = (a[1] * a[6] - a[0] * a[7]) / det; inv.m[8] = (a[0] * a[4] - a[1] * a[3]) / det;
  int64_t hw = (int64_t)h * w, total = hw * n;
  lab2rgb_u8_k<<<ceil_div(total, 256), 256, 0, st>>>(lab, rgb, hw, total, inv);
  count_launch();
  return check_launch("lab2rgb_u8");
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
