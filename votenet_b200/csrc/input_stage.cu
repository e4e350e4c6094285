// Input stage on the device — SURVEY.md §8(f) rank 4: point subsampling + axis flip to the upright camera frame +
// training augmentation, one fused pass (HBM-bound: 12 B gathered + 12 (+4) B written per kept point).
//
// Reference (host numpy, one scene at a time): /root/reference/dataset.py:185-190 — keep POINT_NUM randomly chosen points
// and map upright-depth coordinates to the upright camera frame (sunutils.py:70-77: cam X,Y,Z = depth X,-Z,Y);
// dataset.py:219-231,302-308 — with probability 1/2 each negate x / negate z, rotate about y by a uniform angle in
// +-5 degrees (sunutils.py:133-139), scale by a uniform factor in 1 +- 0.1.  The random draws stay on the host (they are
// the reference's np.random stream): the kernel takes the chosen indices and the per-cloud draws, so that identical
// draws give identical clouds.  Arithmetic in double like numpy, rounded to float once at the end.
#include "common.cuh"

namespace vnb {

__global__ void prepare_input_kernel(int n_raw, int n, const float* __restrict__ raw, const int* __restrict__ choice,
                                     const unsigned char* __restrict__ flip_x, const unsigned char* __restrict__ flip_z,
                                     const double* __restrict__ roty_angle, const double* __restrict__ scale,
                                     float* __restrict__ xyz, float* __restrict__ height, double floor_y) {
  const int cloud = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = choice != nullptr ? choice[(size_t)cloud * n + i] : i;     // dataset.py:185-186
  const float* p = raw + ((size_t)cloud * n_raw + k) * 3;
  double x = p[0], y = -(double)p[2], z = p[1];                            // sunutils.py:75-76
  if (flip_x != nullptr && flip_x[cloud]) x = -x;                          // dataset.py:303-304
  if (flip_z != nullptr && flip_z[cloud]) z = -z;                          // :305-306
  if (roty_angle != nullptr) {                                             // :307, sunutils.py:133-139
    const double a = roty_angle[cloud], c = cos(a), s = sin(a);
    const double rx = c * x + s * z, rz = -s * x + c * z;
    x = rx; z = rz;
  }
  if (scale != nullptr) { const double f = scale[cloud]; x *= f; y *= f; z *= f; }   // :308
  float* o = xyz + ((size_t)cloud * n + i) * 3;
  o[0] = (float)x; o[1] = (float)y; o[2] = (float)z;
  if (height != nullptr) height[(size_t)cloud * n + i] = (float)(floor_y - y);   // BASELINE's (xyz + height) input feature
}

}  // namespace vnb

using namespace vnb;

extern "C" int vnb_prepare_input(int b, int n_raw, int n, const float* raw_upright_depth, const int* choice,
                                 const unsigned char* flip_x, const unsigned char* flip_z, const double* roty_angle,
                                 const double* scale, float* xyz, float* height, double floor_y, void* stream) {
  VNB_REQUIRE(b >= 0 && n >= 0 && n_raw >= 0, "prepare_input: bad shape");
  VNB_REQUIRE(choice != nullptr || n <= n_raw, "prepare_input: without a choice array n must not exceed n_raw");
  if (b == 0 || n == 0) return VNB_OK;
  prepare_input_kernel<<<dim3((n + 255) / 256, b), 256, 0, as_stream(stream)>>>(n_raw, n, raw_upright_depth, choice, flip_x, flip_z,
                                                                              roty_angle, scale, xyz, height, floor_y);
  return check_launch("prepare_input");
}
