// TEST INFRASTRUCTURE — not product code.
// C-ABI shim around the reference's OWN GPU launchers, compiled UNMODIFIED from
// /root/reference/tf_ops/sampling/tf_sampling_g.cu and /root/reference/tf_ops/grouping/tf_grouping_g.cu
// (neither file includes TensorFlow).  Built here by oracle/Makefile into oracle/_ref/libvotenet_ref_gpu.so
// with the reference's flags (nvcc -O2, default -fmad=true) for sm_100a; the .so travels to the GPU box,
// where the `-m gpu` parity tests compare the product kernels against it bit for bit.
#include <cuda_runtime.h>

// prototypes exactly as the reference declares them (tf_sampling.cpp:94,125 ; tf_grouping.cpp:66,142)
void farthestpointsamplingLauncher(int b, int n, int m, const float* inp, float* temp, int* out);
void gatherpointLauncher(int b, int n, int m, const float* inp, const int* idx, float* out);
void queryBallPointLauncher(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2,
                            int* idx, int* pts_cnt);
void groupPointLauncher(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out);
void scatteraddpointLauncher(int b, int n, int m, const float* out_g, const int* idx, float* inp_g);          // tf_sampling.cpp:150
void groupPointGradLauncher(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx,
                            float* grad_points);                                                                // tf_grouping.cpp:173

extern "C" {
// all pointers are device pointers; launches go to the legacy default stream like the reference's.
int ref_gpu_fps(int b, int n, int m, const float* inp, float* temp /* (32,n) */, int* out) {
  farthestpointsamplingLauncher(b, n, m, inp, temp, out);
  return (int)cudaDeviceSynchronize();
}
int ref_gpu_gather_point(int b, int n, int m, const float* inp, const int* idx, float* out) {
  gatherpointLauncher(b, n, m, inp, idx, out);
  return (int)cudaDeviceSynchronize();
}
int ref_gpu_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2,
                             int* idx, int* pts_cnt) {
  queryBallPointLauncher(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt);
  return (int)cudaDeviceSynchronize();
}
int ref_gpu_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out) {
  groupPointLauncher(b, n, c, m, nsample, points, idx, out);
  return (int)cudaDeviceSynchronize();
}
int ref_gpu_gather_point_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g) {
  cudaMemset(inp_g, 0, sizeof(float) * (size_t)b * n * 3);  // tf_sampling.cpp:174
  scatteraddpointLauncher(b, n, m, out_g, idx, inp_g);
  return (int)cudaDeviceSynchronize();
}
int ref_gpu_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx,
                             float* grad_points) {
  cudaMemset(grad_points, 0, sizeof(float) * (size_t)b * n * c);  // tf_grouping.cpp:203
  groupPointGradLauncher(b, n, c, m, nsample, grad_out, idx, grad_points);
  return (int)cudaDeviceSynchronize();
}
}
