// TEST INFRASTRUCTURE — not product code.
// C-ABI shim around the UNMODIFIED reference source tf_ops/3d_interpolation/tf_interpolate.cpp, which is
// pulled in textually from /root/reference (include path set by oracle/Makefile) and compiled against the
// TensorFlow stand-in in oracle/tf_stub/.  Runs the reference's own ThreeNNOp / ThreeInterpolateOp Compute().
#include "tf_interpolate.cpp"  // -I/root/reference/tf_ops/3d_interpolation

extern "C" {

// returns 0 on success, 1 if the reference op rejected the shapes (OP_REQUIRES failure)
int ref_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx) {
  OpKernelConstruction c;
  ThreeNNOp op(&c);
  OpKernelContext ctx;
  ctx.inputs.push_back(Tensor(xyz1, TensorShape{b, n, 3}));
  ctx.inputs.push_back(Tensor(xyz2, TensorShape{b, m, 3}));
  op.Compute(&ctx);
  if (!ctx.status.ok()) return 1;
  memcpy(dist, ctx.outputs[0].raw(), sizeof(float) * (size_t)b * n * 3);
  memcpy(idx, ctx.outputs[1].raw(), sizeof(int) * (size_t)b * n * 3);
  return 0;
}

int ref_three_interpolate(int b, int m, int c, int n, const float* points, const int* idx, const float* weight,
                          float* out) {
  OpKernelConstruction cc;
  ThreeInterpolateOp op(&cc);
  OpKernelContext ctx;
  ctx.inputs.push_back(Tensor(points, TensorShape{b, m, c}));
  ctx.inputs.push_back(Tensor(idx, TensorShape{b, n, 3}));
  ctx.inputs.push_back(Tensor(weight, TensorShape{b, n, 3}));
  op.Compute(&ctx);
  if (!ctx.status.ok()) return 1;
  memcpy(out, ctx.outputs[0].raw(), sizeof(float) * (size_t)b * n * c);
  return 0;
}

int ref_three_interpolate_grad(int b, int n, int c, int m, const float* grad_out, const int* idx, const float* weight,
                               float* grad_points) {
  OpKernelConstruction cc;
  ThreeInterpolateGradOp op(&cc);
  OpKernelContext ctx;
  std::vector<float> points((size_t)b * m * c, 0.f);  // input 0 is only inspected for its shape (tf_interpolate.cpp:229-233)
  ctx.inputs.push_back(Tensor(points.data(), TensorShape{b, m, c}));
  ctx.inputs.push_back(Tensor(idx, TensorShape{b, n, 3}));
  ctx.inputs.push_back(Tensor(weight, TensorShape{b, n, 3}));
  ctx.inputs.push_back(Tensor(grad_out, TensorShape{b, n, c}));
  op.Compute(&ctx);
  if (!ctx.status.ok()) return 1;
  memcpy(grad_points, ctx.outputs[0].raw(), sizeof(float) * (size_t)b * m * c);
  return 0;
}

}  // extern "C"
