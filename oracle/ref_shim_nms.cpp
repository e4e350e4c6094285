// TEST INFRASTRUCTURE — not product code.
// C-ABI shim around the UNMODIFIED reference source tf_ops/3d_nms/tf_nms3d.cpp (pulled in textually from
// /root/reference; include path set by oracle/Makefile; TensorFlow replaced by oracle/tf_stub/).
#include <cstring>
#include "tf_nms3d.cpp"  // -I/root/reference/tf_ops/3d_nms

extern "C" {

// BEV polygon-clip area of two (8,3) corner boxes — reference `intersection` (tf_nms3d.cpp:122-175).
float ref_intersection2d(const float* box1, const float* box2) { return intersection(box1, box2); }
float ref_area2d(const float* box) { return area2d(box); }
float ref_area3d(const float* box) { return area3d(box); }

// reference `IOUGreaterThanThreshold` (tf_nms3d.cpp:178-192) on boxes laid out (b, nboxes, 8, 3)
int ref_iou_greater(const float* boxes, int b, int i, int j, int nboxes, float thr) {
  TTypes<float, 4>::ConstTensor t{boxes, 0};
  return IOUGreaterThanThreshold(t, b, i, j, nboxes, thr) ? 1 : 0;
}

// Full op: NonMaxSuppression3DOp::Compute. Returns number of selected rows (>=0), or -1 if the op rejected
// its inputs (OP_REQUIRES). out_idx must hold b*k*2 ints.  `thr_is_scalar`=0 feeds a rank-1 threshold to
// exercise the reference's scalar check.
int ref_nms3d(int b, int k, const float* bbox, const float* scores, const float* objectiveness, float thr,
              int* out_idx) {
  OpKernelConstruction c;
  NonMaxSuppression3DOp<CPUDevice> op(&c);
  OpKernelContext ctx;
  ctx.inputs.push_back(Tensor(bbox, TensorShape{b, k, 8, 3}));
  ctx.inputs.push_back(Tensor(scores, TensorShape{b, k}));
  ctx.inputs.push_back(Tensor(objectiveness, TensorShape{b, k, 2}));
  ctx.inputs.push_back(Tensor(&thr, TensorShape{}));
  op.Compute(&ctx);
  if (!ctx.status.ok()) return -1;
  int rows = (int)ctx.outputs[0].shape().dim_size(0);
  memcpy(out_idx, ctx.outputs[0].raw(), sizeof(int) * (size_t)rows * 2);
  return rows;
}

}  // extern "C"
