// TEST INFRASTRUCTURE — not product code.
//
// A ~200-line stand-in for the handful of TensorFlow C++ symbols that the reference's CPU op sources
// (tf_ops/3d_interpolation/tf_interpolate.cpp, tf_ops/3d_nms/tf_nms3d.cpp) touch, so that those files can
// be compiled UNMODIFIED, from where they lie under /root/reference, into oracle/_ref/ without TensorFlow.
// Nothing here implements any arithmetic of the hot path: it only provides Tensor/OpKernel plumbing so the
// reference's own Compute() bodies run on caller-supplied host buffers.
#pragma once
// (TensorFlow headers transitively provide these standard headers; the reference relies on that.)
#include <algorithm>
#include <cmath>
#include <deque>
#include <iostream>
#include <cstdint>
#include <cstddef>
#include <functional>
#include <initializer_list>
#include <string>
#include <vector>

#ifndef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif
#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif

namespace Eigen {
struct ThreadPoolDevice {};
struct GpuDevice {};
}  // namespace Eigen

namespace tensorflow {

typedef long long int64;

class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& msg) : ok_(false), msg_(msg) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }

 private:
  bool ok_;
  std::string msg_;
};

namespace errors {
template <typename... Args>
inline Status InvalidArgument(const char* msg, Args...) { return Status(std::string(msg)); }
inline Status InvalidArgument(const std::string& msg) { return Status(msg); }
}  // namespace errors

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> d) : d_(d) {}
  explicit TensorShape(const std::vector<int64>& d) : d_(d) {}
  int dims() const { return (int)d_.size(); }
  int64 dim_size(int i) const { return d_[i]; }
  int64 num_elements() const { int64 n = 1; for (auto v : d_) n *= v; return n; }

 private:
  std::vector<int64> d_;
};

struct TensorShapeUtils {
  static bool IsScalar(const TensorShape& s) { return s.dims() == 0; }
};

template <typename T>
struct FlatView {
  T* p;
  size_t n;
  T& operator()(size_t i) const { return p[i]; }
  T* data() const { return p; }
  size_t size() const { return n; }
};
template <typename T>
struct ScalarView {
  T* p;
  T& operator()() const { return *p; }
};

template <typename T, int N = 1>
struct TTypes {
  typedef FlatView<T> Tensor;
  typedef FlatView<const T> ConstTensor;
  typedef FlatView<T> Flat;
  typedef FlatView<const T> ConstFlat;
};

// A dense host tensor: either wraps a caller buffer (inputs) or owns its bytes (outputs).
class Tensor {
 public:
  Tensor() : ext_(nullptr) {}
  Tensor(const void* ext, const TensorShape& s) : shape_(s), ext_(const_cast<void*>(ext)) {}
  Tensor(size_t elem_bytes, const TensorShape& s) : shape_(s), own_(elem_bytes * (size_t)s.num_elements() + 16), ext_(nullptr) {}
  int dims() const { return shape_.dims(); }
  const TensorShape& shape() const { return shape_; }
  int64 NumElements() const { return shape_.num_elements(); }
  void* raw() const { return ext_ ? ext_ : (void*)own_.data(); }

  template <typename T> FlatView<T> flat() { return FlatView<T>{(T*)raw(), (size_t)NumElements()}; }
  template <typename T> FlatView<const T> flat() const { return FlatView<const T>{(const T*)raw(), (size_t)NumElements()}; }
  template <typename T, int N> FlatView<T> tensor() { return FlatView<T>{(T*)raw(), (size_t)NumElements()}; }
  template <typename T, int N> FlatView<const T> tensor() const { return FlatView<const T>{(const T*)raw(), (size_t)NumElements()}; }
  template <typename T> ScalarView<const T> scalar() const { return ScalarView<const T>{(const T*)raw()}; }

 private:
  TensorShape shape_;
  mutable std::vector<char> own_;
  void* ext_;
};

class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char*, T*) const { return Status::OK(); }
  void CtxFailure(const Status& s) { status = s; }
  void CtxFailure(const char*, int, const Status& s) { status = s; }
  Status status;
};

class OpKernelContext {
 public:
  const Tensor& input(int i) const { return inputs[i]; }
  // The reference only ever allocates float32 / int32 outputs; both are 4 bytes wide.
  Status allocate_output(int i, const TensorShape& s, Tensor** out) {
    if ((int)outputs.size() <= i) outputs.resize(i + 1);
    outputs[i] = Tensor(4, s);
    *out = &outputs[i];
    return Status::OK();
  }
  template <typename DT> Status allocate_temp(DT, const TensorShape& s, Tensor* out) { *out = Tensor(4, s); return Status::OK(); }
  void CtxFailure(const Status& s) { status = s; }
  void CtxFailure(const char*, int, const Status& s) { status = s; }
  void SetStatus(const Status& s) { status = s; }
  std::vector<Tensor> inputs;
  std::vector<Tensor> outputs;
  Status status;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

#define OP_REQUIRES(CTX, EXP, STATUS)          \
  do {                                          \
    if (!(EXP)) {                               \
      (CTX)->CtxFailure((STATUS));              \
      return;                                   \
    }                                           \
  } while (0)

#define OP_REQUIRES_OK(CTX, ...)                         \
  do {                                                   \
    ::tensorflow::Status _s(__VA_ARGS__);                \
    if (!_s.ok()) {                                      \
      (CTX)->CtxFailure(_s);                             \
      return;                                            \
    }                                                    \
  } while (0)

namespace shape_inference {
struct DimensionHandle {};
struct ShapeHandle {};
struct DimensionOrConstant {
  DimensionOrConstant(DimensionHandle) {}
  DimensionOrConstant(long long) {}
};
class InferenceContext {
 public:
  static constexpr int64 kUnknownDim = -1;
  ShapeHandle input(int) { return ShapeHandle(); }
  void set_output(int, ShapeHandle) {}
  Status WithRank(ShapeHandle, int, ShapeHandle*) { return Status::OK(); }
  DimensionHandle Dim(ShapeHandle, int) { return DimensionHandle(); }
  ShapeHandle MakeShape(const std::vector<DimensionOrConstant>&) { return ShapeHandle(); }
  template <typename T> Status GetAttr(const char*, T*) { return Status::OK(); }
};
}  // namespace shape_inference

class OpDefBuilder {
 public:
  explicit OpDefBuilder(const char*) {}
  OpDefBuilder& Input(const char*) { return *this; }
  OpDefBuilder& Output(const char*) { return *this; }
  OpDefBuilder& Attr(const char*) { return *this; }
  OpDefBuilder& SetShapeFn(std::function<Status(shape_inference::InferenceContext*)>) { return *this; }
};

static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

class Name {
 public:
  explicit Name(const char*) {}
  Name& Device(const char*) { return *this; }
  Name& HostMemory(const char*) { return *this; }
};

#define TF_STUB_CAT2(a, b) a##b
#define TF_STUB_CAT(a, b) TF_STUB_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::OpDefBuilder TF_STUB_CAT(_tf_stub_op_, __COUNTER__) = ::tensorflow::OpDefBuilder(name)
#define REGISTER_KERNEL_BUILDER(kb, ...) static int TF_STUB_CAT(_tf_stub_kb_, __COUNTER__) = ((void)(kb), 0)

}  // namespace tensorflow
