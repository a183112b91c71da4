// tf_mock.h -- TEST INFRASTRUCTURE: a stand-in for the handful of TensorFlow classes the op-kernel shim under
// integration/tf/ touches (Status, Tensor, TensorShape, OpKernel(Context/Construction), LookupInterface, mutex,
// REGISTER_OP / REGISTER_KERNEL_BUILDER), with the member names and argument orders of TensorFlow 2.15's public
// headers, so that the shim can be type-checked and EXECUTED without TensorFlow (which this image does not have).
// Tensors live in host memory: the shim is linked with the emulated libdetable (tests/emu/), whose "device" is the host.
// Not a TensorFlow replacement; nothing outside tests/ includes it.
#ifndef TESTS_TF_MOCK_H_
#define TESTS_TF_MOCK_H_

#include <cstdint>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <sstream>
#include <string>
#include <vector>

namespace Eigen {
struct half { uint16_t x; };
struct GpuDevice {
  void* stream() const { return nullptr; }
};
}  // namespace Eigen

namespace tensorflow {

using int8 = int8_t;
using int32 = int32_t;
using int64 = int64_t;   // TF: long long on Linux; the shim only reinterpret_casts to int64_t*
using string = std::string;
using tstring = std::string;
struct bfloat16 { uint16_t x; };

// ---- Status / errors ------------------------------------------------------------------------------------------
namespace error { enum Code { OK = 0, INVALID_ARGUMENT = 3, NOT_FOUND = 5, RESOURCE_EXHAUSTED = 8, UNIMPLEMENTED = 12, INTERNAL = 13 }; }
class Status {
 public:
  Status() {}
  Status(error::Code c, const std::string& m) : code_(c), msg_(m) {}
  bool ok() const { return code_ == error::OK; }
  error::Code code() const { return code_; }
  const std::string& message() const { return msg_; }
  std::string ToString() const { return ok() ? "OK" : msg_; }
 private:
  error::Code code_ = error::OK;
  std::string msg_;
};
inline Status OkStatus() { return Status(); }
namespace errors {
template <class... A> std::string Cat(const A&... a) { std::ostringstream o; int d[] = {0, ((o << a), 0)...}; (void)d; return o.str(); }
template <class... A> Status InvalidArgument(const A&... a) { return Status(error::INVALID_ARGUMENT, Cat(a...)); }
template <class... A> Status Internal(const A&... a) { return Status(error::INTERNAL, Cat(a...)); }
template <class... A> Status NotFound(const A&... a) { return Status(error::NOT_FOUND, Cat(a...)); }
template <class... A> Status ResourceExhausted(const A&... a) { return Status(error::RESOURCE_EXHAUSTED, Cat(a...)); }
template <class... A> Status Unimplemented(const A&... a) { return Status(error::UNIMPLEMENTED, Cat(a...)); }
}  // namespace errors
#define TF_RETURN_IF_ERROR(expr) do { ::tensorflow::Status _s = (expr); if (!_s.ok()) return _s; } while (0)
#define OP_REQUIRES(CTX, EXP, STATUS) do { if (!(EXP)) { (CTX)->CtxFailure((STATUS)); return; } } while (0)
#define OP_REQUIRES_OK(CTX, ...) do { ::tensorflow::Status _s(__VA_ARGS__); if (!_s.ok()) { (CTX)->CtxFailure(_s); return; } } while (0)

// ---- dtypes / shapes / tensors ---------------------------------------------------------------------------------
enum DataType { DT_INVALID = 0, DT_FLOAT = 1, DT_DOUBLE = 2, DT_INT32 = 3, DT_INT8 = 6, DT_INT64 = 9, DT_BOOL = 10, DT_BFLOAT16 = 14, DT_HALF = 19, DT_RESOURCE = 20 };
template <class T> struct DataTypeToEnum;
#define TF_MOCK_DT(T, E) template <> struct DataTypeToEnum<T> { static DataType v() { return E; } static constexpr DataType value = E; }
TF_MOCK_DT(float, DT_FLOAT); TF_MOCK_DT(double, DT_DOUBLE); TF_MOCK_DT(int32, DT_INT32); TF_MOCK_DT(int8, DT_INT8);
TF_MOCK_DT(int64, DT_INT64); TF_MOCK_DT(bool, DT_BOOL); TF_MOCK_DT(bfloat16, DT_BFLOAT16); TF_MOCK_DT(Eigen::half, DT_HALF);
inline size_t DataTypeSize(DataType d) {
  switch (d) { case DT_FLOAT: case DT_INT32: return 4; case DT_DOUBLE: case DT_INT64: return 8; case DT_HALF: case DT_BFLOAT16: return 2; default: return 1; }
}

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> d) : dims_(d) {}
  int dims() const { return (int)dims_.size(); }
  int64 dim_size(int i) const { return dims_[i]; }
  int64 num_elements() const { int64 n = 1; for (int64 d : dims_) n *= d; return n; }
  void AddDim(int64 d) { dims_.push_back(d); }
  void AppendShape(const TensorShape& s) { for (int64 d : s.dims_) dims_.push_back(d); }
  void RemoveLastDims(int n) { dims_.resize(dims_.size() - n); }
  std::string DebugString() const { std::ostringstream o; o << "["; for (size_t i = 0; i < dims_.size(); ++i) o << (i ? "," : "") << dims_[i]; o << "]"; return o.str(); }
  bool operator==(const TensorShape& o) const { return dims_ == o.dims_; }
 private:
  std::vector<int64> dims_;
};
struct TensorShapeUtils {
  static bool IsScalar(const TensorShape& s) { return s.dims() == 0; }
  static bool IsVector(const TensorShape& s) { return s.dims() == 1; }
};

template <class T> struct FlatView {   // what Tensor::flat<T>() / matrix<T>() hand out: .data() and .size()
  T* p; int64 n;
  T* data() const { return p; }
  int64 size() const { return n; }
  T& operator()(int64 i) const { return p[i]; }
};
template <class T> struct ScalarView { T* p; T& operator()() const { return *p; } };

class Tensor {
 public:
  Tensor() : dtype_(DT_INVALID) {}
  Tensor(DataType dt, const TensorShape& s) : dtype_(dt), shape_(s), buf_(std::make_shared<std::vector<uint64_t>>((s.num_elements() * DataTypeSize(dt) + 7) / 8 + 1)) {}
  DataType dtype() const { return dtype_; }
  const TensorShape& shape() const { return shape_; }
  int64 NumElements() const { return shape_.num_elements(); }
  int64 dim_size(int i) const { return shape_.dim_size(i); }
  int dims() const { return shape_.dims(); }
  void* data() const { return buf_ ? (void*)buf_->data() : nullptr; }
  template <class T> FlatView<T> flat() const { return FlatView<T>{(T*)data(), NumElements()}; }
  template <class T> FlatView<T> matrix() const { return flat<T>(); }
  template <class T> ScalarView<T> scalar() const { return ScalarView<T>{(T*)data()}; }
 private:
  DataType dtype_;
  TensorShape shape_;
  std::shared_ptr<std::vector<uint64_t>> buf_;
};

struct AllocatorAttributes {
  void set_on_host(bool) {}
  void set_gpu_compatible(bool) {}
};

// ---- mutex -----------------------------------------------------------------------------------------------------
class mutex : public std::shared_timed_mutex {};
using mutex_lock = std::unique_lock<std::shared_timed_mutex>;
using tf_shared_lock = std::shared_lock<std::shared_timed_mutex>;

// ---- attrs / node def ----------------------------------------------------------------------------------------------
struct AttrValue { int64 i = 0; float f = 0; std::string s; TensorShape shape; };
struct NodeDef { std::map<std::string, AttrValue> attr; };
inline Status MissingAttr(const std::string& n) { return errors::NotFound("No attr named '", n, "' in NodeDef"); }
inline Status GetNodeAttr(const NodeDef& d, const std::string& n, int64* v) { auto it = d.attr.find(n); if (it == d.attr.end()) return MissingAttr(n); *v = it->second.i; return OkStatus(); }
inline Status GetNodeAttr(const NodeDef& d, const std::string& n, int* v) { auto it = d.attr.find(n); if (it == d.attr.end()) return MissingAttr(n); *v = (int)it->second.i; return OkStatus(); }
inline Status GetNodeAttr(const NodeDef& d, const std::string& n, float* v) { auto it = d.attr.find(n); if (it == d.attr.end()) return MissingAttr(n); *v = it->second.f; return OkStatus(); }
inline Status GetNodeAttr(const NodeDef& d, const std::string& n, std::string* v) { auto it = d.attr.find(n); if (it == d.attr.end()) return MissingAttr(n); *v = it->second.s; return OkStatus(); }
inline Status GetNodeAttr(const NodeDef& d, const std::string& n, TensorShape* v) { auto it = d.attr.find(n); if (it == d.attr.end()) return MissingAttr(n); *v = it->second.shape; return OkStatus(); }

class OpKernelConstruction {
 public:
  explicit OpKernelConstruction(const NodeDef& d) : def_(d) {}
  const NodeDef& def() const { return def_; }
  template <class T> Status GetAttr(const std::string& n, T* v) const { return GetNodeAttr(def_, n, v); }
  void CtxFailure(const Status& s) { status_ = s; }
  const Status& status() const { return status_; }
 private:
  NodeDef def_;
  Status status_;
};

class OpKernelContext;
class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction* c) : def_(c->def()) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* ctx) = 0;
  const NodeDef& def() const { return def_; }
 private:
  NodeDef def_;
};

namespace core {
class RefCounted { public: virtual ~RefCounted() {} void Ref() {} void Unref() {} };
struct ScopedUnref { explicit ScopedUnref(RefCounted*) {} };
}  // namespace core
class ResourceBase : public core::RefCounted {
 public:
  virtual std::string DebugString() const { return "resource"; }
  virtual int64 MemoryUsed() const { return 0; }
};

namespace lookup {
// tensorflow/core/framework/lookup_interface.h (TF 2.15): the virtuals a table resource implements
class LookupInterface : public ResourceBase {
 public:
  virtual size_t size() const = 0;
  virtual Status Find(OpKernelContext* ctx, const Tensor& keys, Tensor* values, const Tensor& default_value) = 0;
  virtual Status Insert(OpKernelContext* ctx, const Tensor& keys, const Tensor& values) = 0;
  virtual Status Remove(OpKernelContext* ctx, const Tensor& keys) = 0;
  virtual Status ExportValues(OpKernelContext* ctx) = 0;
  virtual Status ImportValues(OpKernelContext* ctx, const Tensor& keys, const Tensor& values) = 0;
  virtual DataType key_dtype() const = 0;
  virtual DataType value_dtype() const = 0;
  virtual TensorShape key_shape() const = 0;
  virtual TensorShape value_shape() const = 0;
};
}  // namespace lookup

class OpKernelContext {
 public:
  template <class D> const D& eigen_device() const { static D d; return d; }
  int device_ordinal() const { return 0; }
  Status allocate_output(const std::string& name, const TensorShape& s, Tensor** out, AllocatorAttributes = AllocatorAttributes()) {
    auto it = out_dtypes.find(name);
    if (it == out_dtypes.end()) return errors::InvalidArgument("unknown output ", name);
    outputs[name] = Tensor(it->second, s);
    *out = &outputs[name];
    return OkStatus();
  }
  Status allocate_temp(DataType dt, const TensorShape& s, Tensor* out) {
    *out = Tensor(dt, s);
    return OkStatus();
  }
  const Tensor& input(int i) const { return inputs.at(i); }
  void CtxFailure(const Status& s) { status_ = s; }
  const Status& status() const { return status_; }
  // test-side state
  std::map<std::string, DataType> out_dtypes;   // the op's declared outputs
  std::map<std::string, Tensor> outputs;
  std::vector<Tensor> inputs;
  lookup::LookupInterface* table = nullptr;     // what input "table_handle" resolves to
 private:
  Status status_;
};
inline Status GetLookupTable(const std::string&, OpKernelContext* ctx, lookup::LookupInterface** t) {
  if (!ctx->table) return errors::InvalidArgument("table_handle does not hold a table");
  *t = ctx->table;
  return OkStatus();
}

// ---- registration (records what was registered so a test can instantiate kernels by op name) -----------------------
struct OpDefBuilderMock {
  std::string name; std::vector<std::string> inputs, outputs, attrs;
  explicit OpDefBuilderMock(const std::string& n) : name(n) {}
  OpDefBuilderMock& Input(const std::string& s) { inputs.push_back(s); return *this; }
  OpDefBuilderMock& Output(const std::string& s) { outputs.push_back(s); return *this; }
  OpDefBuilderMock& Attr(const std::string& s) { attrs.push_back(s); return *this; }
};
struct Registry {
  std::map<std::string, OpDefBuilderMock> ops;
  std::map<std::string, std::function<OpKernel*(OpKernelConstruction*)>> kernels;
  static Registry& Get() { static Registry r; return r; }
};
struct OpRegistrar { OpRegistrar(const OpDefBuilderMock& b) { Registry::Get().ops.emplace(b.name, b); } };
struct KernelDefMock {
  std::string op;
  KernelDefMock& Device(const char*) { return *this; }
  KernelDefMock& HostMemory(const char*) { return *this; }
  template <class T> KernelDefMock& TypeConstraint(const char*) { return *this; }
};
inline KernelDefMock Name(const char* n) { KernelDefMock k; k.op = n; return k; }
struct KernelRegistrar {
  KernelRegistrar(const KernelDefMock& k, std::function<OpKernel*(OpKernelConstruction*)> f) { Registry::Get().kernels[k.op] = f; }
};
#define DEVICE_GPU "GPU"
#define TF_MOCK_CAT2(a, b) a##b
#define TF_MOCK_CAT(a, b) TF_MOCK_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::OpRegistrar TF_MOCK_CAT(_op_reg_, __COUNTER__) = ::tensorflow::OpDefBuilderMock(name)
#define REGISTER_KERNEL_BUILDER(kdef, ...) \
  static ::tensorflow::KernelRegistrar TF_MOCK_CAT(_k_reg_, __COUNTER__)((kdef), [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new __VA_ARGS__(c); })

}  // namespace tensorflow
#endif  // TESTS_TF_MOCK_H_
