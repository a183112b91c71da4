// oracle/ref_wrapper.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C-ABI around the reference's OWN vendored libcuckoo engine, compiled from the
// sources where they lie under /root/reference (include path only; nothing is copied):
//   tensorflow_recommenders_addons/dynamic_embedding/core/lib/cuckoo/cuckoohash_map.hh
// (incl. TFRA's insert_or_accum / accumrase_fn, cuckoohash_map.hh:620-633, 756-765).
//
// What is restated here (the only TensorFlow-dependent layer of the CPU path, with
// TTypes<V,2>::Tensor replaced by raw pointers):
//   * ValueArray<V,DIM>::operator+=            lookup_table_op_cpu.h:42-52
//   * HybridHash<int64> (murmur3 fmix64)       lookup_table_op_cpu.h:90-101
//   * TableWrapperOptimized<K,V,DIM>           lookup_table_op_cpu.h:148-263
//       find / find+exists / insert_or_assign / insert_or_accum / dump / size / clear / erase
//   * TableWrapperDefault (dim > 100)          lookup_table_op_cpu.h:265-389   (std::vector rows)
//   * LaunchTensors{Find,FindWithExists,Insert,Accum}<CPUDevice>  cuckoo_hashtable_op.cc:39-182
//       contiguous key ranges over worker threads (TF Shard()), is_full_default rule :48-50
//   * Remove = single-threaded erase loop      cuckoo_hashtable_op.cc:268-276
//
// Built by oracle/Makefile into oracle/_ref/libtfra_cuckoo_ref.so (git-ignored, travels to
// the GPU box with the snapshot).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <array>
#include <algorithm>

#include "tensorflow_recommenders_addons/dynamic_embedding/core/lib/cuckoo/cuckoohash_map.hh"

namespace {

using int64 = long long;

// lookup_table_op_cpu.h:42-52
template <class V, size_t DIM>
class ValueArray final : public std::array<V, DIM> {
 public:
  inline ValueArray<V, DIM>& operator+=(const ValueArray<V, DIM>& rhs) noexcept {
    for (size_t i = 0; i < DIM; i++) (*this)[i] += rhs[i];
    return *this;
  }
};

// lookup_table_op_cpu.h:54-64 (DefaultValueArray over InlinedVector -> std::vector here)
template <class V>
class DynValueArray final : public std::vector<V> {
 public:
  inline DynValueArray<V>& operator+=(const DynValueArray<V>& rhs) noexcept {
    for (size_t i = 0; i < this->size(); i++) (*this)[i] = ((*this)[i]) + rhs[i];
    return *this;
  }
};

// lookup_table_op_cpu.h:90-101
struct HybridHashI64 {
  inline std::size_t operator()(int64 const& key) const noexcept {
    uint64_t k = static_cast<uint64_t>(key);
    k ^= k >> 33;
    k *= UINT64_C(0xff51afd7ed558ccd);
    k ^= k >> 33;
    k *= UINT64_C(0xc4ceb9fe1a85ec53);
    k ^= k >> 33;
    return static_cast<std::size_t>(k);
  }
};

struct TableBase {
  virtual ~TableBase() {}
  virtual bool insert_or_assign(int64 key, const float* values, int64 index) = 0;
  virtual bool insert_or_accum(int64 key, const float* vod, bool exist, int64 index) = 0;
  virtual bool find(int64 key, float* out, const float* defaults, bool full_default,
                    int64 index) = 0;
  virtual size_t dump(int64* keys, float* values, size_t offset, size_t length) = 0;
  virtual size_t size() const = 0;
  virtual void clear() = 0;
  virtual bool erase(int64 key) = 0;
  int64 dim = 0;
};

// lookup_table_op_cpu.h:148-263
template <size_t DIM>
struct TableOptimized final : TableBase {
  using ValueType = ValueArray<float, DIM>;
  using Table = cuckoohash_map<int64, ValueType, HybridHashI64>;
  Table* table_;
  explicit TableOptimized(size_t init_size) {
    table_ = new Table(init_size);
    dim = DIM;
  }
  ~TableOptimized() override { delete table_; }
  bool insert_or_assign(int64 key, const float* values, int64 index) override {
    ValueType v;
    std::copy_n(values + index * (int64)DIM, DIM, v.begin());
    return table_->insert_or_assign(key, v);
  }
  bool insert_or_accum(int64 key, const float* vod, bool exist, int64 index) override {
    ValueType v;
    std::copy_n(vod + index * (int64)DIM, DIM, v.begin());
    return table_->insert_or_accum(key, v, exist);
  }
  bool find(int64 key, float* out, const float* defaults, bool full_default,
            int64 index) override {
    ValueType v;
    bool exist = table_->find(key, v);
    if (exist) {
      std::copy_n(v.begin(), DIM, out + index * (int64)DIM);
    } else {
      for (size_t j = 0; j < DIM; j++)
        out[index * (int64)DIM + j] = full_default ? defaults[index * (int64)DIM + j] : defaults[j];
    }
    return exist;
  }
  size_t dump(int64* keys, float* values, size_t offset, size_t length) override {
    auto lt = table_->lock_table();
    auto lt_size = lt.size();
    if (offset > lt_size || lt_size == 0) return 0;
    auto b = lt.begin();
    for (size_t i = 0; i < offset; ++i) ++b;
    auto e = b;
    if (offset + length >= lt_size) {
      e = lt.end();
    } else {
      for (size_t i = 0; i < length; ++i) ++e;
    }
    size_t n = 0;
    for (auto it = b; it != e; ++it, ++keys, values += DIM, ++n) {
      *keys = it->first;
      std::copy_n(it->second.begin(), DIM, values);
    }
    return n;
  }
  size_t size() const override { return table_->size(); }
  void clear() override { table_->clear(); }
  bool erase(int64 key) override { return table_->erase(key); }
};

// lookup_table_op_cpu.h:265-389
struct TableDefault final : TableBase {
  using ValueType = DynValueArray<float>;
  using Table = cuckoohash_map<int64, ValueType, HybridHashI64>;
  Table* table_;
  TableDefault(size_t init_size, int64 d) {
    table_ = new Table(init_size);
    dim = d;
  }
  ~TableDefault() override { delete table_; }
  bool insert_or_assign(int64 key, const float* values, int64 index) override {
    ValueType v;
    v.reserve(dim);
    for (int64 j = 0; j < dim; j++) v.push_back(values[index * dim + j]);
    return table_->insert_or_assign(key, v);
  }
  bool insert_or_accum(int64 key, const float* vod, bool exist, int64 index) override {
    ValueType v;
    v.reserve(dim);
    for (int64 j = 0; j < dim; j++) v.push_back(vod[index * dim + j]);
    return table_->insert_or_accum(key, v, exist);
  }
  bool find(int64 key, float* out, const float* defaults, bool full_default,
            int64 index) override {
    ValueType v;
    bool exist = table_->find(key, v);
    if (exist) {
      std::copy_n(v.begin(), dim, out + index * dim);
    } else {
      for (int64 j = 0; j < dim; j++)
        out[index * dim + j] = full_default ? defaults[index * dim + j] : defaults[j];
    }
    return exist;
  }
  size_t dump(int64* keys, float* values, size_t offset, size_t length) override {
    auto lt = table_->lock_table();
    auto lt_size = lt.size();
    if (offset > lt_size || lt_size == 0) return 0;
    auto b = lt.begin();
    for (size_t i = 0; i < offset; ++i) ++b;
    auto e = b;
    if (offset + length >= lt_size) {
      e = lt.end();
    } else {
      for (size_t i = 0; i < length; ++i) ++e;
    }
    size_t n = 0;
    for (auto it = b; it != e; ++it, ++keys, values += dim, ++n) {
      *keys = it->first;
      std::copy_n(it->second.begin(), dim, values);
    }
    return n;
  }
  size_t size() const override { return table_->size(); }
  void clear() override { table_->clear(); }
  bool erase(int64 key) override { return table_->erase(key); }
};

// CreateTable macro ladder, lookup_table_op_cpu.h:403-468: DIM<=100 -> optimized, else default.
// Only the dims the tests/bench use are instantiated (each is a full libcuckoo instantiation).
TableBase* create_table(size_t init_size, int64 dim) {
  switch (dim) {
#define CASE(D) \
  case D:       \
    return new TableOptimized<D>(init_size);
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(8) CASE(10) CASE(16) CASE(32) CASE(64) CASE(100)
#undef CASE
    default:
      return new TableDefault(init_size, dim);
  }
}

// Worker pool standing in for TF's intra-op pool + Shard() (cuckoo_hashtable_op.cc:60-63):
// the key range is cut into contiguous blocks, one per worker.
class Pool {
 public:
  explicit Pool(int n) : n_(n), stop_(false), gen_(0), pending_(0) {
    for (int i = 1; i < n_; i++) th_.emplace_back([this, i] { loop(i); });
  }
  ~Pool() {
    {
      std::unique_lock<std::mutex> l(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return n_; }
  void shard(int64 total, const std::function<void(int64, int64)>& fn) {
    if (n_ <= 1 || total < 2) {
      fn(0, total);
      return;
    }
    int64 block = (total + n_ - 1) / n_;
    {
      std::unique_lock<std::mutex> l(m_);
      fn_ = &fn;
      total_ = total;
      block_ = block;
      pending_ = n_ - 1;
      ++gen_;
    }
    cv_.notify_all();
    fn(0, std::min(block, total));
    std::unique_lock<std::mutex> l(m_);
    done_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void loop(int id) {
    uint64_t seen = 0;
    while (true) {
      const std::function<void(int64, int64)>* fn;
      int64 total, block;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        fn = fn_;
        total = total_;
        block = block_;
      }
      int64 b = std::min<int64>(id * block, total), e = std::min<int64>(b + block, total);
      if (b < e) (*fn)(b, e);
      {
        std::unique_lock<std::mutex> l(m_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  int n_;
  bool stop_;
  uint64_t gen_;
  int pending_;
  const std::function<void(int64, int64)>* fn_ = nullptr;
  int64 total_ = 0, block_ = 0;
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> th_;
};

struct Handle {
  TableBase* t;
  Pool* pool;
};

}  // namespace

extern "C" {

void* ref_create(long long dim, size_t init_size, int num_threads) {
  if (init_size == 0) init_size = 1024 * 8;  // cuckoo_hashtable_op.cc:199-205
  if (num_threads < 1) num_threads = 1;
  Handle* h = new Handle;
  h->t = create_table(init_size, dim);
  h->pool = new Pool(num_threads);
  return h;
}

void ref_destroy(void* p) {
  Handle* h = (Handle*)p;
  delete h->pool;
  delete h->t;
  delete h;
}

int ref_is_optimized(void* p) { return ((Handle*)p)->t->dim <= 100 ? 1 : 0; }

// LaunchTensorsFind / LaunchTensorsFindWithExists, cuckoo_hashtable_op.cc:39-106
void ref_find(void* p, const long long* keys, long long n, const float* defaults,
              int full_default, float* out, unsigned char* exists) {
  Handle* h = (Handle*)p;
  TableBase* t = h->t;
  h->pool->shard(n, [=](int64 b, int64 e) {
    for (int64 i = b; i < e; ++i) {
      bool ex = t->find(keys[i], out, defaults, full_default != 0, i);
      if (exists) exists[i] = ex ? 1 : 0;
    }
  });
}

// LaunchTensorsInsert, cuckoo_hashtable_op.cc:111-150
void ref_insert(void* p, const long long* keys, const float* values, long long n) {
  Handle* h = (Handle*)p;
  TableBase* t = h->t;
  h->pool->shard(n, [=](int64 b, int64 e) {
    for (int64 i = b; i < e; ++i) t->insert_or_assign(keys[i], values, i);
  });
}

// LaunchTensorsAccum, cuckoo_hashtable_op.cc:155-182
void ref_accum(void* p, const long long* keys, const float* vod, const unsigned char* exists,
               long long n) {
  Handle* h = (Handle*)p;
  TableBase* t = h->t;
  h->pool->shard(n, [=](int64 b, int64 e) {
    for (int64 i = b; i < e; ++i) t->insert_or_accum(keys[i], vod, exists[i] != 0, i);
  });
}

// Remove, cuckoo_hashtable_op.cc:268-276 (single-threaded)
void ref_remove(void* p, const long long* keys, long long n) {
  TableBase* t = ((Handle*)p)->t;
  for (int64 i = 0; i < n; ++i) t->erase(keys[i]);
}

void ref_clear(void* p) { ((Handle*)p)->t->clear(); }

size_t ref_size(void* p) { return ((Handle*)p)->t->size(); }

// ExportValues -> dump(keys, values, 0, size), cuckoo_hashtable_op.cc:293-308
size_t ref_export(void* p, long long* keys, float* values, size_t offset, size_t length) {
  return ((Handle*)p)->t->dump(keys, values, offset, length);
}

int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
