// Stand-in for google-benchmark, TEST INFRASTRUCTURE ONLY (see gtest stub).
// BM_* bodies compile but are never registered; oracle/ref_build times the
// reference through its own driver instead.
#ifndef ORACLE_STUB_BENCHMARK_H_
#define ORACLE_STUB_BENCHMARK_H_
#include <cstdint>
namespace benchmark {
class State {
 public:
  explicit State(int64_t arg = 1, int64_t iters = 1) : arg_(arg), left_(iters) {}
  int64_t range(int) const { return arg_; }
  struct It {
    State* s;
    bool operator!=(const It&) const { return s->left_ > 0; }
    void operator++() { --s->left_; }
    int operator*() const { return 0; }
  };
  It begin() { return It{this}; }
  It end() { return It{this}; }
  void SetItemsProcessed(int64_t) {}
  void SetBytesProcessed(int64_t) {}
 private:
  int64_t arg_, left_;
};
struct Builder {
  Builder* RangeMultiplier(int) { return this; }
  Builder* Range(int64_t, int64_t) { return this; }
  Builder* DenseRange(int64_t, int64_t, int64_t = 1) { return this; }
  Builder* Arg(int64_t) { return this; }
  Builder* Unit(int) { return this; }
};
inline Builder* Register(void (*)(State&)) { static Builder b; return &b; }
template <class T> inline void DoNotOptimize(T const& v) { asm volatile("" : : "r,m"(v) : "memory"); }
template <class T> inline void DoNotOptimize(T& v) { asm volatile("" : "+r,m"(v) : : "memory"); }
inline void Initialize(int*, char**) {}
inline void RunSpecifiedBenchmarks() {}
constexpr int kMillisecond = 0;
}  // namespace benchmark
#define ORACLE_BCAT_(a, b) a##b
#define ORACLE_BCAT(a, b) ORACLE_BCAT_(a, b)
#define BENCHMARK(fn) \
  [[maybe_unused]] static ::benchmark::Builder* ORACLE_BCAT(bm_reg_, __LINE__) = ::benchmark::Register(fn)
#endif
