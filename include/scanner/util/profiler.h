// scanner/util/profiler.h -- interval/counter recorder handed to kernels via set_profiler
// (reference scanner/util/profiler.h, profiler.inl:22-37).  Same calls (add_interval,
// increment), same labels on the hot path ("evaluate:<op>", "op_marshal", "get_frames", ...);
// records stay in memory and are exposed through the engine's stats instead of the reference's
// per-node binary dump.
#pragma once
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "scanner/util/common.h"

namespace scanner {

using timepoint_t = std::chrono::time_point<std::chrono::high_resolution_clock>;
inline timepoint_t now() { return std::chrono::high_resolution_clock::now(); }
inline double nano_since(timepoint_t t) {
  return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(now() - t).count();
}

enum class ProfilerLevel { Debug = 0, Info = 1, Important = 2 };

class Profiler {
 public:
  struct TaskRecord {
    std::string key;
    i64 start_ns;
    i64 end_ns;
    i32 worker;  // pipeline instance that recorded it (thread_worker() of the calling thread)
  };
  // id the engine gives each pipeline-instance thread; -1 outside of one
  static i32& thread_worker() {
    static thread_local i32 w = -1;
    return w;
  }
  explicit Profiler(timepoint_t base = now()) : base_(base) {}

  void add_interval(const std::string& key, timepoint_t start, timepoint_t end,
                    ProfilerLevel level = ProfilerLevel::Info) {
    if ((int)level < min_level()) return;
    auto ns = [this](timepoint_t t) {
      return (i64)std::chrono::duration_cast<std::chrono::nanoseconds>(t - base_).count();
    };
    std::lock_guard<std::mutex> g(mu_);
    totals_ns_[key] += ns(end) - ns(start);
    counts_[key] += 1;
    if (keep_records_) records_.push_back({key, ns(start), ns(end), thread_worker()});
  }
  void increment(const std::string& key, i64 value) {
    std::lock_guard<std::mutex> g(mu_);
    counters_[key] += value;
  }
  std::map<std::string, i64> counters() const {
    std::lock_guard<std::mutex> g(mu_);
    return counters_;
  }
  std::map<std::string, i64> interval_totals_ns() const {
    std::lock_guard<std::mutex> g(mu_);
    return totals_ns_;
  }
  std::map<std::string, i64> interval_counts() const {
    std::lock_guard<std::mutex> g(mu_);
    return counts_;
  }
  const std::vector<TaskRecord>& records() const { return records_; }
  void keep_records(bool k) { keep_records_ = k; }
  static int& min_level() {
    static int lvl = (int)ProfilerLevel::Info;
    return lvl;
  }

 private:
  timepoint_t base_;
  mutable std::mutex mu_;
  std::map<std::string, i64> counters_, totals_ns_, counts_;
  std::vector<TaskRecord> records_;
  bool keep_records_ = false;
};

class ProfileBlock {
 public:
  ProfileBlock(Profiler* p, const std::string& key) : p_(p), key_(key), start_(now()) {}
  ~ProfileBlock() {
    if (p_) p_->add_interval(key_, start_, now());
  }

 private:
  Profiler* p_;
  std::string key_;
  timepoint_t start_;
};

}  // namespace scanner
