// evaluate.h -- the evaluate stage: drives kernels over work packets with Scanner's row semantics
// (reference scanner/engine/evaluate_worker.cpp:408-1327 EvaluateWorker, runtime.cpp:141-189
// copy_or_ref_elements).  One instance per pipeline instance; single-threaded by contract.
#pragma once
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "graph.h"
#include "scanner/api/kernel.h"

namespace scanner {
namespace internal {

// Rows of one column travelling between stages.
struct ColumnBatch {
  DeviceHandle device = CPU_DEVICE;
  Elements elements;
  std::vector<i64> row_ids;
};

// same address space -> extra refs on the same payloads; otherwise ONE block on the target and one
// async copy per contiguous run (reference runtime.cpp:141-189).
Elements copy_or_ref_elements(DeviceHandle src_device, DeviceHandle dst_device, const Elements& in);
void delete_elements(DeviceHandle device, Elements& elements);

// Shared by the pipeline instances of one run: fetch_resources() of an op runs on exactly one of
// its kernel instances, every instance's setup_with_resources() runs after it has returned
// (reference evaluate_worker.cpp:493-550: worker instance 0 fetches, the others wait on a
// condition variable).
struct ResourceGate {
  std::mutex mu;
  std::condition_variable cv;
  std::map<i32, int> state;  // op index -> 0 untouched, 1 fetching, 2 fetched, 3 failed
  std::map<i32, std::string> error;
};

class EvaluateWorker {
 public:
  // gpu_id: the GPU this pipeline instance owns (-1 = none; GPU kernels are then an error).
  // gate: shared by the instances of the run (nullptr: this worker fetches for itself).
  EvaluateWorker(const Graph& graph, const GraphAnalysis& analysis, i32 gpu_id, i32 node_id,
                 Profiler* profiler, ResourceGate* gate = nullptr);
  ~EvaluateWorker();

  // Instantiate every kernel: validate(), fetch_resources(), setup_with_resources()
  // (reference evaluate_worker.cpp:452-550).
  Result init();

  // reset() + new_stream(args) on every kernel, install the task's row sets (:581-708).
  Result new_task(const JobParams& job, const std::vector<i64>& rows_per_op,
                  const std::vector<TaskStream>& task_streams);

  // One work packet: rows of every Source op (ownership of the elements passes in).  Returns the
  // rows each Sink accepted during this packet, keyed by sink op index; the caller owns them.
  Result feed(std::map<i32, ColumnBatch>& source_columns, std::map<i32, ColumnBatch>& sink_columns);

  // Task is over: every op must have consumed all its valid input rows (:586-596).
  Result end_task();

  DeviceHandle gpu_device() const { return DeviceHandle(DeviceType::GPU, gpu_id_); }

 private:
  struct OpState {
    std::unique_ptr<BaseKernel> kernel;
    DeviceHandle device = CPU_DEVICE;
    std::vector<DeviceHandle> in_dev, out_dev;
    std::unique_ptr<DomainSampler> sampler;
    TaskStream ts;
    i64 domain_rows = 0;                // rows of the input domain (REPEAT_EDGE clamp)
    std::vector<size_t> in_idx;         // per input: next expected position in valid_input_rows
    size_t next_compute = 0;            // next position in compute_input_rows
    size_t next_out = 0;                // next position in valid_output_rows
    std::vector<std::deque<Element>> cache;   // per input: elements not yet retired
    std::vector<std::deque<i64>> cache_rows;
  };

  void clear_caches();
  const Element* find_cached(const OpState& st, size_t input, i64 row) const;

  const Graph& graph_;
  const GraphAnalysis& an_;
  i32 gpu_id_;
  i32 node_id_;
  Profiler* profiler_;
  ResourceGate* gate_;
  std::vector<OpState> state_;
};

}  // namespace internal
}  // namespace scanner
