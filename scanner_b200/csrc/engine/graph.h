// graph.h -- the op DAG a job runs, its static analysis and the per-task row algebra.
// Restates what the reference computes in scanner/engine/dag_analysis.cpp:
//   * domain sizes per op           (determine_input_rows_to_slices, :468-782)
//   * batch / stencil / warmup defaults (populate_analysis_info, :1014-1041)
//   * column liveness               (perform_liveness_analysis, :1145-1326)
//   * per-task required rows        (derive_stencil_requirements, :1328-1743) -> TaskStream
// and the domain samplers of scanner/engine/sampler.cpp:33-498 (All, Strided, StridedRanges,
// Gather, SpaceNull, SpaceRepeat) and the Slice partitioners (:500-770 Strided, StridedRange,
// Gather).  Slice/Unslice follow the reference's restrictions (dag_analysis.cpp:70-72,151-160,
// 571-714): one slice level, every slice of a job has the same number of groups, an Unslice
// feeds only sinks.
// Host-side integer bookkeeping only.
#pragma once
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "registry.h"
#include "scanner/util/common.h"

namespace scanner {
namespace internal {

// ---------------------------------------------------------------------------------------------
class DomainSampler {
 public:
  virtual ~DomainSampler() {}
  virtual Result validate() const = 0;
  // rows of the upstream domain needed to produce `downstream_rows`
  virtual Result get_upstream_rows(const std::vector<i64>& downstream_rows,
                                   std::vector<i64>& upstream_rows) const = 0;
  virtual Result get_num_downstream_rows(i64 num_upstream_rows, i64& num_downstream_rows) const = 0;
  // for the upstream rows actually available: which downstream rows they produce and from which
  // position of `upstream_rows` (-1 = null element, Space only)
  virtual Result get_downstream_rows(const std::vector<i64>& upstream_rows,
                                     std::vector<i64>& downstream_rows,
                                     std::vector<i64>& downstream_upstream_mapping) const = 0;
};

// name in {"All","Strided","StridedRanges","Gather","SpaceNull","SpaceRepeat"}; args are the
// proto3 bytes of the matching message in the reference's scanner/sampler_args.proto.
// Internal names used for the per-task view of a sliced job: "SliceOffset" (args: two
// little-endian i64 base, count: downstream rows [base, base+count) <-> upstream [0, count)).
Result make_domain_sampler(const std::string& name, const std::vector<u8>& args,
                           std::unique_ptr<DomainSampler>& out);

// Slice partitioner: the groups of upstream rows a Slice op cuts its input into.
// name in {"Strided","StridedRange","StridedRanges","Gather"} (reference sampler.cpp:505-770).
Result make_partition_groups(const std::string& name, const std::vector<u8>& args, i64 num_rows,
                             std::vector<std::vector<i64>>& groups);

// ---------------------------------------------------------------------------------------------
enum class OpKind { Source, Sample, Space, Kernel, Sink };
// Slice / Unslice are Sample ops with a role: per task they act as a Gather of the group's rows
// and as a row offset back into the concatenated output.
enum class SliceRole { None, Slice, Unslice };

struct OpInput {
  i32 op_index;
  std::string column;
};

struct GraphOp {
  OpKind kind = OpKind::Kernel;
  std::string name;               // registered op name ("Histogram"), or Input/Sample/Space/Output
  std::vector<OpInput> inputs;
  proto::DeviceType device_type = proto::CPU;
  std::vector<u8> args;           // KernelConfig.args
  i32 batch = -1;                 // -1: kernel's preferred batch
  std::vector<i32> stencil;       // empty: op's preferred stencil
  i32 warmup = -1;                // -1: op's registered warmup
  std::vector<std::string> output_columns;  // filled by analysis for Kernel ops
  // Source: column type of its single output; Sink: name of the stored column
  proto::ColumnType column_type = proto::Bytes;
  std::string sink_column_name;
  SliceRole slice_role = SliceRole::None;
};

// Per-job (one input stream -> one output stream) bindings.
struct JobParams {
  std::map<i32, i64> source_rows;                                   // source op -> #rows
  std::map<i32, std::pair<std::string, std::vector<u8>>> samplers;  // Sample/Space op -> (fn,args)
  std::map<i32, std::vector<u8>> stream_args;                       // kernel op -> new_stream args
  // sliced jobs (reference SliceList arguments): Slice op -> partitioner; ops inside the slice may
  // carry one sampler / one new_stream argument per slice group (index = group)
  std::map<i32, std::pair<std::string, std::vector<u8>>> partitioners;
  std::map<i32, std::vector<std::pair<std::string, std::vector<u8>>>> group_samplers;
  std::map<i32, std::vector<std::vector<u8>>> group_stream_args;
};

// What domain_sizes works out for a sliced job.
struct SliceInfo {
  i32 groups = 0;                                          // 0: the job is not sliced
  std::vector<std::vector<i64>> rows_per_op;               // [group][op], slice-local sizes inside the slice
  std::map<i32, std::vector<std::vector<i64>>> slice_rows; // Slice op -> [group] -> upstream rows
  std::vector<i64> out_base;                               // [group] first output row of the group (+ total at the end)
};

// What one op must consume / compute / emit for one task (reference runtime.h:67-79).
struct TaskStream {
  std::vector<i64> valid_input_rows;
  std::vector<i64> compute_input_rows;
  std::vector<i64> valid_output_rows;
};

struct GraphAnalysis {
  // resolved per-op execution parameters
  std::vector<i32> batch;
  std::vector<std::vector<i32>> stencil;
  std::vector<i32> warmup;
  std::vector<bool> bounded, unbounded;
  // (producer op, column idx) consumed last by op index `last_use`; -1 = never read
  std::vector<std::vector<i32>> last_use;        // [op][output column]
  std::vector<std::vector<i32>> input_col_index; // [op][input] -> producer's output column idx
};

class Graph {
 public:
  std::vector<GraphOp> ops;  // topological order; sources first is NOT required

  // Validates the DAG against the registries and fills GraphAnalysis
  // (reference validate_jobs_and_ops :43-466, populate_analysis_info, liveness).
  Result analyze(GraphAnalysis& out);

  // True when every kernel that (transitively through Sample/Space ops) reads the output of
  // `source_op` registered `layout` for that input column -- the decode stage may then deliver
  // decoder-native surfaces instead of RGB24 (scanner-b200 extension, frame.h FrameLayout).
  bool consumers_accept_layout(i32 source_op, FrameLayout layout) const;

  // Rows each op produces for this job (domain sizes).  For a sliced job `slices` (if given) gets
  // the per-group sizes; rows_per_op then holds group sums for the ops inside the slice.
  Result domain_sizes(const JobParams& job, std::vector<i64>& rows_per_op, SliceInfo* slices = nullptr) const;

  // Parameters as one task of slice group `group` sees them: Slice ops gather the group's rows,
  // Unslice ops offset into the concatenated output, per-group samplers / stream args selected.
  void slice_view(const JobParams& job, const SliceInfo& slices, i32 group, JobParams& view_params,
                  std::vector<i64>& view_rows) const;

  std::vector<i32> slice_level;  // filled by analyze(): 0 outside, 1 between Slice and Unslice

  // Declared type of output column `column` of op `op_index` (Source: its column_type; Kernel: the
  // op's registration; Sample / Space / Slice ops pass their input's type through).  What the save
  // stage stores a sink as is decided by this, never by the rows that happen to arrive.
  proto::ColumnType column_type_of(i32 op_index, const std::string& column) const;

  // Back-propagate `output_rows` (rows of every sink for this task) to every op
  // (derive_stencil_requirements).  task_streams[i] corresponds to ops[i]; for Source ops
  // valid_output_rows are the rows to load.
  Result derive_task_streams(const GraphAnalysis& an, const JobParams& job,
                             const std::vector<i64>& rows_per_op,
                             const std::vector<i64>& output_rows,
                             std::vector<TaskStream>& task_streams) const;
};

}  // namespace internal
}  // namespace scanner
