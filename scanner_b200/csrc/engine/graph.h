// graph.h -- the op DAG a job runs, its static analysis and the per-task row algebra.
// Restates what the reference computes in scanner/engine/dag_analysis.cpp:
//   * domain sizes per op           (determine_input_rows_to_slices, :468-782)
//   * batch / stencil / warmup defaults (populate_analysis_info, :1014-1041)
//   * column liveness               (perform_liveness_analysis, :1145-1326)
//   * per-task required rows        (derive_stencil_requirements, :1328-1743) -> TaskStream
// and the domain samplers of scanner/engine/sampler.cpp:33-498 (All, Strided, StridedRanges,
// Gather, SpaceNull, SpaceRepeat).  Slice/Unslice are out of scope.
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
Result make_domain_sampler(const std::string& name, const std::vector<u8>& args,
                           std::unique_ptr<DomainSampler>& out);

// ---------------------------------------------------------------------------------------------
enum class OpKind { Source, Sample, Space, Kernel, Sink };

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
};

// Per-job (one input stream -> one output stream) bindings.
struct JobParams {
  std::map<i32, i64> source_rows;                                   // source op -> #rows
  std::map<i32, std::pair<std::string, std::vector<u8>>> samplers;  // Sample/Space op -> (fn,args)
  std::map<i32, std::vector<u8>> stream_args;                       // kernel op -> new_stream args
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

  // Rows each op produces for this job (domain sizes).
  Result domain_sizes(const JobParams& job, std::vector<i64>& rows_per_op) const;

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
