// graph.cpp -- samplers, DAG analysis and per-task row derivation (see graph.h for the map to
// the reference's dag_analysis.cpp / sampler.cpp).
#include "graph.h"

#include <algorithm>
#include <cmath>

#include "sampler_args.pb.h"

namespace scanner {
namespace internal {

namespace {

Result ok() {
  Result r;
  r.set_success(true);
  return r;
}

i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }

// "All" -- identity (reference sampler.cpp:33-76)
class AllSampler : public DomainSampler {
 public:
  Result validate() const override { return ok(); }
  Result get_upstream_rows(const std::vector<i64>& d, std::vector<i64>& u) const override {
    u = d;
    return ok();
  }
  Result get_num_downstream_rows(i64 n, i64& out) const override {
    out = n;
    return ok();
  }
  Result get_downstream_rows(const std::vector<i64>& u, std::vector<i64>& d,
                             std::vector<i64>& m) const override {
    d = u;
    for (size_t i = 0; i < u.size(); ++i) m.push_back((i64)i);
    return ok();
  }
};

// "Strided" (reference sampler.cpp:78-138): downstream row r <- upstream r*stride
class StridedSampler : public DomainSampler {
 public:
  explicit StridedSampler(const std::vector<u8>& args) {
    valid_.set_success(true);
    if (!args_.ParseFromArray(args.data(), (int)args.size())) {
      RESULT_ERROR(&valid_, "StridedSampler provided with invalid protobuf args");
    } else if (args_.stride() <= 0) {
      RESULT_ERROR(&valid_, "Strided sampler stride (%ld) must be greater than zero", (long)args_.stride());
    }
  }
  Result validate() const override { return valid_; }
  Result get_upstream_rows(const std::vector<i64>& d, std::vector<i64>& u) const override {
    for (i64 r : d) u.push_back(r * args_.stride());
    return ok();
  }
  Result get_num_downstream_rows(i64 n, i64& out) const override {
    out = ceil_div(n, args_.stride());
    return ok();
  }
  Result get_downstream_rows(const std::vector<i64>& u, std::vector<i64>& d,
                             std::vector<i64>& m) const override {
    for (size_t i = 0; i < u.size(); ++i)
      if (u[i] % args_.stride() == 0) {
        d.push_back(u[i] / args_.stride());
        m.push_back((i64)i);
      }
    return ok();
  }

 private:
  Result valid_;
  StridedSamplerArgs args_;
};

// "StridedRanges" (reference sampler.cpp:140-263): concatenation of [start,end) ranges, each
// sampled with the stride.
class StridedRangesSampler : public DomainSampler {
 public:
  explicit StridedRangesSampler(const std::vector<u8>& args) {
    valid_.set_success(true);
    if (!args_.ParseFromArray(args.data(), (int)args.size())) {
      RESULT_ERROR(&valid_, "StridedRange sampler provided with invalid protobuf args");
      return;
    }
    if (args_.stride() <= 0) {
      RESULT_ERROR(&valid_, "StridedRange stride (%ld) must be greater than zero", (long)args_.stride());
      return;
    }
    if (args_.starts_size() != args_.ends_size()) {
      RESULT_ERROR(&valid_, "StridedRange starts and ends not the same size");
      return;
    }
    i64 offset = 0;
    for (int i = 0; i < args_.starts_size(); ++i) {
      if (args_.starts(i) > args_.ends(i)) {
        RESULT_ERROR(&valid_, "StridedRange start (%ld) should not be after end (%ld)",
                     (long)args_.starts(i), (long)args_.ends(i));
        return;
      }
      offsets_.push_back(offset);
      offset += ceil_div(args_.ends(i) - args_.starts(i), args_.stride());
    }
    offsets_.push_back(offset);
  }
  Result validate() const override { return valid_; }
  Result get_upstream_rows(const std::vector<i64>& d, std::vector<i64>& u) const override {
    Result r = ok();
    for (i64 row : d) {
      // first range whose end offset exceeds `row`
      auto it = std::upper_bound(offsets_.begin() + 1, offsets_.end(), row);
      if (it == offsets_.end() || row < 0) {
        RESULT_ERROR(&r, "StridedRange received out of bounds request for row %ld (max requestable row is %ld).",
                     (long)row, (long)offsets_.back());
        return r;
      }
      const size_t idx = (size_t)(it - offsets_.begin()) - 1;
      u.push_back(args_.starts((int)idx) + (row - offsets_[idx]) * args_.stride());
    }
    return r;
  }
  Result get_num_downstream_rows(i64 n, i64& out) const override {
    out = 0;
    int i = 0;
    for (; i < args_.ends_size(); ++i) {
      if (n < args_.ends(i)) break;
      out += ceil_div(args_.ends(i) - args_.starts(i), args_.stride());
    }
    if (i != args_.ends_size() && n > args_.starts(i)) out += ceil_div(n - args_.starts(i), args_.stride());
    return ok();
  }
  Result get_downstream_rows(const std::vector<i64>& u, std::vector<i64>& d,
                             std::vector<i64>& m) const override {
    i64 offset = 0;
    int range = 0;
    for (size_t i = 0; i < u.size(); ++i) {
      const i64 r = u[i];
      while (range < args_.ends_size() && !(r >= args_.starts(range) && r < args_.ends(range))) {
        if (r < args_.starts(range)) break;  // row lies in a gap before this range
        offset += ceil_div(args_.ends(range) - args_.starts(range), args_.stride());
        ++range;
      }
      if (range == args_.ends_size()) break;
      if (r < args_.starts(range)) continue;
      const i64 rel = r - args_.starts(range);
      if (rel % args_.stride() == 0) {
        d.push_back(offset + rel / args_.stride());
        m.push_back((i64)i);
      }
    }
    return ok();
  }

 private:
  Result valid_;
  StridedRangeSamplerArgs args_;
  std::vector<i64> offsets_;
};

// "Gather" (reference sampler.cpp:265-334): downstream row i <- upstream rows[i]
class GatherSampler : public DomainSampler {
 public:
  explicit GatherSampler(const std::vector<u8>& args) {
    valid_.set_success(true);
    if (!args_.ParseFromArray(args.data(), (int)args.size())) {
      RESULT_ERROR(&valid_, "Gather sampler provided with invalid protobuf args");
      return;
    }
    i64 off = 0;
    for (i64 r : args_.rows()) index_[r] = off++;
  }
  Result validate() const override { return valid_; }
  Result get_upstream_rows(const std::vector<i64>& d, std::vector<i64>& u) const override {
    Result r = ok();
    for (i64 row : d) {
      if (row < 0 || row >= args_.rows_size()) {
        RESULT_ERROR(&r, "Gather sampler received out of bounds request for row %ld (max requestable row is %d).",
                     (long)row, args_.rows_size());
        return r;
      }
      u.push_back(args_.rows((int)row));
    }
    return r;
  }
  Result get_num_downstream_rows(i64 n, i64& out) const override {
    out = 0;
    for (i64 r : args_.rows()) {
      if (r >= n) break;
      ++out;
    }
    return ok();
  }
  Result get_downstream_rows(const std::vector<i64>& u, std::vector<i64>& d,
                             std::vector<i64>& m) const override {
    for (size_t i = 0; i < u.size(); ++i) {
      auto it = index_.find(u[i]);
      if (it != index_.end()) {
        d.push_back(it->second);
        m.push_back((i64)i);
      }
    }
    return ok();
  }

 private:
  Result valid_;
  GatherSamplerArgs args_;
  std::map<i64, i64> index_;
};

// "SpaceNull" / "SpaceRepeat" (reference sampler.cpp:337-454): every upstream row becomes
// `spacing` downstream rows: the element followed by nulls (Null) or repeats (Repeat).
class SpaceSampler : public DomainSampler {
 public:
  SpaceSampler(const std::vector<u8>& args, bool repeat) : repeat_(repeat) {
    valid_.set_success(true);
    SpaceNullSamplerArgs a;  // both messages are {int64 spacing = 1}
    if (!a.ParseFromArray(args.data(), (int)args.size())) {
      RESULT_ERROR(&valid_, "Space sampler provided with invalid protobuf args");
      return;
    }
    spacing_ = a.spacing();
    if (spacing_ <= 0) RESULT_ERROR(&valid_, "Space sampler spacing (%ld) must be greater than zero", (long)spacing_);
  }
  Result validate() const override { return valid_; }
  Result get_upstream_rows(const std::vector<i64>& d, std::vector<i64>& u) const override {
    std::set<i64> req;
    for (i64 r : d) req.insert(r / spacing_);
    u.assign(req.begin(), req.end());
    return ok();
  }
  Result get_num_downstream_rows(i64 n, i64& out) const override {
    out = n * spacing_;
    return ok();
  }
  Result get_downstream_rows(const std::vector<i64>& u, std::vector<i64>& d,
                             std::vector<i64>& m) const override {
    for (size_t i = 0; i < u.size(); ++i) {
      const i64 base = u[i] * spacing_;
      d.push_back(base);
      m.push_back((i64)i);
      for (i64 k = 1; k < spacing_; ++k) {
        d.push_back(base + k);
        m.push_back(repeat_ ? (i64)i : -1);
      }
    }
    return ok();
  }

 private:
  Result valid_;
  bool repeat_;
  i64 spacing_ = 1;
};

}  // namespace

Result make_domain_sampler(const std::string& name, const std::vector<u8>& args,
                           std::unique_ptr<DomainSampler>& out) {
  if (name == "All") out.reset(new AllSampler());
  else if (name == "Strided") out.reset(new StridedSampler(args));
  else if (name == "StridedRanges" || name == "StridedRange") out.reset(new StridedRangesSampler(args));
  else if (name == "Gather") out.reset(new GatherSampler(args));
  else if (name == "SpaceNull") out.reset(new SpaceSampler(args, false));
  else if (name == "SpaceRepeat") out.reset(new SpaceSampler(args, true));
  else {
    Result r;
    RESULT_ERROR(&r, "DomainSampler %s not found.", name.c_str());
    return r;
  }
  return out->validate();
}

// ---------------------------------------------------------------------------------------------
Result Graph::analyze(GraphAnalysis& an) {
  Result r;
  const size_t n = ops.size();
  an = GraphAnalysis();
  an.batch.assign(n, 1);
  an.stencil.assign(n, {0});
  an.warmup.assign(n, 0);
  an.bounded.assign(n, false);
  an.unbounded.assign(n, false);
  an.last_use.resize(n);
  an.input_col_index.resize(n);
  size_t n_sources = 0, n_sinks = 0;

  for (size_t i = 0; i < n; ++i) {
    GraphOp& op = ops[i];
    switch (op.kind) {
      case OpKind::Source:
        ++n_sources;
        if (!op.inputs.empty()) {
          RESULT_ERROR(&r, "Source op %zu cannot have inputs", i);
          return r;
        }
        if (op.output_columns.empty()) op.output_columns = {op.column_type == proto::Video ? "frame" : "column"};
        break;
      case OpKind::Sample:
      case OpKind::Space:
      case OpKind::Sink:
        if (op.inputs.size() != 1) {
          RESULT_ERROR(&r, "%s op %zu must have exactly one input", op.name.c_str(), i);
          return r;
        }
        if (op.kind == OpKind::Sink) ++n_sinks;
        else op.output_columns = {op.inputs[0].column};
        break;
      case OpKind::Kernel: {
        const OpInfo* info = get_op_registry()->get_op_info(op.name);
        if (!info) {
          RESULT_ERROR(&r, "Op %s is not registered.", op.name.c_str());
          return r;
        }
        const KernelFactory* kf = get_kernel_registry()->get_kernel(op.name, op.device_type);
        if (!kf) {
          RESULT_ERROR(&r, "Op %s at index %zu requested kernel with device type %s but no such kernel exists.",
                       op.name.c_str(), i, op.device_type == proto::GPU ? "GPU" : "CPU");
          return r;
        }
        if (!info->variadic_inputs && op.inputs.size() != info->input_columns.size()) {
          RESULT_ERROR(&r, "Op %s at index %zu expects %zu input columns, but received %zu", op.name.c_str(), i,
                       info->input_columns.size(), op.inputs.size());
          return r;
        }
        if (op.inputs.empty()) {
          RESULT_ERROR(&r, "Op %s at index %zu has no inputs", op.name.c_str(), i);
          return r;
        }
        op.output_columns.clear();
        for (auto& c : info->output_columns) op.output_columns.push_back(c.name);
        // defaults (reference populate_analysis_info :1014-1041)
        if (!op.stencil.empty()) {
          if (!info->can_stencil) {
            RESULT_ERROR(&r, "Op %s at index %zu specified stencil but that Op was not declared to support stenciling. "
                             "Add .stencil() to the Op declaration to support stenciling.", op.name.c_str(), i);
            return r;
          }
          an.stencil[i] = op.stencil;
        } else {
          an.stencil[i] = info->preferred_stencil;
        }
        std::sort(an.stencil[i].begin(), an.stencil[i].end());
        if (op.batch != -1) {
          if (!kf->can_batch && op.batch > 1) {
            RESULT_ERROR(&r, "Op %s at index %zu specified a batch size but the Kernel for that Op was not declared to "
                             "support batching. Add .batch() to the Kernel declaration to support batching.",
                         op.name.c_str(), i);
            return r;
          }
          an.batch[i] = std::max(1, op.batch);
        } else {
          an.batch[i] = kf->can_batch ? std::max(1, kf->preferred_batch_size) : 1;
        }
        an.bounded[i] = info->has_bounded_state;
        an.unbounded[i] = info->has_unbounded_state;
        an.warmup[i] = op.warmup != -1 ? op.warmup : info->warmup;
        break;
      }
    }
    // inputs must reference earlier ops and existing columns
    for (const OpInput& in : op.inputs) {
      if (in.op_index < 0 || (size_t)in.op_index >= i) {
        RESULT_ERROR(&r, "Op %s (%zu) referenced input index %d. Ops must be specified in topo sort order.",
                     op.name.c_str(), i, in.op_index);
        return r;
      }
      const GraphOp& prod = ops[in.op_index];
      if (prod.kind == OpKind::Sink) {
        RESULT_ERROR(&r, "Op %s (%zu) reads from a Sink", op.name.c_str(), i);
        return r;
      }
      auto it = std::find(prod.output_columns.begin(), prod.output_columns.end(), in.column);
      if (it == prod.output_columns.end()) {
        RESULT_ERROR(&r, "Op %s at index %zu requested column %s from input Op %s at index %d but that Op does not "
                         "have the requested column.", op.name.c_str(), i, in.column.c_str(), prod.name.c_str(),
                     in.op_index);
        return r;
      }
      an.input_col_index[i].push_back((i32)(it - prod.output_columns.begin()));
    }
  }
  if (n_sources == 0 || n_sinks == 0) {
    RESULT_ERROR(&r, "A graph needs at least one Source and one Sink (found %zu / %zu)", n_sources, n_sinks);
    return r;
  }
  // liveness: last consumer of every produced column
  for (size_t i = 0; i < n; ++i) an.last_use[i].assign(ops[i].output_columns.size(), -1);
  for (size_t i = 0; i < n; ++i)
    for (size_t k = 0; k < ops[i].inputs.size(); ++k) {
      i32& lu = an.last_use[ops[i].inputs[k].op_index][an.input_col_index[i][k]];
      lu = std::max(lu, (i32)i);
    }
  r.set_success(true);
  return r;
}

bool Graph::consumers_accept_layout(i32 source_op, FrameLayout layout) const {
  std::vector<i32> work = {source_op};
  std::set<i32> seen;
  bool any = false;
  while (!work.empty()) {
    const i32 producer = work.back();
    work.pop_back();
    if (!seen.insert(producer).second) continue;
    for (size_t i = 0; i < ops.size(); ++i) {
      const GraphOp& op = ops[i];
      for (size_t k = 0; k < op.inputs.size(); ++k) {
        if (op.inputs[k].op_index != producer) continue;
        if (op.kind == OpKind::Sample || op.kind == OpKind::Space) {
          work.push_back((i32)i);
          continue;
        }
        if (op.kind != OpKind::Kernel) return false;  // a sink stores the frames: keep them RGB24
        const OpInfo* info = get_op_registry()->get_op_info(op.name);
        const KernelFactory* kf = get_kernel_registry()->get_kernel(op.name, op.device_type);
        if (!info || !kf || info->variadic_inputs || k >= info->input_columns.size()) return false;
        auto it = kf->input_layouts.find(info->input_columns[k].name);
        if (it == kf->input_layouts.end() || it->second != layout) return false;
        any = true;
      }
    }
  }
  return any;
}

Result Graph::domain_sizes(const JobParams& job, std::vector<i64>& rows) const {
  Result r;
  rows.assign(ops.size(), 0);
  for (size_t i = 0; i < ops.size(); ++i) {
    const GraphOp& op = ops[i];
    if (op.kind == OpKind::Source) {
      auto it = job.source_rows.find((i32)i);
      if (it == job.source_rows.end()) {
        RESULT_ERROR(&r, "Job does not bind source op %zu", i);
        return r;
      }
      rows[i] = it->second;
      continue;
    }
    const i64 in_rows = rows[op.inputs[0].op_index];
    if (op.kind == OpKind::Sample || op.kind == OpKind::Space) {
      auto it = job.samplers.find((i32)i);
      if (it == job.samplers.end()) {
        RESULT_ERROR(&r, "Job does not provide sampling args for op %zu", i);
        return r;
      }
      std::unique_ptr<DomainSampler> s;
      Result sr = make_domain_sampler(it->second.first, it->second.second, s);
      if (!sr.success()) return sr;
      sr = s->get_num_downstream_rows(in_rows, rows[i]);
      if (!sr.success()) return sr;
    } else {
      for (const OpInput& in : op.inputs)
        if (rows[in.op_index] != in_rows) {
          RESULT_ERROR(&r, "Op %s (%zu) has inputs with different numbers of rows (%ld vs %ld)", op.name.c_str(), i,
                       (long)in_rows, (long)rows[in.op_index]);
          return r;
        }
      rows[i] = in_rows;
    }
  }
  // all sinks of a job must receive the same number of rows (reference dag_analysis.cpp:636-644)
  i64 sink_rows = -1;
  for (size_t i = 0; i < ops.size(); ++i)
    if (ops[i].kind == OpKind::Sink) {
      if (sink_rows != -1 && rows[i] != sink_rows) {
        RESULT_ERROR(&r, "Sinks of one job receive different numbers of rows (%ld vs %ld)", (long)sink_rows,
                     (long)rows[i]);
        return r;
      }
      sink_rows = rows[i];
    }
  r.set_success(true);
  return r;
}

Result Graph::derive_task_streams(const GraphAnalysis& an, const JobParams& job,
                                  const std::vector<i64>& rows_per_op,
                                  const std::vector<i64>& output_rows,
                                  std::vector<TaskStream>& streams) const {
  Result r;
  const size_t n = ops.size();
  streams.assign(n, TaskStream());
  std::vector<std::set<i64>> required_out(n);
  for (size_t i = 0; i < n; ++i)
    if (ops[i].kind == OpKind::Sink) required_out[i].insert(output_rows.begin(), output_rows.end());

  for (size_t idx = n; idx-- > 0;) {
    const GraphOp& op = ops[idx];
    std::vector<i64> downstream(required_out[idx].begin(), required_out[idx].end());
    std::vector<i64> new_rows, compute_rows;
    switch (op.kind) {
      case OpKind::Source:
      case OpKind::Sink:
        new_rows = downstream;
        break;
      case OpKind::Sample:
      case OpKind::Space: {
        std::unique_ptr<DomainSampler> s;
        const auto& sa = job.samplers.at((i32)idx);
        Result sr = make_domain_sampler(sa.first, sa.second, s);
        if (!sr.success()) return sr;
        sr = s->get_upstream_rows(downstream, new_rows);
        if (!sr.success()) return sr;
        std::sort(new_rows.begin(), new_rows.end());
        new_rows.erase(std::unique(new_rows.begin(), new_rows.end()), new_rows.end());
        break;
      }
      case OpKind::Kernel: {
        std::set<i64> current;
        if (an.bounded[idx]) {
          // warmup predecessors are computed (and dropped later): reference :1608-1620
          for (i64 row : downstream)
            for (i64 w = 0; w <= an.warmup[idx]; ++w)
              if (row - w >= 0) current.insert(row - w);
        } else if (an.unbounded[idx]) {
          if (!downstream.empty())
            for (i64 row = 0; row <= downstream.back(); ++row) current.insert(row);  // :1622-1626
        } else {
          current.insert(downstream.begin(), downstream.end());
        }
        compute_rows.assign(current.begin(), current.end());
        std::set<i64> stencil_rows;
        const i64 domain = rows_per_op[op.inputs[0].op_index];
        for (i64 row : current)
          for (i32 s : an.stencil[idx]) {
            const i64 q = row + s;
            if (q >= 0 && q < domain) stencil_rows.insert(q);  // boundary: clip to the domain (:1653-1657)
          }
        new_rows.assign(stencil_rows.begin(), stencil_rows.end());
        break;
      }
    }
    if (compute_rows.empty()) compute_rows = new_rows;
    for (const OpInput& in : op.inputs) required_out[in.op_index].insert(new_rows.begin(), new_rows.end());
    streams[idx].valid_input_rows = std::move(new_rows);
    streams[idx].compute_input_rows = std::move(compute_rows);
    streams[idx].valid_output_rows = std::move(downstream);
  }
  r.set_success(true);
  return r;
}

}  // namespace internal
}  // namespace scanner
