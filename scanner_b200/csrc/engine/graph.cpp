// graph.cpp -- samplers, DAG analysis and per-task row derivation (see graph.h for the map to
// the reference's dag_analysis.cpp / sampler.cpp).
#include "graph.h"

#include <algorithm>
#include <cmath>
#include <cstring>

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
    // the evaluate stage walks output rows with an ascending cursor and `index_` maps an upstream
    // row to ONE downstream row: rows must be strictly ascending (what the Gather partitioner also
    // demands); say so here instead of failing later with "Not enough rows in argument 0"
    i64 off = 0, prev = -1;
    for (i64 r : args_.rows()) {
      if (r < 0 || r <= prev) {
        RESULT_ERROR(&valid_, "Gather sampler rows must be non-negative and strictly ascending (row %ld follows %ld "
                              "at position %ld)", (long)r, (long)prev, (long)off);
        return;
      }
      prev = r;
      index_[r] = off++;
    }
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

// Unslice inside one task: downstream (concatenated) rows [base, base + count) <-> the group's
// local rows [0, count).  args: two little-endian i64 (base, count).  Internal.
class SliceOffsetSampler : public DomainSampler {
 public:
  explicit SliceOffsetSampler(const std::vector<u8>& args) {
    valid_.set_success(args.size() == 16);
    if (args.size() == 16) {
      memcpy(&base_, args.data(), 8);
      memcpy(&count_, args.data() + 8, 8);
    } else {
      valid_.set_msg("SliceOffset sampler needs 16 bytes of arguments");
    }
  }
  Result validate() const override { return valid_; }
  Result get_upstream_rows(const std::vector<i64>& d, std::vector<i64>& u) const override {
    Result r = ok();
    for (i64 row : d) {
      if (row < base_ || row >= base_ + count_) {
        RESULT_ERROR(&r, "row %ld is outside slice group [%ld, %ld)", (long)row, (long)base_, (long)(base_ + count_));
        return r;
      }
      u.push_back(row - base_);
    }
    return r;
  }
  Result get_num_downstream_rows(i64 n, i64& out) const override {
    out = base_ + std::min(n, count_);
    return ok();
  }
  Result get_downstream_rows(const std::vector<i64>& u, std::vector<i64>& d, std::vector<i64>& m) const override {
    for (size_t i = 0; i < u.size(); ++i)
      if (u[i] >= 0 && u[i] < count_) {
        d.push_back(u[i] + base_);
        m.push_back((i64)i);
      }
    return ok();
  }

 private:
  Result valid_;
  i64 base_ = 0, count_ = 0;
};

}  // namespace

Result make_partition_groups(const std::string& name, const std::vector<u8>& args, i64 num_rows,
                             std::vector<std::vector<i64>>& groups) {
  Result r = ok();
  groups.clear();
  if (name == "Strided") {  // reference sampler.cpp:505-583
    StridedPartitionerArgs a;
    if (!a.ParseFromArray(args.data(), (int)args.size())) {
      RESULT_ERROR(&r, "Strided partitioner provided with invalid protobuf args");
      return r;
    }
    if (a.stride() <= 0 || a.group_size() <= 0) {
      RESULT_ERROR(&r, "Strided partitioner stride (%ld) and group size (%ld) must be greater than 0", (long)a.stride(),
                   (long)a.group_size());
      return r;
    }
    const i64 strided = (num_rows + a.stride() - 1) / a.stride();
    for (i64 s = 0; s < strided; s += a.group_size()) {
      groups.emplace_back();
      for (i64 i = s; i < std::min(strided, s + a.group_size()); ++i) groups.back().push_back(i * a.stride());
    }
  } else if (name == "StridedRange" || name == "StridedRanges") {  // :585-700
    StridedRangePartitionerArgs a;
    if (!a.ParseFromArray(args.data(), (int)args.size())) {
      RESULT_ERROR(&r, "StridedRange partitioner provided with invalid protobuf args");
      return r;
    }
    if (a.stride() <= 0) {
      RESULT_ERROR(&r, "StridedRange stride (%ld) must be greater than zero", (long)a.stride());
      return r;
    }
    if (a.starts_size() != a.ends_size()) {
      RESULT_ERROR(&r, "StridedRange starts and ends not the same size");
      return r;
    }
    for (int i = 0; i < a.starts_size(); ++i) {
      if (a.starts(i) > a.ends(i)) {
        RESULT_ERROR(&r, "StridedRange start (%ld) should not be after end (%ld)", (long)a.starts(i), (long)a.ends(i));
        return r;
      }
      if (a.ends(i) > num_rows) {
        RESULT_ERROR(&r, "StridedRange end (%ld) should be less than table num rows (%ld)", (long)a.ends(i),
                     (long)num_rows);
        return r;
      }
      groups.emplace_back();
      for (i64 row = a.starts(i); row < a.ends(i); row += a.stride()) groups.back().push_back(row);
    }
  } else if (name == "Gather") {  // :702-770
    GatherPartitionerArgs a;
    if (!a.ParseFromArray(args.data(), (int)args.size())) {
      RESULT_ERROR(&r, "Gather partitioner provided with invalid protobuf args");
      return r;
    }
    for (const GatherList& g : a.groups()) {
      groups.emplace_back(g.rows().begin(), g.rows().end());
      i64 prev = -1;
      for (i64 row : groups.back()) {
        if (row < 0 || row >= num_rows) {
          RESULT_ERROR(&r, "Gather partitioner row %ld is outside [0, %ld)", (long)row, (long)num_rows);
          return r;
        }
        if (row <= prev) {  // rows flow through the pipeline in table order
          RESULT_ERROR(&r, "Gather partitioner rows of one group must be ascending (%ld after %ld)", (long)row,
                       (long)prev);
          return r;
        }
        prev = row;
      }
    }
  } else {
    RESULT_ERROR(&r, "Partitioner %s not found.", name.c_str());
    return r;
  }
  if (groups.empty()) RESULT_ERROR(&r, "Partitioner %s produced no slice groups", name.c_str());
  return r;
}

Result make_domain_sampler(const std::string& name, const std::vector<u8>& args,
                           std::unique_ptr<DomainSampler>& out) {
  if (name == "All") out.reset(new AllSampler());
  else if (name == "SliceOffset") out.reset(new SliceOffsetSampler(args));
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
proto::ColumnType Graph::column_type_of(i32 op_index, const std::string& column) const {
  for (size_t hops = 0; hops <= ops.size(); ++hops) {
    if (op_index < 0 || (size_t)op_index >= ops.size()) return proto::Bytes;
    const GraphOp& op = ops[(size_t)op_index];
    if (op.kind == OpKind::Source) return op.column_type;
    if (op.kind == OpKind::Kernel) {
      const OpInfo* info = get_op_registry()->get_op_info(op.name);
      if (info)
        for (const ColumnDesc& c : info->output_columns)
          if (c.name == column) return c.type;
      return proto::Bytes;
    }
    if (op.inputs.empty()) return proto::Bytes;
    op_index = op.inputs[0].op_index;  // Sample / Space / Sink: one input, same column
  }
  return proto::Bytes;
}

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
  // slice levels (reference dag_analysis.cpp:94-271): one level, inputs of an op share it, an
  // Unslice only feeds sinks, sinks are unsliced
  slice_level.assign(n, 0);
  for (size_t i = 0; i < n; ++i) {
    const GraphOp& op = ops[i];
    if (op.inputs.empty()) continue;
    const i32 in_level = slice_level[op.inputs[0].op_index];
    for (const OpInput& in : op.inputs) {
      if (slice_level[in.op_index] != in_level) {
        RESULT_ERROR(&r, "Input Op %s (%d) specified as input to %s Op (%zu) has a different slice level. Ops within "
                         "a slice level should only receive inputs from other Ops at the same slice level.",
                     ops[in.op_index].name.c_str(), in.op_index, op.name.c_str(), i);
        return r;
      }
      if (ops[in.op_index].slice_role == SliceRole::Unslice && op.kind != OpKind::Sink) {
        RESULT_ERROR(&r, "Unslice Op specified as input to %s Op. Scanner currently only supports Output Ops consuming "
                         "the results of an Unslice Op.", op.name.c_str());
        return r;
      }
    }
    i32 level = in_level;
    if (op.slice_role == SliceRole::Slice) {
      if (in_level > 0) {
        RESULT_ERROR(&r, "Nested slicing not currently supported.");
        return r;
      }
      level = 1;
    } else if (op.slice_role == SliceRole::Unslice) {
      if (in_level == 0) {
        RESULT_ERROR(&r, "Unslice received inputs that have not been sliced.");
        return r;
      }
      level = 0;
    } else if (op.kind == OpKind::Sink && in_level != 0) {
      RESULT_ERROR(&r, "Final output columns are sliced. Final outputs must be unsliced.");
      return r;
    }
    slice_level[i] = level;
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

Result Graph::domain_sizes(const JobParams& job, std::vector<i64>& rows, SliceInfo* slices) const {
  Result r;
  rows.assign(ops.size(), 0);
  SliceInfo local;
  SliceInfo& si = slices ? *slices : local;
  si = SliceInfo();
  // per-group sizes of the ops inside the slice: grows[g][op]
  std::vector<std::vector<i64>>& grows = si.rows_per_op;
  auto level = [&](size_t i) { return i < slice_level.size() ? slice_level[i] : 0; };
  auto sampler_of = [&](size_t i, i32 group, std::unique_ptr<DomainSampler>& s) -> Result {
    Result e;
    auto git = job.group_samplers.find((i32)i);
    if (group >= 0 && git != job.group_samplers.end()) {
      if (git->second.size() != 1 && (i32)git->second.size() != si.groups) {
        RESULT_ERROR(&e, "A job specified %zu samplers but there are %d slice groups for %s Op at %zu.",
                     git->second.size(), si.groups, ops[i].name.c_str(), i);
        return e;
      }
      const auto& sa = git->second[git->second.size() == 1 ? 0 : (size_t)group];
      return make_domain_sampler(sa.first, sa.second, s);
    }
    auto it = job.samplers.find((i32)i);
    if (it == job.samplers.end()) {
      RESULT_ERROR(&e, "Job does not provide sampling args for op %zu", i);
      return e;
    }
    return make_domain_sampler(it->second.first, it->second.second, s);
  };
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
    const i32 in0 = op.inputs[0].op_index;
    if (op.slice_role == SliceRole::Slice) {
      auto it = job.partitioners.find((i32)i);
      if (it == job.partitioners.end()) {
        RESULT_ERROR(&r, "Job does not provide a partitioner for Slice op %zu", i);
        return r;
      }
      std::vector<std::vector<i64>> groups;
      Result pr = make_partition_groups(it->second.first, it->second.second, rows[in0], groups);
      if (!pr.success()) return pr;
      if (si.groups != 0 && si.groups != (i32)groups.size()) {
        RESULT_ERROR(&r, "A job specified one slice with %d groups and another slice with %zu groups. Scanner "
                         "currently does not support multiple slices with different numbers of groups in the same job.",
                     si.groups, groups.size());
        return r;
      }
      if (si.groups == 0) {
        si.groups = (i32)groups.size();
        grows.assign(groups.size(), std::vector<i64>(ops.size(), 0));
      }
      i64 total = 0;
      for (size_t g = 0; g < groups.size(); ++g) {
        grows[g][i] = (i64)groups[g].size();
        total += grows[g][i];
      }
      rows[i] = total;
      si.slice_rows[(i32)i] = std::move(groups);
      continue;
    }
    if (op.slice_role == SliceRole::Unslice) {
      i64 total = 0;
      si.out_base.clear();
      for (i32 g = 0; g < si.groups; ++g) {
        si.out_base.push_back(total);
        total += grows[(size_t)g][in0];
      }
      si.out_base.push_back(total);
      rows[i] = total;
      continue;
    }
    if (level(i) > 0) {  // inside the slice: every group on its own
      i64 total = 0;
      for (i32 g = 0; g < si.groups; ++g) {
        const i64 in_rows = grows[(size_t)g][in0];
        i64 out = in_rows;
        if (op.kind == OpKind::Sample || op.kind == OpKind::Space) {
          std::unique_ptr<DomainSampler> s;
          Result sr = sampler_of(i, g, s);
          if (!sr.success()) return sr;
          sr = s->get_num_downstream_rows(in_rows, out);
          if (!sr.success()) return sr;
        } else {
          for (const OpInput& in : op.inputs)
            if (grows[(size_t)g][in.op_index] != in_rows) {
              RESULT_ERROR(&r, "Op %s (%zu) has inputs with different numbers of rows in slice group %d (%ld vs %ld)",
                           op.name.c_str(), i, g, (long)in_rows, (long)grows[(size_t)g][in.op_index]);
              return r;
            }
        }
        grows[(size_t)g][i] = out;
        total += out;
      }
      rows[i] = total;
      continue;
    }
    const i64 in_rows = rows[in0];
    if (op.kind == OpKind::Sample || op.kind == OpKind::Space) {
      std::unique_ptr<DomainSampler> s;
      Result sr = sampler_of(i, -1, s);
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
  // level-0 ops of a sliced job keep their global sizes in every group's view
  for (i32 g = 0; g < si.groups; ++g)
    for (size_t i = 0; i < ops.size(); ++i)
      if (level(i) == 0 && ops[i].slice_role != SliceRole::Slice) grows[(size_t)g][i] = rows[i];
  r.set_success(true);
  return r;
}

void Graph::slice_view(const JobParams& job, const SliceInfo& si, i32 group, JobParams& view, std::vector<i64>& vrows) const {
  view = job;
  vrows = si.rows_per_op[(size_t)group];
  for (size_t i = 0; i < ops.size(); ++i) {
    const GraphOp& op = ops[i];
    if (op.slice_role == SliceRole::Slice) {
      GatherSamplerArgs ga;  // the group's rows, as a Gather over the upstream domain
      for (i64 row : si.slice_rows.at((i32)i)[(size_t)group]) ga.add_rows(row);
      const std::string bytes = ga.SerializeAsString();
      view.samplers[(i32)i] = {"Gather", std::vector<u8>(bytes.begin(), bytes.end())};
    } else if (op.slice_role == SliceRole::Unslice) {
      const i64 base = si.out_base[(size_t)group], count = si.out_base[(size_t)group + 1] - base;
      std::vector<u8> a(16);
      memcpy(a.data(), &base, 8);
      memcpy(a.data() + 8, &count, 8);
      view.samplers[(i32)i] = {"SliceOffset", a};
    } else {
      auto gs = job.group_samplers.find((i32)i);
      if (gs != job.group_samplers.end() && !gs->second.empty())
        view.samplers[(i32)i] = gs->second[gs->second.size() == 1 ? 0 : (size_t)group];
      auto ga = job.group_stream_args.find((i32)i);
      if (ga != job.group_stream_args.end() && !ga->second.empty())
        view.stream_args[(i32)i] = ga->second[ga->second.size() == 1 ? 0 : (size_t)group];
    }
  }
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
