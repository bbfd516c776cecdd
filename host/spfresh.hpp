// spfresh.hpp — C++ host layer above the C ABI of libspfresh_b200.so (include/spfresh_b200.h).
//
// The reference (jairad26/spfresh) is a Rust crate and this image has no Rust toolchain, so the
// host side of the hot path is written in C++ against the reference's own interface: the same
// types, method names, argument meaning, control flow and error behaviour as
//   src/distances/distance.rs            DistanceMetric + the three metrics
//   src/clustering/hierarchical.rs       InitializationMethod, ClusteringParams, Cluster,
//                                        HierarchicalClustering::{fit, labels, ...}
//   src/spann/config.rs                  Config::{from_file, validate, to_clustering_params}
//   src/spann/spann_builder.rs           SpannIndexBuilder::{new, with_data, build, load}
//   src/spann/spann_index.rs             SpannIndex::{find_k_nearest_neighbor_spann, ...}
//   src/spann/posting_lists.rs           PointData
//   src/spann/lire/operations.rs         Split, Reassign (the parts that are distance work)
// Every distance, mean, argmin and scan runs on the B200 through the C ABI; this header only
// sequences the calls (fit(), the bisect work-list) and owns the random decisions.  Rust
// `Result::Err` / `expect` / `unwrap` panics become `spfresh::Error`.  Header-only, C++17.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <functional>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/spfresh_b200.h"

namespace spfresh {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void check(int rc) {
  if (rc < 0) throw Error(std::string("spfresh_b200: ") + spf_last_error());
}

// Borrowed views, like ndarray's ArrayView1 / ArrayView2 (row-major, `stride` elements per row).
struct ArrayView1 {
  const float* data = nullptr;
  size_t len = 0;
};
struct ArrayView2 {
  const float* data = nullptr;
  size_t rows = 0, cols = 0, stride = 0;
  ArrayView2() = default;
  ArrayView2(const float* p, size_t r, size_t c, size_t s = 0) : data(p), rows(r), cols(c), stride(s ? s : c) {}
  ArrayView1 row(size_t i) const { return ArrayView1{data + i * stride, cols}; }
};

// One CUDA device + stream (spf_ctx).  Shared by the objects built on it.
class Context {
 public:
  explicit Context(int device = 0) { check(spf_ctx_create(device, &h_)); }
  ~Context() { if (h_) spf_ctx_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  spf_ctx* handle() const { return h_; }
  static std::shared_ptr<Context> shared(int device = 0) { return std::make_shared<Context>(device); }

 private:
  spf_ctx* h_ = nullptr;
};

// The rows of the borrowed ArrayView2, resident in HBM for the lifetime of this object.
class DeviceDataset {
 public:
  DeviceDataset(std::shared_ptr<Context> ctx, ArrayView2 rows) : ctx_(std::move(ctx)), n_(rows.rows), d_(rows.cols) {
    check(spf_dataset_upload(ctx_->handle(), rows.data, rows.rows, (uint32_t)rows.cols, rows.stride, &h_));
  }
  ~DeviceDataset() { if (h_) spf_dataset_free(h_); }
  DeviceDataset(const DeviceDataset&) = delete;
  DeviceDataset& operator=(const DeviceDataset&) = delete;
  spf_dataset* handle() const { return h_; }
  size_t rows() const { return n_; }
  size_t dim() const { return d_; }
  const std::shared_ptr<Context>& ctx() const { return ctx_; }

 private:
  std::shared_ptr<Context> ctx_;
  spf_dataset* h_ = nullptr;
  size_t n_, d_;
};

// ---------------------------------------------------------------------------------------------
// src/distances/distance.rs
// ---------------------------------------------------------------------------------------------
namespace distances {

// distance.rs:7-10 plus the routing hint SURVEY.md §8(b) adds to the trait (`kind`), so batched
// callers can tell the device which metric this object is.
class DistanceMetric {
 public:
  virtual ~DistanceMetric() = default;
  virtual int kind() const = 0;
  virtual const char* name() const = 0;
  // Per-pair seam, kept for API compatibility (evaluated on the device).  ndarray-stats returns
  // Err on empty input or a shape mismatch and the reference unwraps it (:19,30,41) → Error.
  float compute(const Context& ctx, ArrayView1 a, ArrayView1 b) const {
    if (a.len != b.len || a.len == 0)
      throw Error("called `Result::unwrap()` on an `Err` value: shape mismatch or empty input");
    float out = 0.f;
    check(spf_distance_pairs(ctx.handle(), kind(), a.data, b.data, (uint32_t)a.len, 1, &out));
    return out;
  }
};
struct SquaredEuclideanDistance : DistanceMetric {   // distance.rs:14-21
  int kind() const override { return SPF_METRIC_EUCLIDEAN; }
  const char* name() const override { return "Euclidean"; }
};
struct ManhattanDistance : DistanceMetric {          // distance.rs:25-32
  int kind() const override { return SPF_METRIC_MANHATTAN; }
  const char* name() const override { return "Manhattan"; }
};
struct ChebyshevDistance : DistanceMetric {          // distance.rs:36-43
  int kind() const override { return SPF_METRIC_CHEBYSHEV; }
  const char* name() const override { return "Chebyshev"; }
};

}  // namespace distances

// ---------------------------------------------------------------------------------------------
// src/clustering/hierarchical.rs
// ---------------------------------------------------------------------------------------------
namespace clustering {

constexpr float BOUNDARY_THRESHOLD = 1.1f;   // hierarchical.rs:55
constexpr uint32_t KMPP_BATCH = 256;          // k-means++ rounds per spf_kmpp_rounds call (one host sync per batch)

enum class InitializationMethod { Random, KMeansPlusPlus };   // hierarchical.rs:13-16

// The random decisions the reference takes from rand::SmallRng, in its order.  rand 0.9's stream
// is not restated: production callers plug their own generator in, tests script
// the decisions.
class RandomSource {
 public:
  virtual ~RandomSource() = default;
  virtual std::vector<uint64_t> choose_multiple(uint64_t n, uint64_t k) = 0;   // (0..n).choose_multiple(rng, k) :204
  virtual uint64_t choose_index(uint64_t n) = 0;                               // (0..n).choose / slice.choose :111,253
  virtual double uniform01() = 0;                                              // the draw behind choose_weighted :285
};

// Default generator: SplitMix64 (the seeding function of SmallRng::seed_from_u64; the streams
// themselves are NOT those of rand 0.9).
class SplitMixRandomSource : public RandomSource {
 public:
  explicit SplitMixRandomSource(uint64_t seed = 0x9e3779b97f4a7c15ull) : s_(seed) {}
  uint64_t next() {
    uint64_t z = (s_ += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  uint64_t choose_index(uint64_t n) override { return n ? next() % n : 0; }
  double uniform01() override { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  std::vector<uint64_t> choose_multiple(uint64_t n, uint64_t k) override {   // reservoir sampling
    if (k > n) k = n;
    std::vector<uint64_t> r(k);
    for (uint64_t i = 0; i < k; ++i) r[i] = i;
    for (uint64_t i = k; i < n; ++i) {
      const uint64_t j = next() % (i + 1);
      if (j < k) r[j] = i;
    }
    return r;
  }

 private:
  uint64_t s_;
};

// Replays explicit decisions (tests, parity runs).
class ScriptedRandomSource : public RandomSource {
 public:
  std::vector<uint64_t> multiple;                       // answer of choose_multiple
  std::function<uint64_t(uint64_t)> index;              // answer of choose_index(n)
  std::vector<double> u01;                              // answers of uniform01, in order
  std::vector<uint64_t> choose_multiple(uint64_t, uint64_t) override { return multiple; }
  uint64_t choose_index(uint64_t n) override { return index ? index(n) : 0; }
  double uniform01() override {
    if (iu_ >= u01.size()) throw Error("ScriptedRandomSource: out of uniform draws");
    return u01[iu_++];
  }

 private:
  size_t iu_ = 0;
};

struct ClusteringParams {                              // hierarchical.rs:18-24
  std::shared_ptr<distances::DistanceMetric> distance_metric;
  InitializationMethod initialization_method = InitializationMethod::Random;
  std::optional<size_t> desired_cluster_size;
  size_t initial_k = 0;
  std::optional<uint64_t> rng_seed;
  std::shared_ptr<RandomSource> random_source;         // overrides rng_seed when set
};

struct Cluster {                                       // hierarchical.rs:26-41
  std::optional<size_t> centroid_idx;
  std::vector<uint64_t> points;
  size_t depth = 0;
  Cluster() = default;
  Cluster(size_t c, std::vector<uint64_t> p, size_t dep) : centroid_idx(c), points(std::move(p)), depth(dep) {}
};

// hierarchical.rs:43-391 with the batched seams routed to the GPU.
class HierarchicalClustering {
 public:
  std::vector<Cluster> clusters;
  size_t max_splits = 1000000;   // the reference loops forever on duplicate-heavy clusters (:74-105)

  HierarchicalClustering(ClusteringParams params, ArrayView2 data, std::shared_ptr<Context> ctx = nullptr)
      : params_(std::move(params)), data_(data) {
    if (!ctx) ctx = Context::shared();
    dataset_ = std::make_shared<DeviceDataset>(std::move(ctx), data);
  }
  ~HierarchicalClustering() { drop_last_assign(); }

  const std::shared_ptr<DeviceDataset>& dataset() const { return dataset_; }

  void fit() {                                         // hierarchical.rs:65-71
    initialize_clusters(params_.initial_k);
    assign_points();
    update_centroids();
    subdivide_clusters();
  }

  void initialize_clusters(size_t k) {                 // hierarchical.rs:192-197
    if (params_.initialization_method == InitializationMethod::Random) initialize_clusters_randomly(k);
    else initialize_clusters_kmeans_plus_plus(k);
  }

  void initialize_clusters_randomly(size_t k) {        // hierarchical.rs:200-210 (host side, RNG only)
    for (uint64_t i : get_rng()->choose_multiple(dataset_->rows(), k)) clusters.emplace_back((size_t)i, std::vector<uint64_t>(), 0);
  }

  void initialize_clusters_kmeans_plus_plus(size_t k) {   // hierarchical.rs:249-293
    auto rng = get_rng();
    const uint64_t n = dataset_->rows();
    const uint64_t first = rng->choose_index(n);                              // :253-255
    clusters.emplace_back((size_t)first, std::vector<uint64_t>(), 0);
    spf_kmpp* s = nullptr;
    check(spf_kmpp_begin(dataset_->handle(), metric(), first, &s));
    try {
      std::vector<double> draws;                                              // drawn ahead, not yet used
      std::vector<uint64_t> rows(KMPP_BATCH);
      size_t left = k > 0 ? k - 1 : 0;
      while (left > 0) {                                                      // :259
        const size_t want = std::min(left, (size_t)KMPP_BATCH);
        while (draws.size() < want) draws.push_back(rng->uniform01());
        uint32_t done = 0;
        const int rc = spf_kmpp_rounds(s, draws.data(), (uint32_t)want, rows.data(), &done);   // :260-286 on the device
        check(rc);
        for (uint32_t i = 0; i < done; ++i) clusters.emplace_back((size_t)rows[i], std::vector<uint64_t>(), 0);
        draws.erase(draws.begin(), draws.begin() + done + (rc > 0 ? 1 : 0));  // the failing round consumed its draw too
        left -= done;
        if (rc > 0) {                                                         // :287-290 uniform fallback
          const uint64_t chosen = rng->choose_index(n);
          check(spf_kmpp_push(s, chosen));
          clusters.emplace_back((size_t)chosen, std::vector<uint64_t>(), 0);
          left -= 1;
        }
      }
    } catch (...) {
      spf_kmpp_free(s);
      throw;
    }
    spf_kmpp_free(s);
  }

  // hierarchical.rs:295-364.  centroids: (row_idx, depth) pairs; returns one member list per centroid.
  std::vector<std::vector<uint64_t>> assign_points_to_clusters(const std::vector<uint64_t>& point_indices,
                                                               const std::vector<std::pair<size_t, size_t>>& centroids) {
    std::vector<uint64_t> rows(centroids.size());
    for (size_t i = 0; i < centroids.size(); ++i) rows[i] = centroids[i].first;
    spf_assign_result* r = nullptr;
    check(spf_assign(dataset_->handle(), metric(), point_indices.data(), point_indices.size(), rows.data(),
                     (uint32_t)rows.size(), BOUNDARY_THRESHOLD, SPF_ASSIGN_DEFAULT, &r));
    auto lists = fetch_lists(r);
    spf_assign_free(r);
    return lists;
  }

  void assign_points() {                               // hierarchical.rs:368-390
    std::vector<uint64_t> rows = centroid_rows();
    drop_last_assign();
    check(spf_assign(dataset_->handle(), metric(), nullptr, dataset_->rows(), rows.data(), (uint32_t)rows.size(),
                     BOUNDARY_THRESHOLD, SPF_ASSIGN_DEFAULT, &last_assign_));   // kept on the device for update_centroids
    auto lists = fetch_lists(last_assign_);
    for (size_t i = 0; i < clusters.size(); ++i) clusters[i].points = std::move(lists[i]);
  }

  void update_centroids() {                            // hierarchical.rs:138-181
    std::vector<uint64_t> old_rows = centroid_rows(), fresh(clusters.size());
    if (last_assign_ && spf_assign_clusters(last_assign_) == clusters.size()) {
      check(spf_update_medoids_from(dataset_->handle(), metric(), last_assign_, old_rows.data(), fresh.data(), nullptr));
      drop_last_assign();
    } else {
      std::vector<uint64_t> offsets(clusters.size() + 1, 0), members;
      for (size_t i = 0; i < clusters.size(); ++i) {
        offsets[i + 1] = offsets[i] + clusters[i].points.size();
        members.insert(members.end(), clusters[i].points.begin(), clusters[i].points.end());
      }
      check(spf_update_medoids(dataset_->handle(), metric(), offsets.data(), members.data(), (uint32_t)clusters.size(),
                               old_rows.data(), fresh.data(), nullptr));
    }
    for (size_t i = 0; i < clusters.size(); ++i) clusters[i].centroid_idx = (size_t)fresh[i];
  }

  void subdivide_clusters() {                          // hierarchical.rs:74-105
    if (!params_.desired_cluster_size) throw Error("desired_cluster_size is not set");   // .unwrap() in the reference
    const size_t desired = *params_.desired_cluster_size;
    size_t i = 0, splits = 0;
    while (i < clusters.size()) {
      if (clusters[i].points.size() > desired) {
        if (++splits > max_splits)
          throw Error("subdivide_clusters: split limit reached (the reference would loop forever here)");
        const std::vector<uint64_t> pts = std::move(clusters[i].points);
        auto sub = create_subclusters(pts, clusters[i].depth + 1);
        clusters[i] = std::move(sub.first);            // :95
        clusters.push_back(std::move(sub.second));     // :98
      } else {
        ++i;
      }
    }
  }

  std::pair<Cluster, Cluster> create_subclusters(const std::vector<uint64_t>& points, size_t new_depth) {   // :107-135
    auto rng = get_rng();
    const uint64_t c1 = points[rng->choose_index(points.size())];                                   // :111
    uint64_t c2 = 0;
    check(spf_farthest(dataset_->handle(), metric(), c1, points.data(), points.size(), &c2));      // :112-126
    auto lists = assign_points_to_clusters(points, {{(size_t)c1, new_depth}, {(size_t)c2, new_depth}});   // :129
    return {Cluster((size_t)c1, std::move(lists[0]), new_depth), Cluster((size_t)c2, std::move(lists[1]), new_depth)};
  }

  // hierarchical.rs:215-246: per point, among the clusters it belongs to, the one whose centroid is
  // strictly nearest, scanning clusters in order starting from label 0.
  std::vector<size_t> labels() {
    const size_t n = dataset_->rows(), d = dataset_->dim();
    std::vector<size_t> lab(n, 0);
    const Context& ctx = *dataset_->ctx();
    std::vector<float> a, b, c, this_d, old_d;
    for (size_t ci = 0; ci < clusters.size(); ++ci) {
      const auto& pts = clusters[ci].points;
      if (pts.empty()) continue;
      a.resize(pts.size() * d); b.resize(pts.size() * d); c.resize(pts.size() * d);
      this_d.resize(pts.size()); old_d.resize(pts.size());
      for (size_t i = 0; i < pts.size(); ++i) {
        const float* x = data_.row(pts[i]).data;
        const float* cen = data_.row(*clusters[ci].centroid_idx).data;
        const float* old = data_.row(*clusters[lab[pts[i]]].centroid_idx).data;
        std::copy(x, x + d, a.begin() + i * d);
        std::copy(cen, cen + d, b.begin() + i * d);
        std::copy(old, old + d, c.begin() + i * d);
      }
      check(spf_distance_pairs(ctx.handle(), metric(), a.data(), b.data(), (uint32_t)d, pts.size(), this_d.data()));
      check(spf_distance_pairs(ctx.handle(), metric(), a.data(), c.data(), (uint32_t)d, pts.size(), old_d.data()));
      for (size_t i = 0; i < pts.size(); ++i)
        if (this_d[i] < old_d[i]) lab[pts[i]] = ci;
    }
    return lab;
  }

 private:
  int metric() const { return params_.distance_metric->kind(); }

  std::shared_ptr<RandomSource> get_rng() const {     // hierarchical.rs:184-189
    if (params_.random_source) return params_.random_source;
    // a fresh generator per call, like SmallRng::seed_from_u64(seed) in every caller
    return std::make_shared<SplitMixRandomSource>(params_.rng_seed ? *params_.rng_seed : 0x2545f4914f6cdd1dull);
  }

  std::vector<uint64_t> centroid_rows() const {
    std::vector<uint64_t> rows(clusters.size());
    for (size_t i = 0; i < clusters.size(); ++i) rows[i] = clusters[i].centroid_idx.value_or(0);
    return rows;
  }

  static std::vector<std::vector<uint64_t>> fetch_lists(const spf_assign_result* r) {
    const uint32_t k = spf_assign_clusters(r);
    std::vector<uint64_t> offsets((size_t)k + 1), members(spf_assign_total(r));
    check(spf_assign_fetch(r, nullptr, nullptr, offsets.data(), members.data()));
    std::vector<std::vector<uint64_t>> lists(k);
    for (uint32_t c = 0; c < k; ++c) lists[c].assign(members.begin() + offsets[c], members.begin() + offsets[c + 1]);
    return lists;
  }

  void drop_last_assign() {
    if (last_assign_) spf_assign_free(last_assign_);
    last_assign_ = nullptr;
  }

  ClusteringParams params_;
  ArrayView2 data_;
  std::shared_ptr<DeviceDataset> dataset_;
  spf_assign_result* last_assign_ = nullptr;
};

}  // namespace clustering

// ---------------------------------------------------------------------------------------------
// src/spann/{config,posting_lists,spann_index,spann_builder}.rs
// ---------------------------------------------------------------------------------------------
namespace spann {

struct PointData {                                     // posting_lists.rs:7-11
  uint64_t point_id = 0;
  std::vector<float> vector;
};

struct ClusteringParamsConfig {                        // config.rs:7-12
  std::string distance_metric, initialization_method;
  size_t initial_k = 0;
};

struct Config {                                        // config.rs:14-19
  ClusteringParamsConfig clustering_params;
  std::optional<std::string> data_file, output_path;

  // config.rs:52-57.  The three-level YAML of examples/example_config.yaml: `key: value` lines, the
  // clustering parameters nested one level under `clustering_params:`.
  static Config from_file(const std::string& file_path) {
    std::ifstream f(file_path);
    if (!f) throw Error("cannot open " + file_path);
    Config c;
    std::string line;
    while (std::getline(f, line)) {
      const size_t hash = line.find('#');
      if (hash != std::string::npos) line.erase(hash);
      const size_t colon = line.find(':');
      if (colon == std::string::npos) continue;
      auto trim = [](std::string s) {
        const char* ws = " \t\r\n\"'";
        const size_t b = s.find_first_not_of(ws), e = s.find_last_not_of(ws);
        return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
      };
      const std::string key = trim(line.substr(0, colon)), val = trim(line.substr(colon + 1));
      if (val.empty()) continue;                       // a section header such as `clustering_params:`
      if (key == "distance_metric") c.clustering_params.distance_metric = val;
      else if (key == "initialization_method") c.clustering_params.initialization_method = val;
      else if (key == "initial_k") c.clustering_params.initial_k = (size_t)std::stoull(val);
      else if (key == "data_file") c.data_file = val;
      else if (key == "output_path") c.output_path = val;
    }
    c.validate();
    return c;
  }

  void validate() const {                              // config.rs:59-87
    const auto& m = clustering_params.distance_metric;
    if (m != "Euclidean" && m != "Manhattan" && m != "Chebyshev") throw Error("Unsupported distance metric: " + m);
    const auto& i = clustering_params.initialization_method;
    if (i != "Random" && i != "KMeansPlusPlus") throw Error("Unsupported initialization method: " + i);
    if (clustering_params.initial_k == 0) throw Error("initial_k must be greater than 0");
  }

  clustering::ClusteringParams to_clustering_params() const {   // config.rs:90-113
    clustering::ClusteringParams p;
    const auto& m = clustering_params.distance_metric;
    if (m == "Euclidean") p.distance_metric = std::make_shared<distances::SquaredEuclideanDistance>();
    else if (m == "Manhattan") p.distance_metric = std::make_shared<distances::ManhattanDistance>();
    else if (m == "Chebyshev") p.distance_metric = std::make_shared<distances::ChebyshevDistance>();
    else throw Error("Unsupported distance metric: " + m);
    p.initialization_method = clustering_params.initialization_method == "Random"
                                  ? clustering::InitializationMethod::Random
                                  : clustering::InitializationMethod::KMeansPlusPlus;
    p.initial_k = clustering_params.initial_k;
    return p;
  }
};

// spann_index.rs:17-197.  The kd-tree over the centroids is replaced by an exact batched probe
// (same result: exact k-NN by squared L2, ascending) and the per-cluster files by lists in HBM.
class SpannIndex {
 public:
  SpannIndex(std::string posting_lists_dir, std::shared_ptr<Context> ctx)
      : dir_(std::move(posting_lists_dir)), ctx_(std::move(ctx)) {}
  ~SpannIndex() { if (idx_) spf_index_free(idx_); }
  SpannIndex(const SpannIndex&) = delete;
  SpannIndex& operator=(const SpannIndex&) = delete;
  SpannIndex(SpannIndex&& o) noexcept : dir_(std::move(o.dir_)), ctx_(std::move(o.ctx_)), idx_(o.idx_), d_(o.d_),
                                        centroids_(std::move(o.centroids_)) { o.idx_ = nullptr; }

  // spann_index.rs:56-114: pack the clusters' members into HBM-resident posting lists.
  void create_posting_lists(const DeviceDataset& ds, const std::vector<clustering::Cluster>& clusters) {
    std::vector<uint64_t> offsets(clusters.size() + 1, 0), members, rows(clusters.size());
    for (size_t i = 0; i < clusters.size(); ++i) {
      offsets[i + 1] = offsets[i] + clusters[i].points.size();
      members.insert(members.end(), clusters[i].points.begin(), clusters[i].points.end());
      rows[i] = clusters[i].centroid_idx.value_or(0);
    }
    if (idx_) { spf_index_free(idx_); idx_ = nullptr; }
    check(spf_index_pack(ds.handle(), offsets.data(), members.data(), rows.data(), (uint32_t)clusters.size(), 0,
                         (uint32_t)clusters.size(), &idx_));
    d_ = ds.dim();
    centroids_.resize(clusters.size() * d_);
    check(spf_dataset_fetch_rows(ds.handle(), rows.data(), rows.size(), centroids_.data()));
  }

  void save_posting_list() const {                     // spann_index.rs:45-53 → posting_lists.rs:108-113
    if (!idx_) throw Error("Posting list is not available");
    check(spf_index_save_dir(idx_, dir_.c_str()));
  }

  // Sidecar for the dense centroid matrix: the reference keeps centroids only inside
  // output.kdtree, kiddo's private layout.  u64 rows, u64 cols, f32 data.
  void save_centroids(const std::string& path) const {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw Error("cannot write " + path);
    const uint64_t shape[2] = {d_ ? centroids_.size() / d_ : 0, d_};
    f.write(reinterpret_cast<const char*>(shape), sizeof(shape));
    f.write(reinterpret_cast<const char*>(centroids_.data()), (std::streamsize)(centroids_.size() * sizeof(float)));
  }

  // spann_index.rs:32-43.  The dense centroid matrix of the GPU probe comes from the caller
  // (`centroids`, nlists x d in list-id order) or from the centroids.bin sidecar build() writes; the
  // reference keeps centroids only inside output.kdtree (kiddo's private layout, not read here), so a
  // directory written by the reference alone needs the caller's matrix.  The posting-list files
  // themselves are compatible in both directions.
  void load_posting_list(const std::string& path, const std::vector<float>* centroids = nullptr, size_t d = 0) {
    uint64_t shape[2] = {0, 0};
    if (centroids) {
      if (d == 0 || centroids->size() % d) throw Error("load_posting_list: centroids must be nlists x d");
      centroids_ = *centroids;
      shape[0] = centroids->size() / d;
      shape[1] = d;
    } else {
      std::ifstream f(path + "/centroids.bin", std::ios::binary);
      if (!f)
        throw Error("cannot read " + path + "/centroids.bin: the reference stores centroids only inside output.kdtree; "
                    "pass the centroid matrix to load_posting_list() or write the sidecar with save_centroids()");
      f.read(reinterpret_cast<char*>(shape), sizeof(shape));
      if (!f || shape[1] == 0 || shape[0] > (1ull << 32) || shape[1] > (1ull << 20)) throw Error("corrupt centroids.bin");
      centroids_.resize(shape[0] * shape[1]);
      f.read(reinterpret_cast<char*>(centroids_.data()), (std::streamsize)(centroids_.size() * sizeof(float)));
      if (!f) throw Error("truncated centroids.bin");
    }
    d_ = shape[1];
    if (idx_) { spf_index_free(idx_); idx_ = nullptr; }
    check(spf_index_load_dir(ctx_->handle(), path.c_str(), centroids_.data(), (uint32_t)shape[0], (uint32_t)d_, &idx_));
  }

  // Batched sibling of find_k_nearest_neighbor_spann: one optional result list per query.
  std::vector<std::optional<std::vector<PointData>>> find_k_nearest_neighbors_batch(ArrayView2 queries, size_t k,
                                                                                   size_t nprobe = 0,
                                                                                   float prune_factor = 1.2f) const {
    if (!idx_) throw Error("Posting list is not available");              // .expect() in the reference (:153-158)
    if (queries.cols != d_) throw Error("Query length mismatch");         // .expect() (:160-162)
    std::vector<float> q(queries.rows * d_);
    for (size_t i = 0; i < queries.rows; ++i) std::copy(queries.row(i).data, queries.row(i).data + d_, q.begin() + i * d_);
    std::vector<uint64_t> ids(queries.rows * k);
    std::vector<float> dists(queries.rows * k), vec(queries.rows * k * d_);
    std::vector<uint32_t> counts(queries.rows);
    check(spf_search_batch(idx_, q.data(), queries.rows, (uint32_t)k, (uint32_t)nprobe, prune_factor, ids.data(),
                           dists.data(), counts.data(), vec.data(), nullptr));
    std::vector<std::optional<std::vector<PointData>>> out(queries.rows);
    for (size_t i = 0; i < queries.rows; ++i) {
      if (counts[i] == 0) continue;                                       // None (:183-186)
      std::vector<PointData> r(counts[i]);
      for (uint32_t j = 0; j < counts[i]; ++j) {
        r[j].point_id = ids[i * k + j];
        r[j].vector.assign(vec.begin() + (i * k + j) * d_, vec.begin() + (i * k + j + 1) * d_);
      }
      out[i] = std::move(r);
    }
    return out;
  }

  std::optional<std::vector<PointData>> find_k_nearest_neighbor_spann(ArrayView1 query, size_t k) const {   // :148-197
    return find_k_nearest_neighbors_batch(ArrayView2(query.data, 1, query.len), k)[0];
  }

  spf_index* handle() const { return idx_; }
  size_t dim() const { return d_; }

 private:
  std::string dir_;
  std::shared_ptr<Context> ctx_;
  spf_index* idx_ = nullptr;
  size_t d_ = 0;
  std::vector<float> centroids_;   // dense nlists x d (stands in for the kd-tree)
};

// spann_builder.rs:8-75
class SpannIndexBuilder {
 public:
  explicit SpannIndexBuilder(Config config, std::shared_ptr<Context> ctx = nullptr,
                             std::shared_ptr<clustering::RandomSource> random_source = nullptr)
      : config_(std::move(config)), ctx_(std::move(ctx)), random_source_(std::move(random_source)) {}

  SpannIndexBuilder& with_data(ArrayView2 data) {      // spann_builder.rs:20-23
    data_ = data;
    return *this;
  }

  // spann_builder.rs:25-64; N is the const generic of the reference (the expected dimension).
  SpannIndex build(size_t N, std::vector<clustering::Cluster>* clusters_out = nullptr) {
    if (!data_) throw Error("No data provided (in-memory or file)");
    if (data_->cols != N) {
      std::ostringstream m;
      m << "Data dimension mismatch: expected " << N << ", got " << data_->cols;
      throw Error(m.str());
    }
    clustering::ClusteringParams params = config_.to_clustering_params();
    params.desired_cluster_size = (size_t)std::llround((double)data_->rows * 0.18);    // :48-49 f64::round
    params.random_source = random_source_;
    if (!ctx_) ctx_ = Context::shared();
    clustering::HierarchicalClustering hc(params, *data_, ctx_);
    hc.fit();
    if (!config_.output_path) throw Error("Output path is not specified");
    SpannIndex index(*config_.output_path, ctx_);
    index.create_posting_lists(*hc.dataset(), hc.clusters);
    try {                                              // `let _ =` in the reference: errors are dropped
      index.save_posting_list();
      index.save_centroids(*config_.output_path + "/centroids.bin");
    } catch (const Error&) {
    }
    if (clusters_out) *clusters_out = hc.clusters;
    return index;
  }

  SpannIndex load(size_t /*N*/) {                      // spann_builder.rs:66-75
    if (!config_.output_path) throw Error("Output path is not specified");
    if (!ctx_) ctx_ = Context::shared();
    SpannIndex index(*config_.output_path, ctx_);
    index.load_posting_list(*config_.output_path);
    return index;
  }

 private:
  Config config_;
  std::optional<ArrayView2> data_;
  std::shared_ptr<Context> ctx_;
  std::shared_ptr<clustering::RandomSource> random_source_;
};

}  // namespace spann

// ---------------------------------------------------------------------------------------------
// src/spann/lire/operations.rs — the two operations that sit on the hot path's kernels
// ---------------------------------------------------------------------------------------------
namespace lire {

using VectorList = std::vector<std::pair<size_t, std::vector<float>>>;   // (vector_id, vector_data)

// operations.rs:8-120.  Tie rules of the reference: `max_by` keeps the LAST maximum (second centroid),
// `dist1 <= dist2` sends ties to the first partition.
class Split {
 public:
  size_t posting_id;
  VectorList vectors;
  std::shared_ptr<distances::DistanceMetric> distance_metric;
  std::pair<size_t, size_t> new_posting_ids;

  Split(size_t posting, VectorList vecs, std::shared_ptr<distances::DistanceMetric> metric,
        std::pair<size_t, size_t> new_ids, std::shared_ptr<Context> ctx = nullptr)
      : posting_id(posting), vectors(std::move(vecs)), distance_metric(std::move(metric)), new_posting_ids(new_ids),
        ctx_(ctx ? std::move(ctx) : Context::shared()) {}

  // operations.rs:33-58: (vectors[0], the farthest of vectors[1..] from it)
  std::pair<std::vector<float>, std::vector<float>> select_initial_centroids() {
    if (vectors.size() < 2) throw Error("Not enough vectors to split");
    upload();
    // the fold keeps the earliest maximum, so it runs over the members in reverse order
    std::vector<uint64_t> members;
    for (size_t i = vectors.size() - 1; i >= 1; --i) members.push_back(i);
    float dist = 0.f;
    uint64_t row = 0;
    check(spf_farthest_from(ds_->handle(), distance_metric->kind(), vectors[0].second.data(), UINT64_MAX, members.data(),
                            members.size(), &dist, &row));
    if (row == UINT64_MAX) row = vectors.size() - 1;    // every distance is 0: max_by returns the last element
    return {vectors[0].second, vectors[(size_t)row].second};
  }

  // operations.rs:61-82
  std::pair<VectorList, VectorList> assign_vectors(const std::vector<float>& c1, const std::vector<float>& c2) {
    upload();
    std::vector<float> cen(c1);
    cen.insert(cen.end(), c2.begin(), c2.end());
    spf_assign_result* r = nullptr;
    check(spf_assign_vectors(ds_->handle(), distance_metric->kind(), nullptr, vectors.size(), cen.data(), 2, 1.0f,
                             SPF_ASSIGN_NO_CSR, &r));
    std::vector<uint32_t> best(vectors.size());
    const int rc = spf_assign_fetch(r, best.data(), nullptr, nullptr, nullptr);
    spf_assign_free(r);
    check(rc);
    std::pair<VectorList, VectorList> out;
    for (size_t i = 0; i < vectors.size(); ++i) (best[i] == 0 ? out.first : out.second).push_back(vectors[i]);
    return out;
  }

  std::vector<size_t> execute() {                         // operations.rs:86-101
    auto c = select_initial_centroids();
    partitions = assign_vectors(c.first, c.second);
    return get_affected_partitions();
  }
  bool validate() const {                                 // operations.rs:103-112
    return vectors.size() >= 2 && new_posting_ids.first != posting_id && new_posting_ids.second != posting_id &&
           new_posting_ids.first != new_posting_ids.second;
  }
  std::vector<size_t> get_affected_partitions() const { return {posting_id, new_posting_ids.first, new_posting_ids.second}; }

  std::pair<VectorList, VectorList> partitions;          // what execute() computed (the reference drops it)

 private:
  void upload() {
    if (ds_) return;
    const size_t d = vectors[0].second.size();
    std::vector<float> rows(vectors.size() * d);
    for (size_t i = 0; i < vectors.size(); ++i) {
      if (vectors[i].second.size() != d) throw Error("vectors of different length");
      std::copy(vectors[i].second.begin(), vectors[i].second.end(), rows.begin() + i * d);
    }
    ds_ = std::make_shared<DeviceDataset>(ctx_, ArrayView2(rows.data(), vectors.size(), d));
  }
  std::shared_ptr<Context> ctx_;
  std::shared_ptr<DeviceDataset> ds_;
};

// operations.rs:222-300: the nearest candidate centroid, `min_by` keeps the FIRST minimum.
class Reassign {
 public:
  size_t vector_id, from_posting;
  std::vector<float> vector;
  VectorList candidate_postings;                          // (posting_id, centroid)
  std::shared_ptr<distances::DistanceMetric> distance_metric;
  uint64_t version;

  Reassign(size_t id, std::vector<float> vec, size_t from, VectorList candidates,
           std::shared_ptr<distances::DistanceMetric> metric, uint64_t ver, std::shared_ptr<Context> ctx = nullptr)
      : vector_id(id), from_posting(from), vector(std::move(vec)), candidate_postings(std::move(candidates)),
        distance_metric(std::move(metric)), version(ver), ctx_(ctx ? std::move(ctx) : Context::shared()) {}

  size_t find_best_posting() const {                      // operations.rs:253-276
    if (candidate_postings.empty()) throw Error("No candidate postings available");
    const size_t d = vector.size(), m = candidate_postings.size();
    std::vector<float> a(m * d), b(m * d), dist(m);
    for (size_t j = 0; j < m; ++j) {
      std::copy(vector.begin(), vector.end(), a.begin() + j * d);
      std::copy(candidate_postings[j].second.begin(), candidate_postings[j].second.end(), b.begin() + j * d);
    }
    check(spf_distance_pairs(ctx_->handle(), distance_metric->kind(), a.data(), b.data(), (uint32_t)d, m, dist.data()));
    size_t best = 0;
    for (size_t j = 1; j < m; ++j)
      if (dist[j] < dist[best]) best = j;
    return candidate_postings[best].first;
  }

 private:
  std::shared_ptr<Context> ctx_;
};

}  // namespace lire
}  // namespace spfresh
