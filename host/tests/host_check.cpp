// host_check — drives the C++ host layer (host/spfresh.hpp) for the parity tests.
//
//   host_check --abi                      prints the ABI version (no GPU needed)
//   host_check <scenario.txt> <out.txt>   runs fit() (+ index build, batched search, save / load)
//                                         with scripted random decisions and writes the result
//
// Scenario lines (`key values...`): data <file.f32> <n> <d> | metric <name> | init <Random|KMeansPlusPlus>
// | initial_k <k> | desired <size> | multiple <count> <rows...> | pick <num> <den> (choose_index(n) =
// n*num/den) | first <row> (first k-means++ pick) | u01 <count> <draws...> | queries <file.f32> <nq>
// | k <topk> | nprobe <n> | out_dir <dir>.  tests/test_host_cpp.py compares the output with the oracle.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../spfresh.hpp"

using namespace spfresh;

static std::vector<float> read_f32(const std::string& path, size_t count) {
  std::vector<float> v(count);
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error("cannot open " + path);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(count * sizeof(float)));
  if (!f) throw Error("short read on " + path);
  return v;
}

int main(int argc, char** argv) {
  if (argc == 2 && !strcmp(argv[1], "--abi")) {
    printf("%d\n", spf_abi_version());
    return 0;
  }
  if (argc == 6 && !strcmp(argv[1], "--lire")) {
    // host_check --lire <vectors.f32> <n> <d> <out.txt>: Split over the n vectors (ids = row numbers,
    // squared L2) and Reassign of vector 1 against vectors 2..9 as candidate centroids
    try {
      const size_t n = std::stoull(argv[3]), d = std::stoull(argv[4]);
      const std::vector<float> rows = read_f32(argv[2], n * d);
      lire::VectorList vl;
      for (size_t i = 0; i < n; ++i) vl.push_back({i, std::vector<float>(rows.begin() + i * d, rows.begin() + (i + 1) * d)});
      auto metric = std::make_shared<distances::SquaredEuclideanDistance>();
      lire::Split sp(7, vl, metric, {8, 9});
      auto cen = sp.select_initial_centroids();
      auto parts = sp.assign_vectors(cen.first, cen.second);
      std::ofstream out(argv[5]);
      size_t far = 0;
      for (size_t i = 1; i < n; ++i) if (vl[i].second == cen.second) far = i;    // last vector equal to c2
      out << "far " << far << "\npartition1 " << parts.first.size();
      for (auto& p : parts.first) out << " " << p.first;
      out << "\npartition2 " << parts.second.size();
      for (auto& p : parts.second) out << " " << p.first;
      lire::VectorList cands(vl.begin() + 2, vl.begin() + 10);
      lire::Reassign re(1, vl[1].second, 0, cands, metric, 1);
      out << "\nbest " << re.find_best_posting() << "\n";
      return 0;
    } catch (const std::exception& e) {
      fprintf(stderr, "host_check: %s\n", e.what());
      return 1;
    }
  }
  if (argc != 3) {
    fprintf(stderr, "usage: host_check <scenario.txt> <out.txt> | --lire <vectors.f32> <n> <d> <out.txt> | --abi\n");
    return 2;
  }
  try {
    std::ifstream sc(argv[1]);
    if (!sc) throw Error(std::string("cannot open ") + argv[1]);
    std::string data_path, query_path, metric = "Euclidean", init = "Random", out_dir;
    size_t n = 0, d = 0, nq = 0, initial_k = 1, desired = 0, topk = 10, nprobe = 0;
    uint64_t pick_num = 0, pick_den = 1, first = 0;
    std::vector<uint64_t> multiple;
    std::vector<double> u01;
    std::string line;
    while (std::getline(sc, line)) {
      std::istringstream ls(line);
      std::string key;
      if (!(ls >> key)) continue;
      if (key == "data") ls >> data_path >> n >> d;
      else if (key == "queries") ls >> query_path >> nq;
      else if (key == "metric") ls >> metric;
      else if (key == "init") ls >> init;
      else if (key == "initial_k") ls >> initial_k;
      else if (key == "desired") ls >> desired;
      else if (key == "k") ls >> topk;
      else if (key == "nprobe") ls >> nprobe;
      else if (key == "out_dir") ls >> out_dir;
      else if (key == "pick") ls >> pick_num >> pick_den;
      else if (key == "first") ls >> first;
      else if (key == "multiple") { size_t c; ls >> c; multiple.resize(c); for (auto& v : multiple) ls >> v; }
      else if (key == "u01") { size_t c; ls >> c; u01.resize(c); for (auto& v : u01) ls >> v; }
    }
    const std::vector<float> rows = read_f32(data_path, n * d);
    auto ctx = Context::shared(0);

    spann::Config cfg;
    cfg.clustering_params = {metric, init, initial_k};
    if (!out_dir.empty()) cfg.output_path = out_dir;
    cfg.validate();
    clustering::ClusteringParams params = cfg.to_clustering_params();
    params.desired_cluster_size = desired;
    auto rs = std::make_shared<clustering::ScriptedRandomSource>();
    rs->multiple = multiple;
    rs->u01 = u01;
    bool first_used = init != "KMeansPlusPlus";
    rs->index = [=](uint64_t m) mutable -> uint64_t {
      if (!first_used) { first_used = true; return first; }   // the first k-means++ pick (:253-255)
      return m * pick_num / pick_den;
    };
    params.random_source = rs;

    clustering::HierarchicalClustering hc(params, ArrayView2(rows.data(), n, d), ctx);
    hc.fit();
    std::ofstream out(argv[2]);
    out << "clusters " << hc.clusters.size() << "\n";
    for (const auto& c : hc.clusters) {
      out << *c.centroid_idx << " " << c.depth << " " << c.points.size();
      for (uint64_t p : c.points) out << " " << p;
      out << "\n";
    }
    const std::vector<size_t> lab = hc.labels();
    out << "labels " << lab.size();
    for (size_t v : lab) out << " " << v;
    out << "\n";

    if (nq) {
      const std::vector<float> q = read_f32(query_path, nq * d);
      spann::SpannIndex index(out_dir, ctx);
      index.create_posting_lists(*hc.dataset(), hc.clusters);
      auto res = index.find_k_nearest_neighbors_batch(ArrayView2(q.data(), nq, d), topk, nprobe);
      out << "search " << nq << "\n";
      for (const auto& r : res) {
        out << (r ? r->size() : 0);
        if (r) for (const auto& p : *r) out << " " << p.point_id;
        out << "\n";
      }
      // the single-query entry point and the vectors it returns
      auto one = index.find_k_nearest_neighbor_spann(ArrayView1{q.data(), d}, topk);
      bool ok = (bool)one == (bool)res[0];
      if (one && res[0]) {
        ok = one->size() == res[0]->size();
        for (size_t i = 0; ok && i < one->size(); ++i) {
          const auto& p = (*one)[i];
          ok = p.point_id == (*res[0])[i].point_id &&
               !memcmp(p.vector.data(), rows.data() + p.point_id * d, d * sizeof(float));
        }
      }
      out << "single_query_same " << (ok ? 1 : 0) << "\n";
      if (!out_dir.empty()) {           // save in the reference's on-disk layout, load, ask again
        index.save_posting_list();
        index.save_centroids(out_dir + "/centroids.bin");
        spann::SpannIndex loaded = spann::SpannIndexBuilder(cfg, ctx).load(d);
        auto res2 = loaded.find_k_nearest_neighbors_batch(ArrayView2(q.data(), nq, d), topk, nprobe);
        bool same = res2.size() == res.size();
        for (size_t i = 0; same && i < res.size(); ++i) {
          same = (bool)res[i] == (bool)res2[i];
          if (same && res[i]) {
            same = res[i]->size() == res2[i]->size();
            for (size_t j = 0; same && j < res[i]->size(); ++j)
              same = (*res[i])[j].point_id == (*res2[i])[j].point_id && (*res[i])[j].vector == (*res2[i])[j].vector;
          }
        }
        out << "loaded_same " << (same ? 1 : 0) << "\n";
      }
    }
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "host_check: %s\n", e.what());
    return 1;
  }
}
