// examples/build_index.rs of the reference, on the C++ host layer: build an index over the 6 x 2
// toy matrix and ask for the nearest neighbour of (1, 2).  Expected: point_id 0, vector [1.0, 2.0].
#include <cstdio>

#include "../spfresh.hpp"

int main(int argc, char** argv) {
  using namespace spfresh;
  try {
    spann::Config config = spann::Config::from_file(argc > 1 ? argv[1] : "host/examples/example_config.yaml");
    const float data[12] = {1.0f, 2.0f, 1.5f, 2.5f, 8.0f, 8.0f, 8.5f, 8.5f, 4.0f, 4.0f, 4.5f, 4.5f};
    spann::SpannIndex index = spann::SpannIndexBuilder(config).with_data(ArrayView2(data, 6, 2)).build(2);
    const float query[2] = {1.0f, 2.0f};
    auto result = index.find_k_nearest_neighbor_spann(ArrayView1{query, 2}, 1);
    if (!result) { printf("None\n"); return 0; }
    for (const auto& p : *result) printf("PointData { point_id: %llu, vector: [%.1f, %.1f] }\n",
                                         (unsigned long long)p.point_id, p.vector[0], p.vector[1]);
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "Failed to build SPANN index: %s\n", e.what());
    return 1;
  }
}
