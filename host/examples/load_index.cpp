// examples/load_index.rs of the reference, on the C++ host layer: load the index build_index wrote.
#include <cstdio>

#include "../spfresh.hpp"

int main(int argc, char** argv) {
  using namespace spfresh;
  try {
    spann::Config config = spann::Config::from_file(argc > 1 ? argv[1] : "host/examples/example_config.yaml");
    spann::SpannIndex index = spann::SpannIndexBuilder(config).load(2);
    const float query[2] = {1.0f, 2.0f};
    auto result = index.find_k_nearest_neighbor_spann(ArrayView1{query, 2}, 1);
    if (!result) { printf("None\n"); return 0; }
    printf("Nearest neighbour: point_id:%llu and vector:[%.1f, %.1f]\n", (unsigned long long)(*result)[0].point_id,
           (*result)[0].vector[0], (*result)[0].vector[1]);
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "Failed to load SPANN index: %s\n", e.what());
    return 1;
  }
}
