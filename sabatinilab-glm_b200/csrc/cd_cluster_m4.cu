// Instantiations of the cluster coordinate-descent kernel for groups of 4 model(s) (see cd_cluster.cuh).
#include "cd_cluster.cuh"

namespace sglm {
namespace cdc {
int launch_m4(const Args &a, int K) { return launch_group<4>(a, K); }
}  // namespace cdc
}  // namespace sglm
