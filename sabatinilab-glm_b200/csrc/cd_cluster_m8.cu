// Instantiations of the cluster coordinate-descent kernel for groups of 8 model(s) (see cd_cluster.cuh).
#include "cd_cluster.cuh"

namespace sglm {
namespace cdc {
int launch_m8(const Args &a, int K) { return launch_group_wide<8>(a, K); }
}  // namespace cdc
}  // namespace sglm
