// placeholder until the HalfCheetah kernels land
#include "mujoco.cuh"
namespace epb {
struct MjcPool { int dummy; };
MjcPool* mjc_pool_create(int, int, int, double, double, double) { return nullptr; }
void mjc_pool_destroy(MjcPool*) {}
int mjc_state_reals(const MjcPool*) { return 0; }
cudaError_t mjc_launch_step(MjcPool*, const StateView&, const OutView&, const double*, const int32_t*, int, int, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t mjc_launch_rollout(MjcPool*, const StateView&, const OutView&, const double*, int, cudaStream_t) { return cudaErrorNotSupported; }
}
