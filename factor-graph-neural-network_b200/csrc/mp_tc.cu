// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace fgnn {
bool tc_supported(const fgnn_mp_args*) { return false; }
size_t tc_workspace_bytes(const fgnn_mp_args*) { return 0; }
int launch_mp_tc(const MpParams&, const fgnn_mp_args*, cudaStream_t) { return FGNN_ERR_UNSUPPORTED; }
}
