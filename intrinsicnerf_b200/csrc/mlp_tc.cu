#include "common.cuh"
namespace inrf {
int launch_mlp_tc(const MlpArgs& a, cudaStream_t st) { set_error("tc path not built yet"); return INRF_EUNSUPPORTED; }
}
