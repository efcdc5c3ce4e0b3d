// placeholder until the tcgen05 chunk kernel lands
#include "gdr_common.cuh"
namespace gdkvm {
bool chunked_supports(const GdkvmGdrParams&) { return false; }
int launch_chunked(const GdkvmGdrParams&, cudaStream_t) { return (int)cudaErrorNotSupported; }
}
