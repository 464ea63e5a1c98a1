// Contiguous short lanes — placeholder until the bit-sliced kernel lands.
#include "nds_internal.h"
namespace nds {
bool short_bitslice_supported(const nds_ctx *, const DeviceCall &) { return false; }
cudaError_t launch_short_bitslice(nds_ctx *, const DeviceCall &) { return cudaErrorNotSupported; }
}  // namespace nds
