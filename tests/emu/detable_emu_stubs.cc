// Symbols of the translation units that are NOT part of the emulated build (host_api.cu, sharded.cu, fused.cu) but are
// referenced by table.cu / evict.cu.
#include "cuda_emu.h"
#include "cuda_runtime_emu.h"

#include "../../recommenders_addons_b200/csrc/host.h"

namespace det {
void host_pipe_free(det_table*) {}
}  // namespace det
