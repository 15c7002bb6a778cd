// k_binning.cu — K1..K5: lattice, voxel keys, leaves, plane-fit rotation, claim + projection.
#include "gpc_device.cuh"
#include "gpc_internal.h"

namespace gpc {

struct BinningWork {
    int unused = 0;
};

void binning_free(BinningWork* w) { delete w; }

}  // namespace gpc
