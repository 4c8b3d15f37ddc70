// Structured-hex value kernel (tiled gather).  Placeholder until the tile kernel lands: the
// general atomic kernel in assemble.cu is used for every mesh.
#include <cstdlib>

#include "smfem_internal.cuh"

struct Material {
    double d11, lam, mu;
};

bool values_tile_enabled() { return false; }

void values_assemble_tile(smfem_ctx *, smfem_mesh *, smfem_matrix *, Material) {
    throw SmfemError(SMFEM_ERR_UNSUPPORTED, "tile kernel not built");
}
