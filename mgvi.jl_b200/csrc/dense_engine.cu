// Dense-linear engine (BASELINE.json configs[3]) -- placeholder until implemented.
#include "engine.h"
namespace mgvi {
Engine* make_dense_engine(mgvib200_ctx*, const mgvib200_model_desc&) {
    throw RtError{"DENSE_LINEAR model: not implemented yet", MGVIB200_ERR_UNSUPPORTED};
}
}  // namespace mgvi
