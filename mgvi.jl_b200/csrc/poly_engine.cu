// Polynomial-fit engine (test/test_models/model_polyfit.jl) -- placeholder until implemented.
#include "engine.h"
namespace mgvi {
Engine* make_poly_engine(mgvib200_ctx*, const mgvib200_model_desc&) {
    throw RtError{"POLY model: not implemented yet", MGVIB200_ERR_UNSUPPORTED};
}
}  // namespace mgvi
