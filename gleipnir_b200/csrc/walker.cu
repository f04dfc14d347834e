// Walker launch for the built-in likelihood models (kernel template: walker.cuh).
#include "walker.cuh"

namespace gl {

int walk_launch(const WalkParams &P, const ModelDev &M, cudaStream_t st) {
    if (M.kind == GL_MODEL_EXTERNAL) {
        GL_REQUIRE(M.ext_walk != nullptr, "external model without a walker entry point");
        return M.ext_walk(&P, &M, (void *)st);
    }
    WalkLauncher L{P, M, st};
    return dispatch_model(M.kind, L);
}

}  // namespace gl
