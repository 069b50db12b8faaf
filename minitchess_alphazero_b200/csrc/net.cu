// net.cu -- built-in policy/value network (placeholder until the tcgen05 tower lands).
#include "engine.cuh"

namespace mcaz {
int network_create(az_engine*) { return fail(MCAZ_ESTATE, "built-in network not available in this build"); }
void network_destroy(az_engine*) {}
int network_set_weights(az_engine*, const float*) { return fail(MCAZ_ESTATE, "built-in network not available"); }
int network_forward(az_engine*, const uint8_t*, const float*, const uint8_t*, int, float*, float*) {
    return fail(MCAZ_ESTATE, "built-in network not available");
}
}  // namespace mcaz
