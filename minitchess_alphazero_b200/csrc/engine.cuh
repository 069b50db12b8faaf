// engine.cuh -- host-side engine object behind the az_* entry points.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "common.cuh"
#include "mcts_core.cuh"

namespace mcaz {
struct Network;  // net.cu
}

struct az_engine {
    az_config cfg;
    az::View v;                  // device pointers + parameters, passed to kernels by value
    cudaStream_t stream = 0;     // legacy default stream: orders with the caller's torch work
    std::vector<void*> allocs;
    mcaz::Scratch scratch;       // staging for host-pointer arguments
    int device = 0;
    bool leaf_pending = false;   // az_select_expand done, az_backup not yet
    // external-evaluator scratch
    double* d_noise = nullptr;       // [G x MC_MAX_MOVES]
    uint8_t* d_noise_used = nullptr; // [G]
    float* d_logits = nullptr;       // [G x 554]  (built-in network output / staging)
    float* d_values = nullptr;       // [G]
    float* d_priors = nullptr;       // [G x MC_MAX_MOVES] staging for injected priors
    // replay
    az_replay_tuple* d_record = nullptr;   // [G x MAX_DEPTH] tuples of the running games
    az_replay_tuple* d_replay = nullptr;   // ring of finished-game tuples
    unsigned long long* d_replay_count = nullptr;
    size_t replay_capacity = 0;
    unsigned long long sim_counter = 0;    // launches of the search (device RNG key only where no per-game serial exists)
    az::CacheEntry* d_cache = nullptr;     // exact evaluation cache (eval_cache_log2 > 0)
    uint32_t cache_mask = 0;
    uint32_t cache_epoch = 1;              // bumped by az_set_weights
    int parity = 0;                        // row counter of the last search launch
    int lookahead_rows = 0;                // look-ahead rows per batch (cfg.lookahead_rows when the cache is on)
    uint32_t* d_pending = nullptr;         // [2] games waiting for a network row after a launch
    int defer_rows = 0;                    // cfg.defer_rows where it applies (built-in network, one leaf per step, no look-ahead rows)
    // az_profile_tree: CUDA events around every search_step_kernel launch
    bool tree_profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tree_events;
    size_t tree_events_used = 0;
    mcaz::Network* net = nullptr;
    uint64_t launches = 0;
};

namespace mcaz {
int engine_check_errors(az_engine* e);   // reads the device error flag (synchronises)
// net.cu
int network_create(az_engine* e);
void network_destroy(az_engine* e);
int network_set_weights(az_engine* e, const float* flat_device);
int network_forward_search(az_engine* e, const az::View& V, float* values);
int network_forward(az_engine* e, const uint8_t* tokens, const float* clocks, const uint8_t* active, int n,
                    float* logits, float* values);
}  // namespace mcaz
