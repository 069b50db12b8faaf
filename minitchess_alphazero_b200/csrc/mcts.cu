// mcts.cu -- kernels and az_* entry points of the batched AlphaZero search engine (sm_100a).
//
// Kernels (one warp per game/tree, grid-stride over games, grid sized to the SM count):
//   select_expand_kernel  PUCT descent + expansion of the leaf (exp/agent.py:54-66,75-88)
//   backup_kernel         legal-logit softmax -> priors, value backup (exp/agent.py:47-52,67-72)
//   search_step_kernel    az_search / az_selfplay inner step: backup of the evaluated leaf, next descents while the
//                         game's budget lasts (finished / cached leaves complete on the spot), dense network rows;
//                         in az_selfplay also the game's own move choice, replay record, move and restart
//   play_kernel / play_device_kernel   the real game line (exp/environment.py:68-82,
//                         exp/agent.py:110-119, exp/callbacks.py:31-54)
//   recycle_kernel        in-place compaction of trees whose old plies can no longer be reached (block per tree)
//   root_stats_kernel / node_stats_kernel   what exp/policy.py:118-121 reads back
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "engine.cuh"

using namespace mcaz;
using az::View;
static_assert(MC_NODE_TERMINAL == az::INFO_TERMINAL && MC_NODE_DECISIVE == az::INFO_DECISIVE, "az_tree_dump info bits");

namespace mcaz {
int num_sms();

__global__ void __launch_bounds__(128) select_expand_kernel(View V, const double* noise, uint8_t* noise_used) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < V.G; g += warps)
        for (int j = 0; j < V.K; ++j) {
            az::select_expand_one(V, g, lane, noise, noise_used, j);
            __syncwarp();
        }
}

// ---- device move choice (throughput mode) -------------------------------------------------------------
// Empties both trees of game g and puts it back on the start position (warp-cooperative).
__device__ __forceinline__ void restart_one(const View& V, int g, int lane, const mc_state& start) {
    uint4* tab = reinterpret_cast<uint4*>(V.ht + (size_t)(2 * g) * V.HC);     // HC is a power of two >= 64
    for (int i = lane; i < 2 * V.HC / 4; i += 32) tab[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    if (lane == 0) {
        V.game_state[g] = start;
        V.game_ply[g] = mc::white_to_move(start) ? 0 : 1;
        V.game_start_ply[g] = V.game_ply[g];
        az::hist_reset(V, g, start);
        for (int t = 2 * g; t < 2 * g + 2; ++t) { V.tree_nodes[t] = 0; V.tree_edges[t] = 0; V.tree_root[t] = az::NONE; }
        for (int j = 0; j < V.K; ++j) V.leaf_kind[g * V.K + j] = az::LEAF_NONE;
        V.game_result[g] = MC_ONGOING;
    }
    __syncwarp();
}

// Play a move taken from the tree's own edge list of the current position (legal by construction) -- what
// az::play_one does on one thread, warp-cooperatively: the result of the new position needs its legal-move count
// (lane = square, as in the expansion) and the fivefold-repetition count over the game line (lanes split the history).
__device__ __forceinline__ void play_chosen_move(const View& V, int g, int lane, int code) {
    const mc_state s = V.game_state[g];
    int fv, tv;
    mc::code_to_view(code, fv, tv);
    const bool white = mc::white_to_move(s);
    const mc_state o = mc::apply_move(s, white ? fv : 29 - fv, white ? tv : 29 - tv);
    const mc::Sets st = mc::sets_of(o);
    const bool mover_white = mc::white_to_move(o);
    const int sq = mover_white ? lane : 29 - lane;
    const int type = (lane < 30 && ((st.own >> sq) & 1u)) ? mc::piece_at(o, sq) : 0;
    int n_moves = mc::popc(az::legal_targets_warp(st, mover_white, type, sq & 31, lane, V.rules));
    for (int k = 16; k > 0; k >>= 1) n_moves += __shfl_xor_sync(0xffffffffu, n_moves, k);
    int res = mc::result_of(o, st, n_moves, V.rules);
    // game line (exp/environment.py:39: board.result() sees the move stack): positions since the last irreversible move
    int n = V.game_hist_len[g];
    __syncwarp();
    if (mc::halfmove(o) == 0) n = 0;
    const az::Board4 key = az::rep_key(o);
    az::Board4* h = V.game_hist + (size_t)g * az::HIST;
    if (lane == 0 && n < az::HIST) h[n] = key;
    if (n < az::HIST) ++n;
    __syncwarp();
    if (res == MC_ONGOING && V.rules.fivefold_repetition) {
        int same = 0;
        for (int i = lane; i < n; i += 32) same += (h[i].x == key.x && h[i].y == key.y && h[i].z == key.z && h[i].w == key.w) ? 1 : 0;
        for (int k = 16; k > 0; k >>= 1) same += __shfl_xor_sync(0xffffffffu, same, k);
        if (same >= 5) res = MC_DRAW;
    }
    if (lane == 0) {
        V.game_hist_len[g] = n;
        V.game_state[g] = o;
        V.game_ply[g] += 1;
        V.tree_root[2 * g] = az::NONE;
        V.tree_root[2 * g + 1] = az::NONE;
        V.game_result[g] = (int8_t)res;
        az::count(V, az::C_MOVES, 1);
        if (res != MC_ONGOING) az::count(V, az::C_GAMES, 1);
    }
    __syncwarp();
}

// Pick a move from the root visit counts, record the replay tuple, play it and back-fill the rewards of a
// finished game (exp/agent.py:110-119, exp/callbacks.py:31-54).  Warp-cooperative; returns false when the
// root has no visits yet (nothing to choose from).
__device__ __forceinline__ bool play_device_one(const View& V, int g, int lane) {
    if (V.game_result[g] != MC_ONGOING) return false;
    const int ply = V.game_ply[g];
    const int t = 2 * g + (ply & 1);
    uint32_t root = V.tree_root[t];
    if (root == az::NONE) root = az::ht_find(V, t, V.game_state[g]);
    if (root == az::NONE) return false;
    const az::NodeHead rh = az::load_head(&V.nodes[(size_t)t * V.NC + root]);
    const int E = (int)(rh.info & 0xffffu);
    const az::Edge* re = V.edges + (size_t)t * V.EC + rh.edge_off;          // the root's edges
    const mc_state s = V.game_state[g];
    unsigned int nsum = 0, nmax = 0;
    for (int i = lane; i < E; i += 32) { unsigned int c = re[i].stat.N; nsum += c; nmax = max(nmax, c); }
    for (int o = 16; o > 0; o >>= 1) { nsum += __shfl_xor_sync(0xffffffffu, nsum, o); nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o)); }
    if (nsum == 0) return false;
    const uint32_t serial = V.move_serial[g];          // the game slot's own move counter keys the RNG
    __syncwarp();
    if (lane == 0) V.move_serial[g] = serial + 1;
    Philox rng(V.seed ^ 0xA5A5A5A5DEADBEEFull, (uint32_t)g, serial, 0u);
    const double u = rng.uniform();     // same on every lane
    int choice = -1;
    if (mc::fullmove(s) < V.tau_change) {
        // sample proportionally to N (np.random.choice(legal, p=pi))
        const double target = u * (double)nsum;
        double acc = 0;
        for (int i = 0; i < E && choice < 0; ++i) { acc += (double)re[i].stat.N; if (target < acc) choice = i; }
        if (choice < 0) choice = E - 1;
    } else {
        int n_best = 0;
        for (int i = 0; i < E; ++i) n_best += (re[i].stat.N == nmax);
        int pick = min((int)(u * n_best), n_best - 1);
        for (int i = 0; i < E; ++i) if (re[i].stat.N == nmax) { if (pick == 0) { choice = i; break; } --pick; }
    }
    const int code = re[choice].link.code;
    // replay tuple of this ply
    az_replay_tuple* rec = V.record + (size_t)g * az::MAX_DEPTH + min(ply - V.game_start_ply[g], az::MAX_DEPTH - 1);
    for (int i = lane; i < E; i += 32) { rec->codes[i] = re[i].link.code; rec->pi[i] = (float)((double)re[i].stat.N / (double)nsum); }
    if (lane == 0) { rec->observation = s; rec->n_legal = (uint16_t)E; rec->action = (uint16_t)code; rec->reward = 0; rec->weights_version = 0u; }
    __syncwarp();
    play_chosen_move(V, g, lane, code);
    const int res = V.game_result[g];
    if (res != MC_ONGOING) {
        // exp/callbacks.py:49-53: the side that moved last gets +reward, alternating backwards; the finished game's
        // tuples go to the replay ring (lanes copy the 604-byte tuples word by word)
        static_assert(sizeof(az_replay_tuple) % 4 == 0, "az_replay_tuple is copied by words");
        const int n_rec = min(V.game_ply[g] - V.game_start_ply[g], az::MAX_DEPTH);
        // the whole game or nothing: reserve its n_rec slots in one step; a game that does not fit is dropped and counted
        // (a partly written game would break the reward chain of its episode)
        unsigned long long base = ~0ull;
        if (lane == 0) {
            unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(V.replay_count);
            while (old + (unsigned long long)n_rec <= V.replay_cap) {
                const unsigned long long got = atomicCAS(V.replay_count, old, old + (unsigned long long)n_rec);
                if (got == old) { base = old; break; }
                old = got;
            }
            if (base == ~0ull) az::count(V, az::C_REPLAY_DROPPED, (unsigned long long)n_rec);
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        const int last_reward = (res == MC_DRAW) ? 0 : 1;
        for (int p = n_rec - 1; p >= 0 && base != ~0ull; --p) {
            az_replay_tuple* r = V.record + (size_t)g * az::MAX_DEPTH + p;
            if (lane == 0) {
                r->reward = (int8_t)(((n_rec - 1 - p) & 1) ? -last_reward : last_reward);
                r->weights_version = V.weights_version;          // the stamp at push time (app/base.py:63-68)
            }
            __syncwarp();
            const uint32_t* src = reinterpret_cast<const uint32_t*>(r);
            uint32_t* out = reinterpret_cast<uint32_t*>(V.replay + base + (unsigned long long)p);
            for (int w = lane; w < (int)(sizeof(az_replay_tuple) / 4); w += 32) out[w] = src[w];
        }
    }
    __syncwarp();
    return true;
}

// One launch of the search for game g (leaves_per_step = 1): finish the simulation whose leaf the network just
// evaluated, then start descents while the game's budget lasts.  A descent that ends on a terminal position
// or on one found in the evaluation cache needs no network row: it is backed up on the spot and the game goes
// straight on (at most free_max descents per launch); the first leaf that needs the network takes a row of the
// batch V.batch and ends the game's turn.  The order of a game's simulations -- all that the reference's
// sequential search depends on -- is unchanged (exp/agent.py:41-45).
template <bool LOOKAHEAD>
__device__ __forceinline__ void search_one(const View& V, int g, int lane, const float* values, const mc_state& start) {
    if (V.check_deferred && V.leaf_kind[g] == az::LEAF_EVAL && (uint32_t)V.slot_row[g] >= V.row_eff[V.parity ^ 1]) {
        // the last pass left this leaf's row out (a short last tile pair, az_config.defer_rows): the same leaf takes a row of
        // this batch and the game waits on -- nothing is backed up, no descent starts, the budget is untouched
        const mc_state s = V.leaf_states[g];
        int row = 0;
        if (lane == 0) { row = (int)atomicAdd(&V.row_count[V.parity], 1u); V.row_slot[row] = g; V.slot_row[g] = row; }
        row = __shfl_sync(0xffffffffu, row, 0);
#if defined(__CUDA_ARCH__)
        az::write_network_row(V, row, lane, s, mc::sets_of(s), mc::white_to_move(s));
#endif
        if (lane == 0 && V.pending_count) atomicAdd(&V.pending_count[V.parity], 1u);
        return;
    }
    az::backup_one(V, g, lane, nullptr, values, nullptr, 0);
    __syncwarp();
    int left = V.new_budget >= 0 ? V.new_budget : V.sims_left[g];
    bool waiting = false;
    for (int it = 0; it < V.free_max; ++it) {
        if (left <= 0) {
            if (!V.async_play) break;
            // az_selfplay: all simulations of this move are backed up -> choose, record, play, maybe restart
            if (play_device_one(V, g, lane) && V.game_result[g] != MC_ONGOING) restart_one(V, g, lane, start);
            left = V.sims_per_move;
        }
        if (V.game_result[g] != MC_ONGOING) break;
        // caller-supplied root noise (az_search_noise): simulation number i of this call mixes row i of the block
        const double* noise = V.noise_block ? V.noise_block + (size_t)(V.noise_budget - left) * V.G * MC_MAX_MOVES : nullptr;
        --left;
        const uint8_t kind = az::select_expand_one<LOOKAHEAD>(V, g, lane, noise, nullptr, 0);
        __syncwarp();
        if (kind != az::LEAF_TERMINAL && kind != az::LEAF_CACHED) { waiting = kind == az::LEAF_EVAL; break; }
        az::backup_one(V, g, lane, nullptr, values, nullptr, 0);
        __syncwarp();
    }
    if (lane == 0) {
        V.sims_left[g] = left;
        // the host stops the batch loop when this stays 0: a game counts while it waits for a row OR still has budget
        // (a launch whose descents all ended on finished / cached leaves leaves nobody waiting with simulations unspent)
        if (V.pending_count && (waiting || (left > 0 && V.game_result[g] == MC_ONGOING))) atomicAdd(&V.pending_count[V.parity], 1u);
    }
}

// az_config.defer_rows: how many rows of the batch just filled the network pass evaluates (one thread; the stem, tower and head
// kernels of the pass and the next search launch all read this one word)
__global__ void cap_rows_kernel(View V) {
    const uint32_t count = min(V.row_count[V.parity], (uint32_t)V.row_cap);
    uint32_t eff = az::rows_to_run(count, (uint32_t)V.defer_thr);
    V.row_eff[V.parity] = eff;
    if (eff < count) { V.counters[az::C_DEFERRED_ROWS] += count - eff; V.counters[az::C_TRIMMED_BATCHES] += 1; }      // (one thread, stream-ordered)
}

// Network input of the look-ahead rows of a batch (row_slot = -1: queued by position only, queue_children): a warp per row,
// lane = board cell, the same code that writes a leaf's own row (write_network_row).
__global__ void __launch_bounds__(128) tokenize_lookahead_kernel(View V) {
    const int rows = min((int)V.row_count[V.parity], V.row_cap), lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        if (V.row_slot[r] >= 0) continue;
#if defined(__CUDA_ARCH__)
        const mc_state s = V.row_state[r];
        az::write_network_row(V, r, lane, s, mc::sets_of(s), mc::white_to_move(s));
#endif
    }
}

// A search that ends while look-ahead rows are queued never evaluates them: take their tags back, so that a later
// expansion queues those positions again instead of meeting each of them as a miss.
__global__ void __launch_bounds__(256) untag_rows_kernel(View V) {
    const int rows = min((int)V.row_count[V.parity], V.row_cap);
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        if (V.row_slot[r] >= 0) continue;
        const mc_state s = V.row_state[r];
        const uint32_t idx = az::cache_hash(s) & V.seen_mask;
#if defined(__CUDA_ARCH__)
        if (V.seen[idx] == az::seen_tag(V, s)) V.seen[idx] = 0u;
#endif
    }
}

// az_search / az_selfplay inner step.  leaves_per_step = K > 1 keeps the fixed form: back up the K descents of
// the previous launch, start K new ones (virtual loss keeps them apart), one row per slot.
template <bool LOOKAHEAD>
__global__ void __launch_bounds__(128, LOOKAHEAD ? 4 : 7) search_step_kernel(View V, const float* values, mc_state start) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    if (V.compact && blockIdx.x == 0 && threadIdx.x == 0) {     // the next launch's counters
        V.row_count[V.parity ^ 1] = 0u;
        if (V.pending_count) V.pending_count[V.parity ^ 1] = 0u;
    }
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < V.G; g += warps) {
        if (V.K == 1) { search_one<LOOKAHEAD>(V, g, lane, values, start); continue; }
        for (int j = 0; j < V.K; ++j) {
            az::backup_one(V, g, lane, nullptr, values, nullptr, j);
            __syncwarp();
        }
        if (V.free_max > 0)
            for (int j = 0; j < V.K; ++j) {
                az::select_expand_one(V, g, lane, nullptr, nullptr, j);
                __syncwarp();
            }
    }
}

__global__ void __launch_bounds__(128) backup_kernel(View V, const float* logits, const float* values, const float* priors) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < V.G; g += warps)
        for (int j = 0; j < V.K; ++j) {
            az::backup_one(V, g, lane, logits, values, priors, j);
            __syncwarp();
        }
}

// (re)start games: position, history, both trees emptied (block per listed game)
__global__ void __launch_bounds__(256) reset_games_kernel(View V, const int32_t* game_ids, int n, const mc_state* states,
                                                          mc_state start) {
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        const int g = game_ids ? game_ids[k] : k;
        if (g < 0 || g >= V.G) continue;
        uint32_t* tab = V.ht + (size_t)(2 * g) * V.HC;
        for (int i = threadIdx.x; i < 2 * V.HC; i += blockDim.x) tab[i] = 0u;
        if (threadIdx.x == 0) {
            mc_state s = states ? states[k] : start;
            V.game_state[g] = s;
            V.game_ply[g] = mc::white_to_move(s) ? 0 : 1;   // white's agent owns tree 0
            V.game_start_ply[g] = V.game_ply[g];
            az::hist_reset(V, g, s);
            V.game_result[g] = (int8_t)az::game_result_of(V, g, s);
            for (int t = 2 * g; t < 2 * g + 2; ++t) { V.tree_nodes[t] = 0; V.tree_edges[t] = 0; V.tree_root[t] = az::NONE; }
            for (int j = 0; j < V.K; ++j) { V.leaf_kind[g * V.K + j] = az::LEAF_NONE; V.needs_eval[g * V.K + j] = 0; V.path_len[g * V.K + j] = 0; }
            V.sims_left[g] = 0;
        }
    }
}

// empty single trees (block per listed tree)
__global__ void __launch_bounds__(256) reset_trees_kernel(View V, const int32_t* tree_ids, int n) {
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        const int t = tree_ids[k];
        if (t < 0 || t >= 2 * V.G) continue;
        uint32_t* tab = V.ht + (size_t)t * V.HC;
        for (int i = threadIdx.x; i < V.HC; i += blockDim.x) tab[i] = 0u;
        if (threadIdx.x == 0) { V.tree_nodes[t] = 0; V.tree_edges[t] = 0; V.tree_root[t] = az::NONE; }
    }
}

__global__ void set_positions_kernel(View V, const int32_t* game_ids, int n, const mc_state* states, const int32_t* tree_of_game) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int g = game_ids ? game_ids[k] : k;
        if (g < 0 || g >= V.G) continue;
        mc_state s = states[k];
        V.game_state[g] = s;
        V.game_ply[g] = tree_of_game ? (tree_of_game[k] & 1) : (mc::white_to_move(s) ? 0 : 1);
        V.game_start_ply[g] = V.game_ply[g];
        az::hist_reset(V, g, s);   // a fresh Board(fen) has no history (exp/agent.py:43)
        V.game_result[g] = (int8_t)az::game_result_of(V, g, s);
        V.tree_root[2 * g] = az::NONE;
        V.tree_root[2 * g + 1] = az::NONE;
        for (int j = 0; j < V.K; ++j) V.leaf_kind[g * V.K + j] = az::LEAF_NONE;
        V.sims_left[g] = 0;
    }
}

__global__ void play_kernel(View V, const int32_t* game_ids, const uint16_t* codes, int n, int8_t* results) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int g = game_ids ? game_ids[k] : k;
        if (g < 0 || g >= V.G) { az::raise(V, az::ERR_ILLEGAL); continue; }
        int st = az::play_one(V, g, codes[k]);
        if (st == 1) az::raise(V, az::ERR_ILLEGAL);
        if (results) results[k] = V.game_result[g];
    }
}

__global__ void game_states_kernel(View V, const int32_t* game_ids, int n, mc_state* states, int8_t* results) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int g = game_ids ? game_ids[k] : k;
        if (g < 0 || g >= V.G) continue;
        if (states) states[k] = V.game_state[g];
        if (results) results[k] = V.game_result[g];
    }
}

// Root statistics (exp/policy.py:118-121): codes, N, Q of the current position in the active tree.
__global__ void __launch_bounds__(128) root_stats_kernel(View V, const int32_t* game_ids, int n, uint16_t* codes,
                                                         uint32_t* visits, double* q, int32_t* n_legal) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += warps) {
        const int g = game_ids ? game_ids[k] : k;
        if (g < 0 || g >= V.G) { if (lane == 0) n_legal[k] = -1; continue; }
        const int t = 2 * g + (V.game_ply[g] & 1);
        uint32_t root = V.tree_root[t];
        if (root == az::NONE) root = az::ht_find(V, t, V.game_state[g]);
        if (root == az::NONE) { if (lane == 0) n_legal[k] = -1; continue; }
        const az::NodeHead h = az::load_head(&V.nodes[(size_t)t * V.NC + root]);
        const int E = (h.info & az::INFO_TERMINAL) ? 0 : (int)(h.info & 0xffffu);
        const az::Edge* re = V.edges + (size_t)t * V.EC + h.edge_off;
        for (int i = lane; i < E; i += 32) {
            const az::EdgeStat st = az::load_stat(&re[i]);
            codes[(size_t)k * MC_MAX_MOVES + i] = re[i].link.code;
            visits[(size_t)k * MC_MAX_MOVES + i] = st.N;
            if (q) q[(size_t)k * MC_MAX_MOVES + i] = st.Q;
        }
        if (lane == 0) n_legal[k] = E;
    }
}

struct NodeStatsOut {
    int found, n_legal, is_terminal;
    double terminal_value;
};

__global__ void node_stats_kernel(View V, int g, int tree, mc_state s, NodeStatsOut* out, uint16_t* codes, uint32_t* visits,
                                  double* q, float* priors) {
    const int lane = threadIdx.x & 31;
    const int t = 2 * g + (tree & 1);
    uint32_t node = az::ht_find(V, t, s);
    if (node == az::NONE) { if (lane == 0) { out->found = 0; out->n_legal = 0; out->is_terminal = 0; out->terminal_value = 0; } return; }
    const az::NodeHead h = az::load_head(&V.nodes[(size_t)t * V.NC + node]);
    const uint32_t info = h.info;
    const bool term = (info & az::INFO_TERMINAL) != 0;
    const int E = term ? 0 : (int)(info & 0xffffu);
    const az::Edge* re = V.edges + (size_t)t * V.EC + h.edge_off;
    for (int i = lane; i < E; i += 32) {
        const az::EdgeStat st = az::load_stat(&re[i]);
        codes[i] = re[i].link.code;
        visits[i] = st.N;
        q[i] = st.Q;
        priors[i] = st.P;
    }
    if (lane == 0) {
        out->found = 1; out->n_legal = E; out->is_terminal = term ? 1 : 0;
        out->terminal_value = term ? ((info & az::INFO_DECISIVE) ? -1.0 : -0.0) : 0.0;
    }
}

// Throughput mode, one move in every game (az_play_device): choose, record, play; then restart finished games.
__global__ void __launch_bounds__(128) play_device_kernel(View V) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < V.G; g += warps) play_device_one(V, g, lane);
}

// restart every finished game from the start position (block per game; clears both hash tables)
__global__ void __launch_bounds__(256) restart_finished_kernel(View V, mc_state start) {
    for (int g = blockIdx.x; g < V.G; g += gridDim.x) {
        if (V.game_result[g] == MC_ONGOING) continue;
        uint32_t* tab = V.ht + (size_t)(2 * g) * V.HC;
        for (int i = threadIdx.x; i < 2 * V.HC; i += blockDim.x) tab[i] = 0u;
        __syncthreads();
        if (threadIdx.x == 0) {
            V.game_state[g] = start;
            V.game_ply[g] = mc::white_to_move(start) ? 0 : 1;
            V.game_start_ply[g] = V.game_ply[g];
            az::hist_reset(V, g, start);
            for (int t = 2 * g; t < 2 * g + 2; ++t) { V.tree_nodes[t] = 0; V.tree_edges[t] = 0; V.tree_root[t] = az::NONE; }
            for (int j = 0; j < V.K; ++j) V.leaf_kind[g * V.K + j] = az::LEAF_NONE;
            V.game_result[g] = MC_ONGOING;
        }
        __syncthreads();
    }
}

// ---- recycling of unreachable plies (SURVEY.md 7.3 point 7) ------------------------------------------
// The reference never prunes its dicts.  Here nodes are stratified by ply (side to move and fullmove number are
// part of the key) and every future root of a game is a descendant of its current position, so a node whose ply is
// not greater than the current position's -- other than that position itself -- can never be reached again, by an
// edge or by transposition: dropping it changes no result.  Block per tree; a tree is compacted in place when it
// could not take `need_nodes` / `need_edges` more.  Order-preserving, so every block of data only moves towards
// the front; the hash-table region serves as the old -> new index map and is rebuilt afterwards.  Runs between
// searches (no simulation pending).
// Networks with sharp priors grow deep trees, and a deep node outlives many moves under that rule; the reference's dicts
// grow without bound there.  A tree that still cannot take the search after the exact compaction is compacted a second
// time (stage 2, counted in C_EVICTED) down to what its edges reach from the game's current position, level by level for as
// many levels as fit beside the coming search; links into a level that was cut read "never taken" again.  A node dropped
// there is expanded anew when the search comes back to it -- the one place where a tree may differ from the reference's,
// taken instead of failing the call with MCAZ_ECAPACITY.
MC_HD int ply_of_meta(uint32_t meta) { return 2 * (int)((meta >> 16) & 0xffu) + ((meta & 1u) ? 0 : 1); }

constexpr int RECYCLE_THREADS = 256;
__global__ void __launch_bounds__(RECYCLE_THREADS) recycle_kernel(View V, int need_nodes, int need_edges) {
    __shared__ uint32_t s_scan[RECYCLE_THREADS / 32][2];
    __shared__ uint32_t s_tot[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int WARPS = RECYCLE_THREADS / 32;
    for (int t = blockIdx.x; t < 2 * V.G; t += gridDim.x) {
      for (int stage = 1; stage <= 2; ++stage) {
        const uint32_t n = V.tree_nodes[t], m = V.tree_edges[t];
        __syncthreads();                                    // everyone has read the counters of this tree
        // stage 1 runs on the worst case of 40 edges per new node (the exact rule costs no result); stage 2, which may, only
        // when the nodes cannot fit or the edges cannot at twice this tree's own mean fan-out
        const uint32_t want_edges = stage == 1 ? (uint32_t)need_edges
                                               : min((uint32_t)need_edges, (uint32_t)need_nodes * (2u * ((m + n) / max(n, 1u)) + 4u));
        if (n + (uint32_t)need_nodes <= (uint32_t)V.NC && m + want_edges <= (uint32_t)V.EC) break;
        const int g = t >> 1;
        const mc_state cur = V.game_state[g];
        const int cur_ply = ply_of_meta(cur.meta);
        const size_t nb = (size_t)t * V.NC, eb = (size_t)t * V.EC;
        uint32_t* map = V.ht + (size_t)t * V.HC;            // [0, NC): old node -> new node; [NC, 2 NC): where new node j's edges were (HC >= 2 NC)
        uint32_t* old_off_of = map + V.NC;
        uint32_t cutoff = 0;
        if (stage == 2) {
            // marks in map[]: 1 on the current position, then level by level (an edge leads one ply down) level + 1 on the children
            // of the nodes of a level -- for as long as the levels fit what the arena can keep beside the coming search
            const uint32_t keep_nodes = (uint32_t)V.NC - min((uint32_t)V.NC, (uint32_t)need_nodes);
            const uint32_t keep_edges = (uint32_t)V.EC - min((uint32_t)V.EC, want_edges);
            for (uint32_t i = tid; i < n; i += RECYCLE_THREADS) {
                const az::Board4 board = az::load_board(&V.nodes[nb + i]);
                const uint32_t meta = V.nodes[nb + i].head.meta;
                map[i] = (meta == cur.meta && board.x == cur.pl0 && board.y == cur.pl1 && board.z == cur.pl2 && board.w == cur.white) ? 1u : 0u;
            }
            uint32_t kept_n = 0, kept_e = 0;
            for (uint32_t level = 1;; ++level) {
                __syncthreads();                            // the marks of this level are written, s_scan is free
                uint32_t a = 0, b = 0;
                for (uint32_t i = tid; i < n; i += RECYCLE_THREADS)
                    if (map[i] == level) {
                        const uint32_t info = V.nodes[nb + i].head.info;
                        ++a;
                        b += (info & az::INFO_TERMINAL) ? 0u : (info & 0xffffu);
                    }
                a = __reduce_add_sync(0xffffffffu, a);
                b = __reduce_add_sync(0xffffffffu, b);
                if (lane == 0) { s_scan[warp][0] = a; s_scan[warp][1] = b; }
                __syncthreads();
                a = 0; b = 0;
                for (int w = 0; w < WARPS; ++w) { a += s_scan[w][0]; b += s_scan[w][1]; }
                if (a == 0 || kept_n + a > keep_nodes || kept_e + b > keep_edges) break;      // this level and everything below it go
                kept_n += a; kept_e += b; cutoff = level;
                for (uint32_t i = tid; i < n; i += RECYCLE_THREADS) {
                    if (map[i] != level) continue;
                    const az::NodeHead h = az::load_head(&V.nodes[nb + i]);
                    if (h.info & az::INFO_TERMINAL) continue;
                    const uint32_t E = h.info & 0xffffu;
                    for (uint32_t k = 0; k < E; ++k) {
                        const uint32_t child = V.edges[eb + h.edge_off + k].link.child;
                        if (child != az::NONE) map[child] = level + 1u;
                    }
                }
            }
            __syncthreads();
        }
        if (tid == 0) { s_tot[0] = 0; s_tot[1] = 0; }
        __syncthreads();
        // ---- pass 1: node headers (chunks of 256 nodes: read, scan, barrier, write).  The scan over the edge counts gives
        // every live node its new first edge at once, so pass 2 can point the links at their children's final places.
        for (uint32_t base = 0; base < n; base += RECYCLE_THREADS) {
            const uint32_t i = base + tid;
            az::Board4 board{}; az::NodeHead head{};
            bool live = false;
            if (i < n) {
                board = az::load_board(&V.nodes[nb + i]); head = az::load_head(&V.nodes[nb + i]);
                const int ply = ply_of_meta(head.meta);
                live = stage == 2 ? (map[i] != 0u && map[i] <= cutoff)
                                  : ply > cur_ply || (head.meta == cur.meta && board.x == cur.pl0 && board.y == cur.pl1 && board.z == cur.pl2 && board.w == cur.white);
            }
            const uint32_t E = live && !(head.info & az::INFO_TERMINAL) ? (head.info & 0xffffu) : 0u;
            // block-wide exclusive scan of (live, E)
            uint32_t a = live ? 1u : 0u, b = E;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t ya = __shfl_up_sync(0xffffffffu, a, o), yb = __shfl_up_sync(0xffffffffu, b, o);
                if (lane >= o) { a += ya; b += yb; }
            }
            if (lane == 31) { s_scan[warp][0] = a; s_scan[warp][1] = b; }
            __syncthreads();
            uint32_t pa = s_tot[0], pb = s_tot[1];
            for (int w = 0; w < warp; ++w) { pa += s_scan[w][0]; pb += s_scan[w][1]; }
            const uint32_t new_i = pa + a - (live ? 1u : 0u), new_off = pb + b - E;
            __syncthreads();
            if (tid == RECYCLE_THREADS - 1) { s_tot[0] = pa + a; s_tot[1] = pb + b; }
            if (i < n) map[i] = live ? new_i : az::NONE;
            if (live) {                                     // new_i <= i and every older chunk is already moved
                az::store_node(&V.nodes[nb + new_i], board, head.meta, new_off, head.info);
                old_off_of[new_i] = head.edge_off;
            }
            __syncthreads();
        }
        const uint32_t n_live = s_tot[0], m_live = s_tot[1];
        __syncthreads();
        // ---- pass 2: edge records, one node per warp and round: all warps read, barrier, all write.  Order-preserving: a node's
        // new place ends where its old one does at the latest, and everything before that has been read by then.
        for (uint32_t base = 0; base < n_live; base += WARPS) {
            const uint32_t j = base + warp;
            uint32_t old_off = 0, new_off = 0, E = 0;
            az::EdgeStat st[3]; az::EdgeLink lk[3];
            if (j < n_live) {
                const az::NodeHead h = az::load_head(&V.nodes[nb + j]);
                old_off = old_off_of[j]; new_off = h.edge_off;
                E = (h.info & az::INFO_TERMINAL) ? 0u : (h.info & 0xffffu);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint32_t i = lane + 32 * k;
                if (i < E) { st[k] = az::load_stat(&V.edges[eb + old_off + i]); lk[k] = az::load_link(&V.edges[eb + old_off + i]); }
            }
            __syncthreads();                                  // this round's reads are done
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint32_t i = lane + 32 * k;
                if (i < E) {
                    az::Edge* e = &V.edges[eb + new_off + i];
                    uint32_t child = lk[k].child, child_off = 0;
                    if (child != az::NONE) {                  // children are deeper: live under the exact rule, their headers final since pass 1;
                        child = map[child];                   // stage 2 may have cut the child's level: the link is then as never taken
                        child_off = child != az::NONE ? V.nodes[nb + child].head.edge_off : 0u;
                    }
                    *reinterpret_cast<uint4*>(&e->stat) = make_uint4((uint32_t)__double2loint(st[k].Q), (uint32_t)__double2hiint(st[k].Q), st[k].N, __float_as_uint(st[k].P));
                    *reinterpret_cast<uint4*>(&e->link) = make_uint4(child, child_off, lk[k].child_info, (uint32_t)lk[k].code | ((uint32_t)lk[k].vl << 16));
                }
            }
            __syncthreads();                                  // this round's writes are done
        }
        __syncthreads();
        // ---- roots, then the hash table
        if (tid == 0) {
            const uint32_t r = V.tree_root[t];
            V.tree_root[t] = (r != az::NONE && r < n) ? map[r] : az::NONE;
            V.tree_nodes[t] = n_live;
            V.tree_edges[t] = m_live;
            atomicAdd(&V.counters[stage == 2 ? az::C_EVICTED : az::C_RECYCLED], (unsigned long long)(n - n_live));
        }
        __syncthreads();
        uint32_t* tab = V.ht + (size_t)t * V.HC;
        for (int i = tid; i < V.HC; i += RECYCLE_THREADS) tab[i] = 0u;
        __syncthreads();
        const uint32_t mask = (uint32_t)V.HC - 1u;
        for (uint32_t i = tid; i < n_live; i += RECYCLE_THREADS) {
            const mc_state s = az::state_of(az::load_board(&V.nodes[nb + i]), V.nodes[nb + i].head.meta);
            uint32_t h = az::hash_state(s) & mask;
            while (atomicCAS(&tab[h], 0u, i + 1u) != 0u) h = (h + 1u) & mask;
        }
        __syncthreads();
      }
    }
}

// n samples of the root noise a root with E edges gets (measurement hook: the sampler itself, outside any search)
__global__ void __launch_bounds__(128) sample_root_noise_kernel(unsigned long long seed, float alpha, int E, int n, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += warps) {
        double gam[3] = {0.0, 0.0, 0.0};
        const double gsum = az::root_noise_gammas(seed, k & 0xffff, (unsigned long long)(k >> 16), E, lane, alpha, gam);
        for (int i = lane, kk = 0; i < E && kk < 3; i += 32, ++kk) out[(size_t)k * E + i] = gam[kk] / gsum;
    }
}

int warp_grid(int n_warps, int block) {
    int per = block / 32;
    int want = (n_warps + per - 1) / per;
    return std::max(1, std::min(want, num_sms() * 16));
}

int engine_check_errors(az_engine* e) {
    int flag = 0;
    MCAZ_CUDA(cudaMemcpyAsync(&flag, e->v.error_flag, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    MCAZ_CUDA(cudaStreamSynchronize(e->stream));
    if (flag == 0) return MCAZ_OK;
    cudaMemsetAsync(e->v.error_flag, 0, sizeof(int), e->stream);
    std::string msg = "engine error flags:";
    if (flag & az::ERR_NODE_CAP) msg += " node arena full";
    if (flag & az::ERR_EDGE_CAP) msg += " edge arena full";
    if (flag & az::ERR_HASH_CAP) msg += " hash table full";
    if (flag & az::ERR_DEPTH) msg += " path deeper than MAX_DEPTH";
    if (flag & az::ERR_ILLEGAL) msg += " illegal move or bad game id in az_play";
    if (flag & 7)
        msg += e->cfg.recycle ? " (a tree needs more than node_capacity even after compaction: raise az_config.node_capacity)"
                              : " (trees are kept whole: reset them per episode -- az_reset_games / az_reset_trees, the reference's "
                                "MonteCarloInit callback -- or raise az_config.node_capacity / max_sims_per_move)";
    return fail((flag & az::ERR_ILLEGAL) && !(flag & 7) ? MCAZ_EINVAL : MCAZ_ECAPACITY, msg);
}

static mc_state start_state() {
    mc_state s;
    mc_state_from_fen("2nbk/2ppp/5/5/PPP2/KBN2 w 0 1", &s);
    return s;
}

template <typename T>
static int dev_alloc(az_engine* e, T** p, size_t n, bool zero = true) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    MCAZ_CUDA(cudaMalloc(&q, bytes));
    if (zero) MCAZ_CUDA(cudaMemset(q, 0, bytes));
    e->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return MCAZ_OK;
}

}  // namespace mcaz

extern "C" {

void az_default_config(az_config* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->n_games = 1;
    c->max_sims_per_move = 36;      // app/base.py:25
    c->cpuct = 1.0f;                // exp/agent.py:96
    c->tau_change = 6;              // exp/agent.py:97
    c->dirichlet_alpha = 0.6f;      // exp/agent.py:82
    c->dirichlet_epsilon = 0.25f;
    c->numpy1_dtype_flow = 0;
    c->device_rng = 0;
    c->seed = 0;
    mc_default_rules(&c->rules);
    c->network = 0;
    c->leaves_per_step = 1;
    c->own_stream = 0;
    c->eval_cache_log2 = 0;
    c->free_sims = 0;
    c->recycle = 0;
    c->lookahead_rows = 0;
    c->fp8_convolutions = 0;
    c->defer_rows = 0;
}

int az_create(const az_config* cfg, az_engine** out) {
    if (!cfg || !out) return fail(MCAZ_EINVAL, "az_create: null argument");
    if (cfg->n_games <= 0 || cfg->max_sims_per_move <= 0) return fail(MCAZ_EINVAL, "az_create: n_games and max_sims_per_move must be positive");
    // a game line of 2 * max_fullmoves plies (+ 2 for a start position with black to move) must fit the per-game records:
    // replay tuples, simulation paths and the repetition history are MAX_DEPTH = HIST = 64 entries each
    if (cfg->rules.max_fullmoves < 1 || 2 * cfg->rules.max_fullmoves + 2 > az::MAX_DEPTH)
        return fail(MCAZ_EINVAL, "az_create: rules.max_fullmoves must be in [1, 31] (per-game records hold 64 plies)");
    if (int rc = require_device()) return rc;
    az_engine* e = new az_engine();
    e->cfg = *cfg;
    cudaGetDevice(&e->device);
    if (cfg->own_stream && cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete e;
        return fail(MCAZ_ECUDA, "az_create: cudaStreamCreateWithFlags failed");
    }
    View& V = e->v;
    V.G = cfg->n_games;
    V.K = cfg->leaves_per_step > 0 ? cfg->leaves_per_step : 1;
    if (V.K > az::MAX_LEAVES) { delete e; return fail(MCAZ_EINVAL, "az_create: leaves_per_step > 16"); }
    // <= 1 new node per simulation, <= 31 searches per tree under the 30-move cap (+ roots)
    // without recycling a tree keeps every node of the game; with it only the plies still ahead (measured ~1.5 x sims live)
    long long nc = cfg->node_capacity > 0 ? cfg->node_capacity : (long long)cfg->max_sims_per_move * (cfg->recycle ? 6 : 31) + 64;
    long long ec = cfg->edge_capacity > 0 ? cfg->edge_capacity : nc * 14;
    if (nc > 0x3fffffff || ec > 0x7fffffff) { delete e; return fail(MCAZ_EINVAL, "az_create: arena too large"); }
    V.NC = (int)nc;
    V.EC = (int)ec;
    int hc = 64;
    while (hc < 2 * V.NC) hc <<= 1;
    V.HC = hc;
    V.cpuct = cfg->cpuct; V.eps = cfg->dirichlet_epsilon; V.alpha = cfg->dirichlet_alpha;
    V.numpy1 = cfg->numpy1_dtype_flow; V.tau_change = cfg->tau_change; V.rules = cfg->rules; V.seed = cfg->seed;
    V.device_rng = 0; V.sim_counter = 0;   // set per launch by az_search
    const size_t G = V.G, T = 2 * G, N = T * V.NC, E = T * V.EC, S = G * V.K;   // S: leaf slots
    // look-ahead rows need the exact cache (their results live nowhere else) and the sequential search
    if (cfg->lookahead_rows < 0 || cfg->lookahead_rows > 65536) { delete e; return fail(MCAZ_EINVAL, "az_create: lookahead_rows must be in [0, 65536]"); }
    if (cfg->defer_rows < 0 || cfg->defer_rows > 255) { delete e; return fail(MCAZ_EINVAL, "az_create: defer_rows must be in [0, 255]"); }
    e->lookahead_rows = (cfg->lookahead_rows > 0 && cfg->eval_cache_log2 > 0 && cfg->network && V.K == 1) ? cfg->lookahead_rows : 0;
    const size_t R = S + (size_t)e->lookahead_rows;                              // rows of the dense batch
    V.row_cap = (int)R; V.spec_rows = 0; V.noise_block = nullptr; V.noise_budget = 0; V.pending_count = nullptr;
    int rc = MCAZ_OK;
#define A(ptr, n) if (!rc) rc = dev_alloc(e, &ptr, (n))
    A(V.game_state, G); A(V.game_result, G); A(V.game_ply, G); A(V.game_hist, G * az::HIST); A(V.game_hist_len, G); A(V.game_start_ply, G);
    A(V.tree_nodes, T); A(V.tree_edges, T); A(V.tree_root, T);
    A(V.nodes, N); A(V.edges, E);
    A(V.ht, T * V.HC);
    A(V.path_len, S); A(V.path_edge, S * az::MAX_DEPTH); A(V.path_node, S * az::MAX_DEPTH);
    A(V.leaf_node, S); A(V.leaf_kind, S); A(V.leaf_value, S);
    A(V.tokens, R * MC_TOKENS); A(V.clocks, R); A(V.needs_eval, S); A(V.leaf_states, S);
    V.row_state = nullptr; V.seen = nullptr; V.seen_mask = 0;
    if (e->lookahead_rows > 0) {
        A(V.row_state, R);
        A(V.seen, (size_t)1 << cfg->eval_cache_log2);
        V.seen_mask = (1u << cfg->eval_cache_log2) - 1u;
    }
    A(e->d_pending, 2);
    V.row_eff = nullptr; V.slot_row = nullptr; V.defer_thr = 0; V.check_deferred = 0;
    e->defer_rows = (cfg->defer_rows > 0 && cfg->network && V.K == 1 && e->lookahead_rows == 0) ? cfg->defer_rows : 0;
    if (e->defer_rows) { A(V.row_eff, 2); A(V.slot_row, S); }
    A(V.counters, AZ_NUM_COUNTERS); A(V.error_flag, 1);
    A(V.sim_serial, G); A(V.move_serial, G); A(V.sims_left, G); A(V.row_count, 2); A(V.row_slot, R);
    V.sims_per_move = cfg->max_sims_per_move; V.new_budget = -1; V.free_max = 1; V.async_play = 0; V.compact = 0; V.parity = 0;
    V.cache = nullptr; V.cache_mask = 0; V.cache_epoch = 1;
    if (cfg->eval_cache_log2 < 0 || cfg->eval_cache_log2 > 28) rc = fail(MCAZ_EINVAL, "az_create: eval_cache_log2 must be in [0, 28]");
    if (!rc && cfg->eval_cache_log2 > 0 && cfg->network && V.K == 1) {
        A(e->d_cache, (size_t)1 << cfg->eval_cache_log2);
        e->cache_mask = (1u << cfg->eval_cache_log2) - 1u;
    }
    A(e->d_noise, G * MC_MAX_MOVES); A(e->d_noise_used, G);
    A(e->d_logits, S * MC_NUM_ACTIONS); A(e->d_values, S); A(e->d_priors, S * MC_MAX_MOVES);
    A(e->d_record, G * az::MAX_DEPTH);
    e->replay_capacity = G * az::MAX_DEPTH;
    A(e->d_replay, e->replay_capacity); A(e->d_replay_count, 1);
#undef A
    V.record = e->d_record; V.replay = e->d_replay; V.replay_count = e->d_replay_count;
    V.replay_cap = (unsigned long long)e->replay_capacity;
    V.weights_version = 0u;
    if (!rc && cfg->network) rc = network_create(e);
    if (rc) { az_destroy(e); return rc; }
    *out = e;
    return az_reset_games(e, nullptr, V.G, nullptr);
}

int az_destroy(az_engine* e) {
    if (!e) return MCAZ_OK;
    cudaDeviceSynchronize();
    if (e->net) network_destroy(e);
    for (auto& ev : e->tree_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    for (void* p : e->allocs) cudaFree(p);
    e->scratch.release();
    if (e->cfg.own_stream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return MCAZ_OK;
}

int az_set_weights(az_engine* e, const float* flat, size_t n) {
    if (!e || !flat) return fail(MCAZ_EINVAL, "az_set_weights: null argument");
    if (n != (size_t)AZ_NUM_WEIGHT_FLOATS) return fail(MCAZ_EINVAL, "az_set_weights: expected AZ_NUM_WEIGHT_FLOATS floats");
    if (!e->net) return fail(MCAZ_ESTATE, "az_set_weights: engine was created with network = 0");
    e->scratch.begin();
    In<float> in;
    if (int rc = in.init(flat, n, e->stream, e->scratch)) return rc;
    int rc = network_set_weights(e, in.ptr);
    if (!rc) MCAZ_CUDA(cudaStreamSynchronize(e->stream));
    if (++e->cache_epoch == 0) e->cache_epoch = 1;   // cached evaluations belong to the old weights (0 = never written)
    e->v.weights_version += 1u;                      // a learner with its own numbering overrides this (az_set_weights_version)
    return rc;
}

int az_set_weights_version(az_engine* e, uint32_t version) {
    if (!e) return fail(MCAZ_EINVAL, "az_set_weights_version: null engine");
    e->v.weights_version = version;
    return MCAZ_OK;
}

int az_reset_games(az_engine* e, const int32_t* game_ids, int n, const mc_state* states) {
    if (!e || n < 0) return fail(MCAZ_EINVAL, "az_reset_games: bad argument");
    if (n == 0) return MCAZ_OK;
    e->scratch.begin();
    In<int32_t> ids; In<mc_state> st;
    if (int rc = ids.init(game_ids, n, e->stream, e->scratch)) return rc;
    if (int rc = st.init(states, n, e->stream, e->scratch)) return rc;
    reset_games_kernel<<<std::min(n, num_sms() * 8), 256, 0, e->stream>>>(e->v, ids.ptr, n, st.ptr, start_state());
    MCAZ_CHECK_LAUNCH();
    e->launches++;
    e->leaf_pending = false;
    MCAZ_CUDA(cudaStreamSynchronize(e->stream));
    return MCAZ_OK;
}

int az_reset_trees(az_engine* e, const int32_t* tree_ids, int n) {
    if (!e || n < 0 || (n > 0 && !tree_ids)) return fail(MCAZ_EINVAL, "az_reset_trees: bad argument");
    if (e->leaf_pending) return fail(MCAZ_ESTATE, "az_reset_trees: a simulation is pending (call az_backup first)");
    if (n == 0) return MCAZ_OK;
    e->scratch.begin();
    In<int32_t> ids;
    if (int rc = ids.init(tree_ids, n, e->stream, e->scratch)) return rc;
    reset_trees_kernel<<<std::min(n, num_sms() * 8), 256, 0, e->stream>>>(e->v, ids.ptr, n);
    MCAZ_CHECK_LAUNCH();
    e->launches++;
    MCAZ_CUDA(cudaStreamSynchronize(e->stream));
    return MCAZ_OK;
}

int az_set_positions(az_engine* e, const int32_t* game_ids, int n, const mc_state* states, const int32_t* tree_of_game) {
    if (!e || n < 0 || (n > 0 && !states)) return fail(MCAZ_EINVAL, "az_set_positions: bad argument");
    if (n == 0) return MCAZ_OK;
    e->scratch.begin();
    In<int32_t> ids; In<mc_state> st; In<int32_t> tg;
    if (int rc = ids.init(game_ids, n, e->stream, e->scratch)) return rc;
    if (int rc = st.init(states, n, e->stream, e->scratch)) return rc;
    if (int rc = tg.init(tree_of_game, n, e->stream, e->scratch)) return rc;
    set_positions_kernel<<<std::max(1, std::min((n + 127) / 128, num_sms() * 8)), 128, 0, e->stream>>>(e->v, ids.ptr, n, st.ptr, tg.ptr);
    MCAZ_CHECK_LAUNCH();
    e->launches++;
    e->leaf_pending = false;
    MCAZ_CUDA(cudaStreamSynchronize(e->stream));
    return MCAZ_OK;
}

int az_select_expand(az_engine* e, const double* root_noise, uint8_t* noise_used) {
    if (!e) return fail(MCAZ_EINVAL, "az_select_expand: null engine");
    if (e->leaf_pending) return fail(MCAZ_ESTATE, "az_select_expand: previous simulation not backed up");
    View V = e->v;
    const double* noise = nullptr;
    if (root_noise) {
        if (is_device_pointer(root_noise)) noise = root_noise;
        else {
            MCAZ_CUDA(cudaMemcpyAsync(e->d_noise, root_noise, (size_t)V.G * MC_MAX_MOVES * sizeof(double), cudaMemcpyHostToDevice, e->stream));
            noise = e->d_noise;
        }
    } else if (e->cfg.device_rng && V.eps > 0.0f) {
        V.device_rng = 1;                     // Dirichlet noise drawn inside the kernel (Philox)
        V.sim_counter = e->sim_counter;
    }
    e->sim_counter++;
    select_expand_kernel<<<warp_grid(V.G, 128), 128, 0, e->stream>>>(V, noise, e->d_noise_used);
    MCAZ_CHECK_LAUNCH();
    e->launches++;
    e->leaf_pending = true;
    if (noise_used) {
        if (is_device_pointer(noise_used)) MCAZ_CUDA(cudaMemcpyAsync(noise_used, e->d_noise_used, V.G, cudaMemcpyDeviceToDevice, e->stream));
        else { MCAZ_CUDA(cudaMemcpyAsync(noise_used, e->d_noise_used, V.G, cudaMemcpyDeviceToHost, e->stream)); MCAZ_CUDA(cudaStreamSynchronize(e->stream)); }
    }
    return MCAZ_OK;
}

int az_leaf_batch(az_engine* e, const uint8_t** tokens, const float** clocks, const uint8_t** needs_eval,
                  const mc_state** leaf_states, int* n_slots) {
    if (!e) return fail(MCAZ_EINVAL, "az_leaf_batch: null engine");
    if (tokens) *tokens = e->v.tokens;
    if (clocks) *clocks = e->v.clocks;
    if (needs_eval) *needs_eval = e->v.needs_eval;
    if (leaf_states) *leaf_states = e->v.leaf_states;
    if (n_slots) *n_slots = e->v.G * e->v.K;
    return MCAZ_OK;
}

int az_backup(az_engine* e, const float* logits, const float* values, const float* priors) {
    if (!e) return fail(MCAZ_EINVAL, "az_backup: null engine");
    if (!e->leaf_pending) return fail(MCAZ_ESTATE, "az_backup: no simulation pending (call az_select_expand first)");
    if (!values || (!logits && !priors)) return fail(MCAZ_EINVAL, "az_backup: values and one of logits/priors are required");
    const View& V = e->v;
    const float *lg = logits, *vl = values, *pr = priors;
    if (logits && !is_device_pointer(logits)) {
        MCAZ_CUDA(cudaMemcpyAsync(e->d_logits, logits, (size_t)V.G * V.K * MC_NUM_ACTIONS * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        lg = e->d_logits;
    }
    if (!is_device_pointer(values)) {
        MCAZ_CUDA(cudaMemcpyAsync(e->d_values, values, (size_t)V.G * V.K * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        vl = e->d_values;
    }
    if (priors && !is_device_pointer(priors)) {
        MCAZ_CUDA(cudaMemcpyAsync(e->d_priors, priors, (size_t)V.G * V.K * MC_MAX_MOVES * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        pr = e->d_priors;
    }
    backup_kernel<<<warp_grid(V.G, 128), 128, 0, e->stream>>>(V, lg, vl, pr);
    MCAZ_CHECK_LAUNCH();
    e->launches++;
    e->leaf_pending = false;
    return MCAZ_OK;
}

int az_eval_backup(az_engine* e) {
    if (!e) return fail(MCAZ_EINVAL, "az_eval_backup: null engine");
    if (!e->net) return fail(MCAZ_ESTATE, "az_eval_backup: engine was created with network = 0");
    if (!e->leaf_pending) return fail(MCAZ_ESTATE, "az_eval_backup: no simulation pending (call az_select_expand first)");
    const View& V = e->v;
    if (int rc = network_forward_search(e, V, e->d_values)) return rc;     // priors land in the new nodes' edges
    backup_kernel<<<warp_grid(V.G, 128), 128, 0, e->stream>>>(V, nullptr, e->d_values, nullptr);
    MCAZ_CHECK_LAUNCH();
    e->launches++;
    e->leaf_pending = false;
    return MCAZ_OK;
}

// View of one az_search / az_selfplay launch: dense rows, the launch's row counter, the evaluation cache.
static View search_view(az_engine* e) {
    View V = e->v;
    V.device_rng = e->cfg.device_rng;
    V.compact = V.K == 1;
    V.parity = V.compact ? (e->parity ^= 1) : 0;
    V.cache = e->d_cache; V.cache_mask = e->cache_mask; V.cache_epoch = e->cache_epoch;
    V.free_max = V.K == 1 ? (e->cfg.free_sims > 0 ? e->cfg.free_sims : 1) : 1;
    V.spec_rows = (V.compact && V.cache) ? e->lookahead_rows : 0;
    return V;
}

static int search_launch(az_engine* e, const View& V) {
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (e->tree_profiling) {
        if (e->tree_events_used == e->tree_events.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            e->tree_events.emplace_back(a, b);
        }
        ev0 = e->tree_events[e->tree_events_used].first; ev1 = e->tree_events[e->tree_events_used].second;
        e->tree_events_used++;
        cudaEventRecord(ev0, e->stream);
    }
    if (V.spec_rows > 0)     // few games: the variant that also queues look-ahead rows
        search_step_kernel<true><<<warp_grid(V.G, 128), 128, 0, e->stream>>>(V, e->d_values, start_state());
    else
        search_step_kernel<false><<<warp_grid(V.G, 128), 128, 0, e->stream>>>(V, e->d_values, start_state());   // priors: written by the policy head
    MCAZ_CHECK_LAUNCH();
    if (ev1) cudaEventRecord(ev1, e->stream);
    e->launches++;
    return MCAZ_OK;
}

// n_batches network batches of the search: step launch -> network forward -> ... -> a closing step launch that
// only backs up.  new_budget >= 0: every game gets that many simulations (az_search); async: games choose and
// play their own moves (az_selfplay).
// (An overlap variant -- the chains of free simulations running on a side stream while the tower evaluates the
// batch -- was measured and dropped: the B200 runs this workload at its power cap, the tree work costs the same
// energy wherever it runs, and the tower slowed down by more than the hidden time; tools/sweep_modes.py.)
// recycle = 1: compact the trees that could not take `need_nodes` more nodes (one per descent).  Called where no
// simulation is pending: after the moves of az_play / az_play_device and at the start of az_search / az_selfplay.
static int maybe_recycle(az_engine* e, long long need) {
    if (!e->cfg.recycle || need <= 0) return MCAZ_OK;
    const int need_nodes = (int)std::min<long long>(need, e->v.NC);
    const int need_edges = (int)std::min<long long>((long long)need_nodes * 40, e->v.EC);
    recycle_kernel<<<std::min(2 * e->v.G, num_sms() * 8), RECYCLE_THREADS, 0, e->stream>>>(e->v, need_nodes, need_edges);
    MCAZ_CHECK_LAUNCH();
    e->launches++;
    return MCAZ_OK;
}

// noise != nullptr: caller-supplied root noise, one row per simulation of the call (az_search_noise).
// Engines with look-ahead rows (few games) chain a game's simulations inside one launch until one needs the network, and
// every such call -- like every call with caller noise -- stops as soon as no game is waiting for a row: the device
// counts the waiting games per launch and the host reads that one word back.
static int run_search(az_engine* e, int n_batches, int new_budget, bool async, int sims_per_move, const double* noise = nullptr) {
    {
        const long long per_game = (long long)n_batches * (e->v.K == 1 ? std::max(1, e->cfg.free_sims) : e->v.K);
        if (int rc = maybe_recycle(e, new_budget >= 0 ? std::min<long long>(per_game, new_budget) : per_game)) return rc;
    }
    const bool chained = e->lookahead_rows > 0 && !async && e->v.K == 1 && new_budget >= 0;
    const bool early_exit = chained || noise != nullptr;
    // deferred rows (az_config.defer_rows, az_selfplay only): every batch but the call's last may leave a short last tile pair to the
    // next one; the last batch is evaluated whole, so that no leaf is pending when the call returns.  A lock-step search
    // (az_search) gives a game exactly the launches its simulations need in the worst case -- there a wait would cost extra batches
    // at the end of the move (measured: 3.43 -> 3.2-3.3 M simulations/s), so it never defers.
    const int defer = (e->defer_rows > 0 && async && !early_exit) ? e->defer_rows : 0;
    int prev_defer = 0;
    for (int s = 0; s <= n_batches; ++s) {
        View V = search_view(e);
        V.new_budget = s == 0 ? new_budget : -1;
        V.async_play = async ? 1 : 0;
        V.sims_per_move = sims_per_move;
        if (noise) { V.noise_block = noise; V.noise_budget = new_budget; }
        if (chained) V.free_max = new_budget + 1;
        if (early_exit) V.pending_count = e->d_pending;
        if (s == n_batches) V.free_max = 0;       // closing launch: back up what the last batch evaluated, start nothing
        V.check_deferred = prev_defer > 0 ? 1 : 0;
        V.defer_thr = (s + 1 < n_batches) ? defer : 0;
        prev_defer = V.defer_thr;
        if (int rc = search_launch(e, V)) return rc;
        if (s == n_batches) break;
        if (V.defer_thr > 0) { cap_rows_kernel<<<1, 1, 0, e->stream>>>(V); MCAZ_CHECK_LAUNCH(); e->launches++; }
        if (early_exit && V.compact) {
            uint32_t waiting = 0;
            MCAZ_CUDA(cudaMemcpyAsync(&waiting, e->d_pending + V.parity, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
            MCAZ_CUDA(cudaStreamSynchronize(e->stream));
            if (waiting == 0) {                   // every game has spent its budget and nothing is left to back up
                if (V.spec_rows > 0) {
                    untag_rows_kernel<<<std::max(1, std::min((V.row_cap + 255) / 256, num_sms())), 256, 0, e->stream>>>(V);
                    MCAZ_CHECK_LAUNCH();
                    e->launches++;
                }
                break;
            }
        }
        if (V.spec_rows > 0) {                        // look-ahead rows were queued by position: their tokens and clocks
            tokenize_lookahead_kernel<<<std::max(1, std::min((V.row_cap + 3) / 4, num_sms() * 4)), 128, 0, e->stream>>>(V);
            MCAZ_CHECK_LAUNCH();
            e->launches++;
        }
        if (int rc = network_forward_search(e, V, e->d_values)) return rc;
    }
    return engine_check_errors(e);
}

int az_search(az_engine* e, int n_sims) {
    if (!e || n_sims < 0) return fail(MCAZ_EINVAL, "az_search: bad argument");
    if (!e->net) return fail(MCAZ_ESTATE, "az_search: engine was created with network = 0 (use az_select_expand / az_backup)");
    if (e->leaf_pending) return fail(MCAZ_ESTATE, "az_search: a simulation is pending (call az_backup first)");
    // Every game starts at least one simulation per batch, so n_sims batches start them all and the closing
    // launch backs up the rest.
    return run_search(e, n_sims, n_sims, false, e->cfg.max_sims_per_move);
}

int az_search_noise(az_engine* e, int n_sims, const double* root_noise) {
    if (!e || n_sims < 0 || (n_sims > 0 && !root_noise)) return fail(MCAZ_EINVAL, "az_search_noise: bad argument");
    if (!e->net) return fail(MCAZ_ESTATE, "az_search_noise: engine was created with network = 0");
    if (e->leaf_pending) return fail(MCAZ_ESTATE, "az_search_noise: a simulation is pending (call az_backup first)");
    if (n_sims == 0) return MCAZ_OK;
    const View& V = e->v;
    const size_t per_sim = (size_t)V.G * MC_MAX_MOVES;
    e->scratch.begin();
    In<double> noise;
    if (int rc = noise.init(root_noise, per_sim * n_sims, e->stream, e->scratch)) return rc;
    if (V.K == 1)
        // the search loop of az_search with the caller's noise: simulation i of every game mixes row i (exact cache, dense
        // rows, look-ahead rows and early exit as configured); same trees as n_sims rounds of select / evaluate / backup
        return run_search(e, n_sims, n_sims, false, e->cfg.max_sims_per_move, noise.ptr);
    const int grid = warp_grid(V.G, 128);
    for (int s = 0; s < n_sims; ++s) {
        // exp/agent.py:54-88 once per game: descent with this simulation's noise, evaluation, backup
        select_expand_kernel<<<grid, 128, 0, e->stream>>>(V, noise.ptr + per_sim * s, e->d_noise_used);
        MCAZ_CHECK_LAUNCH();
        if (int rc = network_forward_search(e, V, e->d_values)) return rc;
        backup_kernel<<<grid, 128, 0, e->stream>>>(V, nullptr, e->d_values, nullptr);
        MCAZ_CHECK_LAUNCH();
        e->launches += 2;
    }
    return engine_check_errors(e);
}

int az_selfplay(az_engine* e, int n_steps, int sims_per_move) {
    if (!e || n_steps < 0 || sims_per_move <= 0) return fail(MCAZ_EINVAL, "az_selfplay: bad argument");
    if (!e->net) return fail(MCAZ_ESTATE, "az_selfplay: engine was created with network = 0");
    if (!e->cfg.device_rng) return fail(MCAZ_ESTATE, "az_selfplay: needs device_rng = 1 (moves are chosen on the device)");
    if (e->v.K != 1) return fail(MCAZ_ESTATE, "az_selfplay: leaves_per_step must be 1");
    if (sims_per_move > e->cfg.max_sims_per_move) return fail(MCAZ_EINVAL, "az_selfplay: sims_per_move exceeds max_sims_per_move (arena size)");
    if (e->leaf_pending) return fail(MCAZ_ESTATE, "az_selfplay: a simulation is pending (call az_backup first)");
    // recycle = 1: the arenas hold a few moves' worth of nodes, and games keep moving inside this call.  Run it in segments
    // in which a tree takes at most sims_per_move new nodes; every segment starts by compacting the trees that could not
    // take that many (run_search) and ends with the closing launch, after which no simulation is pending.  Same games as
    // one long call: only the order of a game's own simulations matters.
    const int fm = std::max(1, e->cfg.free_sims);
    const int seg = e->cfg.recycle ? std::max(1, sims_per_move / fm) : std::max(1, n_steps);
    for (int done = 0; done < n_steps; done += seg)
        if (int rc = run_search(e, std::min(seg, n_steps - done), -1, true, sims_per_move)) return rc;
    return MCAZ_OK;
}

int az_sample_root_noise(uint64_t seed, float alpha, int n_edges, int n, double* out) {
    if (n < 0 || n_edges < 1 || n_edges > MC_MAX_MOVES || !(alpha > 0.f) || (n > 0 && !out)) return fail(MCAZ_EINVAL, "az_sample_root_noise: bad argument");
    if (int rc = require_device()) return rc;
    if (n == 0) return MCAZ_OK;
    cudaStream_t st = 0;
    Scratch& sc = thread_scratch();
    sc.begin();
    Out<double> o;
    if (int rc = o.init(out, (size_t)n * n_edges, st, sc)) return rc;
    sample_root_noise_kernel<<<warp_grid(n, 128), 128, 0, st>>>((unsigned long long)seed, alpha, n_edges, n, o.ptr);
    MCAZ_CHECK_LAUNCH();
    if (int rc = o.finish(st)) return rc;
    MCAZ_CUDA(cudaStreamSynchronize(st));
    return MCAZ_OK;
}

int az_profile_tree(az_engine* e, int on, double* total_ms, int* n_launches) {
    if (!e) return fail(MCAZ_EINVAL, "az_profile_tree: null engine");
    if (total_ms) {
        cudaStreamSynchronize(e->stream);
        double total = 0;
        for (size_t i = 0; i < e->tree_events_used; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, e->tree_events[i].first, e->tree_events[i].second);
            total += ms;
        }
        *total_ms = total;
        if (n_launches) *n_launches = (int)e->tree_events_used;
    }
    e->tree_events_used = 0;
    e->tree_profiling = on != 0;
    return MCAZ_OK;
}

int az_root_stats(az_engine* e, const int32_t* game_ids, int n, uint16_t* codes, uint32_t* visits, double* q, int32_t* n_legal) {
    if (!e || n < 0 || (n > 0 && (!codes || !visits || !n_legal))) return fail(MCAZ_EINVAL, "az_root_stats: bad argument");
    if (n == 0) return MCAZ_OK;
    const bool trace = getenv("MCAZ_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    e->scratch.begin();
    In<int32_t> ids; Out<uint16_t> oc; Out<uint32_t> ov; Out<double> oq; Out<int32_t> on;
    if (int rc = ids.init(game_ids, n, e->stream, e->scratch)) return rc;
    if (int rc = oc.init(codes, (size_t)n * MC_MAX_MOVES, e->stream, e->scratch, true)) return rc;
    if (int rc = ov.init(visits, (size_t)n * MC_MAX_MOVES, e->stream, e->scratch, true)) return rc;
    if (int rc = oq.init(q, q ? (size_t)n * MC_MAX_MOVES : 0, e->stream, e->scratch, true)) return rc;
    if (int rc = on.init(n_legal, n, e->stream, e->scratch)) return rc;
    double t1 = now();
    root_stats_kernel<<<warp_grid(n, 128), 128, 0, e->stream>>>(e->v, ids.ptr, n, oc.ptr, ov.ptr, oq.ptr, on.ptr);
    MCAZ_CHECK_LAUNCH();
    e->launches++;
    if (trace) cudaStreamSynchronize(e->stream);
    double t2 = now();
    if (int rc = oc.finish(e->stream)) return rc;
    if (int rc = ov.finish(e->stream)) return rc;
    if (int rc = oq.finish(e->stream)) return rc;
    if (int rc = on.finish(e->stream)) return rc;
    int rc = engine_check_errors(e);
    double t3 = now();
    if (trace) fprintf(stderr, "[mcaz] root_stats n=%d: staging %.2f ms, kernel %.2f ms, copies+sync %.2f ms\n", n, t1 - t0, t2 - t1, t3 - t2);
    return rc;
}

int az_node_stats(az_engine* e, int game_id, int tree, const mc_state* state, int* found, uint16_t* codes, uint32_t* visits,
                  double* q, float* priors, int32_t* n_legal, int* is_terminal, double* terminal_value) {
    if (!e || !state || !found || game_id < 0 || game_id >= e->v.G) return fail(MCAZ_EINVAL, "az_node_stats: bad argument");
    e->scratch.begin();
    void *p_out, *p_codes, *p_vis, *p_q, *p_p;
    if (int rc = e->scratch.take(sizeof(NodeStatsOut), &p_out)) return rc;
    if (int rc = e->scratch.take(MC_MAX_MOVES * sizeof(uint16_t), &p_codes)) return rc;
    if (int rc = e->scratch.take(MC_MAX_MOVES * sizeof(uint32_t), &p_vis)) return rc;
    if (int rc = e->scratch.take(MC_MAX_MOVES * sizeof(double), &p_q)) return rc;
    if (int rc = e->scratch.take(MC_MAX_MOVES * sizeof(float), &p_p)) return rc;
    NodeStatsOut* d_out = static_cast<NodeStatsOut*>(p_out);
    uint16_t* d_codes = static_cast<uint16_t*>(p_codes);
    uint32_t* d_vis = static_cast<uint32_t*>(p_vis);
    double* d_q = static_cast<double*>(p_q);
    float* d_p = static_cast<float*>(p_p);
    node_stats_kernel<<<1, 32, 0, e->stream>>>(e->v, game_id, tree, *state, d_out, d_codes, d_vis, d_q, d_p);
    g_launches.fetch_add(1);
    e->launches++;
    NodeStatsOut h;
    cudaMemcpyAsync(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost, e->stream);
    if (codes) cudaMemcpyAsync(codes, d_codes, MC_MAX_MOVES * sizeof(uint16_t), cudaMemcpyDeviceToHost, e->stream);
    if (visits) cudaMemcpyAsync(visits, d_vis, MC_MAX_MOVES * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream);
    if (q) cudaMemcpyAsync(q, d_q, MC_MAX_MOVES * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
    if (priors) cudaMemcpyAsync(priors, d_p, MC_MAX_MOVES * sizeof(float), cudaMemcpyDeviceToHost, e->stream);
    cudaError_t err = cudaStreamSynchronize(e->stream);
    if (err != cudaSuccess) return fail(MCAZ_ECUDA, std::string("az_node_stats: ") + cudaGetErrorString(err));
    *found = h.found;
    if (n_legal) *n_legal = h.n_legal;
    if (is_terminal) *is_terminal = h.is_terminal;
    if (terminal_value) *terminal_value = h.terminal_value;
    return MCAZ_OK;
}

int az_tree_dump(az_engine* e, int game_id, int tree, int max_nodes, mc_state* states, uint32_t* info, uint32_t* edge_off, int* n_nodes,
                 int max_edges, uint16_t* codes, uint32_t* visits, double* q, float* priors, int* n_edges) {
    if (!e || game_id < 0 || game_id >= e->v.G || tree < 0 || tree > 1 || max_nodes < 0 || max_edges < 0 || !n_nodes || !n_edges)
        return fail(MCAZ_EINVAL, "az_tree_dump: bad argument");
    const View& V = e->v;
    const size_t t = 2 * (size_t)game_id + (size_t)tree;
    uint32_t n = 0, m = 0;
    MCAZ_CUDA(cudaStreamSynchronize(e->stream));
    MCAZ_CUDA(cudaMemcpy(&n, V.tree_nodes + t, sizeof(n), cudaMemcpyDeviceToHost));
    MCAZ_CUDA(cudaMemcpy(&m, V.tree_edges + t, sizeof(m), cudaMemcpyDeviceToHost));
    *n_nodes = (int)n; *n_edges = (int)m;
    const size_t nn = std::min<size_t>(n, (size_t)max_nodes), mm = std::min<size_t>(m, (size_t)max_edges);
    const size_t nb = t * V.NC, eb = t * V.EC;
    if (nn) {
        std::vector<az::Node> nodes(nn);
        MCAZ_CUDA(cudaMemcpy(nodes.data(), V.nodes + nb, nn * sizeof(az::Node), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < nn; ++i) {
            if (states) states[i] = az::state_of(nodes[i].board, nodes[i].head.meta);
            if (info) info[i] = nodes[i].head.info & (0xffffu | az::INFO_TERMINAL | az::INFO_DECISIVE);
            if (edge_off) edge_off[i] = nodes[i].head.edge_off;
        }
    }
    if (mm && (codes || visits || q || priors)) {
        std::vector<az::Edge> edges(mm);
        MCAZ_CUDA(cudaMemcpy(edges.data(), V.edges + eb, mm * sizeof(az::Edge), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < mm; ++i) {
            if (codes) codes[i] = edges[i].link.code;
            if (visits) visits[i] = edges[i].stat.N;
            if (q) q[i] = edges[i].stat.Q;
            if (priors) priors[i] = edges[i].stat.P;
        }
    }
    return MCAZ_OK;
}

int az_play(az_engine* e, const int32_t* game_ids, const uint16_t* codes, int n, int8_t* results) {
    if (!e || n < 0 || (n > 0 && !codes)) return fail(MCAZ_EINVAL, "az_play: bad argument");
    if (e->leaf_pending) return fail(MCAZ_ESTATE, "az_play: a simulation is pending (call az_backup first)");
    if (n == 0) return MCAZ_OK;
    e->scratch.begin();
    In<int32_t> ids; In<uint16_t> ic; Out<int8_t> orr;
    if (int rc = ids.init(game_ids, n, e->stream, e->scratch)) return rc;
    if (int rc = ic.init(codes, n, e->stream, e->scratch)) return rc;
    if (int rc = orr.init(results, results ? n : 0, e->stream, e->scratch)) return rc;
    play_kernel<<<std::max(1, std::min((n + 127) / 128, num_sms() * 8)), 128, 0, e->stream>>>(e->v, ids.ptr, ic.ptr, n, orr.ptr);
    MCAZ_CHECK_LAUNCH();
    e->launches++;
    if (int rc = maybe_recycle(e, e->cfg.max_sims_per_move)) return rc;     // room for the next move's search
    if (int rc = orr.finish(e->stream)) return rc;
    return engine_check_errors(e);
}

int az_play_device(az_engine* e) {
    if (!e) return fail(MCAZ_EINVAL, "az_play_device: null engine");
    if (e->leaf_pending) return fail(MCAZ_ESTATE, "az_play_device: a simulation is pending");
    const View& V = e->v;
    if (!e->cfg.device_rng) return fail(MCAZ_ESTATE, "az_play_device: needs device_rng = 1");
    play_device_kernel<<<warp_grid(V.G, 128), 128, 0, e->stream>>>(V);
    MCAZ_CHECK_LAUNCH();
    restart_finished_kernel<<<std::min(V.G, num_sms() * 8), 256, 0, e->stream>>>(V, start_state());
    MCAZ_CHECK_LAUNCH();
    e->launches += 2;
    return maybe_recycle(e, e->cfg.max_sims_per_move);                      // room for the next move's search
}

int az_game_states(az_engine* e, const int32_t* game_ids, int n, mc_state* states, int8_t* results) {
    if (!e || n < 0) return fail(MCAZ_EINVAL, "az_game_states: bad argument");
    if (n == 0) return MCAZ_OK;
    e->scratch.begin();
    In<int32_t> ids; Out<mc_state> os; Out<int8_t> orr;
    if (int rc = ids.init(game_ids, n, e->stream, e->scratch)) return rc;
    if (int rc = os.init(states, states ? n : 0, e->stream, e->scratch)) return rc;
    if (int rc = orr.init(results, results ? n : 0, e->stream, e->scratch)) return rc;
    game_states_kernel<<<std::max(1, std::min((n + 127) / 128, num_sms() * 8)), 128, 0, e->stream>>>(e->v, ids.ptr, n, os.ptr, orr.ptr);
    MCAZ_CHECK_LAUNCH();
    e->launches++;
    if (int rc = os.finish(e->stream)) return rc;
    if (int rc = orr.finish(e->stream)) return rc;
    MCAZ_CUDA(cudaStreamSynchronize(e->stream));
    return MCAZ_OK;
}

int az_drain_replay(az_engine* e, az_replay_tuple* out, int max, int* n_out) {
    if (!e || !n_out || max < 0) return fail(MCAZ_EINVAL, "az_drain_replay: bad argument");
    unsigned long long count = 0;
    MCAZ_CUDA(cudaMemcpyAsync(&count, e->d_replay_count, sizeof(count), cudaMemcpyDeviceToHost, e->stream));
    MCAZ_CUDA(cudaStreamSynchronize(e->stream));
    const size_t have = (size_t)std::min<unsigned long long>(count, e->replay_capacity);   // reservations never pass the capacity
    const size_t take = out ? std::min<size_t>(have, (size_t)max) : 0;
    if (take)
        MCAZ_CUDA(cudaMemcpyAsync(out, e->d_replay, take * sizeof(az_replay_tuple),
                                  is_device_pointer(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
    // what was not taken stays queued, in order: move it to the front in pieces no longer than the gap (no overlap)
    const size_t rest = out ? have - take : 0;              // out == NULL discards everything
    for (size_t done = 0; done < rest && take > 0; done += take) {
        const size_t n = std::min(take, rest - done);
        MCAZ_CUDA(cudaMemcpyAsync(e->d_replay + done, e->d_replay + take + done, n * sizeof(az_replay_tuple), cudaMemcpyDeviceToDevice,
                                  e->stream));
    }
    const unsigned long long left = (unsigned long long)(take > 0 ? rest : (out ? have : 0));
    MCAZ_CUDA(cudaMemcpyAsync(e->d_replay_count, &left, sizeof(left), cudaMemcpyHostToDevice, e->stream));
    MCAZ_CUDA(cudaStreamSynchronize(e->stream));
    *n_out = (int)take;
    return MCAZ_OK;
}

int az_counters(az_engine* e, uint64_t* out) {
    if (!e || !out) return fail(MCAZ_EINVAL, "az_counters: bad argument");
    unsigned long long h[AZ_NUM_COUNTERS];
    MCAZ_CUDA(cudaMemcpyAsync(h, e->v.counters, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
    MCAZ_CUDA(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < AZ_NUM_COUNTERS; ++i) out[i] = h[i];
    out[az::C_LAUNCHES] = e->launches;
    return MCAZ_OK;
}

int az_network_forward(az_engine* e, const uint8_t* tokens, const float* clocks, int n, float* logits, float* values) {
    if (!e || n < 0 || (n > 0 && (!tokens || !clocks || !logits || !values))) return fail(MCAZ_EINVAL, "az_network_forward: bad argument");
    if (!e->net) return fail(MCAZ_ESTATE, "az_network_forward: engine was created with network = 0");
    if (n == 0) return MCAZ_OK;
    e->scratch.begin();
    In<uint8_t> it; In<float> ic; Out<float> ol; Out<float> ov;
    if (int rc = it.init(tokens, (size_t)n * MC_TOKENS, e->stream, e->scratch)) return rc;
    if (int rc = ic.init(clocks, n, e->stream, e->scratch)) return rc;
    if (int rc = ol.init(logits, (size_t)n * MC_NUM_ACTIONS, e->stream, e->scratch)) return rc;
    if (int rc = ov.init(values, n, e->stream, e->scratch)) return rc;
    if (int rc = network_forward(e, it.ptr, ic.ptr, nullptr, n, ol.ptr, ov.ptr)) return rc;
    if (int rc = ol.finish(e->stream)) return rc;
    if (int rc = ov.finish(e->stream)) return rc;
    MCAZ_CUDA(cudaStreamSynchronize(e->stream));
    return MCAZ_OK;
}

}  // extern "C"

// ---- learner-side collate (exp/learner.py:23-41) on the device -----------------------------------------
// Packed replay tuples -> the four training tensors collate_fn builds: dense pi [n,554] (float32, pi
// scattered to the legal codes), tokens [n,2,6,5] (int64, Network.process_observation), clock [n,1],
// reward [n,1].  One warp per tuple.
namespace mcaz {
__global__ void __launch_bounds__(128) collate_kernel(const az_replay_tuple* __restrict__ tuples, int n, float* __restrict__ pi,
                                                      long long* __restrict__ tokens, float* __restrict__ clock,
                                                      float* __restrict__ reward) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const az_replay_tuple* t = tuples + i;
        float* row = pi + (size_t)i * MC_NUM_ACTIONS;
        for (int k = lane; k < MC_NUM_ACTIONS; k += 32) row[k] = 0.f;
        __syncwarp();
        const int E = t->n_legal;
        for (int k = lane; k < E; k += 32) row[t->codes[k]] = t->pi[k];
        const mc_state s = t->observation;
        if (lane < 30) {
            const bool white = mc::white_to_move(s);
            const uint32_t occ = s.pl0 | s.pl1 | s.pl2;
            const uint32_t mine = white ? (s.white & occ) : (occ & ~s.white);
            int cell = 5 * (5 - lane / 5) + lane % 5;
            if (!white) cell = 29 - cell;
            const int ty = mc::piece_at(s, cell);
            const bool m = (mine >> cell) & 1u;
            tokens[(size_t)i * MC_TOKENS + lane] = m ? ty : 0;
            tokens[(size_t)i * MC_TOKENS + 30 + lane] = m ? 0 : ty;
        }
        if (lane == 0) {
            const double c = (double)mc::fullmove(s) + (mc::white_to_move(s) ? 0.0 : 0.5);
            clock[i] = (float)(c / 30.0);
            reward[i] = (float)t->reward;
        }
    }
}
}  // namespace mcaz

extern "C" int az_collate(const az_replay_tuple* tuples, int n, float* pi, int64_t* tokens, float* clock, float* reward) {
    if (n < 0 || (n > 0 && (!tuples || !pi || !tokens || !clock || !reward))) return fail(MCAZ_EINVAL, "az_collate: bad argument");
    if (int rc = require_device()) return rc;
    if (n == 0) return MCAZ_OK;
    cudaStream_t st = 0;
    Scratch& sc = thread_scratch();
    sc.begin();
    In<az_replay_tuple> in; Out<float> op; Out<int64_t> ot; Out<float> oc; Out<float> orw;
    if (int rc = in.init(tuples, n, st, sc)) return rc;
    if (int rc = op.init(pi, (size_t)n * MC_NUM_ACTIONS, st, sc)) return rc;
    if (int rc = ot.init(tokens, (size_t)n * MC_TOKENS, st, sc)) return rc;
    if (int rc = oc.init(clock, n, st, sc)) return rc;
    if (int rc = orw.init(reward, n, st, sc)) return rc;
    collate_kernel<<<warp_grid(n, 128), 128, 0, st>>>(in.ptr, n, op.ptr, reinterpret_cast<long long*>(ot.ptr), oc.ptr, orw.ptr);
    MCAZ_CHECK_LAUNCH();
    if (int rc = op.finish(st)) return rc;
    if (int rc = ot.finish(st)) return rc;
    if (int rc = oc.finish(st)) return rc;
    if (int rc = orw.finish(st)) return rc;
    MCAZ_CUDA(cudaStreamSynchronize(st));
    return MCAZ_OK;
}
