// Host build of csrc/mcts_core.cuh behind the same az_* signatures as libmcaz.so, over malloc'd
// arrays and one "lane".  TEST INFRASTRUCTURE: lets `pytest -m "not gpu"` check the tree logic
// (the very source the CUDA kernels compile) bit-exactly against the reference MCTS without a
// GPU.  Lives under tests/ and is never loaded by the product.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mcts_core.cuh"

struct az_engine {
    az_config cfg;
    az::View v;
    std::vector<void*> allocs;
    bool leaf_pending = false;
    std::vector<uint8_t> noise_used;
};

static std::string g_err;
static int fail(int code, const char* msg) { g_err = msg; return code; }

template <typename T>
static void alloc(az_engine* e, T** p, size_t n) {
    *p = static_cast<T*>(calloc(n ? n : 1, sizeof(T)));
    e->allocs.push_back(*p);
}

static mc_state start_state() {
    // 2nbk/2ppp/5/5/PPP2/KBN2 w 0 1
    mc_state s = {0, 0, 0, 0, MC_META(1, 0, 1)};
    auto put = [&](int sq, int t, bool white) {
        if (t & 1) s.pl0 |= 1u << sq;
        if (t & 2) s.pl1 |= 1u << sq;
        if (t & 4) s.pl2 |= 1u << sq;
        if (white) s.white |= 1u << sq;
    };
    put(0, mc::KING, true); put(1, mc::BISHOP, true); put(2, mc::KNIGHT, true);
    put(5, mc::PAWN, true); put(6, mc::PAWN, true); put(7, mc::PAWN, true);
    put(29, mc::KING, false); put(28, mc::BISHOP, false); put(27, mc::KNIGHT, false);
    put(24, mc::PAWN, false); put(23, mc::PAWN, false); put(22, mc::PAWN, false);
    return s;
}

extern "C" {

const char* mcaz_last_error(void) { return g_err.c_str(); }
int mcaz_abi_version(void) { return MCAZ_ABI_VERSION; }
uint64_t mcaz_kernel_launches(void) { return 0; }

void mc_default_rules(mc_rules* r) { r->pawn_double_step = 0; r->promo_multiplicity = 1; r->max_fullmoves = 30; r->insufficient_material = 1; r->fivefold_repetition = 1; }

void az_default_config(az_config* c) {
    memset(c, 0, sizeof(*c));
    c->n_games = 1; c->max_sims_per_move = 36; c->cpuct = 1.0f; c->tau_change = 6;
    c->dirichlet_alpha = 0.6f; c->dirichlet_epsilon = 0.25f;
    c->leaves_per_step = 1;
    mc_default_rules(&c->rules);
}

int az_reset_games(az_engine* e, const int32_t* ids, int n, const mc_state* states);

int az_create(const az_config* cfg, az_engine** out) {
    az_engine* e = new az_engine();
    e->cfg = *cfg;
    az::View& V = e->v;
    V.G = cfg->n_games;
    V.K = cfg->leaves_per_step > 0 ? cfg->leaves_per_step : 1;
    long long nc = cfg->node_capacity > 0 ? cfg->node_capacity : (long long)cfg->max_sims_per_move * 31 + 64;
    long long ec = cfg->edge_capacity > 0 ? cfg->edge_capacity : nc * 14;
    V.NC = (int)nc; V.EC = (int)ec;
    int hc = 64; while (hc < 2 * V.NC) hc <<= 1; V.HC = hc;
    V.cpuct = cfg->cpuct; V.eps = cfg->dirichlet_epsilon; V.alpha = cfg->dirichlet_alpha;
    V.numpy1 = cfg->numpy1_dtype_flow; V.tau_change = cfg->tau_change; V.rules = cfg->rules; V.seed = cfg->seed;
    V.device_rng = 0; V.sim_counter = 0;
    size_t G = V.G, T = 2 * G, N = T * V.NC, E = T * V.EC, S = G * V.K;
    alloc(e, &V.game_state, G); alloc(e, &V.game_result, G); alloc(e, &V.game_ply, G); alloc(e, &V.game_start_ply, G);
    alloc(e, &V.game_hist, G * az::HIST); alloc(e, &V.game_hist_len, G);
    alloc(e, &V.tree_nodes, T); alloc(e, &V.tree_edges, T); alloc(e, &V.tree_root, T);
    alloc(e, &V.nodes, N); alloc(e, &V.edges, E);
    alloc(e, &V.ht, T * V.HC);
    alloc(e, &V.path_len, S); alloc(e, &V.path_edge, S * az::MAX_DEPTH); alloc(e, &V.path_node, S * az::MAX_DEPTH);
    alloc(e, &V.leaf_node, S); alloc(e, &V.leaf_kind, S); alloc(e, &V.leaf_value, S);
    alloc(e, &V.tokens, S * MC_TOKENS); alloc(e, &V.clocks, S); alloc(e, &V.needs_eval, S); alloc(e, &V.leaf_states, S);
    alloc(e, &V.counters, AZ_NUM_COUNTERS); alloc(e, &V.error_flag, 1);
    e->noise_used.assign(G, 0);
    *out = e;
    return az_reset_games(e, nullptr, V.G, nullptr);
}

int az_destroy(az_engine* e) {
    if (!e) return 0;
    for (void* p : e->allocs) free(p);
    delete e;
    return 0;
}

static int check(az_engine* e) {
    int f = *e->v.error_flag;
    *e->v.error_flag = 0;
    if (!f) return 0;
    return fail((f & az::ERR_ILLEGAL) && !(f & 7) ? MCAZ_EINVAL : MCAZ_ECAPACITY, "engine error flag");
}

int az_reset_games(az_engine* e, const int32_t* ids, int n, const mc_state* states) {
    az::View& V = e->v;
    for (int k = 0; k < n; ++k) {
        int g = ids ? ids[k] : k;
        memset(V.ht + (size_t)(2 * g) * V.HC, 0, sizeof(uint32_t) * 2 * V.HC);
        mc_state s = states ? states[k] : start_state();
        V.game_state[g] = s;
        V.game_ply[g] = mc::white_to_move(s) ? 0 : 1;
        V.game_start_ply[g] = V.game_ply[g];
        az::hist_reset(V, g, s);
        V.game_result[g] = (int8_t)az::game_result_of(V, g, s);
        for (int t = 2 * g; t < 2 * g + 2; ++t) { V.tree_nodes[t] = 0; V.tree_edges[t] = 0; V.tree_root[t] = az::NONE; }
        for (int j = 0; j < V.K; ++j) { V.leaf_kind[g * V.K + j] = az::LEAF_NONE; V.needs_eval[g * V.K + j] = 0; V.path_len[g * V.K + j] = 0; }
    }
    e->leaf_pending = false;
    return 0;
}

int az_set_positions(az_engine* e, const int32_t* ids, int n, const mc_state* states, const int32_t* tree_of_game) {
    az::View& V = e->v;
    for (int k = 0; k < n; ++k) {
        int g = ids ? ids[k] : k;
        mc_state s = states[k];
        V.game_state[g] = s;
        V.game_ply[g] = tree_of_game ? (tree_of_game[k] & 1) : (mc::white_to_move(s) ? 0 : 1);
        V.game_start_ply[g] = V.game_ply[g];
        az::hist_reset(V, g, s);
        V.game_result[g] = (int8_t)az::game_result_of(V, g, s);
        V.tree_root[2 * g] = az::NONE; V.tree_root[2 * g + 1] = az::NONE;
        for (int j = 0; j < V.K; ++j) V.leaf_kind[g * V.K + j] = az::LEAF_NONE;
    }
    e->leaf_pending = false;
    return 0;
}

int az_select_expand(az_engine* e, const double* noise, uint8_t* noise_used) {
    if (e->leaf_pending) return fail(MCAZ_ESTATE, "previous simulation not backed up");
    for (int g = 0; g < e->v.G; ++g)
        for (int j = 0; j < e->v.K; ++j) az::select_expand_one(e->v, g, 0, noise, e->noise_used.data(), j);
    if (noise_used) memcpy(noise_used, e->noise_used.data(), e->v.G);
    e->leaf_pending = true;
    return 0;
}

int az_leaf_batch(az_engine* e, const uint8_t** tokens, const float** clocks, const uint8_t** needs_eval,
                  const mc_state** leaf_states, int* n_slots) {
    if (tokens) *tokens = e->v.tokens;
    if (clocks) *clocks = e->v.clocks;
    if (needs_eval) *needs_eval = e->v.needs_eval;
    if (leaf_states) *leaf_states = e->v.leaf_states;
    if (n_slots) *n_slots = e->v.G * e->v.K;
    return 0;
}

int az_backup(az_engine* e, const float* logits, const float* values, const float* priors) {
    if (!e->leaf_pending) return fail(MCAZ_ESTATE, "no simulation pending");
    for (int g = 0; g < e->v.G; ++g)
        for (int j = 0; j < e->v.K; ++j) az::backup_one(e->v, g, 0, logits, values, priors, j);
    e->leaf_pending = false;
    return 0;
}

int az_root_stats(az_engine* e, const int32_t* ids, int n, uint16_t* codes, uint32_t* visits, double* q, int32_t* n_legal) {
    az::View& V = e->v;
    for (int k = 0; k < n; ++k) {
        int g = ids ? ids[k] : k;
        int t = 2 * g + (V.game_ply[g] & 1);
        uint32_t root = V.tree_root[t];
        if (root == az::NONE) root = az::ht_find(V, t, V.game_state[g]);
        if (root == az::NONE) { n_legal[k] = -1; continue; }
        size_t gi = (size_t)t * V.NC + root;
        uint32_t info = V.nodes[gi].head.info;
        int E = (info & az::INFO_TERMINAL) ? 0 : (int)(info & 0xffffu);
        size_t e0 = (size_t)t * V.EC + V.nodes[gi].head.edge_off;
        for (int i = 0; i < E; ++i) {
            codes[(size_t)k * MC_MAX_MOVES + i] = V.edges[e0 + i].link.code;
            visits[(size_t)k * MC_MAX_MOVES + i] = V.edges[e0 + i].stat.N;
            if (q) q[(size_t)k * MC_MAX_MOVES + i] = V.edges[e0 + i].stat.Q;
        }
        n_legal[k] = E;
    }
    return check(e);
}

int az_node_stats(az_engine* e, int g, int tree, const mc_state* s, int* found, uint16_t* codes, uint32_t* visits, double* q,
                  float* priors, int32_t* n_legal, int* is_terminal, double* terminal_value) {
    az::View& V = e->v;
    int t = 2 * g + (tree & 1);
    uint32_t node = az::ht_find(V, t, *s);
    *found = node != az::NONE;
    if (!*found) return 0;
    size_t gi = (size_t)t * V.NC + node;
    uint32_t info = V.nodes[gi].head.info;
    bool term = info & az::INFO_TERMINAL;
    int E = term ? 0 : (int)(info & 0xffffu);
    size_t e0 = (size_t)t * V.EC + V.nodes[gi].head.edge_off;
    for (int i = 0; i < E; ++i) {
        if (codes) codes[i] = V.edges[e0 + i].link.code;
        if (visits) visits[i] = V.edges[e0 + i].stat.N;
        if (q) q[i] = V.edges[e0 + i].stat.Q;
        if (priors) priors[i] = V.edges[e0 + i].stat.P;
    }
    if (n_legal) *n_legal = E;
    if (is_terminal) *is_terminal = term;
    if (terminal_value) *terminal_value = term ? ((info & az::INFO_DECISIVE) ? -1.0 : -0.0) : 0.0;
    return 0;
}

int az_play(az_engine* e, const int32_t* ids, const uint16_t* codes, int n, int8_t* results) {
    if (e->leaf_pending) return fail(MCAZ_ESTATE, "simulation pending");
    for (int k = 0; k < n; ++k) {
        int g = ids ? ids[k] : k;
        if (az::play_one(e->v, g, codes[k]) == 1) az::raise(e->v, az::ERR_ILLEGAL);
        if (results) results[k] = e->v.game_result[g];
    }
    return check(e);
}

int az_game_states(az_engine* e, const int32_t* ids, int n, mc_state* states, int8_t* results) {
    for (int k = 0; k < n; ++k) {
        int g = ids ? ids[k] : k;
        if (states) states[k] = e->v.game_state[g];
        if (results) results[k] = e->v.game_result[g];
    }
    return 0;
}

int az_counters(az_engine* e, uint64_t* out) {
    for (int i = 0; i < AZ_NUM_COUNTERS; ++i) out[i] = e->v.counters[i];
    return 0;
}

}  // extern "C"
