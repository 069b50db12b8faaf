// AddressSanitizer / UndefinedBehaviorSanitizer run of the code the CUDA kernels share with the host: csrc/mcts_core.cuh
// (select / expand / backup, the tree hash table, the game line) and csrc/minitchess.cuh (bitboard rules).
// compute-sanitizer is closed on the GPU pool, so this is the memory-safety check of that shared logic: whole self-play
// games through the az_* calls of the host harness -- wide and narrow arenas (the narrow ones must end in the capacity
// error flag, never in an overrun), virtual loss (leaves_per_step > 1), and the rules on dense synthetic positions.
// TEST INFRASTRUCTURE; built and run by tests/test_sanitizers.py.
#include <cstdio>

#include "mcts_host.cpp"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 32);
}

static int play(int games, int sims, int leaves, int node_cap, bool expect_overflow) {
    az_config cfg;
    az_default_config(&cfg);
    cfg.n_games = games; cfg.max_sims_per_move = sims * leaves; cfg.leaves_per_step = leaves; cfg.node_capacity = node_cap;
    cfg.dirichlet_epsilon = 0.25f;
    az_engine* e = nullptr;
    if (az_create(&cfg, &e)) return 1;
    const int S = games * leaves;
    std::vector<float> values(S), priors((size_t)S * MC_MAX_MOVES);
    std::vector<double> noise((size_t)games * MC_MAX_MOVES);
    std::vector<uint16_t> codes((size_t)games * MC_MAX_MOVES), pick(games);
    std::vector<uint32_t> visits((size_t)games * MC_MAX_MOVES);
    std::vector<int32_t> n_legal(games);
    std::vector<int8_t> results(games);
    bool overflow = false;
    for (int ply = 0; ply < 64 && !overflow; ++ply) {
        for (int s = 0; s < sims && !overflow; ++s) {
            for (auto& x : noise) x = (rnd() % 1000 + 1) / 1000.0;
            az_select_expand(e, noise.data(), nullptr);
            for (int i = 0; i < S; ++i) values[i] = (float)((int)(rnd() % 2001) - 1000) / 1000.0f;
            for (auto& p : priors) p = (float)(rnd() % 1000 + 1) / 1000.0f;
            az_backup(e, nullptr, values.data(), priors.data());
            if (*e->v.error_flag) overflow = true;
        }
        if (az_root_stats(e, nullptr, games, codes.data(), visits.data(), nullptr, n_legal.data())) { overflow = true; break; }
        std::vector<int32_t> ids;
        std::vector<uint16_t> mv;
        az_game_states(e, nullptr, games, nullptr, results.data());
        for (int g = 0; g < games; ++g) {
            if (results[g] != MC_ONGOING || n_legal[g] <= 0) continue;
            uint32_t best = 0; int at = 0;
            for (int i = 0; i < n_legal[g]; ++i) if (visits[(size_t)g * MC_MAX_MOVES + i] >= best) { best = visits[(size_t)g * MC_MAX_MOVES + i]; at = i; }
            ids.push_back(g); mv.push_back(codes[(size_t)g * MC_MAX_MOVES + at]);
        }
        if (ids.empty()) break;
        if (az_play(e, ids.data(), mv.data(), (int)ids.size(), nullptr)) { overflow = true; break; }
    }
    uint64_t c[AZ_NUM_COUNTERS];
    az_counters(e, c);
    printf("games %d sims %d leaves %d node_cap %d: %llu simulations, %llu nodes, %llu moves, overflow %d\n", games, sims, leaves, node_cap,
           (unsigned long long)c[0], (unsigned long long)c[5], (unsigned long long)c[3], (int)overflow);
    az_destroy(e);
    return overflow == expect_overflow ? 0 : 1;
}

static int rules_sweep(int n) {
    mc_rules R;
    mc_default_rules(&R);
    unsigned long long total = 0;
    for (int variant = 0; variant < 3; ++variant) {
        R.pawn_double_step = variant == 1; R.promo_multiplicity = variant == 2 ? 4 : 1;
        for (int i = 0; i < n; ++i) {
            mc_state s = {0, 0, 0, 0, MC_META(rnd() & 1, rnd() % 20, 1 + rnd() % 30)};
            uint32_t used = 0;
            const int pieces = 2 + rnd() % 11;
            for (int k = 0; k < pieces; ++k) {
                int sq = rnd() % 30;
                if ((used >> sq) & 1u) continue;
                used |= 1u << sq;
                int t = k < 2 ? mc::KING : 1 + (int)(rnd() % 5);
                if (t == mc::PAWN && (sq < 5 || sq >= 25)) t = mc::KNIGHT;
                if (t & 1) s.pl0 |= 1u << sq;
                if (t & 2) s.pl1 |= 1u << sq;
                if (t & 4) s.pl2 |= 1u << sq;
                if (k == 0 || (k > 1 && (rnd() & 1))) s.white |= 1u << sq;
            }
            uint16_t codes[MC_MAX_MOVES];
            int res;
            const int E = mc::generate(s, R, codes, &res);
            total += (unsigned long long)E;
            for (int k = 0; k < E; ++k) {
                mc_state o;
                if (mc::step(s, codes[k], R, &o) == 0) {
                    uint8_t tok[MC_TOKENS]; float clk;
                    mc::tokenize(o, tok, &clk);
                }
            }
        }
    }
    printf("rules sweep: %llu moves generated\n", total);
    return total > 0 ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += play(3, 24, 1, 0, false);          // whole games, default arenas
    bad += play(2, 12, 4, 0, false);          // virtual loss: four descents per tree and step
    bad += play(2, 24, 1, 40, true);          // arenas far too small: must stop on the capacity flag, not overrun
    bad += rules_sweep(20000);
    printf(bad ? "FAILED\n" : "sanitize ok\n");
    return bad;
}
