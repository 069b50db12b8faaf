/* mcaz.h -- C ABI of the B200-native MinitChess AlphaZero self-play engine (libmcaz.so).
 *
 * This is the drop-in boundary for the reference's self-play hot path.  The reference is pure
 * Python, so the "FFI" a maintainer binds is ctypes; the entry points
 * below replace, one for one, the Python call sites named beside them (paths relative to the
 * reference tree):
 *
 *   mc_*   stateless MinitChess rules on packed positions  -> exp/environment.py:34-50 (legal
 *          moves, result), :68-82 (step), and the python-chess fork calls at :25,36,39,48,76
 *   az_*   batched AlphaZero search over GPU-resident trees -> exp/agent.py:24-88
 *          (MonteCarloTreeSearch), :110-119 (select_action), exp/policy.py:71-80 (Network
 *          forward), :96-105 (process_observation), :115-122 (get_distribution)
 *
 * Conventions: plain pointers and sizes only.  Unless a parameter says otherwise, every buffer
 * may be a host pointer (pageable or pinned) or a device pointer; the library inspects it with
 * cudaPointerGetAttributes and stages host buffers through its own device scratch.  All
 * functions return 0 on success or a negative MCAZ_E* code; mcaz_last_error() gives the text.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * MCAZ_ENODEV.
 */
#ifndef MCAZ_H
#define MCAZ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCAZ_ABI_VERSION 6   /* 4: az_replay_tuple.weights_version, az_set_weights_version, az_tree_dump, counter 14,
                                az_config.fp8_convolutions, network = 2;  5: az_config.defer_rows, counters 15-16;
                                6: counter 17 (second-stage compaction of recycle = 1) */

/* ---- geometry and action indexing (exp/generate_moves_list.py:5-57, exp/moves_dict.json) */
#define MC_FILES 5
#define MC_RANKS 6
#define MC_SQUARES 30
#define MC_NUM_ACTIONS 554   /* 430 queen-like (from,dir,dist) + 124 knight moves          */
#define MC_MAX_MOVES 96      /* safe upper bound on legal moves in one position (<= 78)     */
#define MC_TOKENS 60         /* Network input: 2 channels x 6 x 5 (exp/policy.py:82-95)     */

/* ---- error codes */
#define MCAZ_OK 0
#define MCAZ_EINVAL (-1)     /* bad argument                                               */
#define MCAZ_ENODEV (-2)     /* no usable CUDA device                                      */
#define MCAZ_ECUDA (-3)      /* a CUDA call failed; see mcaz_last_error()                  */
#define MCAZ_ECAPACITY (-4)  /* a tree arena or hash table overflowed                      */
#define MCAZ_ESTATE (-5)     /* call made in the wrong phase (e.g. backup before select)   */

/* ---- packed position: everything the 4-field FEN holds (exp/environment.py:6)
 * Piece type of square s = bit s of (pl0 | pl1<<1 | pl2<<2), numbered like the reference
 * tokeniser alphabet '0prbnqk' (exp/policy.py:7): 0 empty, 1 pawn, 2 rook, 3 bishop,
 * 4 knight, 5 queen, 6 king.  Square s = 5*rank + file (a1 = 0 ... e6 = 29).               */
typedef struct mc_state {
    uint32_t pl0, pl1, pl2; /* piece-type bit planes over the 30 squares                   */
    uint32_t white;         /* squares holding white pieces                                */
    uint32_t meta;          /* bit 0: 1 = white to move; bits 8-15 halfmove clock;
                               bits 16-23 fullmove number                                  */
} mc_state;

#define MC_META(turn_white, halfmove, fullmove) \
    ((uint32_t)((turn_white) ? 1u : 0u) | ((uint32_t)(halfmove) << 8) | ((uint32_t)(fullmove) << 16))

/* ---- game result of a position, as exp/environment.py:39-45 reads board.result()        */
#define MC_ONGOING 0         /* '*'                                                        */
#define MC_WHITE_WINS 1      /* '1-0'                                                      */
#define MC_BLACK_WINS 2      /* '0-1'                                                      */
#define MC_DRAW 3            /* '1/2-1/2'                                                  */

/* ---- switches for the rules the reference does not pin (SURVEY.md §8c ledger).  Passing
 * NULL where a `const mc_rules*` is expected selects mc_default_rules().                   */
typedef struct mc_rules {
    int32_t pawn_double_step;      /* default 0                                            */
    int32_t promo_multiplicity;    /* default 1 (queen only); 4 = q,r,b,n share one code   */
    int32_t max_fullmoves;         /* default 30: draw once fullmove number > this; an engine
                                      (az_create) takes at most 31: a game line is <= 64 plies */
    int32_t insufficient_material; /* default 1                                            */
    int32_t fivefold_repetition;   /* default 1 (episode-level only: needs history)        */
} mc_rules;

void mc_default_rules(mc_rules* out);

/* ---- library */
int mcaz_abi_version(void);
/* sizeof of the structs of this header as the library was compiled: 0 mc_state, 1 mc_rules, 2 az_config,
 * 3 az_replay_tuple (0 for anything else).  A binding checks its own mirror against these at load time.   */
size_t mcaz_struct_size(int which);
const char* mcaz_last_error(void);
int mcaz_device_count(void);         /* 0 when no CUDA device is usable                    */
int mcaz_set_device(int device);

/* ---- host helpers (string <-> packed; no GPU work) */
int mc_state_from_fen(const char* fen, mc_state* out);
int mc_state_to_fen(const mc_state* s, char* buf, size_t buflen);
/* code -> (from,to) squares as seen on the real board for the given side (1 = white):
 * exp/environment.py:19 MOVES_DICT_INV.  Returns MCAZ_EINVAL for code >= 554.             */
int mc_code_squares(int code, int white_to_move, int* from_sq, int* to_sq);
int mc_squares_code(int from_sq, int to_sq, int white_to_move);   /* -1 if not a move shape */

/* ---- stateless rules kernels (replace python-chess on the path) ------------------------
 * mc_legal_moves: for each of n positions, the sorted legal action codes
 * (codes[i*MC_MAX_MOVES + k], k < counts[i]) and the MC_* result; a finished position
 * still lists its legal moves exactly like exp/environment.py:47-50 does.                  */
int mc_legal_moves(const mc_state* states, int n, const mc_rules* rules,
                   uint16_t* codes, int32_t* counts, int8_t* results);
/* mc_apply: out[i] = states[i] after action codes[i] (pawn reaching the last rank queens,
 * exp/environment.py:72-74).  status[i] = 0 ok, 1 illegal move (out[i] = states[i]),
 * 2 position already finished (exp/environment.py:69-70).                                  */
int mc_apply(const mc_state* states, const uint16_t* codes, int n, const mc_rules* rules,
             mc_state* out, int8_t* status);
/* mc_perft: number of legal move sequences of length `depth` from each root, finished
 * positions having no successors.                                                          */
int mc_perft(const mc_state* roots, int n, int depth, const mc_rules* rules, uint64_t* nodes);
/* mc_tokenize: Network.process_observation (exp/policy.py:82-105): 60 tokens per position in
 * the mover's view (channel 0 = mover's pieces, channel 1 = opponent's) and the clock
 * (fullmove + 0.5*[black to move]) / 30 as float32.                                        */
int mc_tokenize(const mc_state* states, int n, uint8_t* tokens, float* clocks);

/* ---- AlphaZero engine -------------------------------------------------------------------*/
typedef struct az_engine az_engine;

typedef struct az_config {
    int32_t n_games;           /* concurrent games; 2 trees per game (one per colour's agent,
                                  app/base.py:113, exp/agent.py:105-108)                    */
    int32_t max_sims_per_move; /* sizes the per-tree arenas: <= 1 new node per simulation   */
    int32_t node_capacity;     /* nodes per tree; 0 = derive from max_sims_per_move         */
    int32_t edge_capacity;     /* edges per tree; 0 = derive                                */
    float cpuct;               /* exp/agent.py:96 (default 1)                               */
    int32_t tau_change;        /* exp/agent.py:97 (default 6)                               */
    float dirichlet_alpha;     /* exp/agent.py:82 (0.6)                                     */
    float dirichlet_epsilon;   /* exp/agent.py:82 (0.25); 0 disables root noise             */
    int32_t numpy1_dtype_flow; /* Q6: 1 = round P*sqrt(sumN) to f32 as numpy 1.x did        */
    int32_t device_rng;        /* 1 = Philox Dirichlet/sampling on device (throughput mode);
                                  0 = caller supplies noise / picks moves (parity mode)     */
    uint64_t seed;
    mc_rules rules;
    int32_t network;           /* 0 = external evaluator only; 1 = built-in bf16 tcgen05 net; 2 = the same network with the
                                  18 tower convolutions on e4m3 operands (fp32 accumulation, per-channel weight scales,
                                  per-level activation scales calibrated at az_set_weights, bf16 residual stream): about
                                  twice the evaluations per second, priors and values within 1e-2 of fp32 (opt-in) */
    int32_t leaves_per_step;   /* 1 = the reference's sequential search (bit-exact); K > 1 = K descents per
                                  tree and step kept apart by virtual loss (throughput option for few games;
                                  changes visit counts); leaf batch rows = n_games * K, row = game * K + j   */
    int32_t own_stream;        /* 1 = the engine works on its own non-blocking CUDA stream (several engines in one
                                  process then overlap); 0 = the legacy default stream, ordered with the caller's
                                  torch work (needed when device tensors are exchanged every simulation)            */
    int32_t eval_cache_log2;   /* > 0: exact evaluation cache of 2^k entries (192 B each) shared by all trees of the
                                  engine, used by az_search / az_selfplay: a new position whose (board, side to move,
                                  fullmove number) -- all the network sees, exp/policy.py:96-105 -- was evaluated before
                                  with the current weights takes the stored priors and value instead of a network row.
                                  Results are bit-identical with the cache on or off; az_set_weights invalidates it.
                                  0 = off (default)                                                                      */
    int32_t free_sims;         /* descents a game may start per launch of az_search / az_selfplay: simulations that end
                                  on a terminal or cached position need no network row, complete on the spot and the
                                  game goes on to its next simulation in the same launch.  0 = default (1: measured
                                  best on B200 -- longer chains stretch the launch by more than the fuller batch saves)  */
    int32_t recycle;           /* 1: before every az_search / az_selfplay call the trees that could not take the call's
                                  new nodes are compacted: nodes whose ply is not greater than the game's current
                                  position (other than that position) can never be reached again -- their ply is part
                                  of the key -- so no result changes, and node_capacity = 0 then derives a 5x smaller
                                  arena (6 x max_sims_per_move + 64).  A tree that is still too full after that (deep trees
                                  under sharp priors keep most of their nodes) is compacted again down to the nodes its
                                  edges reach from the current position, for as many levels as fit beside the coming
                                  search: what is dropped there (counter 17) is expanded anew if the search comes back
                                  to it -- the call goes on instead of failing with MCAZ_ECAPACITY.  For engines whose positions only move forward
                                  (az_play / az_play_device / az_selfplay); leave 0 when az_set_positions may jump back
                                  to an earlier position of a kept tree (the per-agent facade).  Default 0              */
    int32_t lookahead_rows;    /* > 0 (needs eval_cache_log2 > 0, leaves_per_step = 1): for engines with few games.  A pass
                                  of the network costs the same for one row as for a tile of 256, so every new unfinished
                                  node also queues its children as rows of the batch (at most this many per batch); their
                                  priors and values go to the exact cache only, where the later simulation that expands
                                  such a child finds the very bits a pass would give it then.  az_search / az_search_noise
                                  then run a game's simulations back to back inside one launch until one needs the
                                  network and return as soon as no game waits for a row.  The order of a game's simulations
                                  and every number in its tree are unchanged (exp/agent.py:41-45).  Default 0             */
    int32_t fp8_convolutions;  /* network = 2: how many of the 18 tower convolutions, counted from the first, multiply e4m3
                                  operands -- an even number (whole residual blocks), the rest run in bf16.  0 = default 12:
                                  measured, priors and values then stay within 1e-2 of the fp32 network on random-init,
                                  learner-stepped and BatchNorm-perturbed weights (all 18: 1.2e-2 on the last kind)       */
    int32_t defer_rows;        /* > 0 (az_selfplay; built-in network, leaves_per_step = 1, no look-ahead rows): the network works
                                  on tile pairs of 256 rows, so a dense batch of 3841 leaves costs as much as one of 4096.
                                  When the last tile pair of a batch would hold at most defer_rows rows (and is not the only
                                  one), those leaves are left out of the pass and take rows of the next batch instead: their
                                  games wait one launch longer, nothing else changes -- the order of a game's simulations and
                                  every number in its tree are what they are without it (exp/agent.py:41-45).  az_search
                                  (lock-step: a game has no launch to spare) never defers.  Default 0 = off                */
} az_config;

void az_default_config(az_config* out);
int az_create(const az_config* cfg, az_engine** out);
int az_destroy(az_engine* e);

/* Network weights, fp32, flattened in Network.state_dict() order without the
 * num_batches_tracked counters (exp/policy.py:53-69).  n must be AZ_NUM_WEIGHT_FLOATS.
 * Replaces SimulatePuppet.load_weights (app/base.py:126-129).                              */
#define AZ_NUM_PARAMS 10693458
#define AZ_NUM_BN_STATS 9734
#define AZ_NUM_WEIGHT_FLOATS (AZ_NUM_PARAMS + AZ_NUM_BN_STATS)
int az_set_weights(az_engine* e, const float* flat, size_t n);
/* Version stamp of the weights in use -- LearnPuppet.weights_version (app/base.py:171-174), which the actor sends
 * along with every episode (app/base.py:63-68) so that the learner can drop episodes played with other weights
 * (app/learner.py:51-53).  Every replay tuple carries the stamp in force when its game finished.  az_set_weights
 * adds 1 to the stamp (0 at creation); a learner that numbers its weights itself sets its own value afterwards.  */
int az_set_weights_version(az_engine* e, uint32_t version);

/* (Re)start games: game_ids[i] gets position states[i] (NULL = STARTING_FEN) and two empty
 * trees -- MonteCarloInit.on_episode_begin (exp/callbacks.py:57-62).                        */
int az_reset_games(az_engine* e, const int32_t* game_ids, int n, const mc_state* states);
/* Empty single trees (tree id = 2 * game + k, k = 0 / 1) without touching the game or its other tree: init_mcts() of ONE
 * of the two agents that share an engine (exp/agent.py:105-108).  No simulation may be pending.                          */
int az_reset_trees(az_engine* e, const int32_t* tree_ids, int n);
/* Overwrite the current position of running games without touching their trees (used by the
 * per-agent facade, where the environment owns the game line).                              */
int az_set_positions(az_engine* e, const int32_t* game_ids, int n, const mc_state* states,
                     const int32_t* tree_of_game);

/* -- one simulation, split at the evaluate phase (external evaluator / parity mode) --------
 * az_select_expand: for every active game, one descent from the current root of the tree of
 * the side to move: PUCT select (exp/agent.py:75-88) down to an unvisited or terminal
 * position, expansion of that position (exp/agent.py:57-66).  root_noise is NULL or, per
 * game, MC_MAX_MOVES doubles of Dirichlet noise used if the root is already expanded
 * (exp/agent.py:81-82); noise_used[g] (optional) reports whether it was consumed.
 * Leaves needing a network evaluation are written to the leaf batch.                        */
int az_select_expand(az_engine* e, const double* root_noise, uint8_t* noise_used);
/* The leaf batch of the last az_select_expand: slot i belongs to game i.  needs_eval[i] = 1
 * when (tokens, clock) hold a position to evaluate.  Device pointers owned by the engine.   */
int az_leaf_batch(az_engine* e, const uint8_t** tokens, const float** clocks,
                  const uint8_t** needs_eval, const mc_state** leaf_states, int* n_slots);
/* az_backup: softmax of the legal logits into the new node's priors (exp/agent.py:67-71) and
 * the value backup along the path (exp/agent.py:47-52).  logits [n_slots x 554] and values
 * [n_slots] are float32.  If priors != NULL it supplies per-slot priors [n_slots x
 * MC_MAX_MOVES] directly instead of logits (bit-exact injection for tree-logic parity).     */
int az_backup(az_engine* e, const float* logits, const float* values, const float* priors);

/* az_eval_backup: the evaluate + backup half of one simulation with the built-in network (the leaf batch of
 * the last az_select_expand); lets a caller keep drawing the root noise itself, one simulation at a time,
 * as SimpleAlphaZeroAgent does with numpy's global RNG (exp/agent.py:81-82).                             */
int az_eval_backup(az_engine* e);

/* -- whole searches with the built-in network (throughput mode) ---------------------------*/
int az_search(az_engine* e, int n_sims);

/* n_sims simulations with the built-in network and caller-drawn root noise: simulation s of game g mixes
 * root_noise[(s * n_games + g) * MC_MAX_MOVES ...] into the root priors if the root is expanded by then
 * (exp/agent.py:81-82), exactly like n_sims rounds of az_select_expand(noise_s) + az_eval_backup.  The drop-in
 * MonteCarloTreeSearch.simulate draws the whole block from numpy's global RNG in one call (the same stream as the
 * reference's one dirichlet() per simulation) and makes this single call per move.                              */
int az_search_noise(az_engine* e, int n_sims, const double* root_noise);

/* Continuous self-play (throughput mode, device_rng = 1, leaves_per_step = 1): n_steps network batches.  Every
 * game runs its own loop inside the search kernel -- sims_per_move simulations, then the move choice, replay
 * recording, the move and, when the game is over, a restart (what az_search + az_play_device do for all games
 * in lock-step) -- so a game that needs fewer network rows for a move (terminal or cached leaves) simply moves
 * earlier and the batch stays full.  On return no simulation is pending; a later call carries on where this
 * one stopped.  With the same seed a game slot plays the same games as under az_search + az_play_device.        */
int az_selfplay(az_engine* e, int n_steps, int sims_per_move);

/* Root statistics of the tree of the side to move (exp/policy.py:118-121): sorted legal
 * codes, visit counts N and Q per edge; pi = N / sum(N) is left to the caller.              */
int az_root_stats(az_engine* e, const int32_t* game_ids, int n, uint16_t* codes,
                  uint32_t* visits, double* q, int32_t* n_legal);
/* Statistics of an arbitrary visited position in one tree (MonteCarloTreeSearch.__getitem__,
 * exp/agent.py:38-39): returns 1 in *found when present.  priors may be NULL.               */
int az_node_stats(az_engine* e, int game_id, int tree, const mc_state* state, int* found,
                  uint16_t* codes, uint32_t* visits, double* q, float* priors, int32_t* n_legal,
                  int* is_terminal, double* terminal_value);

/* Play one move in each listed game (exp/environment.py:68-82 on the real game line; fivefold
 * repetition is tracked here).  results[i] is the MC_* result after the move.               */
int az_play(az_engine* e, const int32_t* game_ids, const uint16_t* codes, int n, int8_t* results);
/* Throughput mode: choose (sample while fullmove < tau_change, else argmax with random
 * tie-break, exp/agent.py:113-118), record the replay tuple, play, and restart finished
 * games -- all on device.                                                                   */
int az_play_device(az_engine* e);

int az_game_states(az_engine* e, const int32_t* game_ids, int n, mc_state* states, int8_t* results);

/* Every node of one tree at once -- the whole dicts of MonteCarloTreeSearch (exp/agent.py:25-36: Q, N, P, legal_moves,
 * terminal, visited), in creation order.  HOST output buffers.  Node i: position states[i], info[i] = number of
 * edges | MC_NODE_TERMINAL | MC_NODE_DECISIVE, its edges at [edge_off[i], edge_off[i] + n_edges) of codes / visits /
 * q / priors.  *n_nodes / *n_edges receive the totals; when they exceed max_nodes / max_edges only that many entries
 * are written (call again with larger buffers).  Any of the array pointers may be NULL.                             */
#define MC_NODE_TERMINAL (1u << 16)   /* finished position: no edges; exp/agent.py:59-63 `terminal[fen]`        */
#define MC_NODE_DECISIVE (1u << 17)   /* ... won by the side that moved into it (terminal value -1, else -0)   */
int az_tree_dump(az_engine* e, int game_id, int tree, int max_nodes, mc_state* states, uint32_t* info, uint32_t* edge_off,
                 int* n_nodes, int max_edges, uint16_t* codes, uint32_t* visits, double* q, float* priors, int* n_edges);

/* Replay tuples (exp/callbacks.py:31-54 -> exp/learner.py:23-41), packed per ply.           */
typedef struct az_replay_tuple {
    mc_state observation;            /* position before the move                            */
    uint16_t n_legal;
    uint16_t action;
    int8_t reward;                   /* +1 / 0 / -1 from the mover's point of view, back-filled */
    uint8_t pad[3];
    uint32_t weights_version;        /* az_set_weights_version stamp when the game finished (app/base.py:66) */
    uint16_t codes[MC_MAX_MOVES];
    float pi[MC_MAX_MOVES];
} az_replay_tuple;
/* Moves up to `max` finished-game tuples into out (host or device); *n_out = count.  Tuples beyond `max` stay
 * queued for the next call, in order.  The queue holds 64 * n_games tuples; a finished game is queued whole or, when
 * it does not fit, dropped whole and counted in counter [14] -- so drain at least once per 32 moves of every game
 * (one game is at most 62 plies under the default 30-move cap).                                                   */
int az_drain_replay(az_engine* e, az_replay_tuple* out, int max, int* n_out);

/* Learner-side collate on the device (exp/learner.py:23-41 collate_fn): n packed tuples -> dense pi
 * [n x 554] float32, tokens [n x 60] int64 (= channels [n,2,6,5]), clock [n] and reward [n] float32.    */
int az_collate(const az_replay_tuple* tuples, int n, float* pi, int64_t* tokens, float* clock, float* reward);

/* Counters since creation: [0] simulations, [1] network evaluations, [2] terminal leaves,
 * [3] moves played, [4] games finished, [5] nodes allocated, [6] edges allocated,
 * [7] kernels launched by this library, [8] descents dropped on a pending node (leaves_per_step > 1),
 * [9] simulations whose leaf evaluation came from the evaluation cache ([0] = [1] + [2] + [9]),
 * [10] tree levels descended and [11] edges read on the way (bytes-per-simulation accounting),
 * [12] nodes dropped by the recycler (recycle = 1), [13] network rows whose position was already in the
 * evaluation cache when they were stored (evaluated more than once within one batch),
 * [14] replay tuples dropped because the replay queue was full (whole games; see az_drain_replay),
 * [15] leaf rows a network pass left to the next batch and [16] passes that did so (az_config.defer_rows),
 * [17] nodes dropped by the second-stage compaction of recycle = 1 (not provably unreachable: 0 as long as the
 * trees fit their arenas under the exact rule).                                                               */
#define AZ_NUM_COUNTERS 18
int az_counters(az_engine* e, uint64_t* out);

/* Stand-alone network forward (exp/policy.py:71-80) on n positions with the engine's weights:
 * logits [n x 554], values [n].  Used by the parity tests.                                  */
int az_network_forward(az_engine* e, const uint8_t* tokens, const float* clocks, int n,
                       float* logits, float* values);

/* Measurement hook: with MCAZ_TOWER_STATS=1 in the environment the tower kernel records, per CTA of its last launch, twelve
 * counters {MMA issuer: total cycles, waiting for operands (TMA), waiting for a free accumulator (epilogue); TMA
 * producer: total, waiting for dependencies (previous level), waiting for a free stage; MMA issuer: operand wait at the
 * first stage of a work item, work items; cycles spent issuing e4m3 stages, their number, the same for bf16 stages}.  Copies up to `capacity` words to `out` (host) and returns the number
 * copied (< 0: error).                                                                                                  */
int az_tower_stats(az_engine* e, unsigned long long* out, int capacity);

/* Profiling hook for bench.py: while on, every network forward is bracketed by CUDA events on the
 * engine's stream around its residual tower (18 convolutions: one fused launch, or 18 launches with
 * MCAZ_TOWER=layers; *launches_per_forward says which).  A call with a non-NULL avg_ms_per_tower
 * synchronises, reports the average tower time since the previous call and resets.                 */
int az_profile_network(az_engine* e, int on, double* avg_ms_per_tower, int* n_forwards, int* launches_per_forward);

/* Same for the tree kernel (search_step_kernel: backup + select + expand): total milliseconds and launches since
 * the previous read.                                                                                            */
int az_profile_tree(az_engine* e, int on, double* total_ms, int* n_launches);

/* Measurement hook for the statistical tests: n samples of the root noise that throughput mode (device_rng = 1) mixes into
 * a root with n_edges edges -- Philox4x32-10 -> single-precision Marsaglia-Tsang gamma(alpha) -> normalise, the very
 * device function the search calls -- in place of np.random.dirichlet([alpha] * n_edges) (exp/agent.py:82).
 * out[n x n_edges] float64, host or device.                                                                      */
int az_sample_root_noise(uint64_t seed, float alpha, int n_edges, int n, double* out);

/* Kernels launched by this library in this process (all engines and mc_* calls).               */
uint64_t mcaz_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* MCAZ_H */
