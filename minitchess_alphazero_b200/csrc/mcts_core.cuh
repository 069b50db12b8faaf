// mcts_core.cuh -- AlphaZero tree search over GPU-resident structure-of-arrays trees.
//
// Restates, with the reference's exact arithmetic, exp/agent.py:24-88 (MonteCarloTreeSearch):
//   select   :75-88  u = Q + cpuct*P*sqrt(sum N)/(1+N), first-max argmax, root Dirichlet mix
//   expand   :57-66  unseen position -> terminal record or zeroed Q/N + legal codes
//   evaluate :67-71  softmax of the legal logits -> priors, value
//   backup   :47-52  value = -value; Q = (N*Q + value)/(N+1); N += 1   (float64, this order)
// One warp walks one tree; lanes split the edges of a node.  The same source compiles for the
// host with one "lane" so CPU tests can drive it; the product only launches the kernels.
//
// Layout in HBM (per engine; T = 2*G trees, fixed per-tree arenas of NC nodes / EC edges), records of two 128-bit words:
//   nodes[T*NC]  Node 32 B: {board: piece planes + colour plane | meta, first edge (tree-relative), n_edges | flags, -}
//   edges[T*EC]  Edge 32 B: {stat: Q f64, N u32, P f32 -- all that PUCT reads | link: child node, and, cached from the child's
//                header when the edge is first traversed, the child's first edge and info word, then code u16, virtual loss u16}
//   ht[T*HC] u32  open-addressing table: position -> node (the reference keys its dicts by the
//                 FEN string, so one position reached by two paths is ONE node: a DAG)
// A node's edges are contiguous, lane i of the tree's warp loads edge i with two 128-bit loads, and because the link word
// tells where the child's edges are, a descent costs ONE dependent memory round trip per level (round 1: header, then
// statistics, then child index = three).  The node header is only read at the root, at an untraversed edge (the move has
// to be applied to the parent's position) and by the read-back calls.
#pragma once
#include <math.h>
#include <stdint.h>

#include "minitchess.cuh"
#if defined(__CUDACC__)
#include "philox.cuh"
#endif

namespace az {

constexpr uint32_t NONE = 0xffffffffu;
constexpr int MAX_DEPTH = 64;        // a path cannot be longer than the plies left to the 30-move cap
constexpr int HIST = 64;             // reversible-position history per game (fivefold repetition)
constexpr uint32_t INFO_TERMINAL = 1u << 16;
constexpr uint32_t INFO_DECISIVE = 1u << 17;  // terminal with reward 1.0 (else draw, reward 0.0)
constexpr uint32_t INFO_PENDING = 1u << 18;   // expanded in this step, priors not written yet (leaves_per_step > 1)
constexpr int MAX_LEAVES = 16;                // leaves_per_step limit

// LEAF_CACHED: a new position whose priors and value were found in the exact evaluation cache -- the
// simulation is complete without a network row, like a terminal one.
enum LeafKind : uint8_t { LEAF_NONE = 0, LEAF_EVAL = 1, LEAF_TERMINAL = 2, LEAF_COLLISION = 3, LEAF_CACHED = 4 };
enum Counter : int { C_SIMS = 0, C_EVALS, C_TERMINAL, C_MOVES, C_GAMES, C_NODES, C_EDGES, C_LAUNCHES, C_COLLISIONS, C_CACHED, C_DEPTH, C_PATH_EDGES, C_RECYCLED, C_DUP_ROWS, C_REPLAY_DROPPED, C_DEFERRED_ROWS, C_TRIMMED_BATCHES, C_EVICTED };
enum ErrorBit : int { ERR_NODE_CAP = 1, ERR_EDGE_CAP = 2, ERR_HASH_CAP = 4, ERR_DEPTH = 8, ERR_ILLEGAL = 16 };

struct alignas(16) Board4 { uint32_t x, y, z, w; };
struct alignas(16) NodeHead { uint32_t meta, edge_off, info, pad; };
struct alignas(32) Node { Board4 board; NodeHead head; };
struct alignas(16) EdgeStat { double Q; uint32_t N; float P; };
struct alignas(16) EdgeLink { uint32_t child, child_off, child_info; uint16_t code, vl; };
struct alignas(32) Edge { EdgeStat stat; EdgeLink link; };
static_assert(sizeof(Node) == 32 && sizeof(Edge) == 32 && sizeof(EdgeStat) == 16 && sizeof(EdgeLink) == 16, "tree record layout");

// Exact evaluation cache (SURVEY.md 7.3 point 6): the network sees (tokens, clock) = board, side to move and
// fullmove number -- not the halfmove clock -- so two nodes with that key get bit-identical priors and value,
// whichever tree or game they belong to.  Direct-mapped, newest entry wins; entries carry the weight epoch.
// Every entry is a seqlock (seq odd = being written; a reader accepts only what it read between two equal even
// seq), so readers (tree kernels) and writers (the policy head) need not be ordered by the stream.
constexpr int CACHE_MAX_E = 40;                    // positions with more legal moves are not cached
constexpr uint32_t CACHE_KEY_META = 0x00ff0001u;   // side to move + fullmove number; bits 8-15 of the key word hold n
struct alignas(64) CacheEntry {
    uint32_t seq, pl0, pl1, pl2, white, meta_n, epoch;
    float value;
    float priors[CACHE_MAX_E];
};
static_assert(sizeof(CacheEntry) == 192, "CacheEntry layout");

struct View {
    int G, NC, EC, HC;
    int K;                          // leaves per game per step (1 = the reference's sequential search)
    // real games
    mc_state* game_state; int8_t* game_result; int32_t* game_ply; int32_t* game_start_ply;
    Board4* game_hist; int32_t* game_hist_len;     // [G*HIST] positions since the last irreversible move
    // trees
    uint32_t* tree_nodes; uint32_t* tree_edges; uint32_t* tree_root;
    Node* nodes; Edge* edges;       // [T*NC], [T*EC]; Edge::link.vl = virtual-loss count of descents in flight (K > 1 only)
    uint32_t* ht;
    // per-slot simulation scratch (slot = game * K + leaf index)
    int32_t* path_len; uint32_t* path_edge; uint32_t* path_node;   // [G*MAX_DEPTH]
    uint32_t* leaf_node; uint8_t* leaf_kind; double* leaf_value;
    // leaf batch, one row per slot
    uint8_t* tokens; float* clocks; uint8_t* needs_eval; mc_state* leaf_states;
    unsigned long long* counters; int* error_flag;
    // parameters
    float cpuct; float eps; float alpha; int numpy1; int tau_change; mc_rules rules;
    unsigned long long seed;
    int device_rng;                 // throughput mode: root Dirichlet noise drawn in the select kernel
    unsigned long long sim_counter; // RNG counter of the simulation being run (when sim_serial == nullptr)
    // ---- az_search / az_selfplay bookkeeping (nullptr / 0 elsewhere, e.g. in the host test build)
    unsigned long long* sim_serial; // [G] descents started in this game slot: device RNG key, independent of batching
    uint32_t* move_serial;          // [G] moves chosen on the device: RNG key of the move choice
    int32_t* sims_left;             // [G] descents the game may still start in the running search
    int sims_per_move;              // budget a game gets after each device move (az_selfplay)
    int new_budget;                 // >= 0: the launch first sets sims_left of every game to this (az_search start)
    int free_max;                   // descents a game may start per launch: those that end on a terminal or cached
                                    // position complete on the spot and the game goes straight on to the next
    int async_play;                 // 1: a game whose budget is spent chooses and plays its move inside the search
                                    // kernel and carries on (az_selfplay); 0: it waits for az_play / az_play_device
    // dense leaf rows: the network batch holds only the leaves that need it
    int compact;                    // 1: row = atomic counter (az_search); 0: row = slot (external evaluator API)
    int parity;                     // which of the two row counters this launch fills
    uint32_t* row_count;            // [2]
    int32_t* row_slot;              // [S] row -> slot
    // exact evaluation cache (nullptr = off)
    CacheEntry* cache; uint32_t cache_mask; uint32_t cache_epoch;
    // look-ahead rows (engines with few games): a tower pass costs the same for 1 row as for 256, so every new
    // unfinished node also queues its children as rows of the batch; their (priors, value) only go to the exact cache,
    // where the later simulation that expands such a child finds the very bits the network would give it then
    int row_cap;                    // rows the dense batch holds: slots + look-ahead rows
    int spec_rows;                  // look-ahead rows admitted per batch (0 = off)
    mc_state* row_state;            // [row_cap] position of a look-ahead row (row_slot = -1); the policy head generates its moves
    uint32_t* seen; uint32_t seen_mask;   // tag per cache slot: this position was queued or evaluated (skip it as a child)
    uint32_t* pending_count;        // [2] games that ended this launch waiting for a network row or with budget left (per parity)
    // deferred rows (az_config.defer_rows): the pass evaluates whole tile pairs; a short tail of the batch waits for the next one
    uint32_t* row_eff;              // [2] rows of the batch the pass evaluates (cap_rows_kernel; valid when defer_thr > 0)
    int32_t* slot_row;              // [S] row the slot's pending leaf took
    int defer_thr;                  // > 0: the pass of THIS launch's batch leaves out a last tile pair of <= defer_thr rows
    int check_deferred;             // 1: the pass of the previous batch did so -- a leaf whose row it left out takes a new row
    // caller-supplied root noise of a chained search (az_search_noise): block [noise_budget][G][MC_MAX_MOVES]
    const double* noise_block; int noise_budget;
    // replay recording of the device move choice
    az_replay_tuple* record; az_replay_tuple* replay; unsigned long long* replay_count; unsigned long long replay_cap;
    uint32_t weights_version;       // stamp of the weights in use (az_set_weights_version), written into finished games' tuples
};

#if defined(__CUDA_ARCH__)
#define AZ_LANES 32
#define AZ_SYNCWARP() __syncwarp()
#else
#define AZ_LANES 1
#define AZ_SYNCWARP() ((void)0)
#endif

MC_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
MC_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
MC_HD double ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
MC_HD double dsqrt(double a) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(a);
#else
    return sqrt(a);
#endif
}
MC_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}

// 128-bit accesses to the tree records: one LDG.128 / STG.128 each on the device, plain struct copies on the host
MC_HD NodeHead load_head(const Node* n) {
#if defined(__CUDA_ARCH__)
    const uint4 r = *reinterpret_cast<const uint4*>(&n->head);
    return NodeHead{r.x, r.y, r.z, r.w};
#else
    return n->head;
#endif
}
MC_HD Board4 load_board(const Node* n) {
#if defined(__CUDA_ARCH__)
    const uint4 r = *reinterpret_cast<const uint4*>(&n->board);
    return Board4{r.x, r.y, r.z, r.w};
#else
    return n->board;
#endif
}
MC_HD void store_node(Node* n, const Board4& b, uint32_t meta, uint32_t edge_off, uint32_t info) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(&n->board) = make_uint4(b.x, b.y, b.z, b.w);
    *reinterpret_cast<uint4*>(&n->head) = make_uint4(meta, edge_off, info, 0u);
#else
    n->board = b;
    n->head = NodeHead{meta, edge_off, info, 0u};
#endif
}
MC_HD EdgeStat load_stat(const Edge* e) {
#if defined(__CUDA_ARCH__)
    const uint4 r = *reinterpret_cast<const uint4*>(&e->stat);
    return EdgeStat{__hiloint2double((int)r.y, (int)r.x), r.z, __uint_as_float(r.w)};
#else
    return e->stat;
#endif
}
MC_HD EdgeLink load_link(const Edge* e) {
#if defined(__CUDA_ARCH__)
    const uint4 r = *reinterpret_cast<const uint4*>(&e->link);
    return EdgeLink{r.x, r.y, r.z, (uint16_t)(r.w & 0xffffu), (uint16_t)(r.w >> 16)};
#else
    return e->link;
#endif
}
// A fresh edge: no visits, no prior yet, leads nowhere yet (exp/agent.py:64-65: Q = N = zeros)
MC_HD void store_new_edge(Edge* e, uint16_t code) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(&e->stat) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(&e->link) = make_uint4(0xffffffffu, 0u, 0u, (uint32_t)code);
#else
    e->stat = EdgeStat{0.0, 0u, 0.0f};
    e->link = EdgeLink{0xffffffffu, 0u, 0u, code, 0};
#endif
}

MC_HD void raise(const View& V, int bit) {
#if defined(__CUDA_ARCH__)
    atomicOr(V.error_flag, bit);
#else
    *V.error_flag |= bit;
#endif
}
MC_HD void count(const View& V, int which, unsigned long long n) {
#if defined(__CUDA_ARCH__)
    atomicAdd(&V.counters[which], n);
#else
    V.counters[which] += n;
#endif
}

MC_HD Board4 board_of(const mc_state& s) { return Board4{s.pl0, s.pl1, s.pl2, s.white}; }
MC_HD mc_state state_of(const Board4& b, uint32_t meta) { return mc_state{b.x, b.y, b.z, b.w, meta}; }

// Rows of a dense batch of `count` rows that the network pass evaluates: all of them, or -- when the last tile pair (256 rows)
// would hold at most `thr` rows and is not the only one -- the whole tile pairs only.
MC_HD uint32_t rows_to_run(uint32_t count, uint32_t thr) {
    const uint32_t r = count & 255u;
    return (thr != 0u && count > 256u && r != 0u && r <= thr) ? count - r : count;
}

MC_HD uint32_t hash_state(const mc_state& s) {
    uint32_t h = s.pl0 * 0x9E3779B1u;
    h = (h ^ (h >> 15)) + s.pl1 * 0x85EBCA77u;
    h = (h ^ (h >> 13)) + s.pl2 * 0xC2B2AE3Du;
    h = (h ^ (h >> 16)) + s.white * 0x27D4EB2Fu;
    h = (h ^ (h >> 15)) + s.meta * 0x165667B1u;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}

MC_HD uint32_t cache_hash(const mc_state& s) {
    mc_state k = s;
    k.meta = s.meta & CACHE_KEY_META;
    const uint32_t h = hash_state(k);
    return h ^ (h >> 11) ^ (s.pl0 * 0x2545F491u);   // decorrelated from the per-tree tables
}

// Position -> node of tree t, or NONE.  Scalar: every lane computes the same answer.
MC_HD uint32_t ht_find(const View& V, int t, const mc_state& s) {
    const uint32_t* tab = V.ht + (size_t)t * V.HC;
    uint32_t mask = (uint32_t)V.HC - 1u, h = hash_state(s) & mask;
    for (int probe = 0; probe < V.HC; ++probe) {
        uint32_t e = tab[h];
        if (e == 0) return NONE;
        uint32_t n = e - 1;
        const Node* nd = V.nodes + (size_t)t * V.NC + n;
        const Board4 b = load_board(nd);
        if (b.x == s.pl0 && b.y == s.pl1 && b.z == s.pl2 && b.w == s.white && nd->head.meta == s.meta) return n;
        h = (h + 1) & mask;
    }
    return NONE;
}
// Single writer (lane 0 of the tree's warp).
MC_HD void ht_insert(const View& V, int t, const mc_state& s, uint32_t node) {
    uint32_t* tab = V.ht + (size_t)t * V.HC;
    uint32_t mask = (uint32_t)V.HC - 1u, h = hash_state(s) & mask;
    for (int probe = 0; probe < V.HC; ++probe) {
        if (tab[h] == 0) { tab[h] = node + 1; return; }
        h = (h + 1) & mask;
    }
    raise(V, ERR_HASH_CAP);
}

// Fivefold repetition inside one simulation: the reference builds a fresh Board(fen) per simulation and
// pushes the selected moves, so board.result() at a new leaf sees the positions of this path only
// (python-chess is_fivefold_repetition over the move stack).  A path needs >= 16 plies for that.
MC_HD bool same_position(const Board4& b, uint32_t meta, const mc_state& s) {
    return b.x == s.pl0 && b.y == s.pl1 && b.z == s.pl2 && b.w == s.white && ((meta ^ s.meta) & 1u) == 0;
}
MC_HD int path_repetitions(const View& V, int t, const uint32_t* pnode, int depth, const mc_state& s) {
    int same = 0;
    for (int d = 0; d < depth; ++d) {
        const Node* nd = V.nodes + (size_t)t * V.NC + pnode[d];
        same += same_position(load_board(nd), nd->head.meta, s) ? 1 : 0;
    }
    return same;
}

// Create the node for an unvisited position (exp/agent.py:57-66).  Lane 0 writes; every lane
// gets the node index, its kind and, for a finished position, the value to back up.
MC_HD uint32_t expand(const View& V, int slot, int t, int lane, const mc_state& s, uint8_t* kind, double* value,
                      const uint32_t* pnode, int depth, uint32_t* out_off, uint32_t* out_info) {
    uint32_t node = NONE;
    int is_terminal = 0, decisive = 0;
    *out_off = 0; *out_info = 0;
#if defined(__CUDA_ARCH__)
    if (lane == 0)
#endif
    {
        uint32_t n = V.tree_nodes[t];
        if (n >= (uint32_t)V.NC) {
            raise(V, ERR_NODE_CAP);
        } else {
            uint16_t codes[MC_MAX_MOVES];
            int res;
            int E = mc::generate(s, V.rules, codes, &res);
            if (res == MC_ONGOING && V.rules.fivefold_repetition && depth >= 16 && path_repetitions(V, t, pnode, depth, s) >= 4)
                res = MC_DRAW;
            uint32_t off = V.tree_edges[t];
            size_t gi = (size_t)t * V.NC + n;
            if (res != MC_ONGOING) {
                is_terminal = 1;
                decisive = (res != MC_DRAW);
                E = 0;
            } else if (off + (uint32_t)E > (uint32_t)V.EC) {
                raise(V, ERR_EDGE_CAP);
                E = 0;
                is_terminal = 1;  // keeps the tree consistent; the error flag fails the call
            }
            const uint32_t info = (uint32_t)E | (is_terminal ? INFO_TERMINAL : 0u) | (decisive ? INFO_DECISIVE : 0u) |
                                  ((!is_terminal && V.K > 1) ? INFO_PENDING : 0u);
            store_node(&V.nodes[gi], board_of(s), s.meta, off, info);
            *out_off = off; *out_info = info;
            size_t ge = (size_t)t * V.EC + off;
            for (int k = 0; k < E; ++k) store_new_edge(&V.edges[ge + k], codes[k]);
            V.tree_nodes[t] = n + 1;
            V.tree_edges[t] = off + (uint32_t)E;
            ht_insert(V, t, s, n);
            node = n;
            if (!is_terminal) {
                mc::tokenize(s, V.tokens + (size_t)slot * MC_TOKENS, &V.clocks[slot]);
                V.leaf_states[slot] = s;
            }
            count(V, C_NODES, 1);
            count(V, C_EDGES, (unsigned long long)E);
        }
    }
#if defined(__CUDA_ARCH__)
    node = __shfl_sync(0xffffffffu, node, 0);
    is_terminal = __shfl_sync(0xffffffffu, is_terminal, 0);
    decisive = __shfl_sync(0xffffffffu, decisive, 0);
    __syncwarp();
#endif
    if (node == NONE) { *kind = LEAF_TERMINAL; *value = 0.0; return NONE; }
    if (is_terminal) {
        *kind = LEAF_TERMINAL;
        *value = decisive ? -1.0 : -0.0;  // value = -reward (exp/agent.py:60)
    } else {
        *kind = LEAF_EVAL;
        *value = 0.0;
    }
    return node;
}

#if defined(__CUDACC__)
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int* total) {
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

// Legal targets of the piece on this lane's square (lane = square in the mover's view), warp-cooperative: every lane
// lists the pseudo-legal targets of its own piece, the (piece, target) pairs of the whole position are dealt out one per
// lane, each lane runs the king-safety test of its pair (mc::leaves_king_safe, the predicate mc::legal_targets_by_test
// applies target by target and mc::Guard reproduces per position), and a ballot carries the verdicts back.  A position has 15-30 pseudo-legal moves, so one
// round of ~250 instructions replaces up to ten on the lane that holds the queen.
__device__ __forceinline__ uint32_t legal_targets_warp(const mc::Sets& st, bool white, int type, int sq, int lane, const mc_rules& R) {
    const uint32_t ps = type ? mc::pseudo_targets(st, white, type, sq, R) : 0u;
    const int cnt = mc::popc(ps);
    int tot;
    const int off = warp_excl_scan(cnt, lane, &tot);
    uint32_t tg = 0;
    for (int base = 0; base < tot; base += 32) {
        const int m = base + lane;                         // the pair this lane tests
        int owner = 0;                                     // largest lane whose first pair index is <= m
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int cand = owner + step;
            const int o = __shfl_sync(0xffffffffu, off, cand & 31);
            if (cand < 32 && o <= m) owner = cand;
        }
        const uint32_t owner_ps = __shfl_sync(0xffffffffu, ps, owner);
        const int owner_off = __shfl_sync(0xffffffffu, off, owner);
        const int owner_type = __shfl_sync(0xffffffffu, type, owner);
        const int owner_sq = __shfl_sync(0xffffffffu, sq, owner);
        bool ok = false;
        if (m < tot) {
            const int to = (int)__fns(owner_ps, 0u, m - owner_off + 1);     // the (m - owner_off)-th set bit
            ok = mc::leaves_king_safe(st, white, owner_type, owner_sq, to);
        }
        const uint32_t verdict = __ballot_sync(0xffffffffu, ok);
        // collect the verdicts of this lane's own pairs that fell into this round
        uint32_t rest = ps;
        for (int q = 0; rest; ++q) {
            const int to = mc::lsb(rest);
            rest &= rest - 1;
            const int idx = off + q - base;
            if (idx >= 0 && idx < 32 && ((verdict >> idx) & 1u)) tg |= 1u << to;
        }
    }
    return tg;
}

#endif

#if defined(__CUDA_ARCH__)
// Warp-cooperative expansion (device): lane = square in the mover's view.  Every lane generates the
// legal targets of its own piece (bitboard attack sets), the queen-block and knight-block code counts
// are prefix-summed across the warp, and each lane writes its codes -- already in ascending code order,
// because codes are numbered by view square then (direction, distance) -- and zeroed statistics straight
// into the tree's edge arrays.  Same results as the single-thread mc::generate used on the host.
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {      // L2-coherent load (never a stale L1 line)
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Warp-cooperative move generation of one position (lane = square in the mover's view): legal target sets, the
// prefix sums that place each lane's codes in ascending order, the number of legal moves and the result by the rules.
struct WarpGen {
    mc::Sets st; bool white, knight, pawn; uint32_t tg; int fv, emit_n, off_q, off_n, tot_q, qb, nb, E, res;
};
__device__ __forceinline__ WarpGen warp_generate(const View& V, const mc_state& s, int lane) {
    WarpGen w;
    w.white = mc::white_to_move(s);
    w.st = mc::sets_of(s);
    const int fv = lane;                                   // view square of this lane (30, 31: none)
    w.fv = fv;
    const int sq = w.white ? fv : 29 - fv;
    int type = 0;
    if (fv < 30 && ((w.st.own >> sq) & 1u)) type = mc::piece_at(s, sq);
    w.tg = legal_targets_warp(w.st, w.white, type, sq & 31, lane, V.rules);
    const int r = fv < 30 ? fv / 5 : 0, f = fv < 30 ? fv % 5 : 0;
    // width of this square's slice of the queen / knight code blocks (for the base code) and codes emitted
    int qwidth = 0, nwidth = 0;
    if (fv < 30) {
        qwidth = mc::qcount(r, f);
#pragma unroll
        for (int d = 0; d < 8; ++d) nwidth += mc::n_on(r, f, d) ? 1 : 0;
    }
    w.knight = type == mc::KNIGHT;
    w.pawn = type == mc::PAWN;
    w.emit_n = mc::popc(w.tg);
    if (w.pawn && V.rules.promo_multiplicity > 1)
        w.emit_n += (V.rules.promo_multiplicity - 1) * mc::popc(w.tg & (w.white ? mc::RANK_6 : mc::RANK_1));
    int tot_n, tot_qw, tot_nw, tot_moves;
    w.off_q = warp_excl_scan(w.knight ? 0 : w.emit_n, lane, &w.tot_q);
    w.off_n = warp_excl_scan(w.knight ? w.emit_n : 0, lane, &tot_n);
    w.qb = warp_excl_scan(qwidth, lane, &tot_qw);
    w.nb = 430 + warp_excl_scan(nwidth, lane, &tot_nw);
    warp_excl_scan(mc::popc(w.tg), lane, &tot_moves);
    w.E = w.tot_q + tot_n;
    w.res = mc::result_of(s, w.st, tot_moves, V.rules);
    return w;
}
// This lane's codes, in order, to emit(index within the position's sorted list, code).
template <typename Emit>
__device__ __forceinline__ void warp_emit_codes(const View& V, const WarpGen& w, Emit&& emit) {
    if (w.emit_n <= 0) return;
    const int k = w.knight ? w.tot_q + w.off_n : w.off_q;
    mc::emit_square_codes_indexed(w.fv, w.white, w.knight, w.tg, w.knight ? w.nb : w.qb, w.pawn, V.rules.promo_multiplicity,
                                  [&](int place, uint16_t c) { emit(k + place, c); });
}

// Exact evaluation cache lookup: same (board, side, fullmove, number of legal moves) evaluated before with these
// weights?  On a hit every lane gets the value and the priors of edges lane and lane + 32.
__device__ __forceinline__ bool cache_lookup(const View& V, const mc_state& s, int E, int lane, float* value, float* p0, float* p1) {
    bool hit = false;
    const CacheEntry* c = V.cache + (cache_hash(s) & V.cache_mask);
    const uint32_t s1 = ld_cg_u32(&c->seq);
    if (!(s1 & 1u)) {
        __threadfence();
        const bool key_ok = ld_cg_u32(&c->epoch) == V.cache_epoch && ld_cg_u32(&c->pl0) == s.pl0 && ld_cg_u32(&c->pl1) == s.pl1 &&
                            ld_cg_u32(&c->pl2) == s.pl2 && ld_cg_u32(&c->white) == s.white &&
                            ld_cg_u32(&c->meta_n) == ((s.meta & CACHE_KEY_META) | ((uint32_t)E << 8));
        if (key_ok) {        // the same on every lane unless a writer is at work, which the seq check below catches
            *value = __uint_as_float(ld_cg_u32(reinterpret_cast<const uint32_t*>(&c->value)));
            if (lane < E) *p0 = __uint_as_float(ld_cg_u32(reinterpret_cast<const uint32_t*>(&c->priors[lane])));
            if (lane + 32 < E) *p1 = __uint_as_float(ld_cg_u32(reinterpret_cast<const uint32_t*>(&c->priors[lane + 32])));
            __threadfence();
            hit = ld_cg_u32(&c->seq) == s1;
        }
    }
    return __all_sync(0xffffffffu, hit);
}

// tokens of FEN-order cell i = lane and the clock input: Network.process_observation (exp/policy.py:96-105)
__device__ __forceinline__ void write_network_row(const View& V, int row, int lane, const mc_state& s, const mc::Sets& st, bool white) {
    if (lane < 30) {
        int cell = 5 * (5 - lane / 5) + lane % 5;
        if (!white) cell = 29 - cell;
        const int ty = mc::piece_at(s, cell);
        const bool mine = (st.own >> cell) & 1u;
        V.tokens[(size_t)row * MC_TOKENS + lane] = (uint8_t)(mine ? ty : 0);
        V.tokens[(size_t)row * MC_TOKENS + 30 + lane] = (uint8_t)(mine ? 0 : ty);
    }
    if (lane == 0) {
        const double c = (double)mc::fullmove(s) + (white ? 0.0 : 0.5);
        V.clocks[row] = (float)(c / 30.0);
    }
}

// Tag of a position in the `seen` table (one word per cache slot): "queued or evaluated under these weights".  A stale
// or colliding tag only costs a look-ahead row that is not queued (or one queued twice); results never depend on it.
__device__ __forceinline__ uint32_t seen_tag(const View& V, const mc_state& s) {
    mc_state k = s;
    k.meta = s.meta & CACHE_KEY_META;
    return ((hash_state(k) * 0x2545F491u) ^ (V.cache_epoch * 0x9E3779B1u)) | 0x80000000u;
}

// Look-ahead: queue the children of a new node (its E codes start at edge e0) as rows of the batch, unless they were
// queued or evaluated before or the batch is full.  Lane = child: every lane applies its move, probes the tag table and,
// if its child is new, writes the child's position as a row of the batch; the row's network input (tokens, clock) is made from it
// by tokenize_lookahead_kernel before the pass, and its legal moves are generated later by the policy head, a warp per row -- both
// off this tree's critical path.  Nothing of the tree is touched.
__device__ __forceinline__ void queue_children(const View& V, int lane, const mc_state& s, bool white, size_t e0, int E, bool deep) {
    for (int base = 0; base < E; base += 32) {
        const int i = base + lane;
        bool want = false;
        mc_state cs = s;
        uint32_t idx = 0, tag = 0;
        if (i < E) {
            const uint16_t code = V.edges[e0 + i].link.code;
            if (i == 0 || V.edges[e0 + i - 1].link.code != code) {       // promo_multiplicity > 1 repeats a code
                int fv, tv;
                mc::code_to_view(code, fv, tv);
                cs = mc::apply_move(s, white ? fv : 29 - fv, white ? tv : 29 - tv);
                if (mc::fullmove(cs) <= V.rules.max_fullmoves) {          // else the move cap ends the game there
                    idx = cache_hash(cs) & V.seen_mask;
                    tag = seen_tag(V, cs);
                    want = ld_cg_u32(&V.seen[idx]) != tag;
                }
            }
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, want);
        const int n = mc::popc(ballot);
        if (n == 0) continue;
        // admit while fewer than spec_rows rows are taken: the slots' own rows (atomicAdd, at most G*K) always fit
        int first = 0, take = 0;
        if (lane == 0) {
            uint32_t old = *reinterpret_cast<volatile uint32_t*>(&V.row_count[V.parity]);
            while (old < (uint32_t)V.spec_rows) {
                const uint32_t t = min((uint32_t)n, (uint32_t)V.spec_rows - old);
                const uint32_t got = atomicCAS(&V.row_count[V.parity], old, old + t);
                if (got == old) { first = (int)old; take = (int)t; break; }
                old = got;
            }
        }
        first = __shfl_sync(0xffffffffu, first, 0);
        take = __shfl_sync(0xffffffffu, take, 0);
        const int rank = mc::popc(ballot & ((1u << lane) - 1u));
        const bool queued = want && rank < take;
        if (queued) {
            const int row = first + rank;
            V.row_slot[row] = -1;                        // tokens and clock of a look-ahead row: tokenize_lookahead_kernel, off this path
            V.row_state[row] = cs;
            V.seen[idx] = tag;
        }
        if (take < n) return;                            // batch full
        // A pass is certain (the new node itself needs the network): fill its tile further with the grandchildren.
        // Every lane that queued a child generates that child's moves on its own (scalar code) and queues their
        // positions one by one.
        if (deep && queued) {
            uint16_t codes[MC_MAX_MOVES];
            int res;
            int m = mc::generate(cs, V.rules, codes, &res);
            if (res != MC_ONGOING) m = 0;
            const bool cwhite = mc::white_to_move(cs);
            for (int k = 0; k < m; ++k) {
                if (k > 0 && codes[k] == codes[k - 1]) continue;
                int fv, tv;
                mc::code_to_view(codes[k], fv, tv);
                const mc_state gs = mc::apply_move(cs, cwhite ? fv : 29 - fv, cwhite ? tv : 29 - tv);
                if (mc::fullmove(gs) > V.rules.max_fullmoves) continue;
                const uint32_t gidx = cache_hash(gs) & V.seen_mask, gtag = seen_tag(V, gs);
                if (ld_cg_u32(&V.seen[gidx]) == gtag) continue;
                int row = -1;
                uint32_t old = *reinterpret_cast<volatile uint32_t*>(&V.row_count[V.parity]);
                while (old < (uint32_t)V.spec_rows) {
                    const uint32_t got = atomicCAS(&V.row_count[V.parity], old, old + 1u);
                    if (got == old) { row = (int)old; break; }
                    old = got;
                }
                if (row < 0) break;                      // batch full
                V.row_slot[row] = -1;
                V.row_state[row] = gs;
                V.seen[gidx] = gtag;
            }
        }
        __syncwarp();
    }
}

template <bool LOOKAHEAD>
__device__ __forceinline__ uint32_t expand_warp(const View& V, int slot, int t, int lane, const mc_state& s, uint8_t* kind,
                                                double* value, const uint32_t* pnode, int depth, uint32_t* out_off, uint32_t* out_info) {
    *out_off = 0; *out_info = 0;
    const WarpGen w = warp_generate(V, s, lane);
    const bool white = w.white;
    const mc::Sets& st = w.st;
    int E = w.E;
    int res = w.res;
    if (res == MC_ONGOING && V.rules.fivefold_repetition && depth >= 16) {
        int same = 0;
        for (int d = lane; d < depth; d += 32) {
            const Node* pn = V.nodes + (size_t)t * V.NC + pnode[d];
            same += same_position(load_board(pn), pn->head.meta, s) ? 1 : 0;
        }
        for (int o = 16; o > 0; o >>= 1) same += __shfl_xor_sync(0xffffffffu, same, o);
        if (same >= 4) res = MC_DRAW;
    }
    // lane 0 reserves the node and its edges
    uint32_t node = NONE, off = 0;
    int ok = 1;
    if (lane == 0) {
        const uint32_t n = V.tree_nodes[t];
        off = V.tree_edges[t];
        if (n >= (uint32_t)V.NC) { raise(V, ERR_NODE_CAP); ok = 0; }
        else {
            if (res == MC_ONGOING && off + (uint32_t)E > (uint32_t)V.EC) { raise(V, ERR_EDGE_CAP); ok = 2; }
            node = n;
        }
    }
    node = __shfl_sync(0xffffffffu, node, 0);
    off = __shfl_sync(0xffffffffu, off, 0);
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (ok == 0) { *kind = LEAF_TERMINAL; *value = 0.0; return NONE; }
    const bool terminal = (res != MC_ONGOING) || ok == 2;
    const bool decisive = (res == MC_WHITE_WINS || res == MC_BLACK_WINS);
    if (terminal) E = 0;
    const size_t gi = (size_t)t * V.NC + node;
    // exact evaluation cache: same (board, side, fullmove) evaluated before with these weights?
    bool hit = false;
    float hit_value = 0.f, hit_p0 = 0.f, hit_p1 = 0.f;      // priors of edges lane and lane + 32
    if (!terminal && V.cache && E <= CACHE_MAX_E) hit = cache_lookup(V, s, E, lane, &hit_value, &hit_p0, &hit_p1);
    if (!terminal) {
        const size_t w0 = (size_t)t * V.EC + off;
        warp_emit_codes(V, w, [&](int k, uint16_t c) { store_new_edge(&V.edges[w0 + k], c); });
    }
    if (hit) {
        __syncwarp();            // the zeroed priors above are overwritten by the cached ones
        const size_t w0 = (size_t)t * V.EC + off;
        if (lane < E) V.edges[w0 + lane].stat.P = hit_p0;
        if (lane + 32 < E) V.edges[w0 + lane + 32].stat.P = hit_p1;
    }
    // network row of this leaf: dense (az_search) or the slot itself
    int row = slot;
    const bool needs_net = !terminal && !hit;
    if (needs_net && V.compact) {
        if (lane == 0) { row = (int)atomicAdd(&V.row_count[V.parity], 1u); V.row_slot[row] = slot; if (V.slot_row) V.slot_row[slot] = row; }
        row = __shfl_sync(0xffffffffu, row, 0);
    }
    if (needs_net) write_network_row(V, row, lane, s, st, white);
    const uint32_t info = (uint32_t)E | (terminal ? INFO_TERMINAL : 0u) | ((terminal && decisive) ? INFO_DECISIVE : 0u) |
                          ((needs_net && V.K > 1) ? INFO_PENDING : 0u);
    *out_off = off; *out_info = info;
    if (lane == 0) {
        store_node(&V.nodes[gi], board_of(s), s.meta, off, info);
        V.tree_nodes[t] = node + 1;
        V.tree_edges[t] = off + (uint32_t)E;
        ht_insert(V, t, s, node);
        if (needs_net) V.leaf_states[slot] = s;
        count(V, C_NODES, 1);
        count(V, C_EDGES, (unsigned long long)E);
    }
    __syncwarp();
    if (LOOKAHEAD && !terminal && V.spec_rows > 0 && V.cache) {
        if (lane == 0) V.seen[cache_hash(s) & V.seen_mask] = seen_tag(V, s);      // evaluated now, or found in the cache
        // grandchildren too when a pass is certain and the node sits at most two plies below the root, where the
        // search is dense enough to come by them (deeper down they mostly go unvisited and only cost time)
        queue_children(V, lane, s, white, (size_t)t * V.EC + off, E, needs_net && depth <= 2);
    }
    if (terminal) { *kind = LEAF_TERMINAL; *value = decisive ? -1.0 : -0.0; }
    else if (hit) { *kind = LEAF_CACHED; *value = (double)hit_value; }
    else { *kind = LEAF_EVAL; *value = 0.0; }
    return node;
}
#define AZ_EXPAND expand_warp<LOOKAHEAD>
#else
#define AZ_EXPAND expand
#endif

#if defined(__CUDACC__)
// Throughput mode's root noise (exp/agent.py:81-82 without numpy's stream): Dirichlet(alpha) over the root's E edges = gamma(alpha)
// draws over their sum.  Warp-cooperative: this lane draws the gammas of edges lane, lane + 32, lane + 64 (Philox keyed by
// seed, game slot, the slot's simulation serial and the edge, so a sample does not depend on how descents are batched into
// launches); returns the sum over all edges.  az_sample_root_noise draws from this very function for the statistical tests.
__device__ __forceinline__ double root_noise_gammas(unsigned long long seed, int g, unsigned long long counter, int E, int lane,
                                                    float alpha, double gam[3]) {
    double part = 0.0;
    for (int i = lane, kk = 0; i < E && kk < 3; i += 32, ++kk) {
        mcaz::Philox rng(seed, (uint32_t)g, (uint32_t)counter, (uint32_t)(counter >> 32) ^ ((uint32_t)i << 16));
        gam[kk] = rng.gamma((double)alpha);
        part += gam[kk];
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    return part;
}
#endif

// One simulation of game g down to its leaf (exp/agent.py:54-88 without the backup).
// `noise`: per-game Dirichlet sample [MC_MAX_MOVES] or nullptr.
// Returns the leaf kind (the same on every lane).
template <bool LOOKAHEAD = false>
MC_HD uint8_t select_expand_one(const View& V, int g, int lane, const double* noise, uint8_t* noise_used, int j = 0) {
    const int slot = g * V.K + j;
    if (lane == 0) {
        V.leaf_kind[slot] = LEAF_NONE;
        V.needs_eval[slot] = 0;
        V.path_len[slot] = 0;
        if (noise_used && j == 0) noise_used[g] = 0;
    }
    if (V.game_result[g] != MC_ONGOING) return LEAF_NONE;
    // device RNG counter of this descent: the game's own serial number when the engine keeps one (results then do
    // not depend on how descents are batched into launches), else the caller's global counter
    unsigned long long rng_counter = V.sim_counter;
    if (V.sim_serial) {
        rng_counter = V.sim_serial[g];
        AZ_SYNCWARP();
        if (lane == 0) V.sim_serial[g] = rng_counter + 1;
    }
    const int t = 2 * g + (V.game_ply[g] & 1);
    const size_t nbase = (size_t)t * V.NC, ebase = (size_t)t * V.EC;
    uint32_t* pedge = V.path_edge + (size_t)slot * MAX_DEPTH;
    uint32_t* pnode = V.path_node + (size_t)slot * MAX_DEPTH;
    const bool vloss = V.K > 1;

    uint8_t kind = LEAF_NONE;
    double value = 0.0;
    uint32_t node = V.tree_root[t];
    int depth = 0;
    unsigned int path_edges = 0;       // edges read on the way down (bytes-per-simulation accounting)
    uint32_t info = 0, off = 0;        // header words of `node`: from its own header at the root, from the parent's link below
    if (node == NONE) {
        mc_state s = V.game_state[g];
        node = ht_find(V, t, s);
        if (node == NONE) node = AZ_EXPAND(V, slot, t, lane, s, &kind, &value, pnode, 0, &off, &info);
        if (lane == 0 && node != NONE) V.tree_root[t] = node;
        AZ_SYNCWARP();
    }
    if (kind == LEAF_NONE) {
        const NodeHead h = load_head(&V.nodes[nbase + node]);
        info = h.info; off = h.edge_off;
    }
    while (kind == LEAF_NONE) {
        if (info & INFO_TERMINAL) {  // exp/agent.py:75-77: revisit backs up -terminal[node]
            kind = LEAF_TERMINAL;
            value = (info & INFO_DECISIVE) ? 1.0 : 0.0;
            break;
        }
        if (info & INFO_PENDING) {   // another descent of this step owns the node: drop this one (K > 1 only)
            kind = LEAF_COLLISION;
            value = 0.0;
            break;
        }
        const int E = (int)(info & 0xffffu);
        path_edges += (unsigned int)E;
        const size_t e0 = ebase + off;
        // this lane's edges: both 128-bit words of edge `lane` stay in registers (positions with more than 32 legal moves are rare:
        // their edges 32.. are read where they are used, and again from L1); the host build reads everything in place
#if defined(__CUDA_ARCH__)
        EdgeStat est0{0.0, 0u, 0.0f};
        EdgeLink elk0{NONE, 0u, 0u, 0, 0};
        if (lane < E) { est0 = load_stat(&V.edges[e0 + lane]); elk0 = load_link(&V.edges[e0 + lane]); }
#define AZ_STAT(i, kk) ((kk) == 0 ? est0 : load_stat(&V.edges[e0 + (i)]))
#define AZ_LINK(i, kk) ((kk) == 0 ? elk0 : load_link(&V.edges[e0 + (i)]))
#define AZ_MY_EDGES_BEGIN _Pragma("unroll") for (int kk = 0; kk < 3; ++kk) { const int i = lane + 32 * kk; if (i < E) {
#else
#define AZ_STAT(i, kk) V.edges[e0 + (i)].stat
#define AZ_LINK(i, kk) V.edges[e0 + (i)].link
#define AZ_MY_EDGES_BEGIN for (int i = 0; i < E; ++i) { {
#endif
#define AZ_MY_EDGES_END }}
        // sum of visit counts (exact in float64: small integers)
        unsigned int nsum_u = 0;
        AZ_MY_EDGES_BEGIN nsum_u += AZ_STAT(i, kk).N + (vloss ? (unsigned int)AZ_LINK(i, kk).vl : 0u); AZ_MY_EDGES_END
#if defined(__CUDA_ARCH__)
        for (int o = 16; o > 0; o >>= 1) nsum_u += __shfl_xor_sync(0xffffffffu, nsum_u, o);
#endif
        const double root_n = dsqrt((double)nsum_u);
        bool mix = (depth == 0) && (noise != nullptr) && (V.eps > 0.0f);
#if defined(__CUDA_ARCH__)
        // throughput mode: Dirichlet(alpha) over the root's edges drawn here, a fresh sample every simulation
        double gam[3] = {0.0, 0.0, 0.0}, gsum = 1.0;
        const bool dev_noise = (depth == 0) && (noise == nullptr) && V.device_rng && (V.eps > 0.0f);
        if (dev_noise) {
            gsum = root_noise_gammas(V.seed, g, rng_counter, E, lane, V.alpha, gam);
            mix = true;
        }
#endif
        double best_u = 0.0;
        int best_i = 0x7fffffff;
        AZ_MY_EDGES_BEGIN
            const EdgeStat es = AZ_STAT(i, kk);
            const float p = es.P;
            double x;
            if (mix) {
                // P = (1-eps)*P + eps*dirichlet: float32 product, float64 sum (exp/agent.py:82 under NEP 50)
#if defined(__CUDA_ARCH__)
                const double nz = dev_noise ? gam[kk] / gsum : noise[(size_t)g * MC_MAX_MOVES + i];
#else
                const double nz = noise[(size_t)g * MC_MAX_MOVES + i];
#endif
                double pn = dadd((double)fmul((float)(1.0 - (double)V.eps), p), dmul((double)V.eps, nz));
                x = dmul(dmul((double)V.cpuct, pn), root_n);
            } else if (V.numpy1) {
                // numpy 1.x value-based casting: float32 array * float64 scalar stays float32 (Q6)
                x = (double)fmul(fmul(V.cpuct, p), (float)root_n);
            } else {
                x = dmul((double)fmul(V.cpuct, p), root_n);
            }
            double q_i = es.Q, n_i = (double)es.N;
            if (vloss) {   // descents in flight count as visits that lost (virtual loss)
                const double vl = (double)AZ_LINK(i, kk).vl;
                if (vl > 0.0) { q_i = (n_i * q_i - vl) / (n_i + vl); n_i += vl; }
            }
            const double u = dadd(q_i, ddiv(x, dadd(1.0, n_i)));
            if (best_i == 0x7fffffff || u > best_u) { best_u = u; best_i = i; }  // first max within the lane
        AZ_MY_EDGES_END
#if defined(__CUDA_ARCH__)
        for (int o = 16; o > 0; o >>= 1) {
            double ou = __shfl_xor_sync(0xffffffffu, best_u, o);
            int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (oi != 0x7fffffff && (best_i == 0x7fffffff || ou > best_u || (ou == best_u && oi < best_i))) {
                best_u = ou;
                best_i = oi;
            }
        }
#endif
        if (mix && lane == 0 && noise_used) noise_used[g] = 1;
        if (depth >= MAX_DEPTH) { raise(V, ERR_DEPTH); kind = LEAF_TERMINAL; value = 0.0; break; }
        const size_t e = e0 + (size_t)best_i;
        // where the chosen edge leads: its link word is already in the registers of the lane that owns it
#if defined(__CUDA_ARCH__)
        EdgeLink lk = elk0;
        if (best_i >= 32) lk = load_link(&V.edges[e]);          // every lane reads the same record
        else {
            const int owner = best_i;
            lk.child = __shfl_sync(0xffffffffu, lk.child, owner);
            lk.child_off = __shfl_sync(0xffffffffu, lk.child_off, owner);
            lk.child_info = __shfl_sync(0xffffffffu, lk.child_info, owner);
            lk.code = (uint16_t)__shfl_sync(0xffffffffu, (uint32_t)lk.code, owner);
        }
#else
        const EdgeLink lk = V.edges[e].link;
#endif
#undef AZ_STAT
#undef AZ_LINK
#undef AZ_MY_EDGES_BEGIN
#undef AZ_MY_EDGES_END
        if (lane == 0) {
            pedge[depth] = (uint32_t)(e - ebase);
            pnode[depth] = node;
            if (vloss) V.edges[e].link.vl += 1;
        }
        ++depth;
        uint32_t child = lk.child;
        if (child == NONE) {
            // first traversal of this edge: apply the move to the parent's position and look the result up
            const Node* pn = &V.nodes[nbase + node];
            mc_state ps = state_of(load_board(pn), pn->head.meta);
            int fv, tv;
            mc::code_to_view(lk.code, fv, tv);
            const bool white = mc::white_to_move(ps);
            mc_state cs = mc::apply_move(ps, white ? fv : 29 - fv, white ? tv : 29 - tv);
            child = ht_find(V, t, cs);
            uint32_t c_off = 0, c_info = 0;
            if (child == NONE) { AZ_SYNCWARP(); child = AZ_EXPAND(V, slot, t, lane, cs, &kind, &value, pnode, depth, &c_off, &c_info); }
            else { const NodeHead h = load_head(&V.nodes[nbase + child]); c_off = h.edge_off; c_info = h.info; }    // a transposition
            if (lane == 0 && child != NONE) {
                // the link caches what a descent needs of the child's header (its edge count and terminal flags never change)
                EdgeLink* l = &V.edges[e].link;
                l->child = child; l->child_off = c_off; l->child_info = c_info & ~INFO_PENDING;
            }
            AZ_SYNCWARP();
            if (child == NONE) break;
            info = c_info; off = c_off;
        } else {
            info = lk.child_info; off = lk.child_off;
            if (vloss) info = load_head(&V.nodes[nbase + child]).info;    // INFO_PENDING comes and goes: the header has it
        }
        node = child;
    }
    if (lane == 0) {
        V.path_len[slot] = depth;
        V.leaf_node[slot] = node;
        V.leaf_kind[slot] = kind;
        V.leaf_value[slot] = value;
        V.needs_eval[slot] = (kind == LEAF_EVAL) ? 1 : 0;
        if (kind != LEAF_COLLISION) { count(V, C_DEPTH, (unsigned long long)depth); count(V, C_PATH_EDGES, path_edges); }
    }
    return kind;
}

// Evaluate + backup for game g (exp/agent.py:67-72 and :47-52).
// logits: [G x 554] float32 (softmax over the legal entries, float32) or priors: [G x MC_MAX_MOVES].
MC_HD void backup_one(const View& V, int g, int lane, const float* logits, const float* values, const float* priors, int j = 0) {
    const int slot = g * V.K + j;
    const uint8_t kind = V.leaf_kind[slot];
    if (kind == LEAF_NONE) return;
    const int t = 2 * g + (V.game_ply[g] & 1);
    const size_t nbase = (size_t)t * V.NC, ebase = (size_t)t * V.EC;
    double value = V.leaf_value[slot];
    if (kind == LEAF_EVAL) {
        const uint32_t node = V.leaf_node[slot];
        const NodeHead h = load_head(&V.nodes[nbase + node]);
        const int E = (int)(h.info & 0xffffu);
        const size_t e0 = ebase + h.edge_off;
        if (priors) {
            for (int i = lane; i < E; i += AZ_LANES) V.edges[e0 + i].stat.P = priors[(size_t)slot * MC_MAX_MOVES + i];
        } else if (logits) {
            const float* lg = logits + (size_t)slot * MC_NUM_ACTIONS;
            float m = -INFINITY;
            for (int i = lane; i < E; i += AZ_LANES) m = fmaxf(m, lg[V.edges[e0 + i].link.code]);
#if defined(__CUDA_ARCH__)
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
#endif
            float sum = 0.f;
            for (int i = lane; i < E; i += AZ_LANES) sum += expf(lg[V.edges[e0 + i].link.code] - m);
#if defined(__CUDA_ARCH__)
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#endif
            for (int i = lane; i < E; i += AZ_LANES) V.edges[e0 + i].stat.P = expf(lg[V.edges[e0 + i].link.code] - m) / sum;
        }   // else: the policy head already wrote the priors into the edges (heads_legal_kernel)
        value = (double)values[slot];
        if (V.K > 1 && lane == 0) V.nodes[nbase + node].head.info &= ~INFO_PENDING;
    }
    if (lane == 0) {
        const uint32_t* pedge = V.path_edge + (size_t)slot * MAX_DEPTH;
        const bool vloss = V.K > 1;
        for (int d = V.path_len[slot] - 1; d >= 0; --d) {
            Edge* ed = &V.edges[ebase + pedge[d]];
            if (vloss) ed->link.vl -= 1;
            if (kind == LEAF_COLLISION) continue;      // dropped descent: only its virtual loss is undone
            value = -value;
            const EdgeStat st = load_stat(ed);          // one 128-bit read; Q and N go back (the prior is someone else's to write)
            const double n = (double)st.N;
            ed->stat.Q = ddiv(dadd(dmul(n, st.Q), value), dadd(n, 1.0));
            ed->stat.N = st.N + 1u;
        }
        if (kind == LEAF_COLLISION) count(V, C_COLLISIONS, 1);
        else {
            count(V, C_SIMS, 1);
            count(V, kind == LEAF_EVAL ? C_EVALS : (kind == LEAF_CACHED ? C_CACHED : C_TERMINAL), 1);
        }
        V.leaf_kind[slot] = LEAF_NONE;
    }
}

// ---- the real game line -----------------------------------------------------------------------
MC_HD Board4 rep_key(const mc_state& s) { return Board4{s.pl0, s.pl1, s.pl2, s.white | ((s.meta & 1u) << 31)}; }

// Result of the current position of game g including fivefold repetition over the game line
// (python-chess is_fivefold_repetition over the move stack; exp/environment.py:39).
MC_HD int game_result_of(const View& V, int g, const mc_state& s) {
    uint16_t codes[MC_MAX_MOVES];
    int res;
    mc::generate(s, V.rules, codes, &res);
    if (res == MC_ONGOING && V.rules.fivefold_repetition) {
        const Board4 key = rep_key(s);
        const Board4* h = V.game_hist + (size_t)g * HIST;
        int same = 0, n = V.game_hist_len[g];
        for (int i = 0; i < n; ++i) same += (h[i].x == key.x && h[i].y == key.y && h[i].z == key.z && h[i].w == key.w);
        if (same >= 5) res = MC_DRAW;
    }
    return res;
}

MC_HD void hist_reset(const View& V, int g, const mc_state& s) {
    V.game_hist[(size_t)g * HIST] = rep_key(s);
    V.game_hist_len[g] = 1;
}

// Play `code` in game g (exp/environment.py:68-82).  Returns 0 ok, 1 illegal, 2 finished.  Scalar.
MC_HD int play_one(const View& V, int g, int code) {
    if (V.game_result[g] != MC_ONGOING) return 2;
    mc_state s = V.game_state[g], o;
    int st = mc::step(s, code, V.rules, &o);
    if (st != 0) return st;
    const bool zeroing = mc::halfmove(o) == 0;
    if (zeroing) V.game_hist_len[g] = 0;
    int n = V.game_hist_len[g];
    if (n < HIST) { V.game_hist[(size_t)g * HIST + n] = rep_key(o); V.game_hist_len[g] = n + 1; }
    V.game_state[g] = o;
    V.game_ply[g] += 1;
    V.tree_root[2 * g] = NONE;
    V.tree_root[2 * g + 1] = NONE;
    V.game_result[g] = (int8_t)game_result_of(V, g, o);
    count(V, C_MOVES, 1);
    if (V.game_result[g] != MC_ONGOING) count(V, C_GAMES, 1);
    return 0;
}

}  // namespace az
