// c_api.cu -- library-level entry points of libmcaz.so and host-only helpers (FEN strings,
// action-code tables).  No GPU work here.
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "minitchess.cuh"

namespace mcaz {

static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }
int fail(int code, const std::string& msg) {
    t_last_error = msg;
    return code;
}

Scratch& thread_scratch() {
    static thread_local Scratch sc;
    return sc;
}

int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(MCAZ_ENODEV, "no usable CUDA device (this library has no CPU fallback)");
    }
    return MCAZ_OK;
}

}  // namespace mcaz

using namespace mcaz;

extern "C" {

void mc_default_rules(mc_rules* out) {
    if (!out) return;
    out->pawn_double_step = 0;
    out->promo_multiplicity = 1;
    out->max_fullmoves = 30;
    out->insufficient_material = 1;
    out->fivefold_repetition = 1;
}

int mcaz_abi_version(void) { return MCAZ_ABI_VERSION; }
size_t mcaz_struct_size(int which) {
    switch (which) {
        case 0: return sizeof(mc_state);
        case 1: return sizeof(mc_rules);
        case 2: return sizeof(az_config);
        case 3: return sizeof(az_replay_tuple);
        default: return 0;
    }
}
const char* mcaz_last_error(void) { return t_last_error.c_str(); }

int mcaz_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mcaz_set_device(int device) {
    if (int rc = require_device()) return rc;
    MCAZ_CUDA(cudaSetDevice(device));
    return MCAZ_OK;
}

uint64_t mcaz_kernel_launches(void) { return g_launches.load(); }

// '0prbnqk' numbering (exp/policy.py:7)
static int type_of_char(char c) {
    switch (c | 0x20) {
        case 'p': return mc::PAWN;
        case 'r': return mc::ROOK;
        case 'b': return mc::BISHOP;
        case 'n': return mc::KNIGHT;
        case 'q': return mc::QUEEN;
        case 'k': return mc::KING;
        default: return 0;
    }
}

int mc_state_from_fen(const char* fen, mc_state* out) {
    if (!fen || !out) return fail(MCAZ_EINVAL, "mc_state_from_fen: null argument");
    mc_state s = {0, 0, 0, 0, 0};
    int rank = 5, file = 0;
    const char* p = fen;
    for (; *p && *p != ' '; ++p) {
        char c = *p;
        if (c == '/') {
            if (file != 5) return fail(MCAZ_EINVAL, std::string("bad FEN row width: ") + fen);
            --rank;
            file = 0;
        } else if (c >= '1' && c <= '5') {
            file += c - '0';
        } else {
            int t = type_of_char(c);
            if (!t || rank < 0 || file > 4) return fail(MCAZ_EINVAL, std::string("bad FEN piece: ") + fen);
            uint32_t b = 1u << (5 * rank + file);
            if (t & 1) s.pl0 |= b;
            if (t & 2) s.pl1 |= b;
            if (t & 4) s.pl2 |= b;
            if (c >= 'A' && c <= 'Z') s.white |= b;
            ++file;
        }
    }
    if (rank != 0 || file != 5) return fail(MCAZ_EINVAL, std::string("bad FEN board: ") + fen);
    char turn = 0;
    int half = 0, full = 0;
    if (sscanf(p, " %c %d %d", &turn, &half, &full) != 3 || (turn != 'w' && turn != 'b') || half < 0 || half > 255 ||
        full < 0 || full > 255)
        return fail(MCAZ_EINVAL, std::string("bad FEN fields: ") + fen);
    s.meta = MC_META(turn == 'w', half, full);
    *out = s;
    return MCAZ_OK;
}

int mc_state_to_fen(const mc_state* s, char* buf, size_t buflen) {
    if (!s || !buf) return fail(MCAZ_EINVAL, "mc_state_to_fen: null argument");
    static const char sym[] = ".prbnqk";
    char tmp[64];
    int k = 0;
    for (int rank = 5; rank >= 0; --rank) {
        int run = 0;
        for (int file = 0; file < 5; ++file) {
            int sq = 5 * rank + file, t = mc::piece_at(*s, sq);
            if (t == 0 || t == 7) { ++run; continue; }
            if (run) { tmp[k++] = (char)('0' + run); run = 0; }
            char c = sym[t];
            tmp[k++] = ((s->white >> sq) & 1u) ? (char)(c - 32) : c;
        }
        if (run) tmp[k++] = (char)('0' + run);
        if (rank) tmp[k++] = '/';
    }
    tmp[k] = 0;
    int n = snprintf(buf, buflen, "%s %c %d %d", tmp, (s->meta & 1u) ? 'w' : 'b', mc::halfmove(*s), mc::fullmove(*s));
    if (n < 0 || (size_t)n >= buflen) return fail(MCAZ_EINVAL, "mc_state_to_fen: buffer too small");
    return MCAZ_OK;
}

int mc_code_squares(int code, int white_to_move, int* from_sq, int* to_sq) {
    int fv, tv;
    if (!from_sq || !to_sq || !mc::code_to_view(code, fv, tv)) return fail(MCAZ_EINVAL, "mc_code_squares: bad code");
    *from_sq = white_to_move ? fv : 29 - fv;
    *to_sq = white_to_move ? tv : 29 - tv;
    return MCAZ_OK;
}

int mc_squares_code(int from_sq, int to_sq, int white_to_move) {
    if (from_sq < 0 || from_sq > 29 || to_sq < 0 || to_sq > 29) return -1;
    return white_to_move ? mc::view_to_code(from_sq, to_sq) : mc::view_to_code(29 - from_sq, 29 - to_sq);
}

}  // extern "C"
