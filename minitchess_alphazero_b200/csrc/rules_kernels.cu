// rules_kernels.cu -- stateless MinitChess rules on packed positions, sm_100a.
//
// Replaces on the path: exp/environment.py:34-50 (_update_attributes: result + sorted legal
// codes), :68-82 (step), exp/policy.py:82-105 (tokeniser), and perft-style validation of the
// python-chess fork's move generator.  HBM-bound integer work: one thread per position,
// grid-stride over a grid sized to the SM count; positions are 20-byte records read once,
// outputs written once.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "minitchess.cuh"

namespace mcaz {

static int g_num_sms = 0;

int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

static inline int grid_for(size_t n, int block, int ctas_per_sm) {
    size_t want = (n + block - 1) / block;
    size_t cap = (size_t)num_sms() * ctas_per_sm;
    return (int)std::max<size_t>(1, std::min(want, cap));
}

__global__ void __launch_bounds__(128) legal_moves_kernel(const mc_state* __restrict__ states, int n, mc_rules R,
                                                          uint16_t* __restrict__ codes, int32_t* __restrict__ counts,
                                                          int8_t* __restrict__ results) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        mc_state s = states[i];
        int res;
        int c = mc::generate(s, R, codes + (size_t)i * MC_MAX_MOVES, &res);
        counts[i] = c;
        results[i] = (int8_t)res;
    }
}

__global__ void __launch_bounds__(128) apply_kernel(const mc_state* __restrict__ states, const uint16_t* __restrict__ codes,
                                                    int n, mc_rules R, mc_state* __restrict__ out,
                                                    int8_t* __restrict__ status) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        mc_state s = states[i], o;
        int st = mc::step(s, codes[i], R, &o);
        out[i] = o;
        status[i] = (int8_t)st;
    }
}

__global__ void __launch_bounds__(128) tokenize_kernel(const mc_state* __restrict__ states, int n,
                                                       uint8_t* __restrict__ tokens, float* __restrict__ clocks) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        mc_state s = states[i];
        uint8_t t[MC_TOKENS];
        float c;
        mc::tokenize(s, t, &c);
        uint32_t* dst = reinterpret_cast<uint32_t*>(tokens + (size_t)i * MC_TOKENS);  // 60 B rows: 4-byte aligned
#pragma unroll
        for (int k = 0; k < MC_TOKENS / 4; ++k)
            dst[k] = (uint32_t)t[4 * k] | ((uint32_t)t[4 * k + 1] << 8) | ((uint32_t)t[4 * k + 2] << 16) |
                     ((uint32_t)t[4 * k + 3] << 24);
        clocks[i] = c;
    }
}

// perft, level synchronous: expand a frontier into the next one (positions carry their root id);
// the last level only adds move counts.  Finished positions have no successors.
__global__ void __launch_bounds__(128) perft_expand_kernel(const mc_state* __restrict__ frontier,
                                                           const uint32_t* __restrict__ root_of, int n, mc_rules R,
                                                           int last_level, mc_state* __restrict__ next,
                                                           uint32_t* __restrict__ next_root,
                                                           unsigned long long* __restrict__ next_count,
                                                           unsigned long long capacity,
                                                           unsigned long long* __restrict__ nodes) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        mc_state s = frontier[i];
        uint16_t codes[MC_MAX_MOVES];
        int res;
        int c = mc::generate(s, R, codes, &res);
        if (res != MC_ONGOING || c == 0) continue;
        uint32_t root = root_of[i];
        if (last_level) {
            atomicAdd(&nodes[root], (unsigned long long)c);
            continue;
        }
        unsigned long long base = atomicAdd(next_count, (unsigned long long)c);
        if (base + c > capacity) continue;  // host checks next_count against capacity
        bool white = mc::white_to_move(s);
        int prev = -1;
        for (int k = 0; k < c; ++k) {
            int fv, tv;
            if ((int)codes[k] == prev) {  // promo multiplicity > 1: same 4-char move, one successor each
                next[base + k] = next[base + k - 1];
                next_root[base + k] = root;
                continue;
            }
            prev = codes[k];
            mc::code_to_view(codes[k], fv, tv);
            next[base + k] = mc::apply_move(s, white ? fv : 29 - fv, white ? tv : 29 - tv);
            next_root[base + k] = root;
        }
    }
}

__global__ void iota_kernel(uint32_t* p, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = (uint32_t)i;
}

}  // namespace mcaz

using namespace mcaz;

extern "C" {

int mc_legal_moves(const mc_state* states, int n, const mc_rules* rules, uint16_t* codes, int32_t* counts,
                   int8_t* results) {
    if (n < 0 || (n > 0 && (!states || !codes || !counts || !results))) return fail(MCAZ_EINVAL, "mc_legal_moves: null buffer");
    if (int rc = require_device()) return rc;
    if (n == 0) return MCAZ_OK;
    cudaStream_t st = 0;
    Scratch& sc = thread_scratch();
    sc.begin();
    In<mc_state> in;
    Out<uint16_t> oc;
    Out<int32_t> on;
    Out<int8_t> orr;
    if (int rc = in.init(states, n, st, sc)) return rc;
    if (int rc = oc.init(codes, (size_t)n * MC_MAX_MOVES, st, sc, true)) return rc;
    if (int rc = on.init(counts, n, st, sc)) return rc;
    if (int rc = orr.init(results, n, st, sc)) return rc;
    legal_moves_kernel<<<grid_for(n, 128, 8), 128, 0, st>>>(in.ptr, n, rules_or_default(rules), oc.ptr, on.ptr, orr.ptr);
    MCAZ_CHECK_LAUNCH();
    if (int rc = oc.finish(st)) return rc;
    if (int rc = on.finish(st)) return rc;
    if (int rc = orr.finish(st)) return rc;
    MCAZ_CUDA(cudaStreamSynchronize(st));
    return MCAZ_OK;
}

int mc_apply(const mc_state* states, const uint16_t* codes, int n, const mc_rules* rules, mc_state* out,
             int8_t* status) {
    if (n < 0 || (n > 0 && (!states || !codes || !out || !status))) return fail(MCAZ_EINVAL, "mc_apply: null buffer");
    if (int rc = require_device()) return rc;
    if (n == 0) return MCAZ_OK;
    cudaStream_t st = 0;
    Scratch& sc = thread_scratch();
    sc.begin();
    In<mc_state> in;
    In<uint16_t> ic;
    Out<mc_state> oo;
    Out<int8_t> os;
    if (int rc = in.init(states, n, st, sc)) return rc;
    if (int rc = ic.init(codes, n, st, sc)) return rc;
    if (int rc = oo.init(out, n, st, sc)) return rc;
    if (int rc = os.init(status, n, st, sc)) return rc;
    apply_kernel<<<grid_for(n, 128, 8), 128, 0, st>>>(in.ptr, ic.ptr, n, rules_or_default(rules), oo.ptr, os.ptr);
    MCAZ_CHECK_LAUNCH();
    if (int rc = oo.finish(st)) return rc;
    if (int rc = os.finish(st)) return rc;
    MCAZ_CUDA(cudaStreamSynchronize(st));
    return MCAZ_OK;
}

int mc_tokenize(const mc_state* states, int n, uint8_t* tokens, float* clocks) {
    if (n < 0 || (n > 0 && (!states || !tokens || !clocks))) return fail(MCAZ_EINVAL, "mc_tokenize: null buffer");
    if (int rc = require_device()) return rc;
    if (n == 0) return MCAZ_OK;
    cudaStream_t st = 0;
    Scratch& sc = thread_scratch();
    sc.begin();
    In<mc_state> in;
    Out<uint8_t> ot;
    Out<float> ok;
    if (int rc = in.init(states, n, st, sc)) return rc;
    if (int rc = ot.init(tokens, (size_t)n * MC_TOKENS, st, sc)) return rc;
    if (int rc = ok.init(clocks, n, st, sc)) return rc;
    if (reinterpret_cast<uintptr_t>(ot.ptr) & 3u) return fail(MCAZ_EINVAL, "mc_tokenize: tokens must be 4-byte aligned");
    tokenize_kernel<<<grid_for(n, 128, 8), 128, 0, st>>>(in.ptr, n, ot.ptr, ok.ptr);
    MCAZ_CHECK_LAUNCH();
    if (int rc = ot.finish(st)) return rc;
    if (int rc = ok.finish(st)) return rc;
    MCAZ_CUDA(cudaStreamSynchronize(st));
    return MCAZ_OK;
}

int mc_perft(const mc_state* roots, int n, int depth, const mc_rules* rules, uint64_t* nodes) {
    if (n < 0 || depth < 0 || (n > 0 && (!roots || !nodes))) return fail(MCAZ_EINVAL, "mc_perft: bad argument");
    if (int rc = require_device()) return rc;
    if (n == 0) return MCAZ_OK;
    cudaStream_t st = 0;
    mc_rules R = rules_or_default(rules);
    if (depth == 0) {
        std::vector<uint64_t> ones(n, 1);
        if (is_device_pointer(nodes)) MCAZ_CUDA(cudaMemcpy(nodes, ones.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice));
        else std::memcpy(nodes, ones.data(), n * sizeof(uint64_t));
        return MCAZ_OK;
    }
    Scratch& sc = thread_scratch();
    sc.begin();
    In<mc_state> in;
    Out<uint64_t> on;
    if (int rc = in.init(roots, n, st, sc)) return rc;
    if (int rc = on.init(nodes, n, st, sc, true)) return rc;
    // ping-pong frontiers, grown on demand and kept by the calling thread between calls while they are small
    // (cudaMalloc / cudaFree of a 100 MB frontier cost more than expanding it: 5-30 ms against < 1 ms)
    struct Frontiers {
        mc_state* fr[2] = {nullptr, nullptr};
        uint32_t* rt[2] = {nullptr, nullptr};
        size_t cap[2] = {0, 0};
        unsigned long long* d_count = nullptr;
        int device = -1;
        void release() {
            for (int k = 0; k < 2; ++k) {
                if (fr[k]) cudaFree(fr[k]);
                if (rt[k]) cudaFree(rt[k]);
                fr[k] = nullptr; rt[k] = nullptr; cap[k] = 0;
            }
            if (d_count) cudaFree(d_count);
            d_count = nullptr;
        }
    };
    static thread_local Frontiers F;
    int dev_now = 0;
    cudaGetDevice(&dev_now);
    if (F.device != dev_now) { F.release(); F.device = dev_now; }
    mc_state** fr = F.fr;
    uint32_t** rt = F.rt;
    size_t* cap = F.cap;
    unsigned long long*& d_count = F.d_count;
    int rc = MCAZ_OK;
    auto cleanup = [&]() {
        const size_t kept = (cap[0] + cap[1]) * (sizeof(mc_state) + sizeof(uint32_t));
        if (kept > ((size_t)256 << 20)) F.release();
    };
    auto ensure = [&](int k, size_t need) -> int {
        if (cap[k] >= need) return MCAZ_OK;
        if (fr[k]) cudaFree(fr[k]);
        if (rt[k]) cudaFree(rt[k]);
        fr[k] = nullptr; rt[k] = nullptr; cap[k] = 0;
        MCAZ_CUDA(cudaMalloc(&fr[k], need * sizeof(mc_state)));
        MCAZ_CUDA(cudaMalloc(&rt[k], need * sizeof(uint32_t)));
        cap[k] = need;
        return MCAZ_OK;
    };
    do {
        if ((rc = ensure(0, n))) break;
        if (!d_count && cudaMalloc(&d_count, sizeof(unsigned long long)) != cudaSuccess) { rc = fail(MCAZ_ECUDA, "cudaMalloc"); break; }
        cudaMemcpyAsync(fr[0], in.ptr, (size_t)n * sizeof(mc_state), cudaMemcpyDeviceToDevice, st);
        iota_kernel<<<grid_for(n, 256, 4), 256, 0, st>>>(rt[0], n);
        g_launches.fetch_add(1);
        size_t cur_n = n;
        int cur = 0;
        for (int level = 1; level <= depth && cur_n > 0; ++level) {
            int last = (level == depth);
            size_t need = last ? 1 : cur_n * 40 + 1024;  // generous guess; retried below if too small
            for (;;) {
                if (!last && (rc = ensure(cur ^ 1, need))) break;
                cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st);
                perft_expand_kernel<<<grid_for(cur_n, 128, 8), 128, 0, st>>>(
                    fr[cur], rt[cur], (int)cur_n, R, last, fr[cur ^ 1], rt[cur ^ 1], d_count,
                    (unsigned long long)cap[cur ^ 1], reinterpret_cast<unsigned long long*>(on.ptr));
                g_launches.fetch_add(1);
                unsigned long long produced = 0;
                if (cudaMemcpyAsync(&produced, d_count, sizeof(produced), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                    cudaStreamSynchronize(st) != cudaSuccess) { rc = fail(MCAZ_ECUDA, std::string("perft: ") + cudaGetErrorString(cudaGetLastError())); break; }
                if (last) { cur_n = 0; break; }
                if (produced > cap[cur ^ 1]) {  // frontier did not fit: redo this level with the exact size
                    need = produced;
                    if (last == 0) {  // the last-level adds are not issued on non-last levels, so a retry is safe
                        continue;
                    }
                }
                if (produced > 0x7fffffffULL) { rc = fail(MCAZ_ECAPACITY, "mc_perft: frontier exceeds 2^31 positions"); break; }
                cur_n = (size_t)produced;
                cur ^= 1;
                break;
            }
            if (rc) break;
        }
    } while (0);
    if (!rc) rc = on.finish(st);
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = fail(MCAZ_ECUDA, "mc_perft: sync failed");
    cleanup();
    return rc;
}

}  // extern "C"
