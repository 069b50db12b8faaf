// net.cu -- the policy/value network of exp/policy.py:53-80 as hand-written sm_100a kernels.
//
//   stem_onehot_kernel tokens -> one-hot rows (7 mover's + 7 opponent's token channels of 64) for the stem level
//   tower_tc_kernel    level 0: Embedding(7,4) + Conv3x3(8->256) + BN + ReLU as a one-hot x folded-table MMA (K = 16 per
//                      tap); levels 1-18: the tower convolutions 256->256: tcgen05.mma.cta_group::2 (UMMA 256x256x16 over a
//                      CTA pair, bf16 in, fp32 accumulate in TMEM), operands staged by TMA (SWIZZLE_128B,
//                      K-major), 6-stage mbarrier pipeline, double-buffered TMEM accumulators, fused
//                      bias(+BN) / residual / ReLU / bf16 epilogue; one data-flow ordered launch; works on the
//                      tile pairs that hold rows of the (dense) leaf batch, work items claimed at run time; the epilogue
//                      of the last level also takes the three 1x1 head convolutions (fp32 dot products of the row it
//                      holds) and writes 90 floats per board instead of the 15 KB activation rows
//   heads_kernel       policy head (61->554 linear) and value head (31->256 -> 1, tanh) on those sums, fp32;
//                      heads_legal_kernel: only the leaf's legal logits, softmax into the tree's priors,
//                      evaluation-cache insert (az_search)
//
// Tower data layout in HBM: act[pos 30][board Bpad][channel 256] bf16 (two ping-pong buffers).
// With boards as the GEMM M dimension a 3x3 tap is a plain shift of the *position* index, so the
// implicit GEMM needs no im2col and no halo: for output position p,
//     D[128 boards x 256 cout] = sum over valid taps t, cin:  act[p + t][boards][cin] * W[t][cout][cin]
// and taps that fall off the 6x5 board are simply skipped (208 of 270 tap-positions remain).
// Weights: W[layer 18][tap 9][cout 256][cin 256] bf16 with BatchNorm folded in, 21.2 MB, L2-resident.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.cuh"

namespace mcaz {
int num_sms();

namespace {

constexpr int C = 256;            // tower width
constexpr int NPOS = 30;
constexpr int NLAYERS = 18;       // 9 residual blocks x 2 convolutions
constexpr int NLEVELS = NLAYERS + 1;   // work-item levels of the tower kernel: 0 = stem, 1..18 = the convolutions
constexpr int HEAD_IN = 96;       // floats per board handed to the heads: p0[30], p1[30], -, v[30] at 0 / 30 / 61 (raw 1x1 sums)
constexpr int BLOCK_M = 128;      // boards per tile
constexpr int BLOCK_K = 64;       // bf16 elements = one 128-byte swizzle row
#ifndef TOWER_STAGES
#define TOWER_STAGES 6      // measured: a seventh stage (225 KB of shared memory, no L1 left) slows the epilogue by more than it saves
#endif
constexpr int STAGES = TOWER_STAGES;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB: this CTA's 128 boards x 64 input channels
constexpr int B_BYTES = (C / 2) * BLOCK_K * 2;   // 16 KB: this CTA's half of the 256 output channels
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STEM_K = 16;         // channels of a one-hot stem row (14 live): one K = 16 MMA, one 32-byte swizzle row
constexpr int STEM_TILE_BYTES = BLOCK_M * STEM_K * 2;   // 4 KB: 128 boards (or 128 output channels) x 16 channels
// The e4m3 form runs TWO sets of four epilogue warps, one per TMEM accumulator buffer (items alternate between them): an e4m3 item's
// main loop is half as long as a bf16 one and no longer hides a 256 x 256 epilogue, so each set gets two item times for its item.
#ifndef TOWER_BF16_SETS
#define TOWER_BF16_SETS 1      // measured: two sets in the bf16 form too -- see profiles/r02_fp8_epilogue_ab.txt
#endif
__host__ __device__ constexpr int tower_epilogue_sets(bool fp8) { return fp8 ? 2 : TOWER_BF16_SETS; }
__host__ __device__ constexpr int tower_threads(bool fp8) { return 128 + 128 * tower_epilogue_sets(fp8); }
constexpr int CALIB_ROWS = 2048;                 // positions of the e4m3 tower's calibration pass (az_set_weights)
constexpr int MAX_CHUNK_BOARDS = 8192;           // boards per forward pass (32 tile pairs: three L2 groups)

// flat state_dict offsets (floats), exp/policy.py:56-69 order without num_batches_tracked
constexpr size_t OFF_EMB = 0;
constexpr size_t OFF_STEM = 28;                         // w[256][8][3][3], b, gamma, beta, mean, var
constexpr size_t OFF_TOWER = 19740;
constexpr size_t TOWER_STRIDE = 589824 + 5 * 256;
constexpr size_t OFF_PCONV = 10659612;                  // w[2][256], b2, gamma2, beta2, mean2, var2
constexpr size_t OFF_PLIN = 10660134;                   // w[554][61], b[554]
constexpr size_t OFF_VCONV = 10694482;                  // w[256], b, gamma, beta, mean, var
constexpr size_t OFF_V1 = 10694743;                     // w[256][31], b[256]
constexpr size_t OFF_V2 = 10702935;                     // w[256], b
constexpr float BN_EPS = 1e-5f;

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// Same wait with a suspend-time hint: the thread may stay blocked in hardware for up to `ns` before try_wait returns
// empty-handed (it still wakes as soon as the phase completes), so a long wait costs a handful of instructions
// instead of one spin every ~100 ns -- the waits of this kernel are 32-lane warps, and the run is power-capped.
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t addr = smem_u32(bar), ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(ns)
            : "memory");
    } while (!ok);
}
// ---- CTA-pair (cta_group::2) variants
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even CTA of the pair
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}
// TMA load whose completion bytes are credited to the pair leader's mbarrier
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// e4m3 x e4m3 -> fp32 (kind::f8f6f4, dense): K = 32 per instruction, twice the multiply-adds of the bf16 form per shared-memory byte
__device__ __forceinline__ void umma_fp8_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// four floats -> four e4m3 bytes (round to nearest even, saturating at +-448), lowest address first
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
    uint16_t lo, hi;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(b), "f"(a));     // the first source lands in the upper byte
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(d), "f"(c));
    return (uint32_t)lo | ((uint32_t)hi << 16);
}
// commit: arrive on the barrier at this offset in both CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

// One lane of the (converged) warp; the others skip the block.  With the loop control kept warp-wide, ptxas keeps descriptors,
// addresses and counters in uniform registers and issues UTCHMMA / UTMALDG / UTCBAR straight, instead of wrapping each in a
// loop that makes a lane's operands uniform -- a single thread running ~90 dependent instructions per stage was what bounded
// the round-1 tower (640 cycles per stage against 512 cycles of tensor work).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), LBO in [16,30) (unused for swizzled K-major), SBO = 1024 B (8 rows x
// 128 B) in [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// The stem level's operands: rows of 16 channels = 32 bytes, SWIZZLE_32B (layout type 6), 8-row groups 256 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10),
// both K-major, N >> 3 in [17,23), M >> 4 in [24,29).
// cta_group::2: one instruction spans the CTA pair, M = 256 boards (128 per CTA), N = 256.
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
// kind::f8f6f4: A = B = E4M3 (format 0), D = F32; same shape fields
constexpr uint32_t IDESC_FP8 = (1u << 4) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
constexpr float E4M3_MAX = 448.0f;

__device__ __forceinline__ bool tap_valid(int pos, int tap, int& src) {
    int row = pos / 5 + tap / 3 - 1, col = pos % 5 + tap % 3 - 1;
    src = row * 5 + col;
    return row >= 0 && row < 6 && col >= 0 && col < 5;
}

// ---------------------------------------------------------------------------------- residual tower
// One work item = (level, output position, pair of 128-board tiles), computed by a CTA pair with
// tcgen05.mma.cta_group::2: each CTA stages its own boards (A) and half of the output channels (B), so every
// SM reads and writes half the weight bytes of the single-CTA form -- the shared-memory port, not the tensor
// pipe, is what limits a 1-CTA 128x256 SS-mode MMA.  Roles per CTA: warp 0 lane 0 TMA producer, warp 1 lane 0
// MMA issuer (leader CTA only), warp 2 TMEM allocator, warp 3 work-item scheduler (leader) + dependency watcher,
// warps 4-7 epilogue of this CTA's 128 accumulator rows.  TOWER_STAGES-deep smem ring, double-buffered TMEM accumulators.
//
// Default form: stem + 18 convolutions are ONE persistent launch.  The items of a launch form one list in data-flow order
// (groups of tile pairs that fit the L2 together; within a group level by level, tile pair by tile pair, the positions with
// the most taps first), and the CTA pairs CLAIM items from it at run time with one atomicAdd per item: a pair that runs
// faster (nearer L2 slice, cheaper items) simply takes more.  Item (L, p, tp) may start once the items (L-1, p', tp) for
// the valid taps p' of p have published their outputs (per-item epoch flags in global memory, release/acquire at gpu
// scope); that one rule also covers the write-after-read hazards of the two ping-pong activation buffers.  Because an
// item is only ever claimed by a pair that is running, and its inputs are items claimed earlier, the waits cannot
// deadlock whatever part of the grid is resident: nothing has to be co-resident.  (Round 1 dealt the items out
// ahead of time by estimated cost, one table per batch size; measured, 4-8 % of the kernel went into waits on
// pairs that had fallen behind their share, and a launch next to another kernel could hang.)
// Per-level form (MCAZ_TOWER=layers): the same kernel launched once per level with flags == nullptr --
// kernel boundaries order the levels.  Bit-identical: an item's arithmetic does not depend on who runs it when.
#ifndef SPIN_NS
#define SPIN_NS 40     // back-off of the waits on the dependency watcher: a hot spin costs issue slots and power
#endif
constexpr int TOWER_SMEM = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int TOWER_SMEM_FP8 = TOWER_SMEM;      // the e4m3 form reads its dequantisation factors through L1: a larger carve-out (no L1
                                                // left for the spilled locals of the issue loops) cost more than the table saved
static_assert(TOWER_SMEM_FP8 <= 227 * 1024, "tower shared memory");
constexpr unsigned long long WATCHDOG_CYCLES = 20ull * 1000 * 1000 * 1000;   // ~10 s: a dependency that never arrives
constexpr int ITEM_RING = 16;          // claimed items in flight per CTA (scheduler -> producer / MMA / epilogue); see tower_scheduler
#ifndef TOWER_GROUP_PAIRS
#define TOWER_GROUP_PAIRS 12
#endif
constexpr int GROUP_MAX_PAIRS = TOWER_GROUP_PAIRS;    // tile pairs that walk the levels together (their ping-pong buffers stay in the 126 MB L2)

struct TowerParams {
    float bias[NLEVELS * C];      // [19][256]: stem, then the 18 convolutions.  In the kernel's parameter space (constant
                                  // bank): every epilogue thread reads the same word at the same time
    const float* bias_g;          // the same in global memory, for the out-of-line head epilogue
    const float4* head_w;         // [256]: folded 1x1 filters policy 0, policy 1, value, 0 per input channel
    float* head_in;               // [bpad][HEAD_IN]: written by the last level's epilogue
    __nv_bfloat16* act0;          // layer input of even layers / residual + output of odd layers
    __nv_bfloat16* act1;
    uint32_t* flags;              // [19][n_pairs][30][2] epoch stamps; nullptr = one level per launch, no dependencies
    uint32_t* claim;              // [2]: next item of the list, CTA pairs that have left the kernel (resets both)
    const uint32_t* count;        // device row count of this forward (dense leaf batch) or nullptr: all n_pairs are live
    uint32_t row_base;            // first row of this chunk within the batch that `count` counts
    int bpad;
    int n_pairs;
    uint32_t epoch;
    int first_level;              // levels below this one are not in the list (0; timing experiments only)
    int n_levels;                 // levels in the list from first_level on (19; 1 in the per-level form)
    int fuse_heads;               // 1: the last level's epilogue takes the head convolutions (0 only in timing experiments)
    int group_max;                // tile pairs per L2 group
    uint32_t wait_hint;           // suspend-time hint (ns) of the epilogue warps' waits for an accumulator; 0 = plain try_wait spin
    unsigned long long* stats;    // MCAZ_TOWER_STATS=1: per CTA {MMA issuer: total, waiting for operands, waiting for an accumulator;
                                  // TMA producer: total, waiting for dependencies, waiting for a free stage; MMA issuer: operand wait
                                  // at the first stage of an item, items} in clock cycles; else nullptr
    uint8_t pos_order[32];        // the 30 positions, those with the most valid taps first
    // ---- e4m3 tower (network = 2): operands of the 18 convolutions in fp8, fp32 accumulation, residual stream in bf16
    const float* scale_g;         // [19][256] dequantisation factor per level and output channel: weight scale of the channel x
                                  // activation scale of the level's input (row 0, the stem, is 1)
    float inv_a[NLEVELS];         // 1 / activation scale of each level's output: what the epilogue multiplies by before rounding to e4m3
    uint8_t* actq0;               // e4m3 copy of the block inputs x (levels 0, 2, 4, ...): A operand of a block's first convolution
    uint8_t* actq1;               // e4m3 h between the two convolutions of a block (levels 1, 3, ...); never kept in bf16
    unsigned int* level_absmax;   // calibration launch (bf16 form): per level the largest activation written, as float bits (values >= 0)
    int fp8_levels;               // levels 1 .. fp8_levels (an even number: whole residual blocks) multiply e4m3 operands, the rest bf16
    float l2_budget;              // bytes of the L2's persisting set-aside the activations may fill (0: none granted)
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t* p, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// 256-bit accesses (sm_100: LDG.256 / STG.256; .cg = L2-coherent, never a stale L1 line across levels): a thread's 32 channels of
// a row are 64 contiguous bytes = two whole 32-byte sectors, so the activation rows travel as full sectors in half as many L2
// requests as with 128-bit accesses
__device__ __forceinline__ void ld_cg_v8(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.cg.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p) : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void st_v8(void* p, const uint4& a, const uint4& b) {      // TOWER_L2_MODE 0
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// ---- L2 residency of the activation ping-pong buffers.  A group's activations are written at one level and read back at the
// next: traffic that should live and die in the L2.  Plain stores do not achieve that -- ncu: 740 MB written to and 565 MB read
// from HBM per 2816-row launch, against 27 MB of algorithmic bytes -- because the L2 replaces lines least-recently-used and a
// working set that cycles just above its capacity is LRU's worst case.  So the activation stores of the first `l2_positions`
// output positions carry an evict_last policy (createpolicy): those lines stay in the L2's persisting set-aside
// (cudaLimitPersistingL2CacheSize, sized in network_create) from level to level and launch to launch, get overwritten in place
// and never travel; the positions that would not fit keep the normal policy.  The split is by position, so it is the same lines
// every time (a fractional policy picks lines by a hash the program cannot align with the set-aside).  TOWER_L2_MODE 0 = plain stores.
#ifndef TOWER_L2_MODE
#define TOWER_L2_MODE 1
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_v8_hint(void* p, const uint4& a, const uint4& b, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "l"(pol) : "memory");
}
#if TOWER_L2_MODE
#define ST_ACT(p, a, b) st_v8_hint(p, a, b, l2pol)
#else
#define ST_ACT(p, a, b) st_v8(p, a, b)
#endif

// ---- claimed work items: a ring of 64-bit slots per CTA, slot k % ITEM_RING = (item << 32) | (k + 1).  One aligned
// 64-bit store publishes an item (item and sequence number can never be seen apart), to this CTA and to its peer.
constexpr uint32_t ITEM_END = 0xffffffffu;
__device__ __forceinline__ void ring_put(unsigned long long* ring, uint32_t k, uint32_t item, uint32_t peer) {
    const unsigned long long v = ((unsigned long long)item << 32) | (unsigned long long)(k + 1u);
    const uint32_t addr = smem_u32(&ring[k % ITEM_RING]);
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "st.volatile.shared::cta.b64 [%0], %1;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %2;\n\t"
        "st.volatile.shared::cluster.b64 [ra], %1;\n\t}"
        ::"r"(addr), "l"(v), "r"(peer)
        : "memory");
}
__device__ __forceinline__ uint32_t ring_get(const unsigned long long* ring, uint32_t k) {
    const uint32_t addr = smem_u32(&ring[k % ITEM_RING]);
    unsigned long long v;
    for (;;) {
        asm volatile("ld.volatile.shared::cta.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
        if ((uint32_t)v == k + 1u) break;
        __nanosleep(SPIN_NS);
    }
    return (uint32_t)(v >> 32);
}
// Item number idx of the launch's list -> level << 24 | tile pair << 8 | position, or ITEM_END past the end.
__device__ __forceinline__ uint32_t tower_item(const TowerParams& P, int live_pairs, uint32_t idx) {
    const int n_groups = (live_pairs + P.group_max - 1) / P.group_max;
    for (int gi = 0; gi < n_groups; ++gi) {
        const int g0 = (int)((long long)live_pairs * gi / n_groups), g1 = (int)((long long)live_pairs * (gi + 1) / n_groups);
        const uint32_t per_level = (uint32_t)(g1 - g0) * NPOS, cnt = per_level * (uint32_t)P.n_levels;
        if (idx < cnt) {
            const uint32_t L = (uint32_t)P.first_level + idx / per_level, r = idx % per_level;
            return (L << 24) | ((uint32_t)(g0 + (int)(r / NPOS)) << 8) | (uint32_t)P.pos_order[r % NPOS];
        }
        idx -= cnt;
    }
    return ITEM_END;
}

// Epilogue of the last level for one accumulator row (this thread's board at one position): 8 chunks of 32 channels.
// Kept out of line so that its registers do not weigh on the common epilogue.  The residual row is published before
// the item's MMAs start, so four chunks of it are fetched ahead of the accumulator and refilled as they are used.
// (A template over the calling kernel's form: ptxas 12.9 crashes on one out-of-line function shared by several tcgen05 kernels.)
template <int FORM>
__device__ __noinline__ void epilogue_heads(uint32_t taddr, const float* __restrict__ bias, const float* __restrict__ scale,
                                            const float4* __restrict__ hw, const __nv_bfloat16* res_row, uint64_t* acc_full,
                                            uint32_t acc_phase, uint32_t wait_hint, float* h) {
    constexpr int RD = tower_epilogue_sets((FORM & 2) != 0) == 2 ? 2 : 4;       // chunks of the residual row fetched ahead (the e4m3 form's 384 threads have 168 registers each)
    uint4 res[RD][4];
#pragma unroll
    for (int c = 0; c < RD; ++c) {
        ld_cg_v8(res_row + c * 32, res[c][0], res[c][1]);
        ld_cg_v8(res_row + c * 32 + 16, res[c][2], res[c][3]);
    }
    if (wait_hint) mbar_wait_hint(acc_full, acc_phase, wait_hint); else mbar_wait(acc_full, acc_phase);
    tc_fence_after();
    float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
    for (int c = 0; c < C / 32; ++c) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr + c * 32)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
                const int e = j * 8 + hh * 2;
                const uint32_t r = (&res[c & (RD - 1)][j].x)[hh];
                float a0 = __uint_as_float(v[e]), a1 = __uint_as_float(v[e + 1]);
                if (scale) { a0 *= __ldg(scale + c * 32 + e); a1 *= __ldg(scale + c * 32 + e + 1); }      // e4m3 tower: dequantise
                const float x0 = fmaxf(a0 + __ldg(bias + c * 32 + e) + __uint_as_float(r << 16), 0.f);
                const float x1 = fmaxf(a1 + __ldg(bias + c * 32 + e + 1) + __uint_as_float(r & 0xffff0000u), 0.f);
                const float4 w0 = __ldg(hw + c * 32 + e), w1 = __ldg(hw + c * 32 + e + 1);
                h0 = fmaf(x0, w0.x, h0); h1 = fmaf(x0, w0.y, h1); h2 = fmaf(x0, w0.z, h2);
                h0 = fmaf(x1, w1.x, h0); h1 = fmaf(x1, w1.y, h1); h2 = fmaf(x1, w1.z, h2);
            }
        }
        if (c + RD < C / 32) {
            ld_cg_v8(res_row + (c + RD) * 32, res[c & (RD - 1)][0], res[c & (RD - 1)][1]);
            ld_cg_v8(res_row + (c + RD) * 32 + 16, res[c & (RD - 1)][2], res[c & (RD - 1)][3]);
        }
    }
    h[0] = h0; h[1] = h1; h[2] = h2;
}

// FP8: the 18 convolutions multiply e4m3 operands (map_q0 / map_q1 / map_wq; the bf16 maps are then unused).  CALIB (bf16 form
// only): the epilogue also records the largest activation of every level -- the calibration pass of the e4m3 tower.
template <bool FP8, bool CALIB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tower_threads(FP8), 1)
tower_tc_kernel(const __grid_constant__ CUtensorMap map_act0, const __grid_constant__ CUtensorMap map_act1,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_stem_in,
                const __grid_constant__ CUtensorMap map_stem_w, const __grid_constant__ CUtensorMap map_q0,
                const __grid_constant__ CUtensorMap map_q1, const __grid_constant__ CUtensorMap map_wq,
                const __grid_constant__ TowerParams P) {
    static_assert(!(FP8 && CALIB), "the calibration pass runs the bf16 form");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* acc_full = bars + 2 * STAGES;
    uint64_t* acc_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    __shared__ unsigned long long s_ring[ITEM_RING];   // claimed items (written by the leader's scheduler, here and in the peer)
    __shared__ uint32_t s_deps_ok;            // items [0, s_deps_ok) have all their inputs published (written by warp 3)
    __shared__ uint32_t s_prod_at;            // item the TMA producer of this CTA is loading (flow control of the scheduler)
    __shared__ __align__(16) float s_scale_all[FP8 ? 2 * C : 4];   // e4m3 form: the dequantisation factors of the level each epilogue set is working on

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    // dense leaf batch: only the tile pairs that hold rows are in the list
    int live_pairs = P.n_pairs;
    if (P.count) {
        const uint32_t total = __ldg(P.count), rows = total > P.row_base ? total - P.row_base : 0u;
        live_pairs = (int)min((uint32_t)P.n_pairs, (rows + 2 * BLOCK_M - 1) / (2 * BLOCK_M));
    }

#if TOWER_L2_MODE
    // output positions whose rows stay in the persisting part of the L2: as many as the set-aside holds of one group's two buffers
    int l2_positions = 0;
    {
        const int n_groups = max(1, (live_pairs + P.group_max - 1) / P.group_max);
        const int group_pairs = (live_pairs + n_groups - 1) / n_groups;
        const float per_position = (float)max(group_pairs, 1) * (float)(2 * BLOCK_M) * (float)(C * 2 * 2);    // both buffers, bf16
        l2_positions = min(NPOS, (int)(P.l2_budget / per_position));
    }
    const uint64_t l2pol_last = l2_policy_evict_last(), l2pol_normal = l2_policy_evict_normal();
#endif
    if (threadIdx.x < ITEM_RING) s_ring[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) { s_deps_ok = 0; s_prod_at = 0; }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer (both CTAs): the whole warp runs the loop
        // control, one elected lane issues the loads
        uint32_t stage = 0, phase = 0;
        const bool stats = P.stats != nullptr;
        const long long p_start = stats ? clock64() : 0;
        long long p_deps = 0, p_slot = 0;
        for (uint32_t k = 0;; ++k) {
            const uint32_t item = ring_get(s_ring, k);
            if (item == ITEM_END) break;
            if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&s_prod_at) = k;
            const int pos = item & 0xff, tp = (item >> 8) & 0xffff, L = item >> 24;
            const int tile = 2 * tp + (int)rank;
            if (L > P.first_level && P.flags) {
                // inputs published? (warp 3 polls the global flags ahead of us)
                const long long t0 = stats ? clock64() : 0;
                while (ld_acquire_cta_shared(&s_deps_ok) <= k) __nanosleep(SPIN_NS);
                if (stats) p_deps += clock64() - t0;
                asm volatile("fence.proxy.async;" ::: "memory");   // order the acquired writes before our TMA reads
            }
            // level 0 (stem): one-hot rows x folded embedding/conv table, one 16-channel chunk (32-byte rows) per tap;
            // level L >= 1: convolution L - 1 reads act0 (even) / act1 (odd), four chunks per tap (two of 128 e4m3 channels)
            const bool q8 = FP8 && L >= 1 && L <= P.fp8_levels;       // this level's operands are e4m3: 128 elements per 128-byte row
            const int kchunk = q8 ? 128 : BLOCK_K;
            const int conv = L - 1, chunks = L == 0 ? 1 : C / kchunk;
            const CUtensorMap* map_in = L == 0 ? &map_stem_in : (q8 ? ((conv & 1) ? &map_q1 : &map_q0) : ((conv & 1) ? &map_act1 : &map_act0));
            const CUtensorMap* map_wt = L == 0 ? &map_stem_w : (q8 ? &map_wq : &map_w);
            const int w_base = L == 0 ? 0 : conv * 9;
            const uint32_t tx = L == 0 ? 4 * STEM_TILE_BYTES : 2 * STAGE_BYTES;
            const int row = pos / 5, col = pos - 5 * row;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int rr = row + tap / 3 - 1, cc = col + tap % 3 - 1;
                if (rr < 0 || rr > 5 || cc < 0 || cc > 4) continue;       // the tap falls on zero padding
                const int src = rr * 5 + cc;
                for (int kc = 0; kc < chunks; ++kc) {
                    const long long t1 = stats ? clock64() : 0;
                    mbar_wait(&empty[stage], phase ^ 1u);
                    if (stats) p_slot += clock64() - t1;
                    if (elect_one_sync()) {
                        if (leader) mbar_expect_tx(&full[stage], tx);
                        else mbar_arrive_remote(&full[stage], 0);
                        uint8_t* st = smem + stage * STAGE_BYTES;
                        tma_load_3d_2sm(st, map_in, &full[stage], kc * kchunk, tile * BLOCK_M, src);
                        tma_load_3d_2sm(st + A_BYTES, map_wt, &full[stage], kc * kchunk, (int)rank * (C / 2), w_base + tap);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
        if (stats && lane == 0) {
            unsigned long long* st = P.stats + (size_t)blockIdx.x * 12;
            st[3] = (unsigned long long)(clock64() - p_start); st[4] = (unsigned long long)p_deps; st[5] = (unsigned long long)p_slot;
        }
    } else if (warp == 1 && leader) {
        // ---------------------------------------------------------------- MMA issuer (leader CTA): warp-wide loop, one elected lane issues
        uint32_t stage = 0, phase = 0;
        const bool stats = P.stats != nullptr;
        const long long m_start = stats ? clock64() : 0;
        long long m_full = 0, m_acc = 0, m_first = 0, m_items = 0, m_issue8 = 0, m_issue16 = 0, m_stages8 = 0, m_stages16 = 0;
        const uint64_t desc0 = umma_desc(smem_u32(smem)), desc0_sw32 = umma_desc_sw32(smem_u32(smem));
        for (uint32_t k = 0;; ++k) {
            const uint32_t item = ring_get(s_ring, k);
            if (item == ITEM_END) break;
            const int pos = item & 0xff, L = (int)(item >> 24);
            const bool stem = L == 0;                     // K = 16 per tap: the 14 one-hot channels
            const bool q8 = FP8 && !stem && L <= P.fp8_levels;
            const int row = pos / 5, col = pos - 5 * row;
            const int n_taps = ((row == 0 || row == 5) ? 2 : 3) * ((col == 0 || col == 4) ? 2 : 3);
            const int n_stages = n_taps * (stem ? 1 : (q8 ? C / 128 : C / BLOCK_K));     // all this role has to know of the item
            const uint32_t acc = k & 1;
            const long long t0 = stats ? clock64() : 0;
            mbar_wait(&acc_empty[acc], ((k >> 1) & 1) ^ 1);
            if (stats) { m_acc += clock64() - t0; ++m_items; }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * C;
            uint32_t accumulate = 0;
            for (int st = 0; st < n_stages; ++st) {
                const long long t1 = stats ? clock64() : 0;
                mbar_wait(&full[stage], phase);
                const long long t2 = stats ? clock64() : 0;
                tc_fence_after();
                // descriptors of this stage: start address field + stage * STAGE_BYTES / 16 (A), + A_BYTES / 16 more (B)
                const uint64_t da = (stem ? desc0_sw32 : desc0) + (uint64_t)(stage * (STAGE_BYTES >> 4));
                const uint64_t db = da + (uint64_t)(A_BYTES >> 4);
                if (elect_one_sync()) {
                    if (stem) {
                        umma_bf16_2sm(d_tmem, da, db, IDESC, accumulate);
                    } else if (q8) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) umma_fp8_2sm(d_tmem, da + 2 * kk, db + 2 * kk, IDESC_FP8, kk ? 1u : accumulate);
                    } else {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) umma_bf16_2sm(d_tmem, da + 2 * kk, db + 2 * kk, IDESC, kk ? 1u : accumulate);
                    }
                    umma_commit_2sm(&empty[stage]);
                }
                if (stats) {
                    const long long t3 = clock64();
                    m_full += t2 - t1;
                    if (!accumulate) m_first += t2 - t1;      // the item's first stage: dependency stalls and ring refills show up here
                    if (q8) { m_issue8 += t3 - t2; ++m_stages8; } else if (!stem) { m_issue16 += t3 - t2; ++m_stages16; }
                }
                accumulate = 1;
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            if (elect_one_sync()) umma_commit_2sm(&acc_full[acc]);
        }
        if (stats && lane == 0) {
            unsigned long long* st = P.stats + (size_t)blockIdx.x * 12;
            st[0] = (unsigned long long)(clock64() - m_start); st[1] = (unsigned long long)m_full; st[2] = (unsigned long long)m_acc;
            st[6] = (unsigned long long)m_first; st[7] = (unsigned long long)m_items;
            st[8] = (unsigned long long)m_issue8; st[9] = (unsigned long long)m_stages8; st[10] = (unsigned long long)m_issue16; st[11] = (unsigned long long)m_stages16;
        }
    } else if (warp == 3) {
        // ---------------------------------------------------------------- scheduler (leader) + dependency watcher
        // The leader's lane 0 claims the pair's next item -- one atomicAdd on the launch's list -- as soon as the producer
        // has begun the item before it, and publishes it to both CTAs.  The producer never runs more than a few items
        // ahead of the epilogue (ring stages, two accumulators), so a ring slot is long read by everyone when it is reused.
        // Then item (L, p, tp) reads the previous level's output at the valid taps of p (rows of this CTA's tile): lanes
        // 0-8 each poll one of those flags, so a check costs one L2 round trip and runs ahead of the TMA producer.
        for (uint32_t k = 0;; ++k) {
            uint32_t item = 0;
            if (leader) {
                if (lane == 0) {
                    while (k > *reinterpret_cast<volatile uint32_t*>(&s_prod_at) + 1u) __nanosleep(SPIN_NS);
                    item = tower_item(P, live_pairs, atomicAdd(&P.claim[0], 1u));
                    ring_put(s_ring, k, item, 1u);
                }
                item = __shfl_sync(0xffffffffu, item, 0);
            } else {
                if (lane == 0) item = ring_get(s_ring, k);
                item = __shfl_sync(0xffffffffu, item, 0);
            }
            if (item == ITEM_END) break;
            const int pos = item & 0xff, tp = (item >> 8) & 0xffff, L = item >> 24;
            if (L > P.first_level && P.flags && lane < 9) {
                int src;
                if (tap_valid(pos, lane, src)) {
                    const uint32_t* fl = P.flags + (((size_t)(L - 1) * P.n_pairs + tp) * NPOS + src) * 2 + rank;
                    const long long t0 = clock64();
                    while (ld_acquire_gpu(fl) != P.epoch) {
                        __nanosleep(32);
                        // cannot happen (an item's inputs were claimed before it by pairs that are running): fail the launch
                        // instead of hanging the GPU if it ever does
                        if ((unsigned long long)(clock64() - t0) > WATCHDOG_CYCLES) __trap();
                    }
                }
            }
            __syncwarp();
            if (lane == 0) st_release_cta_shared(&s_deps_ok, k + 1u);
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue (TMEM -> HBM) + publish
        const int q = warp & 3;
        constexpr int SETS = tower_epilogue_sets(FP8), RD = SETS == 2 ? 2 : 4;
        const int set = SETS == 2 ? (warp - 4) >> 2 : 0; // this warp's epilogue set: it takes the items of accumulator buffer `set` (e4m3 form)
        const int bar_done = 1 + set, bar_scale = 3 + set;        // named barriers of the set's 128 threads
        float* s_scale = s_scale_all + (FP8 ? set * C : 0);
        int scale_level = -1;
        for (uint32_t k = 0;; ++k) {
            uint32_t item = 0;
            if (lane == 0) item = ring_get(s_ring, k);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item == ITEM_END) break;
            if (SETS == 2 && (int)(k & 1) != set) continue;       // the other set's item
            const int pos = item & 0xff, tp = (item >> 8) & 0xffff, L = item >> 24;
            const int tile = 2 * tp + (int)rank;
            const uint32_t acc = k & 1, acc_phase = (k >> 1) & 1;
            const bool odd = L >= 2 && (L & 1) == 0;           // second conv of a residual block
            const bool last = L == NLAYERS && P.fuse_heads;    // its output only feeds the three 1x1 head convolutions
            // bf16 form: x lives in act0 (levels 0, 2, 4, ...), h in act1.  e4m3 form: act0 is the bf16 residual stream x, the
            // operands are the e4m3 copies actq0 (x) and actq1 (h)
            // (only where the next level multiplies e4m3; x is always kept in bf16 as well, h only where the next level is bf16)
            const bool is_x = odd || L == 0;
            const bool write_q = FP8 && L + 1 <= P.fp8_levels;
            const bool write_bf16 = is_x || !write_q;
            __nv_bfloat16* out = is_x ? P.act0 : P.act1;
            uint8_t* outq = is_x ? P.actq0 : P.actq1;
            const size_t row_off = ((size_t)pos * P.bpad + (size_t)tile * BLOCK_M + q * 32 + lane) * C;
#if TOWER_L2_MODE
            const uint64_t l2pol = pos < l2_positions ? l2pol_last : l2pol_normal;
#endif
            if (last) {
                // the tower's output row never leaves the SM: bias + residual + ReLU in fp32, then the three 1x1 head
                // filters as dot products over the row this thread holds
                if (P.flags) while (ld_acquire_cta_shared(&s_deps_ok) <= k) __nanosleep(SPIN_NS);
                float h[3];
                epilogue_heads<2 * (int)FP8 + (int)CALIB>(tmem_base + ((uint32_t)(q * 32) << 16) + acc * C, P.bias_g + L * C,
                                                          (FP8 && L <= P.fp8_levels) ? P.scale_g + L * C : nullptr, P.head_w,
                               out + row_off, &acc_full[acc], acc_phase, P.wait_hint, h);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (leader) mbar_arrive(&acc_empty[acc]);
                    else mbar_arrive_remote(&acc_empty[acc], 0);
                }
                float* hin = P.head_in + ((size_t)tile * BLOCK_M + q * 32 + lane) * HEAD_IN;
                hin[pos] = h[0]; hin[NPOS + pos] = h[1]; hin[2 * NPOS + 1 + pos] = h[2];
                asm volatile("bar.sync %0, 128;" ::"r"(bar_done) : "memory");
                if (q == 0 && lane == 0 && P.flags)
                    st_release_gpu(P.flags + (((size_t)L * P.n_pairs + tp) * NPOS + pos) * 2 + rank, P.epoch);
                continue;
            }
            // residual rows (this block's input, written two layers back) are published once the item's
            // dependencies are: fetch them while the MMAs still run
            uint4 res[RD][4];
            if (odd) {
                if (P.flags) while (ld_acquire_cta_shared(&s_deps_ok) <= k) __nanosleep(SPIN_NS);
#pragma unroll
                for (int c = 0; c < RD; ++c) {
                    ld_cg_v8(out + row_off + c * 32, res[c][0], res[c][1]);
                    ld_cg_v8(out + row_off + c * 32 + 16, res[c][2], res[c][3]);
                }
            }
            if (P.wait_hint) mbar_wait_hint(&acc_full[acc], acc_phase, P.wait_hint); else mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
#ifdef TOWER_NO_EPILOGUE        // timing experiment only (wrong results): hand the accumulator back unread
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (leader) mbar_arrive(&acc_empty[acc]); else mbar_arrive_remote(&acc_empty[acc], 0); }
            asm volatile("bar.sync %0, 128;" ::"r"(bar_done) : "memory");
            if (q == 0 && lane == 0 && P.flags) st_release_gpu(P.flags + (((size_t)L * P.n_pairs + tp) * NPOS + pos) * 2 + rank, P.epoch);
            continue;
#endif
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C;
            const int bias0 = L * C;
            const float inv_a = FP8 ? P.inv_a[L] : 0.f;
            if (FP8 && L != scale_level) {
                // e4m3 form: this level's 256 dequantisation factors go through shared memory (64 broadcast LDS.128 per row instead of
                // 256 L1 loads: the epilogue, with half the main loop to hide under, bounds the e4m3 levels).  The four warps walk the
                // same items; the previous item's reads are behind its closing barrier.
                const int t = ((int)threadIdx.x - 128) & 127;
                reinterpret_cast<float2*>(s_scale)[t] = __ldg(reinterpret_cast<const float2*>(P.scale_g + bias0) + t);
                asm volatile("bar.sync %0, 128;" ::"r"(bar_scale) : "memory");
                scale_level = L;
            }
            float seen_max = 0.f;
#pragma unroll
            for (int c = 0; c < C / 32; ++c) {
                uint32_t v[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr + c * 32)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c == C / 32 - 1) {
                    // all TMEM reads of this accumulator are done: hand it back before the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (leader) mbar_arrive(&acc_empty[acc]);
                        else mbar_arrive_remote(&acc_empty[acc], 0);
                    }
                }
                uint4 outv[4];
                uint32_t q8[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t packed[4];
                    float xq[8];
                    float sc[8];
                    if (FP8) {
                        const float4 sa = *reinterpret_cast<const float4*>(&s_scale[c * 32 + j * 8]), sb = *reinterpret_cast<const float4*>(&s_scale[c * 32 + j * 8 + 4]);
                        sc[0] = sa.x; sc[1] = sa.y; sc[2] = sa.z; sc[3] = sa.w; sc[4] = sb.x; sc[5] = sb.y; sc[6] = sb.z; sc[7] = sb.w;
                    }
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const int e = j * 8 + h * 2;
                        float x0 = __uint_as_float(v[e]), x1 = __uint_as_float(v[e + 1]);
                        if (FP8) {      // dequantise the accumulator and add the bias in one rounding (factor 1 on the bf16 levels: the plain sum)
                            x0 = fmaf(x0, sc[2 * h], P.bias[bias0 + c * 32 + e]);
                            x1 = fmaf(x1, sc[2 * h + 1], P.bias[bias0 + c * 32 + e + 1]);
                        } else {
                            x0 += P.bias[bias0 + c * 32 + e];
                            x1 += P.bias[bias0 + c * 32 + e + 1];
                        }
                        if (odd) {
                            const uint32_t r = (&res[c & (RD - 1)][j].x)[h];
                            x0 += __uint_as_float(r << 16);
                            x1 += __uint_as_float(r & 0xffff0000u);
                        }
                        x0 = fmaxf(x0, 0.f);
                        x1 = fmaxf(x1, 0.f);
                        if (CALIB) seen_max = fmaxf(seen_max, fmaxf(x0, x1));
                        __nv_bfloat162 b2 = __floats2bfloat162_rn(x0, x1);
                        packed[h] = *reinterpret_cast<uint32_t*>(&b2);
                        xq[2 * h] = x0 * inv_a; xq[2 * h + 1] = x1 * inv_a;
                    }
                    outv[j] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    if (FP8 && write_q) { q8[2 * j] = pack_e4m3x4(xq[0], xq[1], xq[2], xq[3]); q8[2 * j + 1] = pack_e4m3x4(xq[4], xq[5], xq[6], xq[7]); }
                }
                if (odd && c + RD < C / 32) {
                    ld_cg_v8(out + row_off + (c + RD) * 32, res[c & (RD - 1)][0], res[c & (RD - 1)][1]);
                    ld_cg_v8(out + row_off + (c + RD) * 32 + 16, res[c & (RD - 1)][2], res[c & (RD - 1)][3]);
                }
                if (write_bf16) {
                    ST_ACT(out + row_off + c * 32, outv[0], outv[1]);
                    ST_ACT(out + row_off + c * 32 + 16, outv[2], outv[3]);
                }
                if (FP8 && write_q) ST_ACT(outq + row_off + c * 32, make_uint4(q8[0], q8[1], q8[2], q8[3]), make_uint4(q8[4], q8[5], q8[6], q8[7]));
            }
            if (CALIB) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) seen_max = fmaxf(seen_max, __shfl_xor_sync(0xffffffffu, seen_max, o));
                if (lane == 0) atomicMax(&P.level_absmax[L], __float_as_uint(seen_max));
            }
            // publish: the barrier orders all 128 threads' stores before the (cumulative) gpu-scope release
            asm volatile("bar.sync %0, 128;" ::"r"(bar_done) : "memory");
            if (q == 0 && lane == 0 && P.flags)
                st_release_gpu(P.flags + (((size_t)L * P.n_pairs + tp) * NPOS + pos) * 2 + rank, P.epoch);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
    // the last pair to leave rewinds the list for the next launch on this stream
    if (leader && threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&P.claim[1], 1u) == (gridDim.x >> 1) - 1u) {
            P.claim[0] = 0u;
            P.claim[1] = 0u;
            __threadfence();
        }
    }
}

// ---------------------------------------------------------------------------------- weight preparation
__global__ void prep_tower_kernel(const float* __restrict__ flat, __nv_bfloat16* __restrict__ w, float* __restrict__ bias) {
    // grid: (256 cout, 18 layers), block 256 (cin)
    const int n = blockIdx.x, L = blockIdx.y, k = threadIdx.x;
    const float* base = flat + OFF_TOWER + (size_t)L * TOWER_STRIDE;
    const float *cw = base, *cb = base + 589824, *gamma = cb + 256, *beta = gamma + 256, *mean = beta + 256, *var = mean + 256;
    const float scale = gamma[n] / sqrtf(var[n] + BN_EPS);
    for (int t = 0; t < 9; ++t)
        w[(((size_t)L * 9 + t) * C + n) * C + k] = __float2bfloat16(cw[((size_t)n * C + k) * 9 + t] * scale);
    if (k == 0) bias[L * C + n] = (cb[n] - mean[n]) * scale + beta[n];
}

// stem level weights: Ws[tap][cout][16] bf16.  The stem's input row of a square is one-hot: channel k < 7 = "the mover's
// token is k", channel 7 + k = "the opponent's token is k" (token 0 = none there; Embedding(7,4) gives it a vector too),
// so Embedding + Conv3x3(8->256) + BatchNorm collapse into one table per tap -- the MMA then just adds two of its rows.
__global__ void prep_stem_kernel(const float* __restrict__ flat, __nv_bfloat16* __restrict__ ws, float* __restrict__ bias) {
    const int c = threadIdx.x, t = blockIdx.x;
    const float* sw = flat + OFF_STEM;
    const float *sb = sw + 18432, *gamma = sb + 256, *beta = gamma + 256, *mean = beta + 256, *var = mean + 256;
    const float* emb = flat + OFF_EMB;
    const float scale = gamma[c] / sqrtf(var[c] + BN_EPS);
    for (int k = 0; k < STEM_K; ++k) {
        float acc = 0.f;
        if (k < 14) {
            const int ch = k / 7, tok = k % 7;
            for (int e = 0; e < 4; ++e) acc += sw[((size_t)c * 8 + ch * 4 + e) * 9 + t] * emb[tok * 4 + e];
        }
        ws[((size_t)t * C + c) * STEM_K + k] = __float2bfloat16(acc * scale);
    }
    if (t == 0) bias[c] = (sb[c] - mean[c]) * scale + beta[c];
}

// ---- e4m3 tower: weights.  One block per (output channel, convolution): the folded 3x3x256 filter of the channel over its
// largest magnitude / 448 (per-output-channel scale, multiplied back onto the fp32 accumulator by the epilogue).
__global__ void __launch_bounds__(256) prep_tower_fp8_kernel(const float* __restrict__ flat, uint8_t* __restrict__ wq, float* __restrict__ w_scale) {
    __shared__ float s_max[8];
    const int n = blockIdx.x, L = blockIdx.y, k = threadIdx.x;
    const float* base = flat + OFF_TOWER + (size_t)L * TOWER_STRIDE;
    const float *cw = base, *cb = base + 589824, *gamma = cb + 256, *var = gamma + 768;
    const float fold = gamma[n] / sqrtf(var[n] + BN_EPS);
    float w[9], m = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) { w[t] = cw[((size_t)n * C + k) * 9 + t] * fold; m = fmaxf(m, fabsf(w[t])); }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((k & 31) == 0) s_max[k >> 5] = m;
    __syncthreads();
    m = s_max[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, s_max[i]);
    const float sc = fmaxf(m, 1e-20f) / E4M3_MAX;
#pragma unroll
    for (int t = 0; t < 9; ++t)
        wq[(((size_t)L * 9 + t) * C + n) * C + k] = (uint8_t)__nv_cvt_float_to_fp8(w[t] / sc, __NV_SATFINITE, __NV_E4M3);
    if (k == 0) w_scale[L * C + n] = sc;
}

// ---- e4m3 tower: calibration positions.  One thread per position: a uniformly random legal playout from the start position,
// stopped after (i mod 60) plies or at the end of the game, tokenised like Network.process_observation (exp/policy.py:96-105).
__global__ void __launch_bounds__(128) calib_positions_kernel(mc_state start, mc_rules rules, int n, uint8_t* __restrict__ tokens) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        mc_state s = start;
        Philox rng(0x5EEDCA11B8A7E5ull, (uint32_t)i, 0u, 0u);
        const int plies = i % 60;
        for (int p = 0; p < plies; ++p) {
            uint16_t codes[MC_MAX_MOVES];
            int res;
            const int E = mc::generate(s, rules, codes, &res);
            if (res != MC_ONGOING || E <= 0) break;
            mc_state o;
            if (mc::step(s, codes[min((int)(rng.uniform() * E), E - 1)], rules, &o) != 0) break;
            s = o;
        }
        float clock;
        mc::tokenize(s, tokens + (size_t)i * MC_TOKENS, &clock);
    }
}

struct HeadWeights {
    float* pw;     // [2][256] folded policy conv
    float* pb;     // [2]
    float* vw;     // [256]
    float* vb;     // [1]
    float* plt;    // [61][554] plinear transposed
    float* plb;    // [554]
    float* v1t;    // [31][256] vlinear.0 transposed
    float* v1b;    // [256]
    float* v2;     // [256]
    float* v2b;    // [1]
    float* hw4;    // [256][4]: pw[0][c], pw[1][c], vw[c], 0 -- the tower's last epilogue reads these
};

__global__ void prep_heads_kernel(const float* __restrict__ flat, HeadWeights H) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const float* pc = flat + OFF_PCONV;
    const float *pcb = pc + 512, *pg = pcb + 2, *pbeta = pg + 2, *pm = pbeta + 2, *pv = pm + 2;
    for (int i = tid; i < 512; i += nthr) {
        const int o = i / 256;
        H.pw[i] = pc[i] * (pg[o] / sqrtf(pv[o] + BN_EPS));
    }
    if (tid < 2) H.pb[tid] = (pcb[tid] - pm[tid]) * (pg[tid] / sqrtf(pv[tid] + BN_EPS)) + pbeta[tid];
    const float* vc = flat + OFF_VCONV;
    const float vscale = vc[257] / sqrtf(vc[260] + BN_EPS);
    for (int i = tid; i < 256; i += nthr) H.vw[i] = vc[i] * vscale;
    if (tid == 0) H.vb[0] = (vc[256] - vc[259]) * vscale + vc[258];
    const float* pl = flat + OFF_PLIN;
    for (int i = tid; i < 554 * 61; i += nthr) { const int o = i / 61, j = i % 61; H.plt[j * 554 + o] = pl[i]; }
    for (int i = tid; i < 554; i += nthr) H.plb[i] = pl[554 * 61 + i];
    const float* v1 = flat + OFF_V1;
    for (int i = tid; i < 256 * 31; i += nthr) { const int o = i / 31, j = i % 31; H.v1t[j * 256 + o] = v1[i]; }
    for (int i = tid; i < 256; i += nthr) H.v1b[i] = v1[256 * 31 + i];
    const float* v2 = flat + OFF_V2;
    for (int i = tid; i < 256; i += nthr) H.v2[i] = v2[i];
    if (tid == 0) H.v2b[0] = v2[256];
    for (int i = tid; i < 256; i += nthr) {
        H.hw4[4 * i] = pc[i] * (pg[0] / sqrtf(pv[0] + BN_EPS));
        H.hw4[4 * i + 1] = pc[256 + i] * (pg[1] / sqrtf(pv[1] + BN_EPS));
        H.hw4[4 * i + 2] = vc[i] * vscale;
        H.hw4[4 * i + 3] = 0.f;
    }
}

// ---------------------------------------------------------------------------------- stem input
// One thread per (board, square): the square's two tokens become the 16 leading channels of its one-hot row
// (32 bytes = the whole row).  Rows past the live ones keep
// whatever an earlier batch left there: rows never mix, and nobody reads the results of dead rows.
__global__ void __launch_bounds__(256) stem_onehot_kernel(const uint8_t* __restrict__ tokens, int n, int bpad,
                                                          __nv_bfloat16* __restrict__ stem_in, const uint32_t* __restrict__ count,
                                                          uint32_t row_base) {
    if (count) {
        const uint32_t total = __ldg(count);
        n = min(n, (int)(total > row_base ? total - row_base : 0u));
    }
    const int total_sq = n * NPOS;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total_sq; idx += gridDim.x * blockDim.x) {
        const int board = idx / NPOS, pos = idx - board * NPOS;
        const int mine = tokens[(size_t)board * MC_TOKENS + pos], theirs = 7 + tokens[(size_t)board * MC_TOKENS + NPOS + pos];
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            w[k] = ((mine == 2 * k || theirs == 2 * k) ? 0x3F80u : 0u) | ((mine == 2 * k + 1 || theirs == 2 * k + 1) ? 0x3F800000u : 0u);
        uint4* dst = reinterpret_cast<uint4*>(stem_in + ((size_t)pos * bpad + board) * STEM_K);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// ---------------------------------------------------------------------------------- heads
// One warp per board.  1x1 convolutions: lane = board position (30 of 32 lanes), each lane walks its own
// 512-byte channel row.  Dense layers: lane owns outputs lane + 32k and keeps 18 (policy) / 8 (value)
// independent accumulators; the transposed weight matrices (135 KB + 31 KB) are read through the
// read-only L1 path -- they stay L1/L2 resident, which costs no staging phase and leaves room for
// several CTAs per SM.
// The three 1x1 convolutions 256 -> {policy 0, policy 1, value} were taken by the tower's last epilogue; lane p < 30
// picks up the raw sums of position p from the board's head_in row (384 bytes, coalesced).
__device__ __forceinline__ void heads_load_sums(const float* __restrict__ head_in, int row, int lane, float& d0, float& d1, float& d2) {
    const float* hin = head_in + (size_t)row * HEAD_IN;
    d0 = d1 = d2 = 0.f;
    if (lane < NPOS) { d0 = hin[lane]; d1 = hin[NPOS + lane]; d2 = hin[2 * NPOS + 1 + lane]; }
}

constexpr int HEADS_THREADS = 128;
constexpr int HEADS_WARPS = HEADS_THREADS / 32;

__global__ void __launch_bounds__(HEADS_THREADS, 4)
heads_kernel(const float* __restrict__ head_in, const float* __restrict__ clocks, int n, HeadWeights H,
             float* __restrict__ logits, float* __restrict__ values) {
    __shared__ float s_in[HEADS_WARPS][96];   // per warp: px[60], clock, vx[30], clock
    const float pb0 = __ldg(H.pb), pb1 = __ldg(H.pb + 1), vb = __ldg(H.vb), v2b = __ldg(H.v2b);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* in = s_in[warp];
    for (int board = blockIdx.x * HEADS_WARPS + warp; board < n; board += gridDim.x * HEADS_WARPS) {
        // ---- 1x1 convolutions 256 -> {2, 1}: lane p gets the sums of position p
        float d0, d1, d2;
        heads_load_sums(head_in, board, lane, d0, d1, d2);
        if (lane < NPOS) {
            in[lane] = fmaxf(d0 + pb0, 0.f);
            in[30 + lane] = fmaxf(d1 + pb1, 0.f);
            in[61 + lane] = fmaxf(d2 + vb, 0.f);
        } else if (lane == 30) {
            const float ck = clocks[board];
            in[60] = ck;
            in[91] = ck;
        }
        __syncwarp();
        // ---- policy: 554 logits = plinear([px, clock]); lane owns outputs lane + 32 k
        float acc[18];
#pragma unroll
        for (int k = 0; k < 18; ++k) acc[k] = (lane + 32 * k < MC_NUM_ACTIONS) ? __ldg(H.plb + lane + 32 * k) : 0.f;
        const bool tail = lane < MC_NUM_ACTIONS - 17 * 32;
#pragma unroll 2
        for (int j = 0; j < 61; ++j) {
            const float xj = in[j];
            const float* wr = H.plt + j * 554 + lane;
#pragma unroll
            for (int k = 0; k < 17; ++k) acc[k] += xj * __ldg(wr + 32 * k);
            if (tail) acc[17] += xj * __ldg(wr + 32 * 17);
        }
        float* lg = logits + (size_t)board * MC_NUM_ACTIONS + lane;
#pragma unroll
        for (int k = 0; k < 17; ++k) lg[32 * k] = acc[k];
        if (tail) lg[32 * 17] = acc[17];
        // ---- value: tanh(v2 . relu(v1 [vx, clock] + b1) + b2)
        float hv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hv[k] = __ldg(H.v1b + lane + 32 * k);
#pragma unroll 8
        for (int j = 0; j < 31; ++j) {
            const float xj = in[61 + j];
            const float* wr = H.v1t + j * 256 + lane;
#pragma unroll
            for (int k = 0; k < 8; ++k) hv[k] += xj * __ldg(wr + 32 * k);
        }
        float part = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) part += fmaxf(hv[k], 0.f) * __ldg(H.v2 + lane + 32 * k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) values[board] = tanhf(part + v2b);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------- heads, search form
// Inside az_search only the logits of the leaf's legal moves are ever read (exp/agent.py:68 takes
// p[0][legal_moves].softmax(0)), ~9 of 554.  This variant computes just those, applies the legal-move
// softmax and writes the priors straight into the new node's edges; the value goes to values[g].
// One warp per game slot.
template <bool LOOKAHEAD>
__global__ void __launch_bounds__(HEADS_THREADS, LOOKAHEAD ? 2 : 4)
heads_legal_kernel(const float* __restrict__ head_in, HeadWeights H, az::View V, float* __restrict__ values,
                   int row_base, int chunk_rows) {
    // rows of the batch: dense (row -> slot through row_slot, az_search) or one row per slot with a needs_eval mask;
    // this launch covers rows [row_base, row_base + chunk_rows) of it, which sit in act rows [0, chunk_rows)
    const int total = V.compact ? min((int)__ldg((V.defer_thr > 0 ? V.row_eff : V.row_count) + V.parity), V.row_cap) : V.G * V.K;
    const int n_rows = max(0, min(chunk_rows, total - row_base));
    __shared__ float s_in[HEADS_WARPS][96];
    __shared__ uint16_t s_codes[LOOKAHEAD ? HEADS_WARPS : 1][az::CACHE_MAX_E];
    const float pb0 = __ldg(H.pb), pb1 = __ldg(H.pb + 1), vb = __ldg(H.vb), v2b = __ldg(H.v2b);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* in = s_in[warp];
    for (int r = blockIdx.x * HEADS_WARPS + warp; r < n_rows; r += gridDim.x * HEADS_WARPS) {
        const int slot = V.compact ? V.row_slot[row_base + r] : row_base + r;
        // slot < 0: a look-ahead row -- a position no tree holds yet; its priors and value only go to the cache
        const bool lookahead = LOOKAHEAD && slot < 0;
        if (!V.compact && !V.needs_eval[slot]) continue;
        float d0, d1, d2;
        heads_load_sums(head_in, r, lane, d0, d1, d2);
        if (lane < NPOS) {
            in[lane] = fmaxf(d0 + pb0, 0.f);
            in[30 + lane] = fmaxf(d1 + pb1, 0.f);
            in[61 + lane] = fmaxf(d2 + vb, 0.f);
        } else if (lane == 30) {
            const float ck = V.clocks[row_base + r];
            in[60] = ck;
            in[91] = ck;
        }
        __syncwarp();
        // ---- legal logits -> softmax -> edge_P of the leaf
        int E;
        size_t e0 = 0;
        const uint16_t* codes;
        if (lookahead) {
#if defined(__CUDA_ARCH__)
            // the row was queued by its position alone: generate its legal moves here (lane = square), a warp per row
            const az::WarpGen w = az::warp_generate(V, V.row_state[row_base + r], lane);
            E = w.E;
            if (w.res != MC_ONGOING || E <= 0 || E > az::CACHE_MAX_E) { __syncwarp(); continue; }   // finished there: never evaluated
            uint16_t* sc = s_codes[LOOKAHEAD ? warp : 0];
            az::warp_emit_codes(V, w, [&](int k, uint16_t c) { sc[k] = c; });
            __syncwarp();
            codes = sc;
#else
            E = 0; codes = nullptr;
#endif
        } else {
            const int g = slot / V.K;
            const int t = 2 * g + (V.game_ply[g] & 1);
            const uint32_t node = V.leaf_node[slot];
            const az::NodeHead h = az::load_head(&V.nodes[(size_t)t * V.NC + node]);
            E = (int)(h.info & 0xffffu);
            e0 = (size_t)t * V.EC + h.edge_off;
            codes = nullptr;                       // read from the leaf's edge records below
        }
        float lg[3] = {-INFINITY, -INFINITY, -INFINITY};
        float m = -INFINITY;
        for (int i = lane, kk = 0; i < E && kk < 3; i += 32, ++kk) {
            const int code = lookahead ? codes[i] : V.edges[e0 + i].link.code;
            float acc = __ldg(H.plb + code);
#pragma unroll
            for (int j = 0; j < 61; ++j) acc += in[j] * __ldg(H.plt + j * 554 + code);   // 61 independent gathers in flight
            lg[kk] = acc;
            m = fmaxf(m, acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float sum = 0.f;
        for (int i = lane, kk = 0; i < E && kk < 3; i += 32, ++kk) sum += expf(lg[kk] - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        for (int i = lane, kk = 0; i < E && kk < 3; i += 32, ++kk) {
            lg[kk] = expf(lg[kk] - m) / sum;
            if (!lookahead) V.edges[e0 + i].stat.P = lg[kk];
        }
        // ---- value
        float hv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hv[k] = __ldg(H.v1b + lane + 32 * k);
#pragma unroll 8
        for (int j = 0; j < 31; ++j) {
            const float xj = in[61 + j];
            const float* wr = H.v1t + j * 256 + lane;
#pragma unroll
            for (int k = 0; k < 8; ++k) hv[k] += xj * __ldg(wr + 32 * k);
        }
        float part = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) part += fmaxf(hv[k], 0.f) * __ldg(H.v2 + lane + 32 * k);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        const float value = tanhf(part + v2b);
        if (lane == 0 && !lookahead) values[slot] = value;
        // ---- remember the evaluation (exact cache: key = everything the network saw).  Tree kernels of the other
        // stream may be reading the entry: seqlock (odd seq = being written)
        if (V.cache && E <= az::CACHE_MAX_E) {
            const mc_state s = lookahead ? V.row_state[row_base + r] : V.leaf_states[slot];
            az::CacheEntry* c = V.cache + (az::cache_hash(s) & V.cache_mask);
            // a row whose key is already stored was evaluated twice in this batch (or since its lookup): count them
            if (lane == 0 && c->epoch == V.cache_epoch && c->pl0 == s.pl0 && c->pl1 == s.pl1 && c->pl2 == s.pl2 && c->white == s.white &&
                c->meta_n == ((s.meta & az::CACHE_KEY_META) | ((uint32_t)E << 8)))
                atomicAdd(&V.counters[az::C_DUP_ROWS], 1ull);
            uint32_t old = 1;
            if (lane == 0) old = atomicOr(&c->seq, 1u);
            old = __shfl_sync(0xffffffffu, old, 0);
            if (!(old & 1u)) {                   // else another warp is filling this entry right now: skip
                __threadfence();
                for (int i = lane, kk = 0; i < E && kk < 3; i += 32, ++kk) c->priors[i] = lg[kk];
                if (lane == 0) {
                    c->pl0 = s.pl0; c->pl1 = s.pl1; c->pl2 = s.pl2; c->white = s.white;
                    c->meta_n = (s.meta & az::CACHE_KEY_META) | ((uint32_t)E << 8);
                    c->epoch = V.cache_epoch; c->value = value;
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicExch(&c->seq, old + 2u);
            }
        }
        __syncwarp();
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map_3d(CUtensorMap* map, void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, int elem_bytes = 2) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
            return fail(MCAZ_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * (cuuint64_t)elem_bytes, d0 * d1 * (cuuint64_t)elem_bytes};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(map, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MCAZ_ECUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return MCAZ_OK;
}

}  // namespace

struct Network {
    int capacity = 0;   // boards the activation buffers hold (multiple of 128)
    size_t l2_persist_bytes = 0;       // persisting set-aside of the L2 granted to this context (activation stores, TOWER_L2_MODE)
    __nv_bfloat16 *act[2] = {nullptr, nullptr};
    __nv_bfloat16* w = nullptr;        // [18][9][256][256]
    float* bias = nullptr;             // [18][256]
    __nv_bfloat16* stem_w = nullptr;       // [9 taps][256 cout][16] bf16: folded embedding x stem conv x BN table (14 live channels)
    __nv_bfloat16* stem_in = nullptr;      // [30 pos][capacity][16] bf16 one-hot token rows
    float* head_in = nullptr;              // [capacity][HEAD_IN]
    float* head_pool = nullptr;
    HeadWeights heads{};
    CUtensorMap map_act[2], map_w, map_stem_in, map_stem_w;
    // e4m3 tower (az_config.network = 2)
    bool fp8 = false;
    int fp8_levels = 0;                // levels 1 .. fp8_levels on e4m3 operands (az_config.fp8_convolutions)
    bool calibrating = false;          // forward_chunk runs the bf16 form with the per-level maxima on (calibrate_fp8)
    uint8_t* wq = nullptr;             // [18][9][256][256] e4m3: folded weights over their output channel's scale
    float* w_scale = nullptr;          // [18][256] that scale (largest |weight| of the channel / 448)
    float* scale = nullptr;            // [19][256] dequantisation factors the kernel reads (w_scale x activation scale of the level's input)
    uint8_t* actq[2] = {nullptr, nullptr};   // [30][capacity][256] e4m3 x / h
    unsigned int* level_absmax = nullptr;    // [19] calibration maxima
    uint8_t* calib_tokens = nullptr;   // [CALIB_ROWS][60]
    float act_scale[NLEVELS] = {};     // activation scale of each level's output (calibrated at az_set_weights)
    CUtensorMap map_q[2], map_wq;
    bool have_weights = false;
    uint32_t* tower_flags = nullptr;   // [19][n_pairs][30][2] epoch stamps of the published items
    uint32_t* tower_claim = nullptr;   // [2] next item of the running launch's list, pairs that have left it
    int tower_pairs = -1;
    TowerParams tower_params;          // the kernel's parameter block; its `bias` is the host copy of the folded biases
    uint8_t pos_order[32];
    uint32_t epoch = 0;
    bool per_layer = false;            // false: tower_tc_kernel (one data-flow ordered launch for all levels, default);
                                       // true: the same kernel launched once per level (MCAZ_TOWER=layers)
    // profiling (az_profile_network)
    unsigned long long* stats = nullptr;   // MCAZ_TOWER_STATS=1: [grid][12] wait-cycle counters of the last tower launch
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    size_t events_used = 0;
};

static int net_alloc_acts(az_engine* e, int boards) {
    Network* N = e->net;
    int cap = ((boards + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * (2 * BLOCK_M);   // whole tile pairs
    if (cap <= N->capacity) return MCAZ_OK;
    for (int i = 0; i < 2; ++i) {
        if (N->act[i]) cudaFree(N->act[i]);
        N->act[i] = nullptr;
        MCAZ_CUDA(cudaMalloc(&N->act[i], (size_t)NPOS * cap * C * sizeof(__nv_bfloat16)));
        MCAZ_CUDA(cudaMemset(N->act[i], 0, (size_t)NPOS * cap * C * sizeof(__nv_bfloat16)));
        if (int rc = make_map_3d(&N->map_act[i], N->act[i], C, cap, NPOS, BLOCK_K, BLOCK_M)) return rc;
    }
    if (N->stem_in) cudaFree(N->stem_in);
    if (N->head_in) cudaFree(N->head_in);
    N->stem_in = nullptr; N->head_in = nullptr;
    MCAZ_CUDA(cudaMalloc(&N->stem_in, (size_t)NPOS * cap * STEM_K * sizeof(__nv_bfloat16)));
    MCAZ_CUDA(cudaMemset(N->stem_in, 0, (size_t)NPOS * cap * STEM_K * sizeof(__nv_bfloat16)));
    MCAZ_CUDA(cudaMalloc(&N->head_in, (size_t)cap * HEAD_IN * sizeof(float)));
    MCAZ_CUDA(cudaMemset(N->head_in, 0, (size_t)cap * HEAD_IN * sizeof(float)));
    if (int rc = make_map_3d(&N->map_stem_in, N->stem_in, STEM_K, cap, NPOS, STEM_K, BLOCK_M, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
    if (N->fp8)
        for (int i = 0; i < 2; ++i) {
            if (N->actq[i]) cudaFree(N->actq[i]);
            N->actq[i] = nullptr;
            MCAZ_CUDA(cudaMalloc(&N->actq[i], (size_t)NPOS * cap * C));
            MCAZ_CUDA(cudaMemset(N->actq[i], 0, (size_t)NPOS * cap * C));
            if (int rc = make_map_3d(&N->map_q[i], N->actq[i], C, cap, NPOS, 128, BLOCK_M, CU_TENSOR_MAP_SWIZZLE_128B, 1)) return rc;
        }
    N->capacity = cap;
    return MCAZ_OK;
}

int network_create(az_engine* e) {
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return fail(MCAZ_ENODEV, "the built-in network needs an sm_100 GPU (tcgen05/TMEM/TMA)");
    Network* N = new Network();
    e->net = N;
#if TOWER_L2_MODE
    {   // evict_last lines live in the persisting set-aside of the L2 (a limit of the CUDA context: all the device allows, 79 of
        // 126 MB on a B200; other kernels of the process keep the rest and whatever part of the set-aside is not in use)
        int max_persist = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        size_t want = (size_t)max_persist;
#ifdef MCAZ_TIMING_EXPERIMENTS
        if (const char* pm = getenv("MCAZ_L2_PERSIST_MB")) want = std::min(want, (size_t)atoi(pm) << 20);
#endif
        size_t have = 0;
        if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize);
        cudaGetLastError();                 // a device without the feature (MIG slice): no set-aside, every store keeps the normal policy
        N->l2_persist_bytes = have;
    }
#endif
    MCAZ_CUDA(cudaMalloc(&N->w, (size_t)NLAYERS * 9 * C * C * sizeof(__nv_bfloat16)));
    MCAZ_CUDA(cudaMalloc(&N->bias, (size_t)NLEVELS * C * sizeof(float)));
    MCAZ_CUDA(cudaMalloc(&N->stem_w, (size_t)9 * C * STEM_K * sizeof(__nv_bfloat16)));
    const size_t head_floats = 512 + 4 + 256 + 4 + 61 * 554 + 554 + 31 * 256 + 256 + 256 + 4 + 4 * 256 + 16;
    MCAZ_CUDA(cudaMalloc(&N->head_pool, head_floats * sizeof(float)));
    float* p = N->head_pool;
    N->heads.pw = p; p += 512;
    N->heads.pb = p; p += 4;
    N->heads.vw = p; p += 256;
    N->heads.vb = p; p += 4;
    N->heads.plt = p; p += 61 * 554;
    N->heads.plb = p; p += 554;
    N->heads.v1t = p; p += 31 * 256;
    N->heads.v1b = p; p += 256;
    N->heads.v2 = p; p += 256;
    N->heads.v2b = p; p += 4;
    p += (4 - ((p - N->head_pool) & 3)) & 3;                 // float4 reads
    N->heads.hw4 = p;
    if (int rc = make_map_3d(&N->map_w, N->w, C, C, (uint64_t)NLAYERS * 9, BLOCK_K, C / 2)) return rc;
    if (int rc = make_map_3d(&N->map_stem_w, N->stem_w, STEM_K, C, 9, STEM_K, C / 2, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
    MCAZ_CUDA(cudaFuncSetAttribute(tower_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOWER_SMEM));
    N->fp8 = e->cfg.network == 2;
    if (N->fp8) {
        const int want = e->cfg.fp8_convolutions;
        if (want < 0 || want > NLAYERS || (want & 1)) return fail(MCAZ_EINVAL, "az_create: fp8_convolutions must be an even number in [0, 18]");
        N->fp8_levels = want == 0 ? 12 : want;
        MCAZ_CUDA(cudaFuncSetAttribute(tower_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOWER_SMEM));
        MCAZ_CUDA(cudaFuncSetAttribute(tower_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOWER_SMEM_FP8));
        MCAZ_CUDA(cudaMalloc(&N->wq, (size_t)NLAYERS * 9 * C * C));
        MCAZ_CUDA(cudaMalloc(&N->w_scale, (size_t)NLAYERS * C * sizeof(float)));
        MCAZ_CUDA(cudaMalloc(&N->scale, (size_t)NLEVELS * C * sizeof(float)));
        MCAZ_CUDA(cudaMalloc(&N->level_absmax, NLEVELS * sizeof(unsigned int)));
        MCAZ_CUDA(cudaMalloc(&N->calib_tokens, (size_t)CALIB_ROWS * MC_TOKENS));
        if (int rc = make_map_3d(&N->map_wq, N->wq, C, C, (uint64_t)NLAYERS * 9, 128, C / 2, CU_TENSOR_MAP_SWIZZLE_128B, 1)) return rc;
    } else {
        // the e4m3 maps are kernel arguments of every form: give the unused ones valid contents
        N->map_q[0] = N->map_q[1] = N->map_wq = N->map_w;
    }
    {
        const char* m = getenv("MCAZ_TOWER");      // "layers": one launch per level (bit-identical; for tests and per-level profiles)
        N->per_layer = m && std::strcmp(m, "layers") == 0;
        // positions with the most valid taps first: the list's long items lead each (level, tile pair) run
        int taps_of[NPOS], order[NPOS];
        for (int pos = 0; pos < NPOS; ++pos) {
            const int r = pos / 5, c = pos % 5;
            taps_of[pos] = ((r == 0 || r == 5) ? 2 : 3) * ((c == 0 || c == 4) ? 2 : 3);
            order[pos] = pos;
        }
        std::stable_sort(order, order + NPOS, [&](int a, int b) { return taps_of[a] > taps_of[b]; });
        std::memset(N->pos_order, 0, sizeof(N->pos_order));
        for (int i = 0; i < NPOS; ++i) N->pos_order[i] = (uint8_t)order[i];
        MCAZ_CUDA(cudaMalloc(&N->tower_claim, 2 * sizeof(uint32_t)));
        MCAZ_CUDA(cudaMemset(N->tower_claim, 0, 2 * sizeof(uint32_t)));
    }
    return net_alloc_acts(e, std::min(std::max(std::max(e->v.G * e->v.K, e->v.row_cap), N->fp8 ? CALIB_ROWS : 0), MAX_CHUNK_BOARDS));   // grows on demand
}

void network_destroy(az_engine* e) {
    Network* N = e->net;
    if (!N) return;
#if TOWER_L2_MODE
    if (N->l2_persist_bytes) cudaCtxResetPersistingL2Cache();      // the freed buffers' lines need not persist
#endif
    for (int i = 0; i < 2; ++i) if (N->act[i]) cudaFree(N->act[i]);
    if (N->w) cudaFree(N->w);
    if (N->bias) cudaFree(N->bias);
    if (N->stem_w) cudaFree(N->stem_w);
    if (N->stem_in) cudaFree(N->stem_in);
    if (N->head_in) cudaFree(N->head_in);
    if (N->head_pool) cudaFree(N->head_pool);
    if (N->tower_flags) cudaFree(N->tower_flags);
    if (N->tower_claim) cudaFree(N->tower_claim);
    for (int i = 0; i < 2; ++i) if (N->actq[i]) cudaFree(N->actq[i]);
    if (N->wq) cudaFree(N->wq);
    if (N->w_scale) cudaFree(N->w_scale);
    if (N->scale) cudaFree(N->scale);
    if (N->level_absmax) cudaFree(N->level_absmax);
    if (N->calib_tokens) cudaFree(N->calib_tokens);
    if (N->stats) cudaFree(N->stats);
    for (auto& ev : N->events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    delete N;
    e->net = nullptr;
}

static int calibrate_fp8(az_engine* e, const float* flat);

int network_set_weights(az_engine* e, const float* flat) {
    Network* N = e->net;
    prep_tower_kernel<<<dim3(C, NLAYERS), C, 0, e->stream>>>(flat, N->w, N->bias + C);     // bias row 0 is the stem's
    MCAZ_CHECK_LAUNCH();
    prep_stem_kernel<<<9, C, 0, e->stream>>>(flat, N->stem_w, N->bias);
    MCAZ_CHECK_LAUNCH();
    prep_heads_kernel<<<64, 256, 0, e->stream>>>(flat, N->heads);
    MCAZ_CHECK_LAUNCH();
    e->launches += 3;
    MCAZ_CUDA(cudaMemcpyAsync(N->tower_params.bias, N->bias, sizeof(N->tower_params.bias), cudaMemcpyDeviceToHost, e->stream));
    MCAZ_CUDA(cudaStreamSynchronize(e->stream));       // the tower takes the biases in its parameter block
    N->have_weights = true;
    if (N->fp8) return calibrate_fp8(e, flat);
    return MCAZ_OK;
}

// MCAZ_DEBUG_TOWER (timing experiments, results are then wrong): bit 0 = the last level stores its activations like
// any other instead of taking the head convolutions, bit 1 = no stem level in the list.
// Only in builds made with -DMCAZ_TIMING_EXPERIMENTS (never the shipped library): there the environment also sets the L2
// group size and the epilogue's wait hint.
#ifdef MCAZ_TIMING_EXPERIMENTS
static int tower_debug() {
    static int v = -1;
    if (v < 0) {
        const char* s = getenv("MCAZ_DEBUG_TOWER");
        v = s ? atoi(s) : 0;
        if (v) fprintf(stderr, "libmcaz: MCAZ_DEBUG_TOWER=%d -- timing experiment, NETWORK OUTPUTS ARE WRONG\n", v);
    }
    return v;
}
static int env_or(const char* name, int dflt) { const char* s = getenv(name); return (s && atoi(s) > 0) ? atoi(s) : dflt; }
#else
static constexpr int tower_debug() { return 0; }
static constexpr int env_or(const char*, int dflt) { return dflt; }
#endif

// Epoch flags of the published items, one per (level, tile pair, position, CTA of the pair).
static int tower_flags_for(az_engine* e, int n_pairs) {
    Network* N = e->net;
    if (N->tower_pairs == n_pairs) return MCAZ_OK;
    if (N->tower_flags) cudaFree(N->tower_flags);
    N->tower_flags = nullptr;
    const size_t n_flags = (size_t)NLEVELS * n_pairs * NPOS * 2;
    MCAZ_CUDA(cudaMalloc(&N->tower_flags, n_flags * sizeof(uint32_t)));
    MCAZ_CUDA(cudaMemsetAsync(N->tower_flags, 0, n_flags * sizeof(uint32_t), e->stream));
    N->tower_pairs = n_pairs;
    N->epoch = 0;
    return MCAZ_OK;
}

static int forward_chunk(az_engine* e, const uint8_t* tokens, const float* clocks, int n, float* logits, float* values,
                         const az::View* search_view, int row_base, int sched_rows = 0);

int network_forward(az_engine* e, const uint8_t* tokens, const float* clocks, const uint8_t* /*active*/, int n, float* logits,
                    float* values) {
    Network* N = e->net;
    if (!N->have_weights) return fail(MCAZ_ESTATE, "network weights have not been set (az_set_weights)");
    for (int off = 0; off < n; off += MAX_CHUNK_BOARDS) {
        const int m = std::min(MAX_CHUNK_BOARDS, n - off);
        if (int rc = forward_chunk(e, tokens + (size_t)off * MC_TOKENS, clocks + off, m, logits + (size_t)off * MC_NUM_ACTIONS, values + off,
                                   nullptr, 0))
            return rc;
    }
    return MCAZ_OK;
}

static int forward_chunk(az_engine* e, const uint8_t* tokens, const float* clocks, int n, float* logits, float* values,
                         const az::View* search_view, int row_base, int sched_rows) {
    Network* N = e->net;
    // sched_rows >= n: size the schedule (and buffers) for that many rows, so equal chunks of one batch share a schedule
    const int plan = std::max(n, sched_rows);
    if (int rc = net_alloc_acts(e, plan)) return rc;
    const int n_pairs = (plan + 2 * BLOCK_M - 1) / (2 * BLOCK_M), bpad = N->capacity;
    cudaStream_t st = e->stream;
    // dense leaf batch of az_search: the number of live rows is only known on the device
    // (az_config.defer_rows: the rows the pass evaluates, cap_rows_kernel -- a short last tile pair waits for the next batch)
    const uint32_t* count = (search_view && search_view->compact)
                                ? (search_view->defer_thr > 0 ? search_view->row_eff : search_view->row_count) + search_view->parity : nullptr;
    {
        const long long squares = (long long)n * NPOS;                           // one thread per (board, square)
        int grid = (int)std::max<long long>(1, std::min<long long>((squares + 255) / 256, (long long)num_sms() * 8));
        stem_onehot_kernel<<<grid, 256, 0, st>>>(tokens, n, bpad, N->stem_in, count, (uint32_t)row_base);
        MCAZ_CHECK_LAUNCH();
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (N->profiling) {
        if (N->events_used == N->events.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            N->events.emplace_back(a, b);
        }
        ev0 = N->events[N->events_used].first; ev1 = N->events[N->events_used].second;
        N->events_used++;
        cudaEventRecord(ev0, st);
    }
    TowerParams& T = N->tower_params;     // 20 KB parameter block, passed by value; the biases in it are set by network_set_weights
    std::memcpy(T.pos_order, N->pos_order, sizeof(T.pos_order));
    T.bias_g = N->bias; T.act0 = N->act[0]; T.act1 = N->act[1]; T.head_w = reinterpret_cast<const float4*>(N->heads.hw4); T.head_in = N->head_in;
    T.bpad = bpad; T.n_pairs = n_pairs; T.count = count; T.row_base = (uint32_t)row_base; T.claim = N->tower_claim;
    T.first_level = (tower_debug() & 2) ? 1 : 0; T.n_levels = NLEVELS - T.first_level; T.fuse_heads = (tower_debug() & 1) ? 0 : 1;
    // Groups of tile pairs go through all levels one after the other, so a group's ping-pong buffers stay resident in the
    // 126 MB L2 from level to level: at most 12 pairs per group (3072 boards, 2 x 47 MB), split evenly (16 pairs -> 8 + 8).
    // Small groups leave too few items per level to hide the waits on the previous level (measured on 2 800-row batches:
    // groups of <= 4 pairs 1.40 ms, <= 6: 1.27, <= 8: 1.17, <= 12: 1.165, 16: 1.19).
    T.group_max = env_or("MCAZ_TOWER_GROUP", GROUP_MAX_PAIRS);
    {
        static int want_stats = -1;
        if (want_stats < 0) { const char* ss = getenv("MCAZ_TOWER_STATS"); want_stats = ss && atoi(ss) > 0; }
        if (want_stats && !N->stats) {
            MCAZ_CUDA(cudaMalloc(&N->stats, (size_t)num_sms() * 12 * sizeof(unsigned long long)));
            MCAZ_CUDA(cudaMemset(N->stats, 0, (size_t)num_sms() * 12 * sizeof(unsigned long long)));
        }
        T.stats = N->stats;
    }
    T.wait_hint = (uint32_t)env_or("MCAZ_WAIT_HINT", 2000);     // A/B on one box: 267.2 -> 266.1 ms per 200 batches; 20000 ns: 266.8
    T.scale_g = N->scale; T.actq0 = N->actq[0]; T.actq1 = N->actq[1]; T.level_absmax = N->calibrating ? N->level_absmax : nullptr;
    T.fp8_levels = N->fp8_levels;
    T.l2_budget = (float)env_or("MCAZ_L2_BUDGET_PCT", 100) / 100.f * (float)N->l2_persist_bytes;
    const int grid = 2 * std::max(1, std::min(num_sms() / 2, n_pairs * NPOS));
    auto launch = [&]() {
        if (N->calibrating)
            tower_tc_kernel<false, true><<<grid, tower_threads(false), TOWER_SMEM, st>>>(N->map_act[0], N->map_act[1], N->map_w, N->map_stem_in, N->map_stem_w,
                                                                                N->map_q[0], N->map_q[1], N->map_wq, T);
        else if (N->fp8)
            tower_tc_kernel<true, false><<<grid, tower_threads(true), TOWER_SMEM_FP8, st>>>(N->map_act[0], N->map_act[1], N->map_w, N->map_stem_in,
                                                                                    N->map_stem_w, N->map_q[0], N->map_q[1], N->map_wq, T);
        else
            tower_tc_kernel<false, false><<<grid, tower_threads(false), TOWER_SMEM, st>>>(N->map_act[0], N->map_act[1], N->map_w, N->map_stem_in,
                                                                                 N->map_stem_w, N->map_q[0], N->map_q[1], N->map_wq, T);
    };
    if (N->per_layer) {
        T.flags = nullptr; T.epoch = 0; T.n_levels = 1;
        for (int L = (tower_debug() & 2) ? 1 : 0; L < NLEVELS; ++L) {
            T.first_level = L;
            launch();
            MCAZ_CHECK_LAUNCH();
        }
        e->launches += NLEVELS - 1;
    } else {
        if (int rc = tower_flags_for(e, n_pairs)) return rc;
        T.flags = N->tower_flags; T.epoch = ++N->epoch;
        launch();
        MCAZ_CHECK_LAUNCH();
    }
    if (ev1) cudaEventRecord(ev1, st);
    if (!search_view && !logits) { e->launches += 2; return MCAZ_OK; }      // calibration pass: the tower only
    if (search_view)
    {
        const int grid = std::min(num_sms() * 4, (n + HEADS_WARPS - 1) / HEADS_WARPS);
        if (search_view->spec_rows > 0)
            heads_legal_kernel<true><<<grid, HEADS_THREADS, 0, st>>>(N->head_in, N->heads, *search_view, values, row_base, n);
        else
            heads_legal_kernel<false><<<grid, HEADS_THREADS, 0, st>>>(N->head_in, N->heads, *search_view, values, row_base, n);
    }
    else
        heads_kernel<<<std::min(num_sms() * 4, (n + HEADS_WARPS - 1) / HEADS_WARPS), HEADS_THREADS, 0, st>>>(N->head_in, clocks, n, N->heads, logits, values);
    MCAZ_CHECK_LAUNCH();
    e->launches += 3;
    return MCAZ_OK;
}

// Leaf evaluation inside az_search: the engine's own leaf batch (slot = game), priors written into the tree.
int network_forward_search(az_engine* e, const az::View& V, float* values) {
    Network* N = e->net;
    if (!N->have_weights) return fail(MCAZ_ESTATE, "network weights have not been set (az_set_weights)");
    // more rows than one pass holds: chunks of 8192 rows of the (dense) batch, one after the other
    const int rows = V.compact ? std::max(V.row_cap, V.G * V.K) : V.G * V.K;    // dense batch: slots + look-ahead rows
    const int n_chunks = (rows + MAX_CHUNK_BOARDS - 1) / MAX_CHUNK_BOARDS;
    const int chunk = (((rows + n_chunks - 1) / n_chunks + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * (2 * BLOCK_M);
    for (int base = 0; base < rows; base += chunk) {
        const int m = std::min(chunk, rows - base);
        if (int rc = forward_chunk(e, V.tokens + (size_t)base * MC_TOKENS, V.clocks + base, m, nullptr, values, &V, base, chunk)) return rc;
    }
    return MCAZ_OK;
}

// e4m3 tower, at every az_set_weights: (1) the bf16 form runs CALIB_ROWS positions from random playouts and records the largest
// activation of every level; a level's e4m3 operands are its activations over (1.25 x that maximum / 448), saturating -- a
// static per-level scale, as emulated in tools/fp8_probe.py; (2) the folded weights go to e4m3 over a per-output-channel
// scale; (3) the epilogue's dequantisation factors are the products of the two.
static int calibrate_fp8(az_engine* e, const float* flat) {
    Network* N = e->net;
    cudaStream_t st = e->stream;
    mc_state start;
    mc_state_from_fen("2nbk/2ppp/5/5/PPP2/KBN2 w 0 1", &start);
    calib_positions_kernel<<<(CALIB_ROWS + 127) / 128, 128, 0, st>>>(start, e->v.rules, CALIB_ROWS, N->calib_tokens);
    MCAZ_CHECK_LAUNCH();
    MCAZ_CUDA(cudaMemsetAsync(N->level_absmax, 0, NLEVELS * sizeof(unsigned int), st));
    N->calibrating = true;
    const bool was_profiling = N->profiling;
    N->profiling = false;
    const int rc = forward_chunk(e, N->calib_tokens, nullptr, CALIB_ROWS, nullptr, nullptr, nullptr, 0);
    N->calibrating = false;
    N->profiling = was_profiling;
    if (rc) return rc;
    prep_tower_fp8_kernel<<<dim3(C, NLAYERS), C, 0, st>>>(flat, N->wq, N->w_scale);
    MCAZ_CHECK_LAUNCH();
    e->launches += 2;
    unsigned int absmax[NLEVELS];
    std::vector<float> w_scale((size_t)NLAYERS * C), scale((size_t)NLEVELS * C, 1.0f);
    MCAZ_CUDA(cudaMemcpyAsync(absmax, N->level_absmax, sizeof(absmax), cudaMemcpyDeviceToHost, st));
    MCAZ_CUDA(cudaMemcpyAsync(w_scale.data(), N->w_scale, w_scale.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    MCAZ_CUDA(cudaStreamSynchronize(st));
    for (int L = 0; L < NLEVELS; ++L) {
        float m;
        std::memcpy(&m, &absmax[L], sizeof(m));
        N->act_scale[L] = std::max(1.25f * m, 1e-6f) / E4M3_MAX;
        N->tower_params.inv_a[L] = 1.0f / N->act_scale[L];
    }
    for (int L = 1; L <= N->fp8_levels; ++L)       // level L = convolution L - 1 on the output of level L - 1; bf16 levels keep 1
        for (int n = 0; n < C; ++n) scale[(size_t)L * C + n] = w_scale[(size_t)(L - 1) * C + n] * N->act_scale[L - 1];
    MCAZ_CUDA(cudaMemcpyAsync(N->scale, scale.data(), scale.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    MCAZ_CUDA(cudaStreamSynchronize(st));
    return MCAZ_OK;
}

int network_profile(az_engine* e, int on, double* avg_ms_per_tower, int* n_forwards, int* launches_per_forward) {
    Network* N = e->net;
    if (!N) return fail(MCAZ_ESTATE, "engine has no built-in network");
    if (avg_ms_per_tower) {
        cudaStreamSynchronize(e->stream);
        double total = 0;
        for (size_t i = 0; i < N->events_used; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, N->events[i].first, N->events[i].second);
            total += ms;
        }
        *avg_ms_per_tower = N->events_used ? total / (double)N->events_used : 0.0;
        if (n_forwards) *n_forwards = (int)N->events_used;
    }
    if (launches_per_forward) *launches_per_forward = N->per_layer ? NLEVELS : 1;
    N->events_used = 0;
    N->profiling = on != 0;
    return MCAZ_OK;
}

}  // namespace mcaz

extern "C" int az_tower_stats(az_engine* e, unsigned long long* out, int capacity) {
    if (!e || !e->net || !out) return mcaz::fail(MCAZ_EINVAL, "az_tower_stats: bad argument");
    if (!e->net->stats) return mcaz::fail(MCAZ_ESTATE, "az_tower_stats: run with MCAZ_TOWER_STATS=1");
    const int n = std::min(capacity, mcaz::num_sms() * 12);
    cudaStreamSynchronize(e->stream);
    if (cudaMemcpy(out, e->net->stats, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
        return mcaz::fail(MCAZ_ECUDA, "az_tower_stats: copy failed");
    return n;
}

extern "C" int az_profile_network(az_engine* e, int on, double* avg_ms_per_tower, int* n_forwards, int* launches_per_forward) {
    if (!e) return mcaz::fail(MCAZ_EINVAL, "az_profile_network: null engine");
    return mcaz::network_profile(e, on, avg_ms_per_tower, n_forwards, launches_per_forward);
}
