// Tensor-core tap-GEMM, second generation (see tc2.cuh for the design and the measurements behind it).
//
//   dst[pos(m), n] = epilogue( sum_kb  A_kb[m, 0:32] . W_kb[n, 0:32] ),     a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo
//
// Persistent, warp-specialised, one CTA per SM in clusters of CS CTAs that share the weight tiles.  Default (CS = 2,
// CPB_TC_PAIR=1): the two CTAs form a tcgen05 CTA PAIR -- one elected lane of the leader CTA issues cta_group::2 MMAs
// (M = 256) for both SMs, each CTA stages only its half of the weight tile (see Tc2Cfg), which makes room for a 4th stage:
//   warp 13      producer: per k-block ONE cp.async.bulk.tensor (TMA tensor map: the [positions x 32 floats] box of the
//                activation tile, 128-byte rows, SWIZZLE_128B, out-of-image positions zero-filled by the copy engine) and
//                ONE cp.async.bulk slice of the pre-swizzled, pre-split weight image [b_hi | b_lo], multicast to the cluster; both complete on the
//                stage's `full` mbarrier by byte count.
//   warps 8-11   splitters: a_lo = a - trunc_tf32(a) for the A tile, shared memory -> shared memory
//                (the tensor core ignores the 13 low mantissa bits of a TF32 operand, so the raw tile IS x_hi);
//                fence.proxy.async, then `ready`.
//   warp 12      MMA issuer (one lane): per 8-wide k-step  [main | cross] (+)= a_hi x [b_hi | b_lo]  (one N = 2*BN
//                tcgen05.mma) and  cross += a_lo x b_hi;  tcgen05.commit frees the stage for every CTA of the cluster.
//   warps 0-7    drain (tcgen05.ld of each finished 128-k chunk into fp32 register accumulators: the tensor core
//                accumulates with round-toward-zero, see tc_tapgemm.cu) and the bias / ReLU / ReLU-mask epilogue.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "tc2.cuh"

namespace cpb {

namespace {

using namespace tc;

// PAIR: the two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 stream (M = 256: 128 rows per CTA).  Each CTA then
// stages only ITS half of the weight tile -- rows [r*BN/2, (r+1)*BN/2) of b_hi and of b_lo, adjacent -- and the tensor
// cores read each half once for both SMs: a k-block costs a CTA 16 KB (A) + 16 KB (a_lo) + BN*128 B (weights) of shared
// memory instead of + 2*BN*128 B, which buys the 4th pipeline stage at BN = 128 and cuts the operand reads per MMA
// (profiles/r2_cycle_accounting.md: the 1-CTA kernel saturates the shared-memory data path, not the tensor pipe).
// Accumulator columns of a buffer (h = BN/2):  [m0 | c0 | m1 | c1]  <-  a_hi x {CTA0: [b_hi 0:h | b_lo 0:h], CTA1: [b_hi h:2h | b_lo h:2h]};
// the third product a_lo x b_hi (N = BN, h rows from each CTA) lands on columns [h, h + BN) = [c0 | m1]: both are summed
// into the same outputs by the drain.
template <int BN, bool PAIR>
struct Tc2Cfg {
    static constexpr int B_TILE_BYTES = BN * TBK * 4;
    static constexpr int B_STAGE_BYTES = PAIR ? B_TILE_BYTES : 2 * B_TILE_BYTES;
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + B_STAGE_BYTES;         // [A_hi | A_lo | B_hi | B_lo]
    static constexpr int STAGES = (STAGE_BYTES * 4 <= 200 * 1024) ? 4 : 3;
    static constexpr int SLICE = BN / 2 < 32 ? BN / 2 : 32;                      // columns transposed per epilogue pass
    static constexpr int STAGING_BYTES = 8 * 32 * SLICE * 4;                     // 8 drain warps x [32 rows x SLICE floats]
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024;
    // TSA (BN <= 64): the A operand of both MMAs is read from TENSOR MEMORY, not shared memory: the splitters move each
    // k-block's rows (raw = hi, and lo) into TMEM columns with tcgen05.st, 2 x 32 columns per stage, next to the accumulators
    static constexpr bool TSA = BN <= 64 && !PAIR;
    static constexpr int ACC_COLS = 4 * BN;
    static constexpr int TMEM_NEED = ACC_COLS + (TSA ? STAGES * 64 : 0);
    static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
    static_assert(TMEM_NEED <= 512, "tensor memory budget");
};
constexpr int kPrefetchKb = 12;   // k-blocks the L2 prefetch runs ahead of the shared-memory copies
constexpr int CHUNK_KB = 4;

constexpr int kDrainWarps = 8;
constexpr int kSplitWarp0 = 8;
constexpr int kSplitWarps = 4;
constexpr int kSplitThreads = kSplitWarps * 32;
constexpr int kIssuerWarp = 12;
constexpr int kProducerWarp = 13;
constexpr int kThreads = 448;

int g_cluster = 2;
int g_resident[3] = {0, 0, 0};
int g_enabled = 1;
int g_pair = 1;                  // CTA-pair MMAs (cta_group::2, CPB_TC_PAIR, default on); needs clusters of exactly 2

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask, bool multicast) {
    if (multicast)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA tensor-map load of a 4-D box into shared memory (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask, bool multicast) {
    if (multicast)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"(mask) : "memory");
    else
        umma_commit(bar);
}

// 32 lanes x 32 consecutive 32-bit columns <- 32 registers per thread (this warp's lane quarter)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
          "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]),
          "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]),
          "f"(v[30]), "f"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] x B[smem descriptor]   (A: lane = row, one 32-bit column per k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- CTA-pair (cta_group::2) flavours
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot) {      // one warp of EACH CTA of the pair, same slot offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of every MMA issued so far -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at this offset in CTA `cta` of the cluster.  Plain form on purpose: `.release.cluster` compiles to
// MEMBAR.ALL.GPU (measured: +650 clk per k-block, the splitters became the bottleneck) and is not needed -- what the
// waiter triggers is the ARRIVING CTA's own tensor core reading that CTA's shared / tensor memory, and the arriving
// thread has already fenced those writes (fence.proxy.async / tcgen05.fence::before_thread_sync) before this instruction.
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
#define TC2_PROF(slot) do { if constexpr (PROF) { const long long now_ = clock64(); prof[slot] += now_ - tlast; tlast = now_; } } while (0)

template <int BN, bool PAIR, bool PROF>
__global__ void __launch_bounds__(kThreads, 1)
tc2_tapgemm_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ Tc2Params p, const int mtiles, const int total_st) {
    using Cfg = Tc2Cfg<BN, PAIR>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int B_TILE_BYTES = Cfg::B_TILE_BYTES;
    constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
    static_assert(STAGES <= CHUNK_KB, "late drain relies on the stage ring being no deeper than a chunk");

    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES];     // copy engine (A box + weight slices, by bytes) -> splitters
    __shared__ uint64_t ready_bar[STAGES];    // splitter warps -> issuer: lo tiles written
    __shared__ uint64_t empty_bar[STAGES];    // tensor cores of ALL CTAs of the cluster -> producer
    __shared__ uint64_t chunk_bar[2];         // tensor core -> drain warps
    __shared__ uint64_t drained_bar[2];       // drain warps -> issuer
    __shared__ uint32_t tmem_slot;
    // PROF: per-stage time stamps of the refill cycle (commit -> copies issued -> full -> ready -> MMAs issued)
    __shared__ long long ts_commit[PROF ? 4 : 1], ts_issue[PROF ? 4 : 1], ts_ready[PROF ? 4 : 1];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int ntn = p.N / BN;
    const int CS = p.cluster;
    const int rank = CS > 1 ? (int)cluster_ctarank() : 0;
    const int cl_id = CS > 1 ? (int)cluster_id_x() : (int)blockIdx.x;
    const int cl_n = CS > 1 ? (int)cluster_count_x() : (int)gridDim.x;
    const uint16_t cl_mask = (uint16_t)((1u << CS) - 1u);
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int box_rows = p.bw * p.bh * p.bn;
    const int nkb = p.nkb;

    if (tid == 0) {
#pragma unroll
        // PAIR: ready / drained live in the leader CTA and collect the arrivals of both CTAs; one multicast commit frees a stage in both
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&ready_bar[s], (uint32_t)((PAIR ? 2 : 1) * kSplitWarps));
            mbar_init(&empty_bar[s], PAIR ? 1u : (uint32_t)CS);
        }
        mbar_init(&chunk_bar[0], 1); mbar_init(&chunk_bar[1], 1);
        mbar_init(&drained_bar[0], PAIR ? 2 * kDrainWarps : kDrainWarps); mbar_init(&drained_bar[1], PAIR ? 2 * kDrainWarps : kDrainWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if constexpr (PAIR) tmem_alloc2<Cfg::TMEM_COLS>(&tmem_slot);
        else tmem_alloc<Cfg::TMEM_COLS>(&tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    // super-tile st -> n-tile y (fastest) and m-tile (mgroup * CS + rank); m-tile -> box origin (x fastest)
    auto st_y = [&](int st) { return st % ntn; };
    auto st_mt = [&](int st) { return (st / ntn) * CS + rank; };
    auto mt_origin = [&](int mt, int& x0, int& y0, int& n0) {
        const int tx = mt % p.tiles_x;
        const int r = mt / p.tiles_x;
        x0 = tx * p.bw; y0 = (r % p.tiles_y) * p.bh; n0 = (r / p.tiles_y) * p.bn;     // mt >= mtiles: n0 >= batch (all rows invalid)
    };

    if (warp == kProducerWarp) {
        // ================================ producer ================================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&amap)) : "memory");
            // 1-CTA MMAs: every CTA copies 1/CS of the [b_hi | b_lo] block and multicasts it; PAIR: the CTA's own half block
            // ([b_hi rows of this CTA | b_lo rows of this CTA], tc_weights_kernel raw = 2), no multicast
            const uint32_t slice = (uint32_t)(2 * B_TILE_BYTES / CS);
            const uint32_t tx_bytes = (uint32_t)(((p.debug & 2) ? 0 : box_rows * 128) + Cfg::B_STAGE_BYTES);   // debug 2: no A copies (timing)
            // The activation tiles come from HBM (GBs per layer, nothing is L2-resident): a copy issued when its stage
            // frees up would expose the full DRAM latency to a ring of only 3-4 stages.  A second cursor therefore runs
            // kPrefetchKb k-blocks ahead and pulls the boxes into L2 (cp.async.bulk.prefetch.tensor), no shared memory needed.
            int pst = cl_id, pkb = 0, px0 = 0, py0 = 0, pn0 = 0;
            auto pf_origin = [&]() { if (pst < total_st) { mt_origin(st_mt(pst), px0, py0, pn0); px0 *= p.sx; py0 *= p.sx; } };
            auto pf_step = [&]() {
                if (pst >= total_st) return;
                if (!(p.debug & 8))
                    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                                 ::"l"(reinterpret_cast<uint64_t>(&amap)), "r"((int)p.kb[pkb].c), "r"(px0 + p.kb[pkb].dx), "r"(py0 + p.kb[pkb].dy), "r"(pn0) : "memory");
                if (++pkb == nkb) { pkb = 0; pst += cl_n; pf_origin(); }
            };
            pf_origin();
            for (int i = 0; i < kPrefetchKb; ++i) pf_step();
            int g = 0;
            long long prof[4] = {0, 0, 0, 0}, tlast = 0;
            if constexpr (PROF) tlast = clock64();
            const long long tstart = tlast;
            for (int st = cl_id; st < total_st; st += cl_n) {
                int x0, y0, n0;
                mt_origin(st_mt(st), x0, y0, n0);
                x0 *= p.sx; y0 *= p.sx;
                const char* wt = reinterpret_cast<const char*>(p.wk) + (size_t)st_y(st) * nkb * (2 * B_TILE_BYTES) + (size_t)rank * slice;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % STAGES;
                    pf_step();
                    TC2_PROF(1);
                    if (g >= STAGES) mbar_wait(&empty_bar[s], (uint32_t)((g / STAGES - 1) & 1));
                    TC2_PROF(0);
                    if constexpr (PROF) { if (g >= STAGES) prof[3] += clock64() - ts_commit[s]; }
                    const uint32_t stage = smem_base + s * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], tx_bytes);
                    if (!(p.debug & 2)) tma_load_4d(stage, &amap, p.kb[kb].c, x0 + p.kb[kb].dx, y0 + p.kb[kb].dy, n0, &full_bar[s]);
                    bulk_g2s(stage + 2 * A_TILE_BYTES + (PAIR ? 0u : (uint32_t)rank * slice), wt + (size_t)kb * (2 * B_TILE_BYTES), slice, &full_bar[s], cl_mask,
                             !PAIR && CS > 1);
                    if constexpr (PROF) ts_issue[s] = clock64();
                    TC2_PROF(2);
                }
            }
            if constexpr (PROF) {
                if (blockIdx.x == 0)
                    printf("tc2prof   producer total %lld: wait_empty %lld prefetch %lld issue_copies %lld | commit->producer awake %lld per kb\n", clock64() - tstart, prof[0], prof[1], prof[2],
                           prof[3] / (g > STAGES ? g - STAGES : 1));
            }
        }
        __syncwarp();
    } else if (warp >= kSplitWarp0 && warp < kSplitWarp0 + kSplitWarps) {
        // ================================ splitters ================================
        // a_lo = a - trunc_tf32(a), elementwise on the swizzled image (the layout does not matter): 16-byte chunk
        // tl + 128 * i of the A tile (1024 chunks).  The weight tile arrives pre-split ([b_hi | b_lo], once per step by
        // tc_weights_kernel): every byte a splitter moves competes with the tensor core's operand reads for the 128 B/clk of
        // shared-memory bandwidth, which is what bounds this kernel (profiles/r2_cycle_accounting.md).
        const int tl = tid - kSplitWarp0 * 32;
        int g = 0;
        long long prof[4] = {0, 0, 0, 0}, tlast = 0;
        if constexpr (PROF) tlast = clock64();
        const long long tstart = tlast;
        for (int st = cl_id; st < total_st; st += cl_n) {
            for (int kb = 0; kb < nkb; ++kb, ++g) {
                const int s = g % STAGES;
                mbar_wait(&full_bar[s], (uint32_t)((g / STAGES) & 1));
                TC2_PROF(0);
                if constexpr (PROF) { if (tid == kSplitWarp0 * 32) prof[2] += clock64() - ts_issue[s]; }
                const uint32_t stage = smem_base + s * STAGE_BYTES;
                if constexpr (Cfg::TSA) {
                    // row tl of the tile (8 swizzled 16-byte chunks, conflict-free per quarter-warp) -> registers -> TMEM:
                    // columns [0,32) of the stage's slot hold a_hi (the raw values), [32,64) hold a_lo
                    float v[32], l[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[4 * c]), "=f"(v[4 * c + 1]), "=f"(v[4 * c + 2]), "=f"(v[4 * c + 3])
                                     : "r"(stage + (uint32_t)(tl * 128 + ((c ^ (tl & 7)) << 4))));
#pragma unroll
                    for (int i = 0; i < 32; ++i) { float h; split_tf32(v[i], h, l[i]); }
                    const uint32_t ta = tmem_base + ((uint32_t)((warp - kSplitWarp0) * 32) << 16) + (uint32_t)(Cfg::ACC_COLS + s * 64);
                    tmem_st32(ta, v);
                    tmem_st32(ta + 32u, l);
                    tmem_st_wait();
                    tc_fence_before();
                } else if (!(p.debug & 4)) {
#pragma unroll
                    for (int i = 0; i < A_TILE_BYTES / 16 / kSplitThreads; ++i) {
                        const uint32_t a = stage + (uint32_t)((tl + i * kSplitThreads) * 16);
                        float4 v, h, l;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                        split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + A_TILE_BYTES), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                    }
                }
                fence_async_smem();          // generic-proxy writes (lo tiles) -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR) mbar_arrive_cta(&ready_bar[s], 0u);
                    else mbar_arrive(&ready_bar[s]);
                    if constexpr (PROF) { if (warp == kSplitWarp0 + kSplitWarps - 1) ts_ready[s] = clock64(); }
                }
                TC2_PROF(1);
            }
        }
        if constexpr (PROF) {
            if (blockIdx.x == 0 && tid == kSplitWarp0 * 32)
                printf("tc2prof   splitter total %lld: wait_full %lld split+arrive %lld | copies issued->full (observed) %lld per kb\n", clock64() - tstart, prof[0], prof[1], prof[2] / (g > 0 ? g : 1));
        }
    } else if (warp == kIssuerWarp) {
        // ================================ MMA issuer ================================
        // PAIR: only the leader CTA (rank 0) issues; every instruction drives the tensor cores of both SMs
        if (lane == 0 && (!PAIR || rank == 0)) {
            constexpr uint32_t MDIM = PAIR ? 2u * TBM : (uint32_t)TBM;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((MDIM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((MDIM >> 4) << 24);
            int g = 0, gc = 0;
            long long prof[4] = {0, 0, 0, 0}, tlast = 0, lat_sum = 0;
            if constexpr (PROF) tlast = clock64();
            const long long tstart = tlast;
            for (int st = cl_id; st < total_st; st += cl_n) {
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % STAGES;
                    const uint32_t stage = smem_base + s * STAGE_BYTES;
                    const int b = gc & 1;
                    TC2_PROF(3);
                    mbar_wait(&ready_bar[s], (uint32_t)((g / STAGES) & 1));
                    TC2_PROF(0);
                    long long lat_ready = 0;
                    if constexpr (PROF) lat_ready = clock64() - ts_ready[s];
                    if (kb % CHUNK_KB == 0 && gc >= 2)
                        mbar_wait(&drained_bar[b], (uint32_t)(((gc >> 1) - 1) & 1));
                    TC2_PROF(1);
                    tc_fence_after();
                    const uint64_t a_hi = make_desc(stage);
                    const uint64_t a_lo = make_desc(stage + A_TILE_BYTES);
                    const uint64_t b_hi = make_desc(stage + 2 * A_TILE_BYTES);     // [b_hi | b_lo] adjacent: one N = 2*BN operand
                    const uint32_t d_main = tmem_base + (uint32_t)(b * 2 * BN);
                    const uint32_t d_cross = d_main + (uint32_t)(PAIR ? BN / 2 : BN);   // PAIR: columns [c0 | m1], see Tc2Cfg
                    if constexpr (PAIR) {
                        if (!(p.debug & 1))
#pragma unroll
                        for (int ks = 0; ks < TBK / 8; ++ks) {
                            const uint64_t adv = (uint64_t)(ks * 2);
                            umma2_tf32(d_main, a_hi + adv, b_hi + adv, idesc2, ((kb % CHUNK_KB) | ks) != 0 ? 1u : 0u);
                            umma2_tf32(d_cross, a_lo + adv, b_hi + adv, idesc, 1u);
                        }
                        umma2_commit(&empty_bar[s]);
                        if constexpr (PROF) { ts_commit[s] = clock64(); lat_sum += lat_ready; }
                        if (kb % CHUNK_KB == CHUNK_KB - 1 || kb == nkb - 1) { umma2_commit(&chunk_bar[b]); ++gc; }
                    } else {
                        if constexpr (Cfg::TSA) {
                            const uint32_t ta = tmem_base + (uint32_t)(Cfg::ACC_COLS + s * 64);
                            if (!(p.debug & 1))
#pragma unroll
                            for (int ks = 0; ks < TBK / 8; ++ks) {
                                const uint64_t adv = (uint64_t)(ks * 2);
                                umma_tf32_ts(d_main, ta + (uint32_t)(ks * 8), b_hi + adv, idesc2, ((kb % CHUNK_KB) | ks) != 0 ? 1u : 0u);
                                umma_tf32_ts(d_cross, ta + 32u + (uint32_t)(ks * 8), b_hi + adv, idesc, 1u);
                            }
                        } else if (!(p.debug & 1)) {
#pragma unroll
                            for (int ks = 0; ks < TBK / 8; ++ks) {
                                const uint64_t adv = (uint64_t)(ks * 2);
                                umma_tf32(d_main, a_hi + adv, b_hi + adv, idesc2, ((kb % CHUNK_KB) | ks) != 0 ? 1u : 0u);
                                umma_tf32(d_cross, a_lo + adv, b_hi + adv, idesc, 1u);
                            }
                        }
                        umma_commit_mc(&empty_bar[s], cl_mask, CS > 1);
                        if constexpr (PROF) { ts_commit[s] = clock64(); lat_sum += lat_ready; }
                        if (kb % CHUNK_KB == CHUNK_KB - 1 || kb == nkb - 1) { umma_commit(&chunk_bar[b]); ++gc; }
                    }
                    TC2_PROF(2);
                }
            }
            if constexpr (PROF) {
                if (blockIdx.x == 0)
                    printf("tc2prof BN=%d N=%d nkb=%d quad=%d box=%dx%dx%d supertiles=%d cluster=%d pair=%d stages=%d kb/cta=%d | issuer total %lld: wait_ready %lld wait_drained %lld issue %lld other %lld | ready(last splitter of this CTA)->issuer awake %lld per kb\n",
                           BN, p.N, nkb, p.quad, p.bw, p.bh, p.bn, total_st, CS, PAIR ? 1 : 0, STAGES, g, clock64() - tstart, prof[0], prof[1], prof[2], prof[3], lat_sum / (g > 0 ? g : 1));
            }
        }
        __syncwarp();
    } else if (warp < kDrainWarps) {
        // ================================ accumulator drain + epilogue ================================
        constexpr int HALF_COLS = BN / 2;
        const int q = warp & 3;
        const int half = warp >> 2;
        // this warp's output columns [half*HALF_COLS, +HALF_COLS): main at accumulator column ..., its cross term CROSS_OFF further
        constexpr int CROSS_OFF = PAIR ? HALF_COLS : BN;                       // PAIR: [m0 | c0 | m1 | c1] with h = HALF_COLS
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (PAIR ? 2 * HALF_COLS : HALF_COLS));
        float acc[HALF_COLS];
#pragma unroll
        for (int i = 0; i < HALF_COLS; ++i) acc[i] = 0.f;
        const int bwh = p.bw * p.bh;

        // Epilogue.  After the drain a thread owns ONE output row x HALF_COLS columns; storing that directly makes every
        // warp-level store touch 32 different 128-byte lines with 16 bytes each (measured: 9-16 k clk per tile, more than the
        // tile's MMAs on the short-K layers).  Each warp therefore transposes SLICE = 32 columns at a time through a private
        // shared-memory tile (XOR-swizzled, conflict-free both ways): afterwards 8 lanes cover the 128 contiguous bytes of a
        // row slice and a store / mask-load instruction touches 4 full lines instead of 32 partial ones.
        constexpr int SLICE = Cfg::SLICE;
        constexpr int CPR = SLICE / 4;               // 16-byte chunks per row slice (8, or 4 when BN = 32)
        constexpr int RPI = 32 / CPR;                // rows covered by one warp-level access
        const uint32_t stg = smem_base + STAGES * STAGE_BYTES + (uint32_t)warp * (32 * SLICE * 4);
        // destination of THIS lane's row (q*32 + lane of tile st) for column slice sl: float offset into dst / mask, validity,
        // first output channel (a slice never straddles two quad classes: quad_cb % 32 == 0)
        auto slice_dest = [&](int st, int sl, uint32_t& my_off, bool& my_ok, int& ch0) {
            const int r = q * 32 + lane;
            int x0, y0, n0;
            mt_origin(st_mt(st), x0, y0, n0);
            const int nn = r / bwh;
            const int rr = r - nn * bwh;
            const int yy = rr / p.bw;
            const int ox = x0 + rr - yy * p.bw, oy = y0 + yy, n = n0 + nn;
            my_ok = r < box_rows && ox < p.gw && oy < p.gh && n < p.batch;
            const uint32_t img = (uint32_t)n * (uint32_t)p.dst_img;
            const int col_base = st_y(st) * BN + half * HALF_COLS + sl * SLICE;
            if (p.quad) {
                const int c = col_base >> p.quad_lcb;
                ch0 = col_base & (p.quad_cb - 1);
                const int y = oy * 2 + (c >> 1), x = ox * 2 + (c & 1);
                my_ok = my_ok && y < p.Hd && x < p.Wd;
                my_off = img + (uint32_t)((y * p.Wd + x) * p.dst_pitch + ch0);
            } else {
                ch0 = col_base;
                my_off = img + (uint32_t)((oy * p.Wd + ox) * p.dst_pitch + col_base);
            }
            if (!my_ok) my_off = 0u;
        };
        // The ReLU mask of a data-gradient layer is the saved forward activation (GBs, HBM resident): pull the tile's mask
        // lines into L2 when the tile STARTS draining, so that the epilogue's mask loads do not each pay a DRAM round trip
        // (measured: 11 k clk of epilogue per tile on conv2.dgrad, mostly exposed load latency).
        auto prefetch_mask = [&](int st) {
            if (p.mask == nullptr) return;
#pragma unroll
            for (int sl = 0; sl < HALF_COLS / SLICE; ++sl) {
                uint32_t off; bool ok; int ch0;
                slice_dest(st, sl, off, ok, ch0);
                if (ok) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.mask + off) : "memory");
            }
        };
        auto epilogue = [&](int st) {
            const int c_lane = lane % CPR, r_lane = lane / CPR;
#pragma unroll
            for (int sl = 0; sl < HALF_COLS / SLICE; ++sl) {
                uint32_t my_off; bool my_ok; int ch0;
                slice_dest(st, sl, my_off, my_ok, ch0);
                // row-major [32][SLICE] with the 16-byte chunk index XOR (row % CPR)
#pragma unroll
                for (int c = 0; c < CPR; ++c)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stg + (uint32_t)(lane * SLICE * 4 + ((c ^ (lane & (CPR - 1))) << 4))),
                                 "f"(acc[sl * SLICE + c * 4 + 0]), "f"(acc[sl * SLICE + c * 4 + 1]), "f"(acc[sl * SLICE + c * 4 + 2]), "f"(acc[sl * SLICE + c * 4 + 3]) : "memory");
                __syncwarp();
                float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + ch0 + c_lane * 4));
                constexpr int NIT = 32 / RPI;
                constexpr int GB = 4;                 // rows in flight per lane (mask loads of a batch are issued together)
                float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);      // column sums of this lane's rows (bias gradient)
#pragma unroll
                for (int i0 = 0; i0 < NIT; i0 += GB) {
                    float4 v[GB], mk[GB];
                    uint32_t offs[GB]; bool oks[GB];
#pragma unroll
                    for (int u = 0; u < GB; ++u) {
                        const int row = (i0 + u) * RPI + r_lane;
                        offs[u] = __shfl_sync(0xffffffffu, my_off, row) + (uint32_t)(c_lane * 4);
                        oks[u] = __shfl_sync(0xffffffffu, my_ok ? 1 : 0, row) != 0;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                                     : "r"(stg + (uint32_t)(row * SLICE * 4 + ((c_lane ^ (row & (CPR - 1))) << 4))));
                    }
                    if (p.mask) {
#pragma unroll
                        for (int u = 0; u < GB; ++u)
                            mk[u] = oks[u] ? __ldg(reinterpret_cast<const float4*>(p.mask + offs[u])) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < GB; ++u) {
                        if (!oks[u]) continue;
                        float4 o = v[u];
                        o.x += bias4.x; o.y += bias4.y; o.z += bias4.z; o.w += bias4.w;
                        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        if (p.mask) {
                            o.x = mk[u].x > 0.f ? o.x : 0.f; o.y = mk[u].y > 0.f ? o.y : 0.f;
                            o.z = mk[u].z > 0.f ? o.z : 0.f; o.w = mk[u].w > 0.f ? o.w : 0.f;
                        }
                        *reinterpret_cast<float4*>(p.dst + offs[u]) = o;
                        bs.x += o.x; bs.y += o.y; bs.z += o.z; bs.w += o.w;
                    }
                }
                if (p.colsum != nullptr) {
                    // fold the row groups (lanes with equal c_lane), then the lanes of row group 0 add their 4 columns to this
                    // (CTA, quarter) row of the buffer.  Each address has ONE writer (this warp), its tiles come in a fixed
                    // order and red.add never returns a value: deterministic, and no round trip is exposed.
#pragma unroll
                    for (int m = CPR; m < 32; m <<= 1) {
                        bs.x += __shfl_xor_sync(0xffffffffu, bs.x, m); bs.y += __shfl_xor_sync(0xffffffffu, bs.y, m);
                        bs.z += __shfl_xor_sync(0xffffffffu, bs.z, m); bs.w += __shfl_xor_sync(0xffffffffu, bs.w, m);
                    }
                    if (r_lane == 0) {
                        float* cp = p.colsum + (size_t)(blockIdx.x * 4 + q) * p.N + st_y(st) * BN + half * HALF_COLS + sl * SLICE + c_lane * 4;
                        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(cp), "f"(bs.x) : "memory");
                        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(cp + 1), "f"(bs.y) : "memory");
                        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(cp + 2), "f"(bs.z) : "memory");
                        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(cp + 3), "f"(bs.w) : "memory");
                    }
                }
                __syncwarp();          // the next slice reuses the transposition tile
            }
        };

        const int chunks_per_tile = (nkb + CHUNK_KB - 1) / CHUNK_KB;
        int drained = 0;
        auto drain_cols = [&](uint32_t taddr) {
#pragma unroll
            for (int cc = 0; cc < HALF_COLS; cc += 32) {
                if constexpr (HALF_COLS >= 32) {
                    float v[16], w[16];
                    tmem_ld16_issue(taddr + (uint32_t)cc, v);
                    tmem_ld16_issue(taddr + (uint32_t)(cc + 16), w);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) { acc[cc + i] += v[i]; acc[cc + 16 + i] += w[i]; }
                } else {
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)cc, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc[cc + i] += v[i];
                }
            }
        };
        int stD = cl_id;
        int chunksD = chunks_per_tile;
        long long prof[4] = {0, 0, 0, 0}, tlast = 0;
        if constexpr (PROF) tlast = clock64();
        const long long tstart = tlast;
        while (stD < total_st) {
            const int b = drained & 1;
            if (chunksD == chunks_per_tile) prefetch_mask(stD);
            TC2_PROF(3);
            mbar_wait(&chunk_bar[b], (uint32_t)((drained >> 1) & 1));
            TC2_PROF(0);
            tc_fence_after();
            drain_cols(tmem_lane + (uint32_t)(b * 2 * BN));
            drain_cols(tmem_lane + (uint32_t)(b * 2 * BN + CROSS_OFF));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cta(&drained_bar[b], 0u);
                else mbar_arrive(&drained_bar[b]);
            }
            ++drained;
            TC2_PROF(1);
            if (--chunksD == 0) {
                epilogue(stD);
                TC2_PROF(2);
#pragma unroll
                for (int i = 0; i < HALF_COLS; ++i) acc[i] = 0.f;
                stD += cl_n;
                chunksD = chunks_per_tile;
            }
        }
        if constexpr (PROF) {
            if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 5))
                printf("tc2prof   drain warp %d total %lld: wait_chunk %lld drain %lld epilogue %lld other %lld (chunks %d)\n",
                       warp, clock64() - tstart, prof[0], prof[1], prof[2], prof[3], drained);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();
    if (warp == 0) {
        if constexpr (PAIR) tmem_dealloc2<Cfg::TMEM_COLS>(tmem_base);
        else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
    (void)mtiles;
}

constexpr int bn_slot(int BN) { return BN == 128 ? 2 : (BN == 64 ? 1 : 0); }

template <int BN, bool PAIR, bool PROF>
int32_t launch_t(const CUtensorMap& map, const Tc2Params& p, int mtiles, int total_st, unsigned grid, cudaStream_t stream) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Tc2Cfg<BN, PAIR>::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)p.cluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    CPB_CUDA(cudaLaunchKernelEx(&cfg, tc2_tapgemm_kernel<BN, PAIR, PROF>, map, p, mtiles, total_st));
    CPB_LAUNCHED();
    return CPB_OK;
}

template <int BN>
int32_t launch_bn(const CUtensorMap& map, Tc2Params& p, cudaStream_t stream) {
    p.cluster = g_cluster;
    const long long mtiles = (long long)p.tiles_x * p.tiles_y * p.tiles_n;
    if (mtiles == 0) return CPB_OK;
    const long long mgroups = (mtiles + p.cluster - 1) / p.cluster;
    const long long total_st = mgroups * (p.N / BN);
    const int resident = g_resident[bn_slot(BN)];
    CPB_REQUIRE(total_st < (1ll << 30) && resident > 0, "tc2_tapgemm: bad tile count");
    const unsigned grid = (unsigned)((total_st < resident ? total_st : resident) * p.cluster);
    CPB_REQUIRE(p.colsum == nullptr || grid * 4 <= (unsigned)kTc2ColsumRows, "tc2_tapgemm: grid of %u CTAs exceeds the column-sum buffer", grid);
    if (g_pair) {
        if (p.debug & 16) return launch_t<BN, true, true>(map, p, (int)mtiles, (int)total_st, grid, stream);
        return launch_t<BN, true, false>(map, p, (int)mtiles, (int)total_st, grid, stream);
    }
    if (p.debug & 16) return launch_t<BN, false, true>(map, p, (int)mtiles, (int)total_st, grid, stream);
    return launch_t<BN, false, false>(map, p, (int)mtiles, (int)total_st, grid, stream);
}

template <int BN, bool PAIR>
int32_t init_pair() {
    using Cfg = Tc2Cfg<BN, PAIR>;
    CPB_CUDA(cudaFuncSetAttribute(tc2_tapgemm_kernel<BN, PAIR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    CPB_CUDA(cudaFuncSetAttribute(tc2_tapgemm_kernel<BN, PAIR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(g_cluster * 1024));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)g_cluster; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    CPB_CUDA(cudaOccupancyMaxActiveClusters(&n, tc2_tapgemm_kernel<BN, PAIR, false>, &cfg));
    CPB_REQUIRE(n > 0, "tc2_tapgemm: no resident cluster of %d CTAs possible", g_cluster);
    g_resident[bn_slot(BN)] = n;
    return CPB_OK;
}

template <int BN>
int32_t init_one() {
    return g_pair ? init_pair<BN, true>() : init_pair<BN, false>();
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

// position box (bw, bh, bn) with bw*bh*bn <= 128 that wastes the fewest MMA rows on a gw x gh grid
void pick_box(int gw, int gh, int max_bw, int& bw, int& bh, int& bn) {
    double best = -1.0;
    bw = bh = 1; bn = 128;
    for (int w = 1; w <= gw && w <= 128 && w <= max_bw; ++w)
        for (int h = 1; h <= gh && w * h <= 128; ++h) {
            const int n = 128 / (w * h);
            const int rows = w * h * n;
            const long long covered = (long long)((gw + w - 1) / w) * w * ((gh + h - 1) / h) * h;
            const double eff = (double)gw * gh / (double)covered * rows / 128.0;
            // ties: wider boxes (longer contiguous runs), then taller ones
            if (eff > best + 1e-9 || (eff > best - 1e-9 && (w > bw || (w == bw && h > bh)))) { best = eff; bw = w; bh = h; bn = n; }
        }
}

}  // namespace

int32_t tc2_encode_tiled(CUtensorMap* map, const float* base, int rank, const unsigned long long* dims,
                         const unsigned long long* strides_bytes, const unsigned* box, const unsigned* elem_strides, int atom32) {
    CPB_REQUIRE(g_encode != nullptr, "cuTensorMapEncodeTiled is not available (driver entry point not resolved)");
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides[i]; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    const CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), gd, gs, bx, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE,
                                atom32 == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : (atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] estr [%u,%u,%u,%u]", (int)r, rank,
                  dims[0], dims[1], rank > 2 ? dims[2] : 0ull, rank > 3 ? dims[3] : 0ull, box[0], box[1], rank > 2 ? box[2] : 0u, rank > 3 ? box[3] : 0u,
                  elem_strides[0], elem_strides[1], rank > 2 ? elem_strides[2] : 0u, rank > 3 ? elem_strides[3] : 0u);
        return CPB_ERR_CUDA;
    }
    return CPB_OK;
}

bool tc2_enabled() { return g_enabled != 0 && g_encode != nullptr; }
int tc2_weight_layout() { return g_pair ? 2 : 1; }

int32_t tc2_tapgemm_init() {
    const char* e = getenv("CPB_TC2");
    g_enabled = e ? atoi(e) : 1;
    e = getenv("CPB_TC_CLUSTER");
    g_cluster = e ? atoi(e) : 2;
    CPB_REQUIRE(g_cluster == 1 || g_cluster == 2 || g_cluster == 4 || g_cluster == 8, "CPB_TC_CLUSTER must be 1, 2, 4 or 8");
    e = getenv("CPB_TC_PAIR");
    g_pair = (e ? atoi(e) : 1) != 0 && g_cluster == 2;
    if (g_encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
        else
            (void)cudaGetLastError();
    }
    CPB_TRY(init_one<32>());
    CPB_TRY(init_one<64>());
    CPB_TRY(init_one<128>());
    return CPB_OK;
}

bool tc2_tapgemm_supported(const TapGemmParams& p, int scatter_k) {
    (void)scatter_k;
    if (!tc2_enabled()) return false;
    if (p.ybatch != 1 || p.nclass != 1 || p.wk_hi == nullptr) return false;
    if (!(p.N == 32 || p.N % 64 == 0)) return false;
    const int nkb = p.cls[0].ntaps * (p.C / TBK);
    if (p.C % TBK != 0 || nkb > kTc2MaxKb) return false;
    if (p.quad) return p.N == 4 * p.quad_cb && p.sstride == 1 && p.src_pitch == p.C;
    if (p.sstride == 2) return p.src_pitch % TBK == 0 && p.C == p.cls[0].ntaps * p.src_pitch;   // square k x k window, run = k * Cb
    return p.Hs == 1 && p.Ws == 1 && p.cls[0].ntaps == 1;                                          // dense
}

int32_t launch_tc2_tapgemm(const TapGemmParams& p, int scatter_k, cudaStream_t stream) {
    CPB_REQUIRE(tc2_tapgemm_supported(p, scatter_k), "tc2_tapgemm: unsupported problem (C=%d, N=%d)", p.C, p.N);
    CPB_REQUIRE((long long)p.batch * p.dst_img < (1ll << 31), "tc2_tapgemm: destination too large for 32-bit offsets");
    const TapClass& cls = p.cls[0];
    Tc2Params q;
    memset(&q, 0, sizeof(q));
    q.gw = cls.Wo; q.gh = cls.Ho; q.batch = p.batch;
    q.sx = p.sstride;
    pick_box(q.gw, q.gh, p.sstride == 2 ? 128 : 256, q.bw, q.bh, q.bn);
    q.tiles_x = (q.gw + q.bw - 1) / q.bw;
    q.tiles_y = (q.gh + q.bh - 1) / q.bh;
    q.tiles_n = (q.batch + q.bn - 1) / q.bn;
    q.N = p.N; q.wk = p.wk_hi;
    q.bias = p.bias; q.mask = p.mask; q.dst = p.dst; q.relu = p.relu;
    q.colsum = p.colsum;
    q.quad = p.quad; q.quad_cb = p.quad_cb;
    if (p.quad) {
        CPB_REQUIRE((p.quad_cb & (p.quad_cb - 1)) == 0, "tc2_tapgemm: quad form needs a power-of-two channel count");
        while ((1 << q.quad_lcb) < p.quad_cb) ++q.quad_lcb;
    }
    q.Hd = p.Hd; q.Wd = p.Wd; q.dst_pitch = p.dst_pitch; q.dst_img = p.dst_img;
    q.debug = p.debug;
    const int kbt = p.C / TBK;
    q.nkb = cls.ntaps * kbt;
    for (int t = 0; t < cls.ntaps; ++t)
        for (int kc = 0; kc < kbt; ++kc) {
            Tc2KBlock& b = q.kb[t * kbt + kc];
            if (p.sstride == 2) {                    // gather: tap t = kernel row kh; the run is (kw, cb)
                const int off = kc * TBK;
                b.c = (short)(off % p.src_pitch); b.dx = (short)(off / p.src_pitch); b.dy = (short)t;
            } else {                                 // quad / dense: tap displacement in source pixels
                b.c = (short)(kc * TBK); b.dx = (short)cls.taps[t].dx; b.dy = (short)cls.taps[t].dy;
            }
        }
    // tensor map of the source: dense NHWC {C, W, H, B}
    alignas(64) CUtensorMap map;
    const unsigned long long dims[4] = {(unsigned long long)p.src_pitch, (unsigned long long)p.Ws, (unsigned long long)p.Hs, (unsigned long long)p.batch};
    const unsigned long long strides[3] = {(unsigned long long)p.src_pitch * 4ull, (unsigned long long)p.Ws * p.src_pitch * 4ull, (unsigned long long)p.src_img * 4ull};
    const unsigned es = (unsigned)p.sstride;
    const unsigned box[4] = {(unsigned)TBK, (unsigned)q.bw * es, (unsigned)q.bh * es, (unsigned)q.bn};
    const unsigned estr[4] = {1u, es, es, 1u};
    CPB_TRY(tc2_encode_tiled(&map, p.src, 4, dims, strides, box, estr));
    switch (tc_bn(p.N)) {
        case 128: return launch_bn<128>(map, q, stream);
        case 64: return launch_bn<64>(map, q, stream);
        default: return launch_bn<32>(map, q, stream);
    }
}

}  // namespace cpb
