// Tensor-core weight gradient, second generation (tf Conv2DBackpropFilter for the conv and transposed-conv layers):
//
//   gw[i, j] = sum_{positions (x, y, n)}  big[n, 2y + kh(i), 2x + kw(i), c(i)] * small[n, y, x, j]        i = (kh, kw, c)
//
// as D[128 x BN] += A^T B with the reduction (output positions) on the MMA K axis.  In NHWC memory the channel index --
// i for `big`, j for `small` -- is contiguous, not the position, so both operands are MN-MAJOR for the tensor core:
// a k-block is a BOX of KR positions (bw x bh x bn of the small grid) and every operand tile is a set of column groups
// [KR rows x 32 floats = 128 B], each ONE TMA tensor-map box.  For 32-bit MN-major operands the tensor core accepts
// exactly one shared-memory layout, SWIZZLE_128B_BASE32B: atoms of 4 rows (k) x 128 B with the 32-BYTE chunk index
// XOR-ed with row % 4 -- written by the copy engine with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B -- i.e. the canonical
// ((4,8,m),(4,k)) : ((1,4,LBO),(32,SBO)) in floats with LBO = KRp*128 B (next 32-wide column group) and SBO = 512 B (next
// 4 positions).  (Round 1 tried the 16-byte SWIZZLE_128B atom here and got zeros: the descriptor's layout type must be 1.)  `big` is read through elementStrides {1,2,2,1} (stride-2 windows); the k-blocks
// of all image rows / images of this CTA's split follow each other.  No register transposes, no LDG/STS operand path
// (round 1: LSU pipe 50-68 %, tensor pipe 27-47 %).  x_lo = x - trunc_tf32(x) is derived in shared memory by the splitter
// warps; 3xTF32 products, separate cross-term columns, chunked drain into fp32 registers as in tc2_tapgemm.cu.
// One wave of split-K CTAs; reduce_partials() sums the splits in a fixed order (deterministic).
#include <cuda.h>

#include "tc2.cuh"
#include "wgrad.cuh"

namespace cpb {

namespace {

using namespace tc;

constexpr int kDrainWarps = 8;
constexpr int kSplitWarp0 = 8;
constexpr int kSplitWarps = 4;
constexpr int kSplitThreads = kSplitWarps * 32;
constexpr int kIssuerWarp = 12;
constexpr int kProducerWarp = 13;
constexpr int kThreads = 448;
constexpr int kStages = 3;
constexpr int kWgPrefetchKb = 10;
constexpr int kWgMaxDynSmem = 226 * 1024;     // 227 KB per CTA minus the static barriers

struct Tc2WgParams {
    int bw, bh, bn;             // position box; KR = bw*bh*bn positions per k-block
    int krp;                    // KR padded to a multiple of 8 (rows of a column group in shared memory)
    int tiles_x, tiles_y, tiles_n;
    long long boxes_per_split;  // k-blocks per split
    long long nboxes;
    int I, J;
    int Cb, run;                // big channels, floats per kernel row (k * Cb)
    float* partial;             // [splits][I][J]
    int chunk_kb;               // k-blocks accumulated in TMEM before a drain (~128 positions)
    int debug;                  // 1: swap LBO / SBO (descriptor probe)
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
// MN-major SWIZZLE_128B_BASE32B operand (layout type 1): LBO = bytes between 32-element column groups, SBO = bytes
// between 4-row (k) atoms
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
tc2_wgrad_kernel(const __grid_constant__ CUtensorMap bigmap, const __grid_constant__ CUtensorMap smallmap, const __grid_constant__ Tc2WgParams p) {
    constexpr int NGA = 4;                   // 32-wide column groups of the A (i) tile
    constexpr int NGB = BN / 32;             // ... of the B (j) tile
    constexpr int TMEM_COLS = BN == 128 ? 512 : (BN == 64 ? 256 : 128);

    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[kStages], ready_bar[kStages], empty_bar[kStages];
    __shared__ uint64_t chunk_bar[2], drained_bar[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int i0 = blockIdx.x * TBM;
    const int j0 = blockIdx.y * BN;
    const uint32_t group_bytes = (uint32_t)p.krp * 128u;                  // one column group: KRp rows x 128 B (multiple of 1024)
    const uint32_t stage_bytes = group_bytes * (2 * NGA + 2 * NGB);       // [A_hi x4 | A_lo x4 | B_hi xNGB | B_lo xNGB]
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const long long kb_begin = (long long)blockIdx.z * p.boxes_per_split;
    long long kb_end = kb_begin + p.boxes_per_split;
    if (kb_end > p.nboxes) kb_end = p.nboxes;
    const int nkb = kb_end > kb_begin ? (int)(kb_end - kb_begin) : 0;
    const int nga = (p.I - i0 + 31) / 32 < NGA ? (p.I - i0 + 31) / 32 : NGA;   // valid column groups of this i-tile
    const int CH = p.chunk_kb;
    const int nchunks = (nkb + CH - 1) / CH;

    // rows >= KR of every column group are never written by the copies and must read as zero: clear all stages once
    for (uint32_t o = (uint32_t)tid * 16u; o < stage_bytes * kStages; o += kThreads * 16u)
        asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(smem_base + o), "f"(0.f) : "memory");
    fence_async_smem();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], kSplitWarps); mbar_init(&empty_bar[s], 1); }
        mbar_init(&chunk_bar[0], 1); mbar_init(&chunk_bar[1], 1);
        mbar_init(&drained_bar[0], kDrainWarps); mbar_init(&drained_bar[1], kDrainWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<TMEM_COLS>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == kProducerWarp) {
        // ================================ producer ================================
        if (lane == 0 && nkb > 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&bigmap)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&smallmap)) : "memory");
            const uint32_t box_bytes = (uint32_t)(p.bw * p.bh * p.bn) * 128u;
            // tensor-map coordinates of the A column groups: i -> (kh, kw, c)
            int gc[NGA], gkw[NGA], gkh[NGA];
#pragma unroll
            for (int g = 0; g < NGA; ++g) {
                const int i = i0 + 32 * g;
                const int kh = i / p.run, off = i - kh * p.run;
                gkh[g] = kh; gkw[g] = off / p.Cb; gc[g] = off - gkw[g] * p.Cb;
            }
            // box cursor (x fastest)
            long long b = kb_begin;
            int tx = (int)(b % p.tiles_x);
            long long r = b / p.tiles_x;
            int ty = (int)(r % p.tiles_y);
            int tn = (int)(r / p.tiles_y);
            // second cursor, kWgPrefetchKb k-blocks ahead: L2 prefetch of the boxes (both operands stream from HBM; the
            // 3-stage ring alone cannot cover the DRAM latency)
            int ptx = tx, pty = ty, ptn = tn, pkb = 0;
            auto pf_step = [&]() {
                if (pkb >= nkb) return;
                const int x0 = ptx * p.bw, y0 = pty * p.bh, n0 = ptn * p.bn;
#pragma unroll
                for (int g = 0; g < NGA; ++g)
                    if (g < nga)
                        asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                                     ::"l"(reinterpret_cast<uint64_t>(&bigmap)), "r"(gc[g]), "r"(2 * x0 + gkw[g]), "r"(2 * y0 + gkh[g]), "r"(n0) : "memory");
#pragma unroll
                for (int g = 0; g < NGB; ++g)
                    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                                 ::"l"(reinterpret_cast<uint64_t>(&smallmap)), "r"(j0 + 32 * g), "r"(x0), "r"(y0), "r"(n0) : "memory");
                ++pkb;
                if (++ptx == p.tiles_x) { ptx = 0; if (++pty == p.tiles_y) { pty = 0; ++ptn; } }
            };
            if (!(p.debug & 8))
                for (int i = 0; i < kWgPrefetchKb; ++i) pf_step();
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                if (!(p.debug & 8)) pf_step();
                if (kb >= kStages) mbar_wait(&empty_bar[s], (uint32_t)((kb / kStages - 1) & 1));
                const uint32_t stage = smem_base + (uint32_t)s * stage_bytes;
                mbar_expect_tx(&full_bar[s], box_bytes * (uint32_t)(nga + NGB));
                const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = tn * p.bn;
#pragma unroll
                for (int g = 0; g < NGA; ++g)
                    if (g < nga) tma_load_4d(stage + (uint32_t)g * group_bytes, &bigmap, gc[g], 2 * x0 + gkw[g], 2 * y0 + gkh[g], n0, &full_bar[s]);
#pragma unroll
                for (int g = 0; g < NGB; ++g)
                    tma_load_4d(stage + (uint32_t)(2 * NGA + g) * group_bytes, &smallmap, j0 + 32 * g, x0, y0, n0, &full_bar[s]);
                if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; ++tn; } }
            }
        }
        __syncwarp();
    } else if (warp >= kSplitWarp0 && warp < kSplitWarp0 + kSplitWarps) {
        // ================================ splitters ================================
        const int tl = tid - kSplitWarp0 * 32;
        const uint32_t a_chunks = (uint32_t)NGA * group_bytes / 16u, b_chunks = (uint32_t)NGB * group_bytes / 16u;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kStages;
            mbar_wait(&full_bar[s], (uint32_t)((kb / kStages) & 1));
            const uint32_t stage = smem_base + (uint32_t)s * stage_bytes;
            for (uint32_t c = (uint32_t)tl; c < a_chunks; c += kSplitThreads) {
                const uint32_t a = stage + c * 16u;
                float4 v, h, l;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + NGA * group_bytes), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
            }
            for (uint32_t c = (uint32_t)tl; c < b_chunks; c += kSplitThreads) {
                const uint32_t a = stage + 2 * NGA * group_bytes + c * 16u;
                float4 v, h, l;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + NGB * group_bytes), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready_bar[s]);
        }
    } else if (warp == kIssuerWarp) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            // instruction descriptor: fp32 accumulate, TF32 x TF32, A and B MN-major (bits 15, 16)
            const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TBM >> 4) << 24);
            const uint32_t idesc = idesc_base | ((uint32_t)(BN >> 3) << 17);
            const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * BN) >> 3) << 17);
            const uint32_t lbo = (p.debug & 1) ? 512u : group_bytes;
            const uint32_t sbo = (p.debug & 1) ? group_bytes : 512u;
            const int ksteps = p.krp / 8;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t stage = smem_base + (uint32_t)s * stage_bytes;
                const int chunk = kb / CH;
                const int b = chunk & 1;
                mbar_wait(&ready_bar[s], (uint32_t)((kb / kStages) & 1));
                if (kb % CH == 0 && chunk >= 2) mbar_wait(&drained_bar[b], (uint32_t)(((chunk >> 1) - 1) & 1));
                tc_fence_after();
                const uint32_t d_main = tmem_base + (uint32_t)(b * 2 * BN);
                const uint32_t d_cross = d_main + (uint32_t)BN;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t koff = (uint32_t)ks * 1024u;                       // 8 positions further down every column group
                    const uint64_t a_hi = make_desc_mn(stage + koff, lbo, sbo);
                    const uint64_t a_lo = make_desc_mn(stage + NGA * group_bytes + koff, lbo, sbo);
                    const uint64_t b_hi = make_desc_mn(stage + 2 * NGA * group_bytes + koff, lbo, sbo);   // [b_hi | b_lo] column groups adjacent
                    umma_tf32(d_main, a_hi, b_hi, idesc2, ((kb % CH) | ks) != 0 ? 1u : 0u);
                    umma_tf32(d_cross, a_lo, b_hi, idesc, 1u);
                }
                umma_commit(&empty_bar[s]);
                if (kb % CH == CH - 1 || kb == nkb - 1) umma_commit(&chunk_bar[b]);
                    }
        }
        __syncwarp();
    } else if (warp < kDrainWarps) {
        // ================================ drain + partial store ================================
        constexpr int HALF_COLS = BN / 2;
        const int q = warp & 3;
        const int half = warp >> 2;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HALF_COLS);
        float acc[HALF_COLS];
#pragma unroll
        for (int i = 0; i < HALF_COLS; ++i) acc[i] = 0.f;
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
        for (int c = 0; c < nchunks; ++c) {
            const int b = c & 1;
            mbar_wait(&chunk_bar[b], (uint32_t)((c >> 1) & 1));
            tc_fence_after();
            drain_cols(tmem_lane + (uint32_t)(b * 2 * BN));
            drain_cols(tmem_lane + (uint32_t)(b * 2 * BN + BN));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&drained_bar[b]);
        }
        const int i = i0 + q * 32 + lane;
        if (i < p.I) {
            float* out = p.partial + ((long long)blockIdx.z * p.I + i) * p.J + j0 + half * HALF_COLS;
#pragma unroll
            for (int c = 0; c < HALF_COLS; c += 4)
                *reinterpret_cast<float4*>(out + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_base);
}

int wg_bn(int J) { return J % 128 == 0 ? 128 : (J % 64 == 0 ? 64 : 32); }

// position box for the reduction: divides the grid exactly (no partial boxes), whole images only when several are
// batched, at most kr_max positions; prefer full 8-row groups (no zero rows), then more positions, then wider boxes
void pick_kbox(int gw, int gh, int kr_max, int& bw, int& bh, int& bn) {
    double best = -1.0;
    bw = bh = bn = 1;
    for (int w = 1; w <= gw; ++w) {
        if (gw % w != 0) continue;
        for (int h = 1; h <= gh && w * h <= kr_max; ++h) {
            if (gh % h != 0) continue;
            const int nmax = (w == gw && h == gh) ? kr_max / (w * h) : 1;
            for (int n = 1; n <= nmax; ++n) {
                const int kr = w * h * n;
                const int krp = (kr + 7) / 8 * 8;
                const double score = (double)kr / krp + 0.25 * krp / (double)kr_max;
                if (score > best + 1e-9 || (score > best - 1e-9 && w > bw)) { best = score; bw = w; bh = h; bn = n; }
            }
        }
    }
}

template <int BN>
int32_t wg_launch(const CUtensorMap& bigmap, const CUtensorMap& smallmap, const Tc2WgParams& p, int splits, cudaStream_t stream) {
    const size_t smem = (size_t)p.krp * 128 * (2 * 4 + 2 * (BN / 32)) * kStages + 1024;
    CPB_REQUIRE(smem <= (size_t)kWgMaxDynSmem, "tc2_wgrad: stage too large (%zu bytes)", smem);
    dim3 grid((unsigned)cdiv(p.I, TBM), (unsigned)(p.J / BN), (unsigned)splits);
    tc2_wgrad_kernel<BN><<<grid, kThreads, smem, stream>>>(bigmap, smallmap, p);
    CPB_LAUNCHED();
    return CPB_OK;
}

template <int BN>
int32_t wg_init_one() {
    CPB_CUDA(cudaFuncSetAttribute(tc2_wgrad_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgMaxDynSmem));
    return CPB_OK;
}

}  // namespace

int32_t tc2_wgrad_init() {
    CPB_TRY(wg_init_one<32>());
    CPB_TRY(wg_init_one<64>());
    CPB_TRY(wg_init_one<128>());
    return CPB_OK;
}

bool tc2_wgrad_supported(const WgradParams& w) {
    if (!tc2_enabled()) return false;
    if (w.J % 32 != 0 || w.I < 128 || w.I % 32 != 0) return false;
    if (w.Ho == 1 && w.Wo == 1) return w.ntaps == 1 && w.I == w.run && w.big_pitch == w.I;   // dense: big [M, I], small [M, J]
    return w.sstride == 2 && w.big_pitch % 32 == 0 && w.run == w.ntaps * w.big_pitch && w.I == w.ntaps * w.run;
}

// splits of the one-wave split-K grid
int tc2_wgrad_pick_splits(int I, int J, long long nboxes) {
    const long long tiles = (long long)cdiv(I, TBM) * (J / wg_bn(J));
    long long splits = 148 / tiles;
    if (splits > nboxes) splits = nboxes;
    if (splits < 1) splits = 1;
    return (int)splits;
}

// geometry shared by the caller (partial buffer sizing) and the launcher
void tc2_wgrad_plan(const WgradParams& w, int& bw, int& bh, int& bn, long long& nboxes) {
    const int BN = wg_bn(w.J);
    // stage = KRp * 128 B * (8 + 2*BN/32) must fit 3x in ~200 KB
    const int kr_max = BN == 128 ? 32 : (BN == 64 ? 40 : 48);
    if (w.Ho == 1 && w.Wo == 1) { bw = bh = 1; bn = kr_max / 8 * 8; }
    else pick_kbox(w.Wo, w.Ho, kr_max, bw, bh, bn);
    nboxes = (long long)(w.Wo / bw) * (w.Ho / bh) * ((w.batch + bn - 1) / bn);
}

int32_t launch_tc2_wgrad(const WgradParams& w, cudaStream_t stream) {
    CPB_REQUIRE(tc2_wgrad_supported(w), "tc2_wgrad: unsupported problem (I=%d J=%d)", w.I, w.J);
    Tc2WgParams p;
    memset(&p, 0, sizeof(p));
    tc2_wgrad_plan(w, p.bw, p.bh, p.bn, p.nboxes);
    const int kr = p.bw * p.bh * p.bn;
    p.krp = (kr + 7) / 8 * 8;
    p.tiles_x = w.Wo / p.bw; p.tiles_y = w.Ho / p.bh; p.tiles_n = (w.batch + p.bn - 1) / p.bn;
    p.I = w.I; p.J = w.J; p.Cb = w.big_pitch; p.run = w.run;
    p.partial = w.partial;
    p.chunk_kb = 128 / p.krp > 0 ? 128 / p.krp : 1;
    if (p.chunk_kb < kStages) p.chunk_kb = kStages;            // late drain: the stage ring must not be deeper than a chunk
    p.debug = w.tc_variant;
    const int splits = w.splits;
    CPB_REQUIRE(splits >= 1, "tc2_wgrad: bad split count");
    p.boxes_per_split = (p.nboxes + splits - 1) / splits;
    const bool dense = w.Ho == 1 && w.Wo == 1;
    alignas(64) CUtensorMap bigmap, smallmap;
    {
        // big: dense NHWC {Cb, Wb, Hb, B}; Hb from the image size
        const unsigned long long Hb = dense ? 1ull : (unsigned long long)(w.big_img / ((long long)w.Wb * w.big_pitch));
        const unsigned long long dims[4] = {(unsigned long long)w.big_pitch, (unsigned long long)(dense ? 1 : w.Wb), Hb, (unsigned long long)w.batch};
        const unsigned long long strides[3] = {(unsigned long long)w.big_pitch * 4ull, (unsigned long long)(dense ? 1 : w.Wb) * w.big_pitch * 4ull, (unsigned long long)w.big_img * 4ull};
        const unsigned es = dense ? 1u : 2u;
        const unsigned box[4] = {32u, (unsigned)p.bw * es, (unsigned)p.bh * es, (unsigned)p.bn};
        const unsigned estr[4] = {1u, es, es, 1u};
        CPB_TRY(tc2_encode_tiled(&bigmap, w.big, 4, dims, strides, box, estr, 1));
    }
    {
        const unsigned long long dims[4] = {(unsigned long long)w.J, (unsigned long long)w.Wo, (unsigned long long)w.Ho, (unsigned long long)w.batch};
        const unsigned long long strides[3] = {(unsigned long long)w.J * 4ull, (unsigned long long)w.Wo * w.J * 4ull, (unsigned long long)w.Ho * w.Wo * w.J * 4ull};
        const unsigned box[4] = {32u, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
        const unsigned estr[4] = {1u, 1u, 1u, 1u};
        CPB_TRY(tc2_encode_tiled(&smallmap, w.small, 4, dims, strides, box, estr, 1));
    }
    switch (wg_bn(w.J)) {
        case 128: return wg_launch<128>(bigmap, smallmap, p, splits, stream);
        case 64: return wg_launch<64>(bigmap, smallmap, p, splits, stream);
        default: return wg_launch<32>(bigmap, smallmap, p, splits, stream);
    }
}

}  // namespace cpb
