// Tensor-core weight gradient with the A operand in TENSOR MEMORY (J <= 64: conv2, deconv3 -- the two largest):
//
//   gw[i, j] = sum_{positions (x, y, n)}  big[n, 2y + kh(i), 2x + kw(i), c(i)] * small[n, y, x, j]        i = (kh, kw, c)
//
// as D[128 x BN] += A B with A[i, k] = big(position k, channel i) and B[k, j] = small(position k, channel j), k = positions.
// Every SS-mode 3xTF32 kernel in this library is bound by the 128 B/clk of shared-memory bandwidth (operand staging +
// the tensor core's own operand reads; profiles/r2_cycle_accounting.md), and in the weight gradient the 128-channel A tile
// is two thirds of that traffic.  Here A never touches shared memory:
//   * an A loader thread owns ONE channel i (= its TMEM lane) and reads its value at the 32 positions of a k-block with 32
//     LDG.32 -- across a warp that is one coalesced 128-byte line per position (the 32 channels of a pixel are contiguous in
//     NHWC), so no transpose exists anywhere: the access pattern itself puts the reduction index on the TMEM columns;
//   * the 32 values (= a_hi, the tensor core ignores the 13 low mantissa bits) and their residuals a_lo = a - trunc_tf32(a)
//     go to the stage's 2 x 32 TMEM columns with two tcgen05.st;  the MMAs read A from TMEM (TS form);
//   * B (J <= 64 channels of `small`) is MN-major in shared memory: one TMA tensor-map box [32 positions x 32 floats] per
//     32-wide column group (SWIZZLE_128B_ATOM_32B, see tc2_wgrad.cu), b_lo derived in shared memory by two splitter warps.
// Per 32 positions the shared-memory traffic drops from ~26 KB (register path of tc_wgrad.cu, BN = 64) to ~12 KB and the LSU
// executes 128 LDG.32 warp instructions instead of 64 LDG.128 + 128 STS.128 + the 4x4 register transposes.
//
// Warp roles (512 threads): warps 0-7 A loaders + accumulator drain (set = warp / 4 takes the k-blocks with kb % 2 == set;
// both sets cover lane quarters 0-3), warp 8 MMA issuer, warp 9 B producer (TMA + L2 prefetch), warps 10-11 B splitters,
// warps 12-15 idle.  One wave of split-K CTAs; reduce_partials() sums the splits in a fixed order.
#include <cuda.h>

#include "tc2.cuh"
#include "wgrad.cuh"

namespace cpb {

namespace {

using namespace tc;

constexpr int kLoaderWarps = 8;
constexpr int kIssuerWarp = 8;
constexpr int kProducerWarp = 9;
constexpr int kSplitWarp0 = 10;
constexpr int kSplitWarps = 2;
constexpr int kThreads = 512;
constexpr int kStages = 4;                 // TMEM: 4 x 64 columns of A next to 4 * BN accumulator columns
constexpr int kKR = 32;                    // positions per k-block (= TMEM columns of one A plane)
constexpr int kChunkKb = 4;
constexpr int kPrefetchKb = 8;

struct Tc3WgParams {
    const float* big;
    int bw, bh, bn;                         // position box, bw * bh * bn == 32
    int tiles_x, tiles_y, tiles_n;
    long long boxes_per_split, nboxes;
    int I, J, batch;
    int Cb, run, Wb;                        // big channels, floats per kernel row (k * Cb), big image width
    long long big_img;
    float* partial;
    int debug;                              // 16: per-phase clock64 accounting of the A loaders (CTA 0)
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                 // SWIZZLE_128B_BASE32B: the only layout for MN-major 32-bit operands
    return d;
}
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
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
tc3_wgrad_kernel(const __grid_constant__ CUtensorMap smallmap, const __grid_constant__ Tc3WgParams p) {
    constexpr int NGB = BN / 32;
    constexpr int ACC_COLS = 4 * BN;                          // 2 chunk buffers x (main | cross)
    constexpr int TMEM_COLS = 512;                            // ACC_COLS (<= 256) + kStages * 64 = 512 at BN = 64
    static_assert(ACC_COLS + kStages * 64 <= TMEM_COLS, "tensor memory budget");
    constexpr uint32_t GROUP_BYTES = kKR * 128;               // one B column group: 32 positions x 128 B
    constexpr uint32_t STAGE_BYTES = GROUP_BYTES * 2 * NGB;   // [B_hi x NGB | B_lo x NGB]

    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t b_full[kStages], b_ready[kStages], a_ready[kStages], empty_bar[kStages];
    __shared__ uint64_t chunk_bar[2], drained_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ long long rel_off[kKR];      // window origin of box position q relative to the box origin (floats)
    __shared__ int rel_n[kKR];              // image index of box position q relative to the box origin

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid < kKR) {
        const int bwh = p.bw * p.bh;
        const int nn = tid / bwh, rr = tid - nn * bwh;
        const int yy = rr / p.bw, xx = rr - yy * p.bw;
        rel_n[tid] = nn;
        rel_off[tid] = (long long)nn * p.big_img + ((long long)(2 * yy) * p.Wb + 2 * xx) * p.Cb;
    }
    const int i0 = blockIdx.x * TBM;
    const int j0 = blockIdx.y * BN;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const long long kb_begin = (long long)blockIdx.z * p.boxes_per_split;
    long long kb_end = kb_begin + p.boxes_per_split;
    if (kb_end > p.nboxes) kb_end = p.nboxes;
    const int nkb = kb_end > kb_begin ? (int)(kb_end - kb_begin) : 0;
    const int nchunks = (nkb + kChunkKb - 1) / kChunkKb;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&b_full[s], 1); mbar_init(&b_ready[s], kSplitWarps); mbar_init(&a_ready[s], 4); mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&chunk_bar[0], 1); mbar_init(&chunk_bar[1], 1);
        mbar_init(&drained_bar[0], kLoaderWarps); mbar_init(&drained_bar[1], kLoaderWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<TMEM_COLS>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    // box index -> origin (x fastest)
    auto box_origin = [&](long long b, int& x0, int& y0, int& n0) {
        const int tx = (int)(b % p.tiles_x);
        const long long r = b / p.tiles_x;
        x0 = tx * p.bw; y0 = (int)(r % p.tiles_y) * p.bh; n0 = (int)(r / p.tiles_y) * p.bn;
    };

    if (warp < kLoaderWarps) {
        // ================================ A loaders (+ accumulator drain) ================================
        const int set = warp >> 2;                        // k-blocks kb % 2 == set
        const int q = warp & 3;                           // TMEM lane quarter
        const int i = i0 + q * 32 + lane;                 // this thread's channel
        const bool i_ok = i < p.I;
        long long coloff = 0;                             // float offset of (kh, kw, c) inside a window
        if (i_ok) {
            const int kh = i / p.run, off = i - kh * p.run;
            coloff = (long long)kh * p.Wb * p.Cb + off;
        }
        const float* src = p.big + coloff;
        // box origin -> float offset of its first window origin; position q adds rel_off[q] (valid while n0 + rel_n[q] < batch)
        auto box_base = [&](int x0, int y0, int n0) -> long long {
            return (long long)n0 * p.big_img + ((long long)(2 * y0) * p.Wb + 2 * x0) * p.Cb;
        };

        // ---- drain state (all 8 warps drain: quarter q x column half `set`)
        constexpr int HALF_COLS = BN / 2;
        const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * HALF_COLS);
        float acc[HALF_COLS];
#pragma unroll
        for (int c = 0; c < HALF_COLS; ++c) acc[c] = 0.f;
        int drained = 0;
        auto drain_cols = [&](uint32_t taddr) {
#pragma unroll
            for (int cc = 0; cc < HALF_COLS; cc += 16) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)cc, v);
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[cc + c] += v[c];
            }
        };
        auto drain_one = [&]() {
            const int b = drained & 1;
            mbar_wait(&chunk_bar[b], (uint32_t)((drained >> 1) & 1));
            tc_fence_after();
            drain_cols(tmem_acc + (uint32_t)(b * 2 * BN));
            drain_cols(tmem_acc + (uint32_t)(b * 2 * BN + BN));
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&drained_bar[b]);
            ++drained;
        };

        long long prof[6] = {0, 0, 0, 0, 0, 0}, tlast = clock64();
        const long long tstart = tlast;
        const bool PROF = (p.debug & 16) != 0;
#define TC3_PROF(slot) do { if (PROF) { const long long now_ = clock64(); prof[slot] += now_ - tlast; tlast = now_; } } while (0)
        for (int kb = set; kb < nkb; kb += 2) {
            const int s = kb % kStages;
            TC3_PROF(5);
            // chunks finished two k-block pairs ago are drained while this k-block's loads are in flight
            int x0, y0, n0;
            box_origin(kb_begin + kb, x0, y0, n0);
            float v[kKR];
            {
                const float* bp = src + box_base(x0, y0, n0);
                const int nleft = p.batch - n0;
#pragma unroll
                for (int qq = 0; qq < kKR; ++qq) v[qq] = (i_ok && rel_n[qq] < nleft) ? __ldg(bp + rel_off[qq]) : 0.f;
            }
            TC3_PROF(0);
            // L2 prefetch of the k-block kPrefetchKb ahead (one 128-byte line per position and warp)
            if (kb + kPrefetchKb < nkb) {
                int px, py, pn;
                box_origin(kb_begin + kb + kPrefetchKb, px, py, pn);
                const float* bp = src + box_base(px, py, pn);
                const int nleft = p.batch - pn;
#pragma unroll 8
                for (int qq = 0; qq < kKR; ++qq)
                    if (i_ok && rel_n[qq] < nleft) asm volatile("prefetch.global.L2 [%0];" ::"l"(bp + rel_off[qq]) : "memory");
            }
            TC3_PROF(1);
            while (drained < kb / kChunkKb - 1) drain_one();
            TC3_PROF(2);
            if (kb >= kStages) mbar_wait(&empty_bar[s], (uint32_t)((kb / kStages - 1) & 1));     // the MMAs that read this TMEM slot are done
            TC3_PROF(3);
            tc_fence_after();
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ACC_COLS + s * 64);
            tmem_st32(ta, v);                              // a_hi = raw values
#pragma unroll
            for (int qq = 0; qq < kKR; ++qq) { float h, l; split_tf32(v[qq], h, l); v[qq] = l; }
            tmem_st32(ta + 32u, v);                        // a_lo
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[s]);
            TC3_PROF(4);
        }
        if (PROF && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (warp == 0 || warp == 4))
            printf("tc3prof warp %d nkb %d total %lld: issue_loads %lld prefetch %lld drain %lld wait_empty %lld wait_data+st %lld other %lld\n",
                   warp, nkb, clock64() - tstart, prof[0], prof[1], prof[2], prof[3], prof[4], prof[5]);
        while (drained < nchunks) drain_one();
        // ---- partial[split][i][j]
        if (i_ok) {
            float* out = p.partial + ((long long)blockIdx.z * p.I + i) * p.J + j0 + set * HALF_COLS;
#pragma unroll
            for (int c = 0; c < HALF_COLS; c += 4)
                *reinterpret_cast<float4*>(out + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        }
    } else if (warp == kProducerWarp) {
        // ================================ B producer ================================
        if (lane == 0 && nkb > 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&smallmap)) : "memory");
            auto prefetch_b = [&](int kb) {
                if (kb >= nkb) return;
                int x0, y0, n0;
                box_origin(kb_begin + kb, x0, y0, n0);
#pragma unroll
                for (int g = 0; g < NGB; ++g)
                    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                                 ::"l"(reinterpret_cast<uint64_t>(&smallmap)), "r"(j0 + 32 * g), "r"(x0), "r"(y0), "r"(n0) : "memory");
            };
            for (int kb = 0; kb < kPrefetchKb; ++kb) prefetch_b(kb);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                prefetch_b(kb + kPrefetchKb);
                if (kb >= kStages) mbar_wait(&empty_bar[s], (uint32_t)((kb / kStages - 1) & 1));
                const uint32_t stage = smem_base + (uint32_t)s * STAGE_BYTES;
                mbar_expect_tx(&b_full[s], GROUP_BYTES * NGB);
                int x0, y0, n0;
                box_origin(kb_begin + kb, x0, y0, n0);
#pragma unroll
                for (int g = 0; g < NGB; ++g)
                    tma_load_4d(stage + (uint32_t)g * GROUP_BYTES, &smallmap, j0 + 32 * g, x0, y0, n0, &b_full[s]);
            }
        }
        __syncwarp();
    } else if (warp >= kSplitWarp0 && warp < kSplitWarp0 + kSplitWarps) {
        // ================================ B splitters: b_lo = b - trunc_tf32(b) ================================
        const int tl = tid - kSplitWarp0 * 32;
        constexpr uint32_t CHUNKS = NGB * GROUP_BYTES / 16;
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kStages;
            mbar_wait(&b_full[s], (uint32_t)((kb / kStages) & 1));
            const uint32_t stage = smem_base + (uint32_t)s * STAGE_BYTES;
#pragma unroll 4
            for (uint32_t c = (uint32_t)tl; c < CHUNKS; c += kSplitWarps * 32) {
                const uint32_t a = stage + c * 16u;
                float4 v, h, l;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + NGB * GROUP_BYTES), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_ready[s]);
        }
    } else if (warp == kIssuerWarp) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            // A: K-major by construction (TMEM lane = row, column = k); B: MN-major shared memory (bit 16)
            const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(TBM >> 4) << 24);
            const uint32_t idesc = idesc_base | ((uint32_t)(BN >> 3) << 17);
            const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * BN) >> 3) << 17);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                const uint32_t stage = smem_base + (uint32_t)s * STAGE_BYTES;
                const int chunk = kb / kChunkKb;
                const int b = chunk & 1;
                const uint32_t par = (uint32_t)((kb / kStages) & 1);
                mbar_wait(&a_ready[s], par);
                mbar_wait(&b_ready[s], par);
                if (kb % kChunkKb == 0 && chunk >= 2) mbar_wait(&drained_bar[b], (uint32_t)(((chunk >> 1) - 1) & 1));
                tc_fence_after();
                const uint32_t d_main = tmem_base + (uint32_t)(b * 2 * BN);
                const uint32_t d_cross = d_main + (uint32_t)BN;
                const uint32_t ta = tmem_base + (uint32_t)(ACC_COLS + s * 64);
#pragma unroll
                for (int ks = 0; ks < kKR / 8; ++ks) {
                    const uint64_t b_hi = make_desc_mn(stage + (uint32_t)ks * 1024u, GROUP_BYTES, 512u);     // [b_hi | b_lo] column groups adjacent
                    umma_tf32_ts(d_main, ta + (uint32_t)(ks * 8), b_hi, idesc2, ((kb % kChunkKb) | ks) != 0 ? 1u : 0u);
                    umma_tf32_ts(d_cross, ta + 32u + (uint32_t)(ks * 8), b_hi, idesc, 1u);
                }
                umma_commit(&empty_bar[s]);
                if (kb % kChunkKb == kChunkKb - 1 || kb == nkb - 1) umma_commit(&chunk_bar[b]);
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// position box with exactly 32 positions that divides the grid in x and y
bool pick_box32(int gw, int gh, int& bw, int& bh, int& bn) {
    int best = -1;
    for (int w = 1; w <= gw && w <= 32; ++w) {
        if (gw % w != 0 || 32 % w != 0) continue;
        for (int h = 1; h <= gh && w * h <= 32; ++h) {
            if (gh % h != 0 || 32 % (w * h) != 0) continue;
            const int score = w * 64 + h;            // wide first (contiguous TMA rows of `small`), then tall
            if (score > best) { best = score; bw = w; bh = h; bn = 32 / (w * h); }
        }
    }
    return best >= 0;
}

template <int BN>
int32_t launch_bn(const CUtensorMap& smallmap, const Tc3WgParams& p, int splits, cudaStream_t stream) {
    const size_t smem = (size_t)kKR * 128 * 2 * (BN / 32) * kStages + 1024;
    dim3 grid((unsigned)cdiv(p.I, TBM), (unsigned)(p.J / BN), (unsigned)splits);
    tc3_wgrad_kernel<BN><<<grid, kThreads, smem, stream>>>(smallmap, p);
    CPB_LAUNCHED();
    return CPB_OK;
}

int g_tc3_enabled = 0;
int g_tc3_probe = 0;      // set once the kernel attributes are configured: the debug entry may use the kernel

}  // namespace

int32_t tc3_wgrad_init() {
    // Opt-in (CPB_TC3_WGRAD=1): parity-green (tests/test_tc_gpu.py runs it through cpb_debug_tc_wgrad) but measured 3x SLOWER
    // than the register-path kernel at B=4096 (deconv3.wgrad 10.1 ms vs 3.3 ms): with one k-block of 32 dependent-latency
    // LDG.32 per loader warp in flight the A side delivers a k-block every ~4.5 k clk against ~0.7 k clk of MMAs.  The TS-form
    // MMA itself is fine (tc2_tapgemm.cu uses it for BN <= 64); what this kernel lacks is a deeper A pipeline.
    const char* e = getenv("CPB_TC3_WGRAD");
    g_tc3_enabled = e ? atoi(e) : 0;
    g_tc3_probe = 1;
    CPB_CUDA(cudaFuncSetAttribute(tc3_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + 1024));
    CPB_CUDA(cudaFuncSetAttribute(tc3_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + 1024));
    return CPB_OK;
}

static bool tc3_shape_ok(const WgradParams& w) {
    if (!tc2_enabled()) return false;
    if (!(w.J == 32 || w.J == 64) || w.I < 128 || w.I % 32 != 0) return false;
    int bw, bh, bn;
    if (w.Ho == 1 && w.Wo == 1) return w.ntaps == 1 && w.I == w.run && w.big_pitch == w.I;
    return w.sstride == 2 && w.big_pitch % 32 == 0 && w.run == w.ntaps * w.big_pitch && w.I == w.ntaps * w.run && pick_box32(w.Wo, w.Ho, bw, bh, bn);
}

bool tc3_wgrad_supported(const WgradParams& w) { return tc2_enabled() && g_tc3_enabled && tc3_shape_ok(w); }
bool tc3_wgrad_available(const WgradParams& w) { return tc2_enabled() && g_tc3_probe && tc3_shape_ok(w); }      // unit tests / debug entry

long long tc3_wgrad_boxes(const WgradParams& w) {
    int bw = 1, bh = 1, bn = 32;
    if (!(w.Ho == 1 && w.Wo == 1)) pick_box32(w.Wo, w.Ho, bw, bh, bn);
    return (long long)(w.Wo / bw) * (w.Ho / bh) * ((w.batch + bn - 1) / bn);
}

int32_t launch_tc3_wgrad(const WgradParams& w, cudaStream_t stream) {
    CPB_REQUIRE(tc3_wgrad_available(w), "tc3_wgrad: unsupported problem (I=%d J=%d)", w.I, w.J);
    CPB_REQUIRE((long long)w.batch * w.big_img < (1ll << 40), "tc3_wgrad: tensor too large");
    Tc3WgParams p;
    memset(&p, 0, sizeof(p));
    const bool dense = w.Ho == 1 && w.Wo == 1;
    p.bw = 1; p.bh = 1; p.bn = 32;
    if (!dense) pick_box32(w.Wo, w.Ho, p.bw, p.bh, p.bn);
    p.tiles_x = w.Wo / p.bw; p.tiles_y = w.Ho / p.bh; p.tiles_n = (w.batch + p.bn - 1) / p.bn;
    p.nboxes = (long long)p.tiles_x * p.tiles_y * p.tiles_n;
    p.boxes_per_split = (p.nboxes + w.splits - 1) / w.splits;
    p.big = w.big; p.I = w.I; p.J = w.J; p.batch = w.batch;
    p.Cb = w.big_pitch; p.run = w.run; p.Wb = dense ? 0 : w.Wb; p.big_img = w.big_img;
    p.partial = w.partial;
    p.debug = w.tc_variant;
    alignas(64) CUtensorMap smallmap;
    const unsigned long long dims[4] = {(unsigned long long)w.J, (unsigned long long)w.Wo, (unsigned long long)w.Ho, (unsigned long long)w.batch};
    const unsigned long long strides[3] = {(unsigned long long)w.J * 4ull, (unsigned long long)w.Wo * w.J * 4ull, (unsigned long long)w.Ho * w.Wo * w.J * 4ull};
    const unsigned box[4] = {32u, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
    const unsigned estr[4] = {1u, 1u, 1u, 1u};
    CPB_TRY(tc2_encode_tiled(&smallmap, w.small, 4, dims, strides, box, estr, 1));
    return w.J == 64 ? launch_bn<64>(smallmap, p, w.splits, stream) : launch_bn<32>(smallmap, p, w.splits, stream);
}

}  // namespace cpb
