// Tensor-core tap-GEMM for sm_100a: the same contraction as tapgemm.cu, computed with tcgen05.mma
// (kind::tf32, fp32 accumulators in TMEM) using the error-compensated 3xTF32 split
//
//     a*b  ~=  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo ,   x_hi = x with the 13 low mantissa bits cleared,  x_lo = x - x_hi
//
// so that results stay fp32-accurate (relative error ~1e-6; a single TF32 pass gives ~3e-4 and would break the
// 1e-5 parity bar).  Operands are staged in shared memory in the UMMA canonical K-major SWIZZLE_128B layout
// (rows of 32 floats = 128 B, 8-row groups of 1024 B, 16-byte chunk index XOR row%8):
//   * A (activations, gathered rows) : LDG.128 -> registers -> hi/lo split -> 2 x STS.128 (swizzled)
//   * B (weights)                    : pre-split, K-major copies written once per step by tc_weights_kernel,
//                                      copied with cp.async (no register staging)
// One elected thread issues 12 MMAs per 32-wide k-block (4 k-steps x 3 products); tcgen05.commit arrives on the
// stage's mbarrier when they retire, which frees the stage for the loaders.
//
// Accumulation.  The tensor core adds into its fp32 accumulator with round-toward-zero, which shrinks a long
// running sum systematically (measured on B200: relative bias -6.5e-9 x K, i.e. -7e-6 at K=1024, -2.6e-5 at
// K=4096; scripts/diag_tc.py).  Two measures bring this back to fp32-FMA level:
//   * the two cross terms (2^-11 of the main term) accumulate in their OWN TMEM tile, so they no longer
//     re-truncate the large accumulator twice per k-step;
//   * the main term is accumulated in TMEM only over chunks of 128 k (4 k-blocks, 16 MMAs) into two ping-pong
//     tiles; each finished chunk is drained with tcgen05.ld and added to per-thread fp32 REGISTER accumulators
//     (round-to-nearest).  The drain of chunk j is issued one chunk late, when its MMAs have long retired, so it
//     never stalls the loaders.
// The register accumulators feed the bias / ReLU / ReLU-mask epilogue directly.
#include "tapgemm.cuh"
#include "tc_common.cuh"

namespace cpb {

namespace {

using namespace tc;

template <int BN>
struct TcCfg {
    static constexpr int B_TILE_BYTES = BN * TBK * 4;
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int STAGES = (STAGE_BYTES * 4 <= 200 * 1024) ? 4 : ((STAGE_BYTES * 3 <= 200 * 1024) ? 3 : 2);
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;   // +1024: manual 1 KB alignment
    static constexpr int TMEM_COLS = BN == 128 ? 512 : (BN == 64 ? 256 : 128);   // 2 main tiles + 1 cross tile, power of 2
};
constexpr int CHUNK_KB = 4;   // k-blocks accumulated inside TMEM before draining to registers (must be >= STAGES)

constexpr int kLoaderThreads = 256;   // warps 0-7: loaders, accumulator drain, epilogue
constexpr int kTcThreads = 288;       // + warp 8: MMA issuer (one elected lane)

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_tapgemm_kernel(const __grid_constant__ TapGemmParams p) {
    using Cfg = TcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int B_TILE_BYTES = Cfg::B_TILE_BYTES;
    constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr int B_CHUNKS = BN * 8;                 // 16-byte chunks per B tile
    constexpr int B_ITERS = (B_CHUNKS + 255) / 256;
    static_assert(STAGES <= CHUNK_KB, "late drain relies on the stage ring being no deeper than a chunk");

    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[STAGES];     // loaders -> issuer: stage holds A(hi,lo) and B(hi,lo) of a k-block
    __shared__ uint64_t empty_bar[STAGES];    // tensor core -> loaders: the MMAs reading the stage retired
    __shared__ uint64_t chunk_bar[2];         // tensor core -> loaders: main tile b holds a finished 128-k chunk
    __shared__ uint64_t drained_bar[2];       // loaders -> issuer: main tile b was added to the register accumulators
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const TapClass& cls = p.cls[blockIdx.z];
    const int Wo = cls.Wo;
    const int HoWo = cls.Ho * Wo;
    const long long M = (long long)p.batch * HoWo;
    const long long m0 = (long long)blockIdx.x * TBM;
    if (m0 >= M) return;                              // uniform per CTA: safe before any barrier / alloc
    const int n0 = blockIdx.y * BN;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], kLoaderThreads); mbar_init(&empty_bar[s], 1); }
        mbar_init(&chunk_bar[0], 1); mbar_init(&chunk_bar[1], 1);
        mbar_init(&drained_bar[0], kLoaderThreads); mbar_init(&drained_bar[1], kLoaderThreads);
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<Cfg::TMEM_COLS>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const int nkb = cls.ntaps * (p.C / TBK);
    const int nchunks = (nkb + CHUNK_KB - 1) / CHUNK_KB;

    if (warp == 8) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
            const uint32_t d_cross = tmem_base + (uint32_t)(2 * BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t stage = smem_base + s * STAGE_BYTES;
                const int chunk = kb / CHUNK_KB;
                mbar_wait(&full_bar[s], (uint32_t)((kb / STAGES) & 1));
                if (kb % CHUNK_KB == 0 && chunk >= 2)      // main tile (chunk & 1) must have been drained (chunk - 2)
                    mbar_wait(&drained_bar[chunk & 1], (uint32_t)(((chunk >> 1) - 1) & 1));
                tc_fence_after();
                const uint64_t a_hi = make_desc(stage);
                const uint64_t a_lo = make_desc(stage + A_TILE_BYTES);
                const uint64_t b_hi = make_desc(stage + 2 * A_TILE_BYTES);
                const uint64_t b_lo = make_desc(stage + 2 * A_TILE_BYTES + B_TILE_BYTES);
                const uint32_t d_main = tmem_base + (uint32_t)((chunk & 1) * BN);
#pragma unroll
                for (int ks = 0; ks < TBK / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 2);      // 32 bytes per k-step, in 16-byte units
                    umma_tf32(d_main, a_hi + adv, b_hi + adv, idesc, ((kb % CHUNK_KB) | ks) != 0 ? 1u : 0u);
                    umma_tf32(d_cross, a_lo + adv, b_hi + adv, idesc, (kb | ks) != 0 ? 1u : 0u);
                    umma_tf32(d_cross, a_hi + adv, b_lo + adv, idesc, 1u);
                }
                umma_commit(&empty_bar[s]);
                if (kb % CHUNK_KB == CHUNK_KB - 1 || kb == nkb - 1) umma_commit(&chunk_bar[chunk & 1]);
                if (kb == nkb - 1) umma_commit(&done_bar);
            }
        }
        __syncwarp();
    } else {
        // ================================ loaders / drain / epilogue ================================
        // A rows: 8 threads cover the 128 bytes of one row, 32 rows per pass, 4 passes
        const int a_chunk = tid & 7;
        long long a_base[4];
        int a_iy[4], a_ix[4];
        bool a_ok[4];
        uint32_t a_soff[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = (tid >> 3) + i * 32;
            const long long m = m0 + r;
            a_ok[i] = m < M;
            const long long mm = a_ok[i] ? m : 0;
            const int n = (int)(mm / HoWo);
            const int rem = (int)(mm - (long long)n * HoWo);
            const int oy = rem / Wo;
            const int ox = rem - oy * Wo;
            a_iy[i] = oy * p.sstride;
            a_ix[i] = ox * p.sstride;
            a_base[i] = (long long)n * p.src_img + ((long long)a_iy[i] * p.Ws + a_ix[i]) * p.src_pitch + a_chunk * 4;
            a_soff[i] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((a_chunk ^ (r & 7)) << 4));
        }
        auto load_a = [&](int tap_idx, int c0, float4* regs) {
            const Tap& t = cls.taps[tap_idx];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bool v = a_ok[i];
                if (p.check) v = v && (unsigned)(a_iy[i] + t.dy) < (unsigned)p.Hs && (unsigned)(a_ix[i] + t.dx) < (unsigned)p.Ws;
                regs[i] = v ? __ldg(reinterpret_cast<const float4*>(p.src + a_base[i] + t.src_off + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // B tiles (pre-split K-major weights): cp.async straight into the swizzled layout, one k-block ahead of A
        int b_tap = 0, b_c = 0;
        auto issue_b = [&](int kbt) {
            const uint32_t stage = smem_base + (kbt % STAGES) * STAGE_BYTES;
            const Tap& t = cls.taps[b_tap];
            const long long woff = t.w_off + (long long)n0 * p.C + b_c;
#pragma unroll
            for (int it = 0; it < B_ITERS; ++it) {
                const int f = tid + it * 256;
                if (B_CHUNKS % 256 == 0 || f < B_CHUNKS) {
                    const int n = f >> 3, c = f & 7;
                    const uint32_t so = (uint32_t)((n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4));
                    const long long go = woff + (long long)n * p.C + c * 4;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage + 2 * A_TILE_BYTES + so), "l"(p.wk_hi + go));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage + 2 * A_TILE_BYTES + B_TILE_BYTES + so), "l"(p.wk_lo + go));
                }
            }
            b_c += TBK;
            if (b_c == p.C) { b_c = 0; ++b_tap; }
        };

        // register accumulators: this thread owns row (q*32 + lane) x columns [half*BN/2, +BN/2)
        constexpr int HALF_COLS = BN / 2;
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int half = warp >> 2;                  // column half handled by this warp
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * HALF_COLS);
        float acc[HALF_COLS];
#pragma unroll
        for (int i = 0; i < HALF_COLS; ++i) acc[i] = 0.f;
        int drained = 0;
        auto drain_one = [&]() {
            const int b = drained & 1;
            mbar_wait(&chunk_bar[b], (uint32_t)((drained >> 1) & 1));
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < HALF_COLS; cc += 16) {
                float v[16];
                tmem_ld16(tmem_lane + (uint32_t)(b * BN + cc), v);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[cc + i] += v[i];
            }
            tc_fence_before();
            mbar_arrive(&drained_bar[b]);
            ++drained;
        };

        float4 areg[4];
        int ld_tap = 0, ld_c = 0;
        load_a(0, 0, areg);
        issue_b(0);
        cp_async_commit();

        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t stage = smem_base + s * STAGE_BYTES;
            // B(kb+1) into the next stage (free once the MMAs of k-block kb+1-STAGES retired); stage s itself was
            // claimed the same way one iteration ago.
            if (kb + 1 < nkb) {
                if (kb + 1 >= STAGES) mbar_wait(&empty_bar[(kb + 1) % STAGES], (uint32_t)(((kb + 1) / STAGES - 1) & 1));
                issue_b(kb + 1);
            }
            cp_async_commit();
            // late drain: chunk j-2 retired long ago (the stage ring is no deeper than a chunk), never blocks
            if (kb % CHUNK_KB == 0) {
                while (drained < kb / CHUNK_KB - 1) drain_one();
            }
            // A tile: split the prefetched fp32 rows into hi / lo and store both (swizzled)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 x = areg[i];
                float4 hi, lo;
                split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage + a_soff[i]), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage + A_TILE_BYTES + a_soff[i]), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
            }
            // prefetch the next k-block's A rows
            ld_c += TBK;
            if (ld_c == p.C) { ld_c = 0; ++ld_tap; }
            if (kb + 1 < nkb) load_a(ld_tap, ld_c, areg);

            cp_async_wait<1>();          // everything but the group just committed: this thread's B(kb) chunks landed
            fence_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(&full_bar[s]);
        }

        // drain what is still in TMEM (last one or two chunks, then the cross-term tile)
        while (drained < nchunks) drain_one();
        mbar_wait(&done_bar, 0);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < HALF_COLS; cc += 16) {
            float v[16];
            tmem_ld16(tmem_lane + (uint32_t)(2 * BN + cc), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[cc + i] += v[i];
        }
        // epilogue from registers: bias / ReLU / mask -> global
        const int row = q * 32 + lane;
        const long long m = m0 + row;
        if (m < M) {
            const int n = (int)(m / HoWo);
            const int rem = (int)(m - (long long)n * HoWo);
            const int oy = rem / Wo;
            const int ox = rem - oy * Wo;
            const int col0 = n0 + half * HALF_COLS;
            const long long img = (long long)n * p.dst_img;
            const long long off_plain = img + ((long long)(oy * p.dstride + cls.py) * p.Wd + (ox * p.dstride + cls.px)) * p.dst_pitch;
            // groups of 4 columns, handled 4 at a time so that the ReLU-mask loads of a batch are all in flight
            // together (one at a time they serialise 16 global round trips per thread)
            constexpr int NG = HALF_COLS / 4;
            constexpr int GB = NG < 4 ? NG : 4;
#pragma unroll
            for (int g0 = 0; g0 < NG; g0 += GB) {
                long long offs[GB];
                int chs[GB];
                bool oks[GB];
                float4 mks[GB];
#pragma unroll
                for (int u = 0; u < GB; ++u) {
                    const int col = col0 + (g0 + u) * 4;
                    oks[u] = true;
                    if (p.quad) {
                        const int c = col / p.quad_cb;
                        chs[u] = col - c * p.quad_cb;
                        const int y = oy * 2 + (c >> 1), x = ox * 2 + (c & 1);
                        oks[u] = y < p.Hd && x < p.Wd;
                        offs[u] = img + ((long long)y * p.Wd + x) * p.dst_pitch + chs[u];
                    } else {
                        chs[u] = col;
                        offs[u] = off_plain + col;
                    }
                }
                if (p.mask) {
#pragma unroll
                    for (int u = 0; u < GB; ++u)
                        mks[u] = oks[u] ? __ldg(reinterpret_cast<const float4*>(p.mask + offs[u])) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < GB; ++u) {
                    if (!oks[u]) continue;
                    const int g = g0 + u;
                    float4 o = make_float4(acc[g * 4 + 0], acc[g * 4 + 1], acc[g * 4 + 2], acc[g * 4 + 3]);
                    if (p.bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + chs[u]));
                        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                    }
                    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (p.mask) {
                        o.x = mks[u].x > 0.f ? o.x : 0.f; o.y = mks[u].y > 0.f ? o.y : 0.f;
                        o.z = mks[u].z > 0.f ? o.z : 0.f; o.w = mks[u].w > 0.f ? o.w : 0.f;
                    }
                    *reinterpret_cast<float4*>(p.dst + offs[u]) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

template <int BN>
int32_t tc_launch(const TapGemmParams& p, cudaStream_t stream) {
    using Cfg = TcCfg<BN>;
    long long max_m = 0;
    for (int c = 0; c < p.nclass; ++c) {
        long long m = (long long)p.batch * p.cls[c].Ho * p.cls[c].Wo;
        if (m > max_m) max_m = m;
    }
    if (max_m == 0) return CPB_OK;
    dim3 grid((unsigned)((max_m + TBM - 1) / TBM), (unsigned)(p.N / BN), (unsigned)p.nclass);
    tc_tapgemm_kernel<BN><<<grid, kTcThreads, Cfg::SMEM_BYTES, stream>>>(p);
    CPB_LAUNCHED();
    return CPB_OK;
}

template <int BN>
int32_t tc_init_one() {
    CPB_CUDA(cudaFuncSetAttribute(tc_tapgemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::SMEM_BYTES));
    return CPB_OK;
}

// weight preparation: per tap, a K-major [N][C] block, split into hi / lo
__global__ void tc_weights_kernel(const float* __restrict__ params, float* __restrict__ dst, const __grid_constant__ TcWeightTable t) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= t.total) return;
    int j = 0;
    while (j < t.njobs - 1 && idx >= t.jobs[j].count) { idx -= t.jobs[j].count; ++j; }
    const TcWeightJob& job = t.jobs[j];
    float x;
    if (job.mode == 0) {
        x = params[job.src_off + idx];
    } else if (job.mode == 2) {
        // quad scatter form: destination [j][i][class*Cb + cb][cs]; class (py,px) uses kernel tap (py+2j, px+2i)
        const int w = (job.k + 1) / 2;
        const int cs = (int)(idx % job.cs);
        long long rest = idx / job.cs;
        const int ncol = (int)(rest % (4 * job.cb));
        rest /= (4 * job.cb);
        const int i = (int)(rest % w), j = (int)(rest / w);
        const int cls = ncol / job.cb, cb = ncol - cls * job.cb;
        const int kh = (cls >> 1) + 2 * j, kw = (cls & 1) + 2 * i;
        x = (kh < job.k && kw < job.k) ? params[job.src_off + (((long long)kh * job.k + kw) * job.cb + cb) * job.cs + cs] : 0.f;
    } else {
        // gather form: destination [kh][cs][kw*Cb + cb]  <-  source [kh][kw][cb][cs]
        const int run = job.k * job.cb;
        const int c = (int)(idx % run);
        const long long rest = idx / run;
        const int cs = (int)(rest % job.cs);
        const int kh = (int)(rest / job.cs);
        const int kw = c / job.cb, cb = c - kw * job.cb;
        x = params[job.src_off + (((long long)kh * job.k + kw) * job.cb + cb) * job.cs + cs];
    }
    const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    dst[job.dst_hi + idx] = hi;
    dst[job.dst_lo + idx] = x - hi;
}

}  // namespace

int32_t tc_tapgemm_init() {
    CPB_TRY(tc_init_one<32>());
    CPB_TRY(tc_init_one<64>());
    CPB_TRY(tc_init_one<128>());
    return CPB_OK;
}

bool tc_tapgemm_supported(const TapGemmParams& p) {
    if (p.quad && (p.N != 4 * p.quad_cb || p.nclass != 1)) return false;
    return p.ybatch == 1 && p.C % TBK == 0 && (p.N == 32 || p.N % 64 == 0) && p.wk_hi != nullptr && p.wk_lo != nullptr;
}

int32_t launch_tc_tapgemm(const TapGemmParams& p, cudaStream_t stream) {
    CPB_REQUIRE(tc_tapgemm_supported(p), "tc_tapgemm: unsupported problem (C=%d, N=%d)", p.C, p.N);
    if (p.N % 128 == 0) return tc_launch<128>(p, stream);
    if (p.N % 64 == 0) return tc_launch<64>(p, stream);
    return tc_launch<32>(p, stream);
}

int32_t launch_tc_weights(const float* params, float* dst, const TcWeightTable& table, cudaStream_t stream) {
    if (table.total == 0) return CPB_OK;
    tc_weights_kernel<<<cdiv(table.total, 256), 256, 0, stream>>>(params, dst, table);
    CPB_LAUNCHED();
    return CPB_OK;
}

}  // namespace cpb
