// Tensor-core weight gradient with the A operand in TENSOR MEMORY (J <= 64: conv2, deconv3 -- the two largest):
//
//   gw[i, j] = sum_{positions (x, y, n)}  big[n, 2y + kh(i), 2x + kw(i), c(i)] * small[n, y, x, j]        i = (kh, kw, c)
//
// as D[128 x BN] += A B with A[i, k] = big(position k, channel i) and B[k, j] = small(position k, channel j), k = positions.
// Every SS-mode 3xTF32 kernel in this library is bound by the 128 B/clk of shared-memory bandwidth (operand staging +
// the tensor core's own operand reads; profiles/r2_cycle_accounting.md), and in the weight gradient the 128-channel A tile
// is two thirds of that traffic.  Here A never touches shared memory:
//   * the copy engine brings the k-block's A values to shared memory as they lie in NHWC memory: per 32 channels one TMA
//     tensor-map box [32 positions x 32 floats] (128-byte rows, no swizzle, stride-2 traversal of the window origins,
//     images past the batch zero-filled) -- 4 boxes = 16 KB per k-block, written once and read once;
//   * an A loader thread owns ONE channel i (= its TMEM lane) and picks its value at the 32 positions with 32 LDS.32 --
//     across a warp one conflict-free 128-byte row per position, so no transpose exists anywhere: the access pattern itself
//     puts the reduction index on the TMEM columns.  (The first version read these values with LDG.32 straight from global
//     memory: 128 scattered 128-byte requests per warp and k-block saturated the LSU's miss path -- ~2.2 k clk per k-block
//     just to ISSUE them, `tc3prof` in profiles/r2_cycle_accounting.md -- 3x slower than the register-transpose kernel.)
//   * the 32 values (= a_hi, the tensor core ignores the 13 low mantissa bits) and their residuals a_lo = a - trunc_tf32(a)
//     go to the stage's 2 x 32 TMEM columns with two tcgen05.st;  the MMAs read A from TMEM (TS form);
//   * B (J <= 64 channels of `small`) is MN-major in shared memory: one TMA tensor-map box [32 positions x 32 floats] per
//     32-wide column group (SWIZZLE_128B_ATOM_32B, see tc2_wgrad.cu), b_lo derived in shared memory by two splitter warps.
// Shared-memory traffic per 32 positions (BN = 64): A 16 KB written + 16 KB read, B 8 + 8 + 8 KB staged and 24 KB of MMA
// operand reads = 80 KB against ~104 KB of the register path, and the LSU executes 128 LDS.32 warp instructions instead of
// 64 LDG.128 + 128 STS.128 + the 4x4 register transposes.
//
// Warp roles (512 threads): warps 0-7 A loaders (set = warp / 4 takes the k-blocks with kb % 2 == set; both sets cover lane
// quarters 0-3), warp 8 MMA issuer, warp 9 producer (one lane: the B boxes and the A run boxes of every k-block), warps 10-11
// B splitters, warps 12-15 accumulator drain (lane quarters 0-3).  Rings: A 6 stages in shared memory + 4 slots in tensor
// memory, B 6 stages -- deep enough to cover the HBM latency of operands that stream exactly once, so there is no separate
// L2 prefetch (neither cp.async.bulk.prefetch.tensor nor prefetch.global.L2 changed the time once the rings were this deep).
// One wave of split-K CTAs; reduce_partials() sums the splits in a fixed order.
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
constexpr int kDrainWarp0 = 12;
constexpr int kDrainWarps = 4;
constexpr int kThreads = 512;
constexpr int kStages = 4;                 // TMEM: 4 x 64 columns of A next to 4 * BN accumulator columns
constexpr int kKR = 32;                    // positions per k-block (= TMEM columns of one A plane)
constexpr int kChunkKb = 4;
constexpr int kPrefetchKb = 10;            // k-blocks the L2 prefetch (drain warps, between chunks) runs ahead of the producer

struct Tc3WgParams {
    const float* big;
    int bw, bh, bn;                         // position box, bw * bh * bn == 32
    int tiles_x, tiles_y, tiles_n;
    long long boxes_per_split, nboxes;
    int I, J, batch;
    int Cb, run, Wb;                        // big channels, floats per kernel row (k * Cb), big image width
    int sx;                                 // big pixels per position step (2: stride-2 windows, 1: dense)
    const float* small;
    int Wo, Ho;                             // small-image grid
    long long big_img;
    float* partial;
    int debug;                              // 16: per-phase clock64 accounting of the A loaders (CTA 0)
};
constexpr uint32_t kAGroupBytes = kKR * 128;                  // one A channel group: 32 positions x 32 floats
constexpr uint32_t kAStageBytes = 4 * kAGroupBytes;           // 128 channels
constexpr int kAStages = 7;                                   // shared-memory ring of the A boxes (deeper than the TMEM ring)
constexpr int kBStages = 7;                                   // shared-memory ring of the B boxes
// tensor maps of `big` by box width: m[w - 1] loads boxes of 32 * w floats per position (w = 1..4 channel groups)
struct Tc3BigMaps {
    alignas(64) CUtensorMap m[4];
};
// the <= 4 channel groups of a 128-row tile form runs that are contiguous in memory (same kernel row): one TMA box per run
struct Tc3Seg {
    int g0, ng, kh, off0;
};
__device__ __forceinline__ int tile_segments(int i0, int I, int run, Tc3Seg* seg) {
    int n = 0;
    for (int g = 0; g < 4; ++g) {
        const int ig = i0 + 32 * g;
        if (ig >= I) break;
        const int kh = ig / run, off = ig - kh * run;
        if (n > 0 && seg[n - 1].kh == kh) { ++seg[n - 1].ng; continue; }
        seg[n].g0 = g; seg[n].ng = 1; seg[n].kh = kh; seg[n].off0 = off;
        ++n;
    }
    return n;
}

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
tc3_wgrad_kernel(const __grid_constant__ CUtensorMap smallmap, const __grid_constant__ Tc3BigMaps bigmaps, const __grid_constant__ Tc3WgParams p) {
    constexpr int NGB = BN / 32;
    constexpr int ACC_COLS = 4 * BN;                          // 2 chunk buffers x (main | cross)
    constexpr int TMEM_COLS = 512;                            // ACC_COLS (<= 256) + kStages * 64 = 512 at BN = 64
    static_assert(ACC_COLS + kStages * 64 <= TMEM_COLS, "tensor memory budget");
    constexpr uint32_t GROUP_BYTES = kKR * 128;               // one B column group: 32 positions x 128 B
    constexpr uint32_t STAGE_BYTES = GROUP_BYTES * 2 * NGB;   // [B_hi x NGB | B_lo x NGB]

    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t b_full[kBStages], b_ready[kBStages], b_empty[kBStages];      // B ring (shared memory)
    __shared__ uint64_t a_full[kAStages], a_free[kAStages];   // A ring (shared memory): boxes landed / read out by the 4 loader warps
    __shared__ uint64_t a_ready[kStages], empty_bar[kStages]; // A ring (tensor memory): slot written / its MMAs retired
    __shared__ uint64_t chunk_bar[2], drained_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ int progress;                                  // k-block the producer has reached (L2 prefetch pacing)

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    if (tid == 0) progress = 0;
    const int i0 = blockIdx.x * TBM;
    const int j0 = blockIdx.y * BN;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base + (uint32_t)kBStages * STAGE_BYTES;    // A stages behind the B stages
    const long long kb_begin = (long long)blockIdx.z * p.boxes_per_split;
    long long kb_end = kb_begin + p.boxes_per_split;
    if (kb_end > p.nboxes) kb_end = p.nboxes;
    const int nkb = kb_end > kb_begin ? (int)(kb_end - kb_begin) : 0;
    const int nchunks = (nkb + kChunkKb - 1) / kChunkKb;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) { mbar_init(&a_ready[s], 4); mbar_init(&empty_bar[s], 1); }
#pragma unroll
        for (int s = 0; s < kBStages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_ready[s], kSplitWarps); mbar_init(&b_empty[s], 1); }
#pragma unroll
        for (int s = 0; s < kAStages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_free[s], 4); }
        mbar_init(&chunk_bar[0], 1); mbar_init(&chunk_bar[1], 1);
        mbar_init(&drained_bar[0], kDrainWarps); mbar_init(&drained_bar[1], kDrainWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc<TMEM_COLS>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    // Box cursor: box index -> origin (x fastest), advanced incrementally -- a 64-bit div/mod chain per k-block in the single
    // producer lane costs more than the k-block's MMAs (it did: the first version of this kernel).
    struct Cursor {
        int tx, ty, tn;
    };
    auto cursor_at = [&](long long b) {
        Cursor c;
        c.tx = (int)(b % p.tiles_x);
        const long long r = b / p.tiles_x;
        c.ty = (int)(r % p.tiles_y); c.tn = (int)(r / p.tiles_y);
        return c;
    };
    auto cursor_next = [&](Cursor& c) {
        if (++c.tx == p.tiles_x) { c.tx = 0; if (++c.ty == p.tiles_y) { c.ty = 0; ++c.tn; } }
    };
    const bool PROF = (p.debug & 16) != 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define TC3_PROF(slot) do { if (PROF) { const long long now_ = clock64(); prof[slot] += now_ - tlast; tlast = now_; } } while (0)

    if (warp < kLoaderWarps) {
        // ================================ A loaders: shared memory -> registers -> tensor memory ================================
        const int set = warp >> 2;                        // k-blocks kb % 2 == set
        const int q = warp & 3;                           // TMEM lane quarter
        const int i = i0 + q * 32 + lane;                 // this thread's channel
        const bool i_ok = i < p.I;                        // whole 32-channel groups are valid or not (I % 32 == 0)
        // where this warp's channel group sits inside an A stage: its run's base, the run's row pitch, the group's column
        uint32_t a_off = 0u, a_pitch = 128u;
        {
            Tc3Seg seg[4];
            const int nseg = tile_segments(i0, p.I, p.run, seg);
            for (int t = 0; t < nseg; ++t)
                if (q >= seg[t].g0 && q < seg[t].g0 + seg[t].ng) {
                    a_off = (uint32_t)seg[t].g0 * kAGroupBytes + (uint32_t)(q - seg[t].g0) * 128u + (uint32_t)lane * 4u;
                    a_pitch = (uint32_t)seg[t].ng * 128u;
                }
        }
        long long prof[4] = {0, 0, 0, 0}, tlast = clock64();
        const long long tstart = tlast;
        for (int kb = set; kb < nkb; kb += 2) {
            const int s = kb % kStages;
            const int sa = kb % kAStages;
            float v[kKR];
            // every warp waits (also one whose channel group lies past I): its arrival on a_free must not overtake the slot's phase
            mbar_wait(&a_full[sa], (uint32_t)((kb / kAStages) & 1));
            TC3_PROF(0);
            if (i_ok) {
                const uint32_t ap = a_base + (uint32_t)sa * kAStageBytes + a_off;
#pragma unroll
                for (int qq = 0; qq < kKR; ++qq) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[qq]) : "r"(ap + (uint32_t)qq * a_pitch));
            } else {
#pragma unroll
                for (int qq = 0; qq < kKR; ++qq) v[qq] = 0.f;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_free[sa]);       // the values are in registers: the copy engine may refill the slot
            TC3_PROF(1);
            if (kb >= kStages) mbar_wait(&empty_bar[s], (uint32_t)((kb / kStages - 1) & 1));     // the MMAs that read this TMEM slot are done
            TC3_PROF(2);
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
            TC3_PROF(3);
        }
        if (PROF && lane == 0 && (warp == 0 || warp == 4))
            printf("tc3prof loader warp %d nkb %d total %lld: wait_a_full %lld lds %lld wait_empty %lld split+st %lld\n",
                   warp, nkb, clock64() - tstart, prof[0], prof[1], prof[2], prof[3]);
    } else if (warp >= kDrainWarp0) {
        // ================================ accumulator drain (warps 12-15 = TMEM lane quarters 0-3) ================================
        // Each finished 128-position chunk is added to fp32 register accumulators (the tensor core accumulates with
        // round-toward-zero, see tc_tapgemm.cu); at the end every thread writes its row of partial[split][i][j0 .. j0 + BN).
        const int q = warp & 3;
        const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16);
        float acc[BN];
#pragma unroll
        for (int c = 0; c < BN; ++c) acc[c] = 0.f;
        // Between chunks these warps pull the operands of the k-blocks up to kPrefetchKb ahead of the producer into L2
        // (prefetch.global.L2, one 128-byte line per lane and instruction; both operands stream from HBM exactly once).
        // They never block on it: `progress` bounds how far ahead they may run, a finished chunk always comes first.
        const int t = tid - kDrainWarp0 * 32;
        constexpr int NT = kDrainWarps * 32;
        constexpr int SLOTS = (kKR * (4 + NGB) + NT - 1) / NT;      // lines per lane and k-block
        int ngroups = 0;
        while (ngroups < 4 && i0 + 32 * ngroups < p.I) ++ngroups;
        const int per_pos = ngroups + NGB;                          // A channel groups + B column groups: one line each per position
        long long rel[SLOTS];                                       // float offset of the lane's line from the box origin of its tensor
        int rel_n[SLOTS];                                           // image of the line relative to the box (-1: no line)
        bool is_b[SLOTS];
        {
            const int bwh = p.bw * p.bh;
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
                const int l = t + k * NT;
                rel_n[k] = -1; rel[k] = 0; is_b[k] = false;
                if (l >= kKR * per_pos) continue;
                const int pos = l / per_pos, g = l - pos * per_pos;
                const int nn = pos / bwh, rr = pos - nn * bwh;
                const int yy = rr / p.bw, xx = rr - yy * p.bw;
                rel_n[k] = nn;
                if (g < ngroups) {
                    const int ig = i0 + 32 * g, kh = ig / p.run;
                    rel[k] = (long long)nn * p.big_img + ((long long)(p.sx * yy + kh) * p.Wb + p.sx * xx) * p.Cb + (ig - kh * p.run);
                } else {
                    is_b[k] = true;
                    rel[k] = (((long long)nn * p.Ho + yy) * p.Wo + xx) * p.J + j0 + 32 * (g - ngroups);
                }
            }
        }
        Cursor pc = cursor_at(kb_begin);
        int pnext = (p.debug & 8) ? nkb : 0;                        // debug 8: no L2 prefetch
        auto prefetch_some = [&]() -> bool {
            if (pnext >= nkb || pnext >= *reinterpret_cast<volatile int*>(&progress) + kPrefetchKb) return false;
            const int x0 = pc.tx * p.bw, y0 = pc.ty * p.bh, n0 = pc.tn * p.bn;
            const long long a_org = (long long)n0 * p.big_img + ((long long)(p.sx * y0) * p.Wb + p.sx * x0) * p.Cb;
            const long long b_org = (((long long)n0 * p.Ho + y0) * p.Wo + x0) * p.J;
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
                if (rel_n[k] < 0 || n0 + rel_n[k] >= p.batch) continue;
                const float* a = is_b[k] ? p.small + b_org + rel[k] : p.big + a_org + rel[k];
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a) : "memory");
            }
            cursor_next(pc);
            ++pnext;
            return true;
        };
        auto chunk_done = [&](int b, uint32_t parity) -> bool {
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&chunk_bar[b])), "r"(parity) : "memory");
            return __all_sync(0xffffffffu, ok != 0);
        };
        long long prof[4] = {0, 0, 0, 0}, tlast = clock64();
        const long long tstart = tlast;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int b = ch & 1;
            while (!chunk_done(b, (uint32_t)((ch >> 1) & 1)))
                if (!prefetch_some()) __nanosleep(64);
            TC3_PROF(0);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < BN; cc += 32) {
                float v[16], w[16], x[16], y[16];
                tmem_ld16_issue(tmem_acc + (uint32_t)(b * 2 * BN + cc), v);
                tmem_ld16_issue(tmem_acc + (uint32_t)(b * 2 * BN + cc + 16), w);
                tmem_ld16_issue(tmem_acc + (uint32_t)(b * 2 * BN + BN + cc), x);
                tmem_ld16_issue(tmem_acc + (uint32_t)(b * 2 * BN + BN + cc + 16), y);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) { acc[cc + c] += v[c] + x[c]; acc[cc + 16 + c] += w[c] + y[c]; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&drained_bar[b]);
            TC3_PROF(1);
        }
        const int i = i0 + q * 32 + lane;
        if (i < p.I) {
            float* out = p.partial + ((long long)blockIdx.z * p.I + i) * p.J + j0;
#pragma unroll
            for (int c = 0; c < BN; c += 4)
                *reinterpret_cast<float4*>(out + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        }
        if (PROF && lane == 0 && warp == kDrainWarp0)
            printf("tc3prof drain nkb %d total %lld: wait_chunk %lld drain %lld\n", nkb, clock64() - tstart, prof[0], prof[1]);
    } else if (warp == kProducerWarp) {
        // ================================ producer: B boxes and A run boxes of every k-block ================================
        if (lane == 0 && nkb > 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&smallmap)) : "memory");
            Tc3Seg seg[4];
            const int nseg = tile_segments(i0, p.I, p.run, seg);
            uint32_t tx = 0;
            for (int t = 0; t < nseg; ++t) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&bigmaps.m[seg[t].ng - 1])) : "memory");
                tx += (uint32_t)seg[t].ng * kAGroupBytes;
            }
            Cursor c = cursor_at(kb_begin);
            long long prof[4] = {0, 0, 0, 0}, tlast = clock64();
            const long long tstart = tlast;
            for (int kb = 0; kb < nkb; ++kb, cursor_next(c)) {
                const int sb = kb % kBStages, sa = kb % kAStages;
                const int x0 = c.tx * p.bw, y0 = c.ty * p.bh, n0 = c.tn * p.bn;
                if (kb >= kBStages) mbar_wait(&b_empty[sb], (uint32_t)((kb / kBStages - 1) & 1));
                *reinterpret_cast<volatile int*>(&progress) = kb;
                TC3_PROF(0);
                mbar_expect_tx(&b_full[sb], GROUP_BYTES * NGB);
#pragma unroll
                for (int g = 0; g < NGB; ++g)
                    tma_load_4d(smem_base + (uint32_t)sb * STAGE_BYTES + (uint32_t)g * GROUP_BYTES, &smallmap, j0 + 32 * g, x0, y0, n0, &b_full[sb]);
                TC3_PROF(2);
                if (kb >= kAStages) mbar_wait(&a_free[sa], (uint32_t)((kb / kAStages - 1) & 1));
                TC3_PROF(1);
                mbar_expect_tx(&a_full[sa], tx);
                for (int t = 0; t < nseg; ++t)
                    tma_load_4d(a_base + (uint32_t)sa * kAStageBytes + (uint32_t)seg[t].g0 * kAGroupBytes, &bigmaps.m[seg[t].ng - 1], seg[t].off0, x0, p.sx * y0 + seg[t].kh,
                                n0, &a_full[sa]);
                TC3_PROF(2);
            }
            *reinterpret_cast<volatile int*>(&progress) = nkb;
            if (PROF) printf("tc3prof producer nkb %d total %lld: wait_b_empty %lld wait_a_free %lld issue %lld\n", nkb, clock64() - tstart, prof[0], prof[1], prof[2]);
        }
        __syncwarp();
    } else if (warp >= kSplitWarp0 && warp < kSplitWarp0 + kSplitWarps) {
        // ================================ B splitters: b_lo = b - trunc_tf32(b) ================================
        const int tl = tid - kSplitWarp0 * 32;
        constexpr uint32_t CHUNKS = NGB * GROUP_BYTES / 16;
        for (int kb = 0; kb < nkb; ++kb) {
            const int sb = kb % kBStages;
            mbar_wait(&b_full[sb], (uint32_t)((kb / kBStages) & 1));
            const uint32_t stage = smem_base + (uint32_t)sb * STAGE_BYTES;
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
            if (lane == 0) mbar_arrive(&b_ready[sb]);
        }
    } else if (warp == kIssuerWarp) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            // A: K-major by construction (TMEM lane = row, column = k); B: MN-major shared memory (bit 16)
            const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(TBM >> 4) << 24);
            const uint32_t idesc = idesc_base | ((uint32_t)(BN >> 3) << 17);
            const uint32_t idesc2 = idesc_base | ((uint32_t)((2 * BN) >> 3) << 17);
            long long prof[4] = {0, 0, 0, 0}, tlast = clock64();
            const long long tstart = tlast;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages, sb = kb % kBStages;
                const uint32_t stage = smem_base + (uint32_t)sb * STAGE_BYTES;
                const int chunk = kb / kChunkKb;
                const int b = chunk & 1;
                TC3_PROF(3);
                mbar_wait(&b_ready[sb], (uint32_t)((kb / kBStages) & 1));
                TC3_PROF(0);
                mbar_wait(&a_ready[s], (uint32_t)((kb / kStages) & 1));
                TC3_PROF(1);
                if (kb % kChunkKb == 0 && chunk >= 2) mbar_wait(&drained_bar[b], (uint32_t)(((chunk >> 1) - 1) & 1));
                TC3_PROF(2);
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
                umma_commit(&b_empty[sb]);
                if (kb % kChunkKb == kChunkKb - 1 || kb == nkb - 1) umma_commit(&chunk_bar[b]);
            }
            if (PROF)
                printf("tc3prof issuer nkb %d total %lld: wait_b_ready %lld wait_a_ready %lld wait_drained %lld issue+other %lld\n", nkb, clock64() - tstart, prof[0], prof[1], prof[2], prof[3]);
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
int32_t launch_bn(const CUtensorMap& smallmap, const Tc3BigMaps& bigmaps, const Tc3WgParams& p, int splits, cudaStream_t stream) {
    const size_t smem = (size_t)kKR * 128 * 2 * (BN / 32) * kBStages + (size_t)kAStageBytes * kAStages + 1024;
    dim3 grid((unsigned)cdiv(p.I, TBM), (unsigned)(p.J / BN), (unsigned)splits);
    tc3_wgrad_kernel<BN><<<grid, kThreads, smem, stream>>>(smallmap, bigmaps, p);
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
    CPB_CUDA(cudaFuncSetAttribute(tc3_wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024 + 1024));
    CPB_CUDA(cudaFuncSetAttribute(tc3_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024 + 1024));
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
    p.small = w.small; p.Wo = dense ? 1 : w.Wo; p.Ho = dense ? 1 : w.Ho;
    p.debug = w.tc_variant;
    alignas(64) CUtensorMap smallmap;
    const unsigned long long dims[4] = {(unsigned long long)w.J, (unsigned long long)w.Wo, (unsigned long long)w.Ho, (unsigned long long)w.batch};
    const unsigned long long strides[3] = {(unsigned long long)w.J * 4ull, (unsigned long long)w.Wo * w.J * 4ull, (unsigned long long)w.Ho * w.Wo * w.J * 4ull};
    const unsigned box[4] = {32u, (unsigned)p.bw, (unsigned)p.bh, (unsigned)p.bn};
    const unsigned estr[4] = {1u, 1u, 1u, 1u};
    CPB_TRY(tc2_encode_tiled(&smallmap, w.small, 4, dims, strides, box, estr, 1));
    // big as a tensor of WINDOW RUNS {f, x, row, image}: f = float inside the kernel row's run starting at window origin x
    // (stride 2 pixels -- consecutive windows overlap, which the tensor map does not mind), rows traversed with stride 2.
    // A box = the 32 window origins of a k-block x (32 w) contiguous floats, unswizzled; one map per box width w.
    Tc3BigMaps bigmaps;
    memset(&bigmaps, 0, sizeof(bigmaps));
    p.sx = dense ? 1 : 2;
    {
        const unsigned long long pitch = (unsigned long long)w.big_pitch;
        const unsigned long long Wb = dense ? 1ull : (unsigned long long)w.Wb;
        const unsigned long long Hb = dense ? 1ull : (unsigned long long)(w.big_img / ((long long)w.Wb * w.big_pitch));
        const unsigned long long xs = dense ? 1ull : (unsigned long long)w.Wo;
        const unsigned long long f_extent = dense ? (unsigned long long)w.I : (Wb - 2ull * (xs - 1ull)) * pitch;       // floats a window may span in its row
        CPB_REQUIRE(f_extent >= (unsigned long long)w.run, "tc3_wgrad: window run leaves the image row");
        const unsigned long long bdims[4] = {f_extent, xs, Hb, (unsigned long long)w.batch};
        const unsigned long long bstrides[3] = {(dense ? pitch : 2ull * pitch) * 4ull, Wb * pitch * 4ull, (unsigned long long)w.big_img * 4ull};
        const unsigned es = (unsigned)p.sx;
        const unsigned bestr[4] = {1u, 1u, es, 1u};
        bool need[4] = {false, false, false, false};
        for (int i0 = 0; i0 < w.I; i0 += TBM) {
            int prev_kh = -1, ng = 0;
            for (int g = 0; g < 4 && i0 + 32 * g < w.I; ++g) {
                const int kh = (i0 + 32 * g) / w.run;
                if (kh != prev_kh && ng > 0) { need[ng - 1] = true; ng = 0; }
                prev_kh = kh; ++ng;
            }
            if (ng > 0) need[ng - 1] = true;
        }
        for (int wd = 1; wd <= 4; ++wd) {
            if (!need[wd - 1]) continue;
            const unsigned bbox[4] = {32u * (unsigned)wd, (unsigned)p.bw, (unsigned)p.bh * es, (unsigned)p.bn};
            CPB_TRY(tc2_encode_tiled(&bigmaps.m[wd - 1], w.big, 4, bdims, bstrides, bbox, bestr, 2));
        }
    }
    return w.J == 64 ? launch_bn<64>(smallmap, bigmaps, p, w.splits, stream) : launch_bn<32>(smallmap, bigmaps, p, w.splits, stream);
}

}  // namespace cpb
