// Second-generation tensor-core kernels (tc2_tapgemm.cu, tc2_wgrad.cu): operands arrive by TMA TENSOR MAPS
// (cp.async.bulk.tensor, SASS UTMALDG) as boxes of [positions x 32 floats] = rows of 128 bytes, SWIZZLE_128B, and the
// second TF32 operand (x - trunc_tf32(x)) is derived INSIDE the kernel from the tile already in shared memory.
//
// Why (round-2 measurements, profiles/r2_cycle_accounting.md): the round-1 kernels fetched BOTH planes (x and x_lo) of both
// operands through L2 with 16-byte cp.async: 48 KB per 32-wide k-block and CTA, i.e. 4-5.3 KB/clk chip-wide against the
// ~6.3 KB/clk the L2 delivers -- they were L2-bandwidth bound at 27-58 % tensor-pipe activity, and every activation /
// gradient tensor existed twice in HBM.  Now a k-block moves 16 KB (A) + BN*128/CS bytes (weights, multicast over the
// CS CTAs of a cluster); the lo planes are gone from HBM altogether.
//
// Tile geometry.  A tile's 128 MMA rows are a BOX of output positions (bw x bh x bn <= 128 positions of the
// [x][y][image] grid, x fastest), because a TMA box is rectangular: row r <-> (x0 + r % bw, y0 + (r / bw) % bh,
// n0 + r / (bw*bh)); rows >= bw*bh*bn are never written by the copy and their results are never stored.
//   * gather form (Conv2D forward, conv2d_transpose data-gradient; stride-2 windows, always in bounds): the tensor map is
//     the dense NHWC source {C, W, H, B} traversed with elementStrides {1,2,2,1}; k-block (kh, kw, channel group c0) is the
//     box at coordinates {c0, 2*x0 + kw, 2*y0 + kh, n0};
//   * quad-fused scatter form (conv2d_transpose forward, Conv2D data-gradient): rows are 2x2 output quads, the k-blocks of
//     window tap (j, i) are the boxes at {c0, x0 - i, y0 - j, n0} of the dense source; positions outside the image are
//     zero-filled by the copy engine (no bounds code in the kernel);
//   * dense layers: {K, 1, 1, B} with bw = bh = 1.
#pragma once
#include <cuda.h>

#include "tapgemm.cuh"
#include "tc_common.cuh"

namespace cpb {

constexpr int kTc2ColsumRows = 4 * 160;   // (CTA, lane quarter) rows of a column-sum buffer: the persistent grid never exceeds 160 CTAs
constexpr int kTc2MaxKb = 160;   // k-blocks per tile (conv4 / deconv1: 4 taps x 512 / 32 = 64; deconv3 quad: 9 x 2)

struct Tc2KBlock {
    short c, dx, dy, pad;        // tensor-map coordinates of the k-block relative to the tile origin
};

struct Tc2Params {
    // tile geometry (positions)
    int bw, bh, bn;              // box extent in x, y, images; rows = bw*bh*bn <= 128
    int sx;                      // source pixels per output step: 2 gather, 1 quad / dense
    int gw, gh, batch;           // position grid and batch
    int tiles_x, tiles_y, tiles_n;
    int nkb;                     // k-blocks per tile
    int N;                       // output columns (multiple of BN)
    const float* wk;             // weight tile images: block (n-tile y, k-block kb) at ((y * nkb + kb) * BN * 32) floats
    // epilogue
    const float* bias;
    const float* mask;
    float* dst;
    int relu;
    int quad, quad_cb, quad_lcb;
    int Hd, Wd, dst_pitch;
    long long dst_img;
    float* colsum;               // see TapGemmParams::colsum
    int cluster;
    int debug;
    Tc2KBlock kb[kTc2MaxKb];
};

// host: driver entry point for cuTensorMapEncodeTiled (resolved through the runtime, no libcuda link dependency)
// atom32 = 0: SWIZZLE_128B (16-byte chunks XOR row % 8; K-major operands); 1: SWIZZLE_128B_ATOM_32B (32-byte chunks XOR
// row % 4) -- the only shared-memory layout the tensor core accepts for MN-major 32-bit (TF32) operands; 2: no swizzle
// (tiles that threads, not the tensor core, read)
int32_t tc2_encode_tiled(CUtensorMap* map, const float* base, int rank, const unsigned long long* dims,
                         const unsigned long long* strides_bytes, const unsigned* box, const unsigned* elem_strides, int atom32 = 0);

int32_t tc2_tapgemm_init();
bool tc2_enabled();
// layout of the weight images the tap-GEMM expects (TcWeightJob.raw): 1 = [hi image | lo image] per block, 2 = CTA-pair order
int tc2_weight_layout();
// Builds the tensor map + schedule for a gather / quad / dense problem described by the round-1 TapGemmParams and
// enqueues it.  p.wk_hi must point at the RAW weight images (tc_weights mode "raw").
int32_t launch_tc2_tapgemm(const TapGemmParams& p, int scatter_k, cudaStream_t stream);
bool tc2_tapgemm_supported(const TapGemmParams& p, int scatter_k);

// ---- weight gradient (tc2_wgrad.cu): MN-major operands by TMA tensor maps; same WgradParams as the round-1 kernels
struct WgradParams;
int32_t tc2_wgrad_init();
bool tc2_wgrad_supported(const WgradParams& w);
void tc2_wgrad_plan(const WgradParams& w, int& bw, int& bh, int& bn, long long& nboxes);
int tc2_wgrad_pick_splits(int I, int J, long long nboxes);
int32_t launch_tc2_wgrad(const WgradParams& w, cudaStream_t stream);      // w.splits from tc2_wgrad_pick_splits; w.tc_variant & 1: descriptor probe

// ---- weight gradient with the A operand in tensor memory (tc3_wgrad.cu), J = 32 / 64
int32_t tc3_wgrad_init();
bool tc3_wgrad_supported(const WgradParams& w);     // enabled (CPB_TC3_WGRAD=1) and the shape fits
bool tc3_wgrad_available(const WgradParams& w);     // the shape fits (unit tests go through cpb_debug_tc_wgrad)
long long tc3_wgrad_boxes(const WgradParams& w);                          // k-blocks (boxes of 32 positions) of the whole reduction
int32_t launch_tc3_wgrad(const WgradParams& w, cudaStream_t stream);      // w.splits <= 148 / tiles, <= boxes

}  // namespace cpb
