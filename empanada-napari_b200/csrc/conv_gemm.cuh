// Implicit-GEMM convolution on tcgen05 / TMEM fed by TMA (sm_100a).
//
// Replaces the cuDNN convolutions that the reference reaches through
// `self.model(image, render_steps, interpolate_ins)` (empanada/inference/engines.py:250),
// i.e. every Conv2d of ResNet-50 / ASPP / decoder / heads
// (empanada/models/encoders/resnet.py:217, decoders/aspp.py:97, heads.py:9).
//
// Data layout: activations NHWC bf16 (channel stride may exceed C so that a conv can write
// straight into a concat buffer), weights [Cout][R*S*Cin] bf16 (K-major), fp32 accumulate.
// One output tile = 128 pixels (TB x TH x TW block of the output map) x BLOCK_N channels.
// The A operand of tap (r,s) is ONE 4-D TMA box of the input tensor shifted by
// (r*dil - pad, s*dil - pad): TMA's out-of-bounds zero fill is the conv padding and its
// element strides are the conv stride, so no im2col buffer ever exists in HBM.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace convgemm {

constexpr int TILE_M = 128;   // output pixels per tile (UMMA M)
constexpr int KCHUNK = 64;    // bf16 channels per pipeline stage (128 B swizzle row)
constexpr int MAX_N = 256;    // UMMA N limit
constexpr int NUM_THREADS = 576;  // warp0 TMA, warp1 MMA, warps2-17 epilogue

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SILU = 2 };

struct Params {
  // problem
  int B, Ho, Wo;         // output map
  int Cin, Cout;         // channels (GEMM K per tap, GEMM N)
  int R, S;              // filter taps
  int stride, dil, pad;  // conv geometry
  // tiling
  int TW, TH, TB;        // TW*TH*TB == 128
  int tiles_x, tiles_y, tiles_b, tiles_n;
  int block_n;           // channels per tile (multiple of 16, <= 256)
  int kchunks;           // ceil(Cin / 64)
  int stages;            // smem pipeline depth
  int b_stationary;      // 1: weights of the current N tile stay resident in smem
  int staged;            // 1: epilogue transposes through smem for coalesced loads/stores
  // epilogue
  __nv_bfloat16* out;    // NHWC, pixel stride out_ld, written at channel offset out_coff
  long long out_ld;
  int out_coff;
  float* out_f32;        // optional fp32 NHWC output (pixel stride out_f32_ld)
  long long out_f32_ld;
  const float* bias;     // [Cout] or nullptr (BN folded on the host)
  long long bias_img_stride;  // 0: shared bias; else bias + b * stride (per-image bias)
  int out_f32_planar;    // fp32 output as [B][Cout][Ho*Wo] instead of NHWC
  // fused 1x1 "head" (Cout -> head_n <= 2) applied to the activated tile: the 256-channel
  // intermediate never reaches HBM (requires tiles_n == 1)
  const float* head_w;   // [head_n][Cout]
  const float* head_b;   // [head_n]
  float* head_out;       // [B][head_n][Ho*Wo] fp32
  int head_n;
  const __nv_bfloat16* residual;  // optional NHWC addend (pixel stride res_ld)
  long long res_ld;
  int act;
  // ConvTranspose2d(k=2, s=2) as a 1x1 GEMM with N = 4 * Cout_real: N tile nt = 2*dy + dx holds
  // the Cout_real channels of output pixel (2y + dy, 2x + dx) (staged epilogue only)
  int shuffle2x2;
  // 3x3 / stride 1 / Cin = 64 "halo" schedule: ONE TMA box per tile holds the (TH+2) x 16-pixel
  // input patch (row pitch 16 pixels = 2 KiB, so every 8-row group of every tap starts at the
  // same swizzle phase); the nine taps are nine UMMA descriptors into that patch instead of nine
  // boxes from L2. 0 = off, 1 / 2 = on (descriptor base_offset 0 / = tap column)
  int halo;
};
constexpr int HALO_PW = 16;                    // patch row pitch in pixels
constexpr int HALO_TW = 8, HALO_TH = 16;       // output tile of the halo schedule

// Host side: builds the two tensor maps and launches. Returns cudaError_t as int.
struct Launch {
  CUtensorMap tmap_a;  // input activation (C, W, H, B)
  CUtensorMap tmap_b;  // weights (K_total, Cout)
  Params p;
  int grid;
  size_t smem;
};

// cuTensorMapEncodeTiled for a bf16 tensor (dims / box innermost first, strides in bytes for
// dims 1..rank-1); zero fill out of bounds. Returns 0 or a negative error. Shared with the
// TMA-fed CUDA-core kernels (depthwise convolution).
int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                     const unsigned long long* strides_bytes, const unsigned* box,
                     const unsigned* elem_strides, int swizzle_128b);

}  // namespace convgemm
