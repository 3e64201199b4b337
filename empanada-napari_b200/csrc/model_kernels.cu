// Non-GEMM layers of the PanopticDeepLab forward pass (piece 1): everything around the tcgen05
// implicit-GEMM convolutions. NHWC bf16 activations, fp32 math, 16-byte vector accesses.
//   stem_kernel        slice gather + normalise + zero pad (volume_dataset.py:37-53,
//                      utils.py:170-201, postprocess.py:26-36) fused into conv1 7x7/2 + BN + ReLU
//                      (encoders/resnet.py:217-220)
//   maxpool_kernel     MaxPool2d(3, 2, 1)                      (encoders/resnet.py:221)
//   dwconv_kernel      depthwise k x k of SeparableConv2d      (blocks.py:15-35)
//   bilinear_kernel    F.interpolate(bilinear, align_corners=True) into a concat slice
//                      (decoders/panoptic_deeplab.py:76)
//   aspp_pool_*        ASPPPooling branch folded into a per-image bias of the ASPP projection
//                      (decoders/aspp.py:30-48,97-102)
//   up2_kernel / topk_* / pr_* : PointRend refinement (point_rend.py:110-137,241-269)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "common.cuh"
#include "conv_gemm.cuh"
#include "sm100_ptx.cuh"

namespace mk {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __bfloat162float(h[i].x); f[2 * i + 1] = __bfloat162float(h[i].y); }
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}


// Input volume element types (numpy integer dtypes accepted by Preprocessor, empanada_napari/
// utils.py:189-201: `image.astype(np.float32)` then (x - mean*max) * 1/(std*max), max = iinfo.max)
__device__ __forceinline__ float load_elem(const void* __restrict__ base, long long i, int elem) {
  switch (elem) {
    case 0: return static_cast<float>(static_cast<const uint8_t*>(base)[i]);
    case 1: return static_cast<float>(static_cast<const int8_t*>(base)[i]);
    case 2: return static_cast<float>(static_cast<const uint16_t*>(base)[i]);
    case 3: return static_cast<float>(static_cast<const int16_t*>(base)[i]);
    case 4: return static_cast<float>(static_cast<const uint32_t*>(base)[i]);
    case 5: return static_cast<float>(static_cast<const int32_t*>(base)[i]);
    case 6: return static_cast<float>(static_cast<const unsigned long long*>(base)[i]);
    case 7: return static_cast<float>(static_cast<const long long*>(base)[i]);
    default: return static_cast<const float*>(base)[i];   // 8: already normalised fp32 (engine API)
  }
}

// ------------------------------------------------------------------ stem
// One thread = one output pixel x 64 channels. 16x16 output tile per CTA, 37x37 input patch.
constexpr int ST = 16;
constexpr int SP = 2 * ST + 5;
__global__ void __launch_bounds__(ST * ST)
stem_kernel(const void* __restrict__ vol, int elem, long long stride_s, long long stride_y,
            long long stride_x, int s0, int h, int w, int H, int W, float mean255, float den,
            const float* __restrict__ wt /*[49][64]*/, const float* __restrict__ bias,
            bf16* __restrict__ out) {
  __shared__ float patch[SP][SP + 1];
  __shared__ __align__(16) float ws[49 * 64];
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * ST, ox0 = blockIdx.x * ST;
  const int Ho = H / 2, Wo = W / 2;
  const long long src0 = static_cast<long long>(s0 + b) * stride_s;
  for (int i = threadIdx.x; i < 49 * 64; i += blockDim.x) ws[i] = wt[i];
  for (int i = threadIdx.x; i < SP * SP; i += blockDim.x) {
    const int py = i / SP, px = i - py * SP;
    const int y = 2 * oy0 - 3 + py, x = 2 * ox0 - 3 + px;
    float v = 0.0f;
    if (y >= 0 && y < h && x >= 0 && x < w)
      v = __fmul_rn(__fsub_rn(load_elem(vol, src0 + y * stride_y + x * stride_x, elem), mean255), den);
    patch[py][px] = v;
  }
  __syncthreads();
  const int ty = threadIdx.x / ST, tx = threadIdx.x - ty * ST;
  const int oy = oy0 + ty, ox = ox0 + tx;
  float acc[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = 0.0f;
  for (int r = 0; r < 7; ++r) {
    for (int s = 0; s < 7; ++s) {
      const float v = patch[2 * ty + r][2 * tx + s];
      const float4* wv = reinterpret_cast<const float4*>(&ws[(r * 7 + s) * 64]);
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 wq = wv[c4];
        acc[4 * c4] = fmaf(v, wq.x, acc[4 * c4]);
        acc[4 * c4 + 1] = fmaf(v, wq.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(v, wq.z, acc[4 * c4 + 2]);
        acc[4 * c4 + 3] = fmaf(v, wq.w, acc[4 * c4 + 3]);
      }
    }
  }
  if (oy < Ho && ox < Wo) {
    bf16* op = out + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * 64;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(acc[8 * g + j] + bias[8 * g + j], 0.0f);
      reinterpret_cast<uint4*>(op)[g] = pack8(f);
    }
  }
}

// ------------------------------------------------------------------ stem + maxpool fused
// conv1 7x7/2 + BN + ReLU followed by MaxPool2d(3, 2, 1) (encoders/resnet.py:217-221) in one
// kernel: the 64-channel half-resolution map (the largest activation of the network) never
// reaches HBM. CTA = 8 x 16 pooled pixels: it needs a 17 x 33 tile of conv1 outputs, which needs
// a 39 x 71 patch of the (gathered, normalised, zero padded) input slice.
// Lane l owns channels (2l, 2l+1). A warp computes strips of 11 conv1 pixels of one tile row:
// per filter row it loads 7 tap pairs (conflict-free 8-byte words) and 27 patch values (broadcast
// 8-byte words) for 11*7*2 FMAs per lane. Same fp32 fmaf chain order as `stem_kernel`
// (filter-row major, tap ascending, bias added last), so both produce identical bits.
// Positions of the conv1 tile outside the map hold 0: every pooling window contains at least one
// real (post-ReLU, >= 0) value, so 0 is equivalent to MaxPool's -inf padding.
constexpr int SPY = 8, SPX = 16;
constexpr int STY = 2 * SPY + 1, STX = 2 * SPX + 1;      // 17 x 33 conv1 tile
constexpr int SIH = 2 * STY + 5, SIW = 2 * STX + 5;      // 39 x 71 input patch
constexpr int SIWP = 72;
constexpr int SSTRIP = 11;
constexpr int stem_pool_smem_bytes() { return STY * STX * 32 * 4 + 49 * 64 * 4 + SIH * SIWP * 8; }
__global__ void __launch_bounds__(256, 2)
stem_pool_kernel(const void* __restrict__ vol, int elem, long long stride_s, long long stride_y,
                 long long stride_x, int s0, int h, int w, int H, int W, float mean255, float den,
                 const float* __restrict__ wt /*[49][64]*/, const float* __restrict__ bias,
                 bf16* __restrict__ out /*[B][H/4][W/4][64]*/) {
  extern __shared__ __align__(16) uint8_t sp_smem[];
  uint32_t* tile = reinterpret_cast<uint32_t*>(sp_smem);           // [STY*STX][32] bf16x2
  float* ws = reinterpret_cast<float*>(tile + STY * STX * 32);     // [49][64]
  // every input value is stored twice, (v, v): one broadcast 8-byte load yields the packed
  // multiplicand of an FFMA2 (two channels per lane) without a register move
  float2* patch = reinterpret_cast<float2*>(ws + 49 * 64);         // [SIH][SIWP]
  const int b = blockIdx.z;
  const int py0 = blockIdx.y * SPY, px0 = blockIdx.x * SPX;
  const int Ho = H / 2, Wo = W / 2, Hp = H / 4, Wp = W / 4;
  const int sy0 = 2 * py0 - 1, sx0 = 2 * px0 - 1;                  // conv1 tile origin
  const int iy0 = 2 * sy0 - 3, ix0 = 2 * sx0 - 3;                  // input patch origin
  const long long src0 = static_cast<long long>(s0 + b) * stride_s;
  for (int i = threadIdx.x; i < 49 * 64; i += blockDim.x) ws[i] = __ldg(wt + i);
  for (int i = threadIdx.x; i < SIH * SIWP; i += blockDim.x) {
    const int py = i / SIWP, px = i - py * SIWP;
    const int y = iy0 + py, x = ix0 + px;
    float v = 0.0f;
    if (px < SIW && y >= 0 && y < h && x >= 0 && x < w)
      v = __fmul_rn(__fsub_rn(load_elem(vol, src0 + y * stride_y + x * stride_x, elem), mean255), den);
    patch[i] = make_float2(v, v);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float2 bb = __ldg(reinterpret_cast<const float2*>(bias) + lane);
  constexpr int NSTRIP = STY * (STX / SSTRIP);
  for (int j = warp; j < NSTRIP; j += 8) {
    const int ly = j / (STX / SSTRIP), lx0 = (j - ly * (STX / SSTRIP)) * SSTRIP;
    float2 acc[SSTRIP];   // packed fp32 FMAs (FFMA2): both channels of the lane per instruction
#pragma unroll
    for (int i = 0; i < SSTRIP; ++i) acc[i] = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int r = 0; r < 7; ++r) {
      float2 wr[7];
#pragma unroll
      for (int t = 0; t < 7; ++t) wr[t] = *reinterpret_cast<const float2*>(ws + (r * 7 + t) * 64 + 2 * lane);
      float2 pv[2 * SSTRIP + 5];
      const float2* prow = patch + (2 * ly + r) * SIWP + 2 * lx0;
#pragma unroll
      for (int i = 0; i < 2 * SSTRIP + 5; ++i) pv[i] = prow[i];
#pragma unroll
      for (int t = 0; t < 7; ++t) {
#pragma unroll
        for (int i = 0; i < SSTRIP; ++i) acc[i] = __ffma2_rn(pv[2 * i + t], wr[t], acc[i]);
      }
    }
    const int sy = sy0 + ly;
#pragma unroll
    for (int i = 0; i < SSTRIP; ++i) {
      const int sx = sx0 + lx0 + i;
      __nv_bfloat162 o = __floats2bfloat162_rn(0.0f, 0.0f);
      if (sy >= 0 && sy < Ho && sx >= 0 && sx < Wo)
        o = __floats2bfloat162_rn(fmaxf(acc[i].x + bb.x, 0.0f), fmaxf(acc[i].y + bb.y, 0.0f));
      tile[(ly * STX + lx0 + i) * 32 + lane] = *reinterpret_cast<uint32_t*>(&o);
    }
  }
  __syncthreads();
  for (int pp = warp; pp < SPY * SPX; pp += 8) {
    const int oyl = pp / SPX, oxl = pp - oyl * SPX;
    const int oy = py0 + oyl, ox = px0 + oxl;
    if (oy >= Hp || ox >= Wp) continue;
    __nv_bfloat162 m = __floats2bfloat162_rn(0.0f, 0.0f);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint32_t u = tile[((2 * oyl + dy) * STX + 2 * oxl + dx) * 32 + lane];
        m = __hmax2(m, *reinterpret_cast<const __nv_bfloat162*>(&u));
      }
    *reinterpret_cast<__nv_bfloat162*>(out + ((static_cast<long long>(b) * Hp + oy) * Wp + ox) * 64 + 2 * lane) = m;
  }
}

// ------------------------------------------------------------------ maxpool 3x3 / 2, pad 1
__global__ void maxpool_kernel(const bf16* __restrict__ in, int B, int Hi, int Wi, int C,
                               bf16* __restrict__ out, int Ho, int Wo) {
  const int cg = C / 8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * Ho * Wo * cg;
  if (i >= total) return;
  const int g = static_cast<int>(i % cg);
  long long pix = i / cg;
  const int ox = static_cast<int>(pix % Wo); pix /= Wo;
  const int oy = static_cast<int>(pix % Ho);
  const int b = static_cast<int>(pix / Ho);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
  for (int dy = 0; dy < 3; ++dy) {
    const int y = 2 * oy - 1 + dy;
    if (y < 0 || y >= Hi) continue;
    for (int dx = 0; dx < 3; ++dx) {
      const int x = 2 * ox - 1 + dx;
      if (x < 0 || x >= Wi) continue;
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * Hi + y) * Wi + x) * C + g * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
    }
  }
  *reinterpret_cast<uint4*>(out + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * C + g * 8) = pack8(m);
}

// ------------------------------------------------------------------ depthwise k x k, stride 1
// CTA = 8 rows x 32 pixels x 64 channels, 8 warps. The (8+K-1) x (32+K-1) x 64-channel input
// patch is staged once in shared memory (128 B per pixel). Lane l of a warp owns channel pair
// (2l, 2l+1): its K*K x 2 filter taps live in registers, every shared load is a conflict-free
// 4-byte word, every global store instruction writes one pixel's 128 contiguous bytes. A warp
// owns a strip of 4 pixels and walks down the 8 output rows: each patch row it loads
// (4+K-1 words) feeds the K output rows that overlap it, held in a ring of K row accumulators,
// so the FMA : load ratio is 2*4*K*K : (4+K-1) per row and the kernel is FMA-pipe bound.
// Accumulation order per output is filter-row major, tap ascending (one fp32 fmaf chain).
// Optional fused producer: channels [0, Cup) of the input are the align_corners=True bilinear
// upsampling of a low-resolution NHWC tensor `up` (the decoder's `F.interpolate` + `torch.cat`,
// decoders/panoptic_deeplab.py:76-77), channels [Cup, C) come from `in`; the concatenated
// tensor is never materialised. The low-resolution source pixels a patch needs are staged in
// shared memory first (DW_SH x DW_SW pixels), then interpolated out of shared memory.
constexpr int DW_TY = 8, DW_TX = 32, DW_PX = 4, DW_CB = 64;
constexpr int DW_SH = 8, DW_SW = 14;
template <int K>
constexpr int dw_smem_bytes() { return ((DW_TY + K - 1) * (DW_TX + K - 1) + DW_SH * DW_SW) * DW_CB * 2; }

template <int K>
__global__ void __launch_bounds__(256, 2)
dwconv_kernel(const bf16* __restrict__ in, long long in_ld, int B, int H, int W, int C,
              const float* __restrict__ wt /*[K*K][C]*/, bf16* __restrict__ out, long long out_ld,
              const bf16* __restrict__ up, int Cup, int Hu, int Wu) {
  constexpr int PAD = (K - 1) / 2;
  constexpr int PH = DW_TY + K - 1, PW = DW_TX + K - 1;
  constexpr int V = DW_CB / 8;                       // 16-byte vectors per pixel
  extern __shared__ __align__(16) uint8_t dw_smem[];
  uint4* patch = reinterpret_cast<uint4*>(dw_smem);  // [PH*PW][V]
  uint4* srcp = patch + PH * PW * V;                 // [DW_SH*DW_SW][V]
  const int cblocks = (C + DW_CB - 1) / DW_CB;
  const int b = blockIdx.z / cblocks;
  const int c0 = (blockIdx.z - b * cblocks) * DW_CB;
  const int y0 = blockIdx.y * DW_TY, x0 = blockIdx.x * DW_TX;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ch = c0 + 2 * lane;
  // this lane's filter taps (issued first: their latency hides under the patch fill)
  float2 wk[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    wk[t] = make_float2(0.0f, 0.0f);
    if (ch < C) wk[t] = __ldg(reinterpret_cast<const float2*>(wt + t * C + ch));
  }
  if (up != nullptr && c0 < Cup) {
    const float sy = (H > 1) ? static_cast<float>(Hu - 1) / static_cast<float>(H - 1) : 0.0f;
    const float sx = (W > 1) ? static_cast<float>(Wu - 1) / static_cast<float>(W - 1) : 0.0f;
    const int ya = max(y0 - PAD, 0), xa = max(x0 - PAD, 0);
    const int ys0 = static_cast<int>(sy * ya), xs0 = static_cast<int>(sx * xa);
    for (int i = threadIdx.x; i < DW_SH * DW_SW * V; i += blockDim.x) {
      const int v = i % V, pix = i / V;
      const int yy = min(ys0 + pix / DW_SW, Hu - 1), xx = min(xs0 + pix % DW_SW, Wu - 1);
      srcp[i] = __ldg(reinterpret_cast<const uint4*>(up + ((static_cast<long long>(b) * Hu + yy) * Wu + xx) * Cup + c0 + v * 8));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PH * PW * V; i += blockDim.x) {
      const int v = i % V, pix = i / V;
      const int px = pix % PW, py = pix / PW;
      const int y = y0 - PAD + py, x = x0 - PAD + px;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const float fy = sy * y, fx = sx * x;
        const int yy0 = static_cast<int>(fy), xx0 = static_cast<int>(fx);
        const int yy1 = min(yy0 + 1, Hu - 1), xx1 = min(xx0 + 1, Wu - 1);
        const float ly = fy - yy0, lx = fx - xx0;
        float a[8], c[8], d[8], e[8], o[8];
        unpack8(srcp[((yy0 - ys0) * DW_SW + (xx0 - xs0)) * V + v], a);
        unpack8(srcp[((yy0 - ys0) * DW_SW + (xx1 - xs0)) * V + v], c);
        unpack8(srcp[((yy1 - ys0) * DW_SW + (xx0 - xs0)) * V + v], d);
        unpack8(srcp[((yy1 - ys0) * DW_SW + (xx1 - xs0)) * V + v], e);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = (1.0f - ly) * ((1.0f - lx) * a[j] + lx * c[j]) + ly * ((1.0f - lx) * d[j] + lx * e[j]);
        val = pack8(o);  // same bf16 rounding point as the materialised concat buffer
      }
      patch[i] = val;
    }
  } else {
    const int coff = c0 - (up != nullptr ? Cup : 0);   // channel offset inside `in`
    constexpr int NLD = (PH * PW * V + 255) / 256;
    constexpr int UNR = 7;
#pragma unroll 1
    for (int u0 = 0; u0 < NLD; u0 += UNR) {
      uint4 val[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int i = threadIdx.x + (u0 + u) * 256;
        const int v = i % V, pix = i / V;
        const int px = pix % PW, py = pix / PW;
        const int y = y0 - PAD + py, x = x0 - PAD + px;
        val[u] = make_uint4(0, 0, 0, 0);
        if (i < PH * PW * V && y >= 0 && y < H && x >= 0 && x < W && c0 + v * 8 < C)
          val[u] = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * H + y) * W + x) * in_ld + coff + v * 8));
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int i = threadIdx.x + (u0 + u) * 256;
        if (i < PH * PW * V) patch[i] = val[u];
      }
    }
  }
  __syncthreads();
  const uint32_t* pw32 = reinterpret_cast<const uint32_t*>(patch) + (warp * DW_PX) * (DW_CB / 2) + lane;
  // packed fp32 FMAs (FFMA2: two channels per instruction, same rounding as two fmaf)
  float2 acc[K][DW_PX];
#pragma unroll
  for (int pr = 0; pr < PH; ++pr) {
    float2 v[DW_PX + K - 1];
#pragma unroll
    for (int c = 0; c < DW_PX + K - 1; ++c) {
      const uint32_t u = pw32[(pr * PW + c) * (DW_CB / 2)];
      v[c] = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
    }
#pragma unroll
    for (int ry = 0; ry < K; ++ry) {
      const int orow = pr - ry;
      if (orow < 0 || orow >= DW_TY) continue;
      const int slot = orow % K;
      if (ry == 0) {
#pragma unroll
        for (int px = 0; px < DW_PX; ++px) acc[slot][px] = make_float2(0.0f, 0.0f);
      }
#pragma unroll
      for (int sx2 = 0; sx2 < K; ++sx2) {
#pragma unroll
        for (int px = 0; px < DW_PX; ++px)
          acc[slot][px] = __ffma2_rn(v[px + sx2], wk[ry * K + sx2], acc[slot][px]);
      }
    }
    const int done = pr - (K - 1);
    if (done >= 0) {
      const int slot = done % K;
      const int y = y0 + done;
      if (y < H && ch < C) {
#pragma unroll
        for (int px = 0; px < DW_PX; ++px) {
          const int x = x0 + warp * DW_PX + px;
          if (x < W)
            *reinterpret_cast<__nv_bfloat162*>(out + ((static_cast<long long>(b) * H + y) * W + x) * out_ld + ch) =
                __floats2bfloat162_rn(acc[slot][px].x, acc[slot][px].y);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ depthwise, TMA-fed persistent
// Same arithmetic and thread mapping as `dwconv_kernel`, but the input patch of a tile is ONE 4-D
// TMA box {64 channels, PW, PH, 1 image} of the NHWC tensor (out-of-bounds zero fill = the conv
// padding and the channel tail), double buffered: a persistent CTA prefetches the patch of its
// next tile while the FMA pipe works on the current one, and no thread spends issue slots on
// address arithmetic or shared-memory stores for the fill.
template <int K>
constexpr int dw_tma_smem_bytes() { return 2 * (DW_TY + K - 1) * (DW_TX + K - 1) * DW_CB * 2 + 128 + 64; }

template <int K>
__global__ void __launch_bounds__(256, 2)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmap, int H, int W, int C,
                  const float* __restrict__ wt /*[K*K][C]*/, bf16* __restrict__ out, long long out_ld,
                  int tiles_x, int tiles_y, int cblocks, int total_tiles) {
  constexpr int PAD = (K - 1) / 2;
  constexpr int PH = DW_TY + K - 1, PW = DW_TX + K - 1;
  constexpr int PATCH_BYTES = PH * PW * DW_CB * 2;
  extern __shared__ uint8_t dwt_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dwt_smem_raw) + 127) & ~static_cast<uintptr_t>(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * PATCH_BYTES);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    sm100::tma_prefetch_desc(&tmap);
    sm100::mbar_init(&bars[0], 1);
    sm100::mbar_init(&bars[1], 1);
    sm100::fence_mbar_init();
  }
  __syncthreads();
  const int spatial = tiles_x * tiles_y;
  int t = blockIdx.x;
  auto issue = [&](int tile, int buf) {
    const int bc = tile / spatial, sp = tile - bc * spatial;
    const int b = bc / cblocks, c0 = (bc - b * cblocks) * DW_CB;
    const int ty = sp / tiles_x, tx = sp - ty * tiles_x;
    sm100::mbar_expect_tx(&bars[buf], PATCH_BYTES);
    sm100::tma_load_4d(base + buf * PATCH_BYTES, &tmap, &bars[buf], c0, tx * DW_TX - PAD, ty * DW_TY - PAD, b);
  };
  if (t < total_tiles && threadIdx.x == 0) issue(t, 0);
  float2 wk[K * K];
  int prev_c0 = -1;
  for (int it = 0; t < total_tiles; ++it, t += gridDim.x) {
    const int buf = it & 1;
    if (t + static_cast<int>(gridDim.x) < total_tiles && threadIdx.x == 0) issue(t + gridDim.x, buf ^ 1);
    const int bc = t / spatial, sp = t - bc * spatial;
    const int b = bc / cblocks, c0 = (bc - b * cblocks) * DW_CB;
    const int ty = sp / tiles_x, tx = sp - ty * tiles_x;
    const int y0 = ty * DW_TY, x0 = tx * DW_TX;
    const int ch = c0 + 2 * lane;
    if (c0 != prev_c0) {
      prev_c0 = c0;
#pragma unroll
      for (int k = 0; k < K * K; ++k) {
        wk[k] = make_float2(0.0f, 0.0f);
        if (ch < C) wk[k] = __ldg(reinterpret_cast<const float2*>(wt + k * C + ch));
      }
    }
    sm100::mbar_wait(&bars[buf], (it >> 1) & 1);
    const uint32_t pw32 = sm100::smem_u32(base + buf * PATCH_BYTES) + ((warp * DW_PX) * (DW_CB / 2) + lane) * 4;
    // output addressing hoisted out of the row loop: one row pointer, four column predicates
    bf16* orow = out + ((static_cast<long long>(b) * H + y0) * W + x0 + warp * DW_PX) * out_ld + ch;
    const long long row_step = static_cast<long long>(W) * out_ld;
    bool xok[DW_PX];
#pragma unroll
    for (int px = 0; px < DW_PX; ++px) xok[px] = (x0 + warp * DW_PX + px < W) && (ch < C);
    // packed fp32 FMAs (FFMA2: two channels per instruction, same rounding as two fmaf)
    float2 acc[K][DW_PX];
#pragma unroll
    for (int pr = 0; pr < PH; ++pr) {
      float2 v[DW_PX + K - 1];
#pragma unroll
      for (int c = 0; c < DW_PX + K - 1; ++c) {
        const uint32_t u = lds_u32(pw32 + (pr * PW + c) * (DW_CB * 2));
        v[c] = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
      }
#pragma unroll
      for (int ry = 0; ry < K; ++ry) {
        const int orw = pr - ry;
        if (orw < 0 || orw >= DW_TY) continue;
        const int slot = orw % K;
        if (ry == 0) {
#pragma unroll
          for (int px = 0; px < DW_PX; ++px) acc[slot][px] = make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int sx2 = 0; sx2 < K; ++sx2) {
#pragma unroll
          for (int px = 0; px < DW_PX; ++px)
            acc[slot][px] = __ffma2_rn(v[px + sx2], wk[ry * K + sx2], acc[slot][px]);
        }
      }
      const int done = pr - (K - 1);
      if (done >= 0) {
        const int slot = done % K;
        if (y0 + done < H) {
#pragma unroll
          for (int px = 0; px < DW_PX; ++px)
            if (xok[px])
              *reinterpret_cast<__nv_bfloat162*>(orow + px * out_ld) = __floats2bfloat162_rn(acc[slot][px].x, acc[slot][px].y);
        }
        orow += row_step;
      }
    }
    __syncthreads();  // every warp is done with this buffer before it is refilled
  }
}

// ------------------------------------------------------------------ BiFPN fast-normalised fusion
// out = (w1 * R(a) + w2 * b [+ w3 * c]) / denom   (decoders/bifpn.py:52-68,106-133), where R is
// the nearest x2 upsampling (top-down), MaxPool2d(3, 2, 1) (bottom-up) or the identity; fp32
// math in the reference's operation order, bf16 NHWC in/out with independent pixel strides.
// mode: 0 identity, 1 nearest-up2 (a is [B][H/2][W/2]), 2 max-pool (a is [B][Ha][Wa]).
__global__ void bifpn_fuse_kernel(const bf16* __restrict__ a, long long a_ld, int mode, int Ha, int Wa,
                                  const bf16* __restrict__ bsrc, long long b_ld,
                                  const bf16* __restrict__ csrc, long long c_ld, float w1, float w2,
                                  float w3, float denom, int B, int H, int W, int C,
                                  bf16* __restrict__ out, long long out_ld) {
  const int cg = C / 8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * H * W * cg;
  if (i >= total) return;
  const int g = static_cast<int>(i % cg);
  long long pix = i / cg;
  const int x = static_cast<int>(pix % W); pix /= W;
  const int y = static_cast<int>(pix % H);
  const int b = static_cast<int>(pix / H);
  float ra[8];
  if (mode == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ra[j] = -INFINITY;
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = 2 * y - 1 + dy;
      if (yy < 0 || yy >= Ha) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = 2 * x - 1 + dx;
        if (xx < 0 || xx >= Wa) continue;
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(a + ((static_cast<long long>(b) * Ha + yy) * Wa + xx) * a_ld + g * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) ra[j] = fmaxf(ra[j], f[j]);
      }
    }
  } else {
    const int yy = mode == 1 ? min(y >> 1, Ha - 1) : y, xx = mode == 1 ? min(x >> 1, Wa - 1) : x;
    unpack8(__ldg(reinterpret_cast<const uint4*>(a + ((static_cast<long long>(b) * Ha + yy) * Wa + xx) * a_ld + g * 8)), ra);
  }
  const long long opix = (static_cast<long long>(b) * H + y) * W + x;
  float fb[8], o[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(bsrc + opix * b_ld + g * 8)), fb);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = __fadd_rn(__fmul_rn(w1, ra[j]), __fmul_rn(w2, fb[j]));
  if (csrc != nullptr) {
    float fc[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(csrc + opix * c_ld + g * 8)), fc);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = __fadd_rn(o[j], __fmul_rn(w3, fc[j]));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = __fdiv_rn(o[j], denom);
  *reinterpret_cast<uint4*>(out + opix * out_ld + g * 8) = pack8(o);
}

// ------------------------------------------------------------------ bilinear, align_corners=True
// CTA = 32 output pixels of one row; 8 lanes cover 64 channels of a pixel with 16-byte vectors
// (128 contiguous bytes per pixel per step) and walk over the channel groups, so the
// interpolation weights and source offsets are computed once per pixel, not once per vector.
__global__ void __launch_bounds__(256)
bilinear_kernel(const bf16* __restrict__ in, long long in_ld, int B, int Hi, int Wi,
                int C, bf16* __restrict__ out, long long out_ld, int out_coff,
                int Ho, int Wo) {
  const int b = blockIdx.z, oy = blockIdx.y;
  const int ox = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int lane8 = threadIdx.x & 7;
  if (ox >= Wo) return;
  (void)B;
  const float sy = (Ho > 1) ? static_cast<float>(Hi - 1) / static_cast<float>(Ho - 1) : 0.0f;
  const float sx = (Wo > 1) ? static_cast<float>(Wi - 1) / static_cast<float>(Wo - 1) : 0.0f;
  const float fy = sy * oy, fx = sx * ox;
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
  const float ly = fy - y0, lx = fx - x0;
  const bf16* base = in + static_cast<long long>(b) * Hi * Wi * in_ld;
  const bf16* p00 = base + (static_cast<long long>(y0) * Wi + x0) * in_ld;
  const bf16* p01 = base + (static_cast<long long>(y0) * Wi + x1) * in_ld;
  const bf16* p10 = base + (static_cast<long long>(y1) * Wi + x0) * in_ld;
  const bf16* p11 = base + (static_cast<long long>(y1) * Wi + x1) * in_ld;
  bf16* op = out + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * out_ld + out_coff;
  for (int g = lane8; g < C / 8; g += 8) {
    float a[8], c[8], d[8], e[8], o[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(p00 + g * 8)), a);
    unpack8(__ldg(reinterpret_cast<const uint4*>(p01 + g * 8)), c);
    unpack8(__ldg(reinterpret_cast<const uint4*>(p10 + g * 8)), d);
    unpack8(__ldg(reinterpret_cast<const uint4*>(p11 + g * 8)), e);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = (1.0f - ly) * ((1.0f - lx) * a[j] + lx * c[j]) + ly * ((1.0f - lx) * d[j] + lx * e[j]);
    *reinterpret_cast<uint4*>(op + g * 8) = pack8(o);
  }
}

// ------------------------------------------------------------------ ASPP image-pool branch
// pooled[b][c] += partial mean over a split of the pixels (fp32, pooled pre-zeroed)
__global__ void avgpool_kernel(const bf16* __restrict__ in, int HW, int C, int splits,
                               float* __restrict__ pooled) {
  const int b = blockIdx.y;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;  // 8-channel group
  if (g * 8 >= C) return;
  const int per = (HW + splits - 1) / splits;
  const int p0 = blockIdx.z * per, p1 = min(HW, p0 + per);
  const bf16* p = in + static_cast<long long>(b) * HW * C + g * 8;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.0f;
  for (int i = p0; i < p1; ++i) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(p + static_cast<long long>(i) * C)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] += f[j];
  }
  const float inv = 1.0f / static_cast<float>(HW);
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(&pooled[b * C + g * 8 + j], s[j] * inv);
}
// v[b][j] = relu(sum_c Wpool[j][c] * pooled[b][c]);  one warp per output
__global__ void gemv_relu_kernel(const float* __restrict__ Wm, const float* __restrict__ x, int K,
                                 int N, int relu, const float* __restrict__ bias0,
                                 float* __restrict__ y) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= N) return;
  float s = 0.0f;
  for (int c = lane; c < K; c += 32) s = fmaf(Wm[static_cast<long long>(j) * K + c], x[b * K + c], s);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (bias0) s += bias0[j];
    y[b * N + j] = relu ? fmaxf(s, 0.0f) : s;
  }
}

// ------------------------------------------------------------------ PointRend
// bilinear x2, align_corners=False (F.interpolate(scale_factor=2)) on (B,h,w) fp32
__global__ void up2_kernel(const float* __restrict__ in, int B, int h, int w, float* __restrict__ out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const int Ho = 2 * h, Wo = 2 * w;
  const long long total = static_cast<long long>(B) * Ho * Wo;
  if (i >= total) return;
  const int ox = static_cast<int>(i % Wo);
  const int oy = static_cast<int>((i / Wo) % Ho);
  const int b = static_cast<int>(i / (static_cast<long long>(Wo) * Ho));
  float fy = (oy + 0.5f) * 0.5f - 0.5f; if (fy < 0.f) fy = 0.f;
  float fx = (ox + 0.5f) * 0.5f - 0.5f; if (fx < 0.f) fx = 0.f;
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float* p = in + static_cast<long long>(b) * h * w;
  out[i] = (1.f - ly) * ((1.f - lx) * p[y0 * w + x0] + lx * p[y0 * w + x1]) +
           ly * ((1.f - lx) * p[y1 * w + x0] + lx * p[y1 * w + x1]);
}

// bilinear x4, align_corners=True on planar fp32 maps: `Interpolate2d(4, 'bilinear',
// align_corners=True)` applied to ctr_hmp / offsets when `interpolate_ins` is set
// (quantization/panoptic_deeplab.py:233-234, blocks.py:74-92). One thread = 4 output pixels of a row.
// one output pixel per thread (see be_resize_linear_u8)
__device__ __forceinline__ void resize_coef(int d, double scale, int n, int clamp_edges, int& s0, int& c0, int& c1) {
  float f = static_cast<float>((d + 0.5) * scale - 0.5);
  int s = static_cast<int>(floorf(f));
  f -= static_cast<float>(s);
  if (clamp_edges) {   // the x pass resets the fraction at the borders; the y pass clips the rows
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n - 1) { f = 0.f; s = n - 1; }
  }
  s0 = s;
  c0 = static_cast<int>(rintf((1.f - f) * 2048.f));
  c1 = static_cast<int>(rintf(f * 2048.f));
}
__global__ void __launch_bounds__(128)
resize_linear_u8_kernel(const uint8_t* __restrict__ vol, long long stride_s, long long stride_y,
                        long long stride_x, int h, int w, int dh, int dw, double scale_y,
                        double scale_x, int area2, uint8_t* __restrict__ out) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y, b = blockIdx.z;
  if (dx >= dw) return;
  const uint8_t* img = vol + static_cast<long long>(b) * stride_s;
  auto at = [&](int y, int x) { return static_cast<int>(__ldg(img + y * stride_y + x * stride_x)); };
  int v;
  if (area2) {
    v = (at(2 * dy, 2 * dx) + at(2 * dy, 2 * dx + 1) + at(2 * dy + 1, 2 * dx) + at(2 * dy + 1, 2 * dx + 1) + 2) >> 2;
  } else {
    int sx, a0, a1, sy, b0, b1;
    resize_coef(dx, scale_x, w, 1, sx, a0, a1);
    resize_coef(dy, scale_y, h, 0, sy, b0, b1);
    const int x1 = min(sx + 1, w - 1);
    const int y0 = min(max(sy, 0), h - 1), y1 = min(max(sy + 1, 0), h - 1);
    const int r0 = at(y0, sx) * a0 + at(y0, x1) * a1;
    const int r1 = at(y1, sx) * a0 + at(y1, x1) * a1;
    v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
  }
  out[(static_cast<long long>(b) * dh + dy) * dw + dx] = static_cast<uint8_t>(v);
}

__global__ void up4_kernel(const float* __restrict__ in, int planes, int h, int w, float* __restrict__ out) {
  const int Ho = 4 * h, Wo = 4 * w;
  const int pl = blockIdx.z, oy = blockIdx.y;
  const int ox0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (ox0 >= Wo) return;
  (void)planes;
  const float sy = (Ho > 1) ? static_cast<float>(h - 1) / static_cast<float>(Ho - 1) : 0.0f;
  const float sx = (Wo > 1) ? static_cast<float>(w - 1) / static_cast<float>(Wo - 1) : 0.0f;
  const float fy = sy * oy;
  const int y0 = static_cast<int>(fy), y1 = min(y0 + 1, h - 1);
  const float ly = fy - y0;
  const float* p0 = in + (static_cast<long long>(pl) * h + y0) * w;
  const float* p1 = in + (static_cast<long long>(pl) * h + y1) * w;
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float fx = sx * (ox0 + j);
    const int x0 = static_cast<int>(fx), x1 = min(x0 + 1, w - 1);
    const float lx = fx - x0;
    o[j] = (1.f - ly) * ((1.f - lx) * __ldg(p0 + x0) + lx * __ldg(p0 + x1)) +
           ly * ((1.f - lx) * __ldg(p1 + x0) + lx * __ldg(p1 + x1));
  }
  *reinterpret_cast<float4*>(out + (static_cast<long long>(pl) * Ho + oy) * Wo + ox0) = make_float4(o[0], o[1], o[2], o[3]);
}

// top-k of uncertainty = -|x|  <=>  k smallest |x|. 3-pass radix select (11/11/10 bits) on the
// fp32 bit pattern of |x|. state per image: [0]=prefix, [1]=remaining k, [2]=ties to take,
// [3]=output cursor, [4]=tie cursor.
__device__ __forceinline__ unsigned absbits(float v) { return __float_as_uint(v) & 0x7fffffffu; }
__global__ void topk_hist_kernel(const float* __restrict__ x, int n, int pass,
                                 const unsigned* __restrict__ state, unsigned* __restrict__ hist) {
  __shared__ unsigned sh[2048];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const unsigned prefix = state[b * 8];
  const int shift = (pass == 0) ? 21 : (pass == 1 ? 10 : 0);
  const unsigned himask = (pass == 0) ? 0u : (pass == 1 ? 0xFFE00000u : 0xFFFFFC00u);
  const unsigned dmask = (pass == 2) ? 0x3FFu : 0x7FFu;
  const float* xb = x + static_cast<long long>(b) * n;
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;  // warp-uniform trip count
    const unsigned k = (i < n) ? absbits(xb[i]) : 0u;
    const bool hit = (i < n) && ((k & himask) == (prefix & himask));
    const unsigned bin = (k >> shift) & dmask;
    // warp-aggregated increment: lanes hitting the same bin elect one leader
    const unsigned active = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const unsigned peers = __match_any_sync(active, bin);
      if ((__ffs(peers) - 1) == (threadIdx.x & 31)) atomicAdd(&sh[bin], __popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[(b * 3 + pass) * 2048 + i], sh[i]);
}
__global__ void __launch_bounds__(1024)
topk_pick_kernel(int pass, unsigned* __restrict__ state, const unsigned* __restrict__ hist) {
  // one CTA per image: parallel inclusive scan over the 2048 bins, then the first bin whose
  // cumulative count reaches `remaining` is the digit of the k-th smallest key
  __shared__ unsigned cum[2048];
  __shared__ unsigned warp_tot[32];
  const int b = blockIdx.x;
  const unsigned remaining = state[b * 8 + 1];
  const unsigned* h = hist + (b * 3 + pass) * 2048;
  const int shift = (pass == 0) ? 21 : (pass == 1 ? 10 : 0);
  const int bins = (pass == 2) ? 1024 : 2048;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const unsigned a0 = (2 * t < bins) ? h[2 * t] : 0u, a1 = (2 * t + 1 < bins) ? h[2 * t + 1] : 0u;
  unsigned v = a0 + a1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp == 0) {
    unsigned w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned u = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += u;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  const unsigned incl = v + (warp ? warp_tot[warp - 1] : 0u);  // inclusive over pairs
  cum[2 * t] = incl - a1;
  cum[2 * t + 1] = incl;
  __syncthreads();
  // bin d is the answer iff cum[d-1] < remaining <= cum[d]
  for (int d = t; d < bins; d += blockDim.x) {
    const unsigned before = d ? cum[d - 1] : 0u;
    if (before < remaining && remaining <= cum[d]) {
      state[b * 8] |= static_cast<unsigned>(d) << shift;
      state[b * 8 + 1] = remaining - before;
      if (pass == 2) state[b * 8 + 2] = remaining - before;
    }
  }
}
__global__ void topk_select_kernel(const float* __restrict__ x, int n, int k, unsigned* __restrict__ state,
                                   int* __restrict__ idx_out) {
  const int b = blockIdx.y;
  const unsigned thr = state[b * 8];
  const unsigned ties = state[b * 8 + 2];
  const float* xb = x + static_cast<long long>(b) * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned key = absbits(xb[i]);
    bool take = key < thr;
    if (key == thr) take = atomicAdd(&state[b * 8 + 4], 1u) < ties;
    if (take) {
      const unsigned pos = atomicAdd(&state[b * 8 + 3], 1u);
      if (pos < static_cast<unsigned>(k)) idx_out[b * k + pos] = i;
    }
  }
}
__global__ void topk_init_kernel(unsigned* state, unsigned* hist, int B, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * 3 * 2048) hist[i] = 0;
  if (i < B * 8) state[i] = ((i & 7) == 1) ? static_cast<unsigned>(k) : 0u;
}

// grid_sample(bilinear, zeros padding, align_corners=False) of the /4 coarse logits and the /4
// feature map at the selected fine-grid points. One warp per point: lanes cover 8 channels each.
// P: [B*k][ldp] bf16 (cols 0..C-1 features, col C coarse, rest zero); also written to P2's
// coarse column so every MLP layer sees the re-concatenated coarse prediction.
__global__ void pr_sample_kernel(const int* __restrict__ idx, int k, int Hf, int Wf,
                                 const float* __restrict__ coarse, const bf16* __restrict__ feat,
                                 int h4, int w4, int C, bf16* __restrict__ P, bf16* __restrict__ P2,
                                 int ldp, float* __restrict__ coarse_pts, int total_pts) {
  const int pt = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pt >= total_pts) return;
  const int b = pt / k;
  const int id = idx[pt];
  const float w_step = 1.0f / static_cast<float>(Wf), h_step = 1.0f / static_cast<float>(Hf);
  const float cx = 0.5f * w_step + w_step * static_cast<float>(id % Wf);
  const float cy = 0.5f * h_step + h_step * static_cast<float>(id / Wf);
  const float gx = 2.0f * cx - 1.0f, gy = 2.0f * cy - 1.0f;
  const float ix = ((gx + 1.0f) * w4 - 1.0f) * 0.5f, iy = ((gy + 1.0f) * h4 - 1.0f) * 0.5f;
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0), x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
  const bool vx0 = x0 >= 0 && x0 < w4, vx1 = x1 >= 0 && x1 < w4, vy0 = y0 >= 0 && y0 < h4, vy1 = y1 >= 0 && y1 < h4;
  const float w00 = (vx0 && vy0) ? wx0 * wy0 : 0.f, w01 = (vx1 && vy0) ? wx1 * wy0 : 0.f;
  const float w10 = (vx0 && vy1) ? wx0 * wy1 : 0.f, w11 = (vx1 && vy1) ? wx1 * wy1 : 0.f;
  const int cx0 = min(max(x0, 0), w4 - 1), cx1 = min(max(x1, 0), w4 - 1);
  const int cy0 = min(max(y0, 0), h4 - 1), cy1 = min(max(y1, 0), h4 - 1);
  const long long pb = static_cast<long long>(b) * h4 * w4;
  const long long o00 = pb + cy0 * w4 + cx0, o01 = pb + cy0 * w4 + cx1, o10 = pb + cy1 * w4 + cx0, o11 = pb + cy1 * w4 + cx1;
  bf16* row = P + static_cast<long long>(pt) * ldp;
  for (int g = lane; g < C / 8; g += 32) {
    float a[8], c[8], d[8], e[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(feat + o00 * C + g * 8), a);
    unpack8(*reinterpret_cast<const uint4*>(feat + o01 * C + g * 8), c);
    unpack8(*reinterpret_cast<const uint4*>(feat + o10 * C + g * 8), d);
    unpack8(*reinterpret_cast<const uint4*>(feat + o11 * C + g * 8), e);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = w00 * a[j] + w01 * c[j] + w10 * d[j] + w11 * e[j];
    *reinterpret_cast<uint4*>(row + g * 8) = pack8(o);
  }
  if (lane == 0) {
    const float cv = w00 * coarse[o00] + w01 * coarse[o01] + w10 * coarse[o10] + w11 * coarse[o11];
    coarse_pts[pt] = cv;
    bf16* row2 = P2 + static_cast<long long>(pt) * ldp;
    row[C] = __float2bfloat16_rn(cv);
    row2[C] = __float2bfloat16_rn(cv);
    for (int j = C + 1; j < ldp; ++j) { row[j] = __float2bfloat16_rn(0.f); row2[j] = __float2bfloat16_rn(0.f); }
  }
}

// predictor (Conv1d(C+1 -> 1)) on the last MLP activation + scatter into the fine logits
__global__ void pr_predict_kernel(const bf16* __restrict__ X, int ldp, int C,
                                  const float* __restrict__ coarse_pts, const float* __restrict__ wp,
                                  float bias, const int* __restrict__ idx, int k, int HWf,
                                  float* __restrict__ sem, int total_pts) {
  const int pt = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pt >= total_pts) return;
  const bf16* row = X + static_cast<long long>(pt) * ldp;
  float s = 0.0f;
  for (int g = lane; g < C / 8; g += 32) {
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(row + g * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(f[j], wp[g * 8 + j], s);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    s += wp[C] * coarse_pts[pt] + bias;
    const int b = pt / k;
    sem[static_cast<long long>(b) * HWf + idx[pt]] = s;
  }
}

}  // namespace mk

extern "C" {

int be_stem(const void* vol, int elem, long long stride_s, long long stride_y, long long stride_x, int s0,
            int B, int h, int w, int H, int W, float mean255, float den, const float* wt,
            const float* bias, __nv_bfloat16* out, cudaStream_t st) {
  dim3 grid((W / 2 + mk::ST - 1) / mk::ST, (H / 2 + mk::ST - 1) / mk::ST, B);
  mk::stem_kernel<<<grid, mk::ST * mk::ST, 0, st>>>(vol, elem, stride_s, stride_y, stride_x, s0, h, w, H, W,
                                                   mean255, den, wt, bias, out);
  return be_check_launch("stem_kernel");
}
int be_stem_pool(const void* vol, int elem, long long stride_s, long long stride_y, long long stride_x,
                 int s0, int B, int h, int w, int H, int W, float mean255, float den, const float* wt,
                 const float* bias, __nv_bfloat16* out, cudaStream_t st) {
  if (H % 4 || W % 4) return be_set_error("stem_pool: padded slice size must be a multiple of 4");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mk::stem_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mk::stem_pool_smem_bytes());
    attr_set = true;
  }
  dim3 grid((W / 4 + mk::SPX - 1) / mk::SPX, (H / 4 + mk::SPY - 1) / mk::SPY, B);
  mk::stem_pool_kernel<<<grid, 256, mk::stem_pool_smem_bytes(), st>>>(vol, elem, stride_s, stride_y, stride_x, s0, h, w,
                                                                      H, W, mean255, den, wt, bias, out);
  return be_check_launch("stem_pool_kernel");
}
int be_maxpool(const __nv_bfloat16* in, int B, int Hi, int Wi, int C, __nv_bfloat16* out, int Ho,
               int Wo, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * Ho * Wo * (C / 8);
  mk::maxpool_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, B, Hi, Wi, C, out, Ho, Wo);
  return be_check_launch("maxpool_kernel");
}
int be_dwconv(const __nv_bfloat16* in, long long in_ld, int B, int H, int W, int C, int k,
              const float* wt, __nv_bfloat16* out, long long out_ld, const __nv_bfloat16* up, int Cup,
              int Hu, int Wu, cudaStream_t st) {
  if (C % 8 || Cup % 8 || in_ld % 8 || out_ld % 2) return be_set_error("dwconv: C must be a multiple of 8");
  // the staged low-resolution source patch is 8 x 14 pixels: needs an upsampling factor >= ~3.3
  if (up != nullptr && (Cup % mk::DW_CB != 0 || 100LL * (Wu - 1) > 31LL * (W - 1) ||
                        100LL * (Hu - 1) > 45LL * (H - 1)))
    return be_set_error("dwconv: fused upsampling needs a scale factor >= 3.3 and Cup % 64 == 0");
  const int cblocks = (C + mk::DW_CB - 1) / mk::DW_CB;
  const int tiles_x = (W + mk::DW_TX - 1) / mk::DW_TX, tiles_y = (H + mk::DW_TY - 1) / mk::DW_TY;
  static bool attr_set = false;
  static int num_sms = 148;
  if (!attr_set) {
    cudaFuncSetAttribute(mk::dwconv_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, mk::dw_smem_bytes<5>());
    cudaFuncSetAttribute(mk::dwconv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mk::dw_smem_bytes<3>());
    cudaFuncSetAttribute(mk::dwconv_tma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, mk::dw_tma_smem_bytes<5>());
    cudaFuncSetAttribute(mk::dwconv_tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, mk::dw_tma_smem_bytes<3>());
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_set = true;
  }
  if (k != 5 && k != 3) return be_set_error("dwconv: only 3x3 and 5x5 kernels are built");
  if (up == nullptr && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    // TMA-fed persistent kernel: patch = one 4-D box of the NHWC input
    CUtensorMap tmap;
    const unsigned long long dims[4] = {static_cast<unsigned long long>(C), static_cast<unsigned long long>(W),
                                        static_cast<unsigned long long>(H), static_cast<unsigned long long>(B)};
    const unsigned long long strides[3] = {static_cast<unsigned long long>(in_ld) * 2,
                                           static_cast<unsigned long long>(in_ld) * 2 * W,
                                           static_cast<unsigned long long>(in_ld) * 2 * W * H};
    const unsigned box[4] = {static_cast<unsigned>(mk::DW_CB), static_cast<unsigned>(mk::DW_TX + k - 1),
                             static_cast<unsigned>(mk::DW_TY + k - 1), 1u};
    const unsigned estr[4] = {1, 1, 1, 1};
    const int rc = convgemm::encode_tmap_bf16(&tmap, in, 4, dims, strides, box, estr, 0);
    if (rc != 0) return be_set_error("dwconv: cuTensorMapEncodeTiled failed");
    const long long total = 1LL * B * cblocks * tiles_x * tiles_y;
    const int grid = static_cast<int>(total < 2LL * num_sms ? total : 2LL * num_sms);
    if (k == 5) mk::dwconv_tma_kernel<5><<<grid, 256, mk::dw_tma_smem_bytes<5>(), st>>>(tmap, H, W, C, wt, out, out_ld, tiles_x, tiles_y, cblocks, static_cast<int>(total));
    else mk::dwconv_tma_kernel<3><<<grid, 256, mk::dw_tma_smem_bytes<3>(), st>>>(tmap, H, W, C, wt, out, out_ld, tiles_x, tiles_y, cblocks, static_cast<int>(total));
    return be_check_launch("dwconv_tma_kernel");
  }
  dim3 grid(tiles_x, tiles_y, B * cblocks);
  if (k == 5) mk::dwconv_kernel<5><<<grid, 256, mk::dw_smem_bytes<5>(), st>>>(in, in_ld, B, H, W, C, wt, out, out_ld, up, Cup, Hu, Wu);
  else mk::dwconv_kernel<3><<<grid, 256, mk::dw_smem_bytes<3>(), st>>>(in, in_ld, B, H, W, C, wt, out, out_ld, up, Cup, Hu, Wu);
  return be_check_launch("dwconv_kernel");
}
int be_bifpn_fuse(const __nv_bfloat16* a, long long a_ld, int mode, int Ha, int Wa,
                  const __nv_bfloat16* b, long long b_ld, const __nv_bfloat16* c, long long c_ld,
                  float w1, float w2, float w3, float denom, int B, int H, int W, int C,
                  __nv_bfloat16* out, long long out_ld, cudaStream_t st) {
  if (C % 8 || a_ld % 8 || b_ld % 8 || c_ld % 8 || out_ld % 8) return be_set_error("bifpn_fuse: channel counts / strides must be multiples of 8");
  if (mode == 1 && (H != 2 * Ha || W != 2 * Wa)) return be_set_error("bifpn_fuse: nearest x2 needs (H, W) == 2 * (Ha, Wa)");
  if (mode == 2 && (H != (Ha + 1) / 2 || W != (Wa + 1) / 2)) return be_set_error("bifpn_fuse: max-pool needs (H, W) == ceil((Ha, Wa) / 2)");
  if (mode == 0 && (H != Ha || W != Wa)) return be_set_error("bifpn_fuse: identity needs equal sizes");
  const long long total = static_cast<long long>(B) * H * W * (C / 8);
  mk::bifpn_fuse_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(a, a_ld, mode, Ha, Wa, b, b_ld, c, c_ld, w1, w2, w3, denom, B, H, W, C, out, out_ld);
  return be_check_launch("bifpn_fuse_kernel");
}
int be_bilinear(const __nv_bfloat16* in, long long in_ld, int B, int Hi, int Wi, int C,
                __nv_bfloat16* out, long long out_ld, int out_coff, int Ho, int Wo, cudaStream_t st) {
  if (C % 8 || in_ld % 8 || out_ld % 8 || out_coff % 8) return be_set_error("bilinear: channel counts / strides must be multiples of 8");
  dim3 grid((Wo + 31) / 32, Ho, B);
  mk::bilinear_kernel<<<grid, 256, 0, st>>>(in, in_ld, B, Hi, Wi, C, out, out_ld, out_coff, Ho, Wo);
  return be_check_launch("bilinear_kernel");
}
// bias_out[b][n] = bias_proj[n] + sum_j Wproj_pool[n][j] * relu(sum_c Wpool[j][c] * mean_pix(in[b,:,c]))
int be_aspp_pool_bias(const __nv_bfloat16* in, int B, int HW, int C, const float* w_pool, int Cmid,
                      const float* w_proj_pool, const float* bias_proj, int N, float* pooled,
                      float* mid, float* bias_out, cudaStream_t st) {
  cudaMemsetAsync(pooled, 0, sizeof(float) * B * C, st);
  const int splits = HW >= 1024 ? 32 : 1;
  mk::avgpool_kernel<<<dim3((C / 8 + 63) / 64, B, splits), 64, 0, st>>>(in, HW, C, splits, pooled);
  mk::gemv_relu_kernel<<<dim3((Cmid + 7) / 8, B), 256, 0, st>>>(w_pool, pooled, C, Cmid, 1, nullptr, mid);
  mk::gemv_relu_kernel<<<dim3((N + 7) / 8, B), 256, 0, st>>>(w_proj_pool, mid, Cmid, N, 0, bias_proj, bias_out);
  return be_check_launch("aspp_pool_bias kernels");
}
// Down-sampling of the slices of a uint8 volume by `resize_by_factor` (empanada/data/utils/
// transforms.py:9-21): cv2.resize(image, (ceil(w/f), ceil(h/f))) with OpenCV's default
// INTER_LINEAR. OpenCV is a third-party dependency that is not vendored in the reference; this
// restates the 8-bit path of its imgproc/resize.cpp: pixel-centre mapping, 11-bit fixed-point
// coefficients (round-half-even), horizontal pass in int32, vertical pass
// ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2)>>2; an exact 2x2 reduction takes the INTER_AREA
// fast path (a+b+c+d+2)>>2, as OpenCV switches to it for INTER_LINEAR.
int be_resize_linear_u8(const uint8_t* vol, long long stride_s, long long stride_y, long long stride_x,
                        int N, int h, int w, int dh, int dw, uint8_t* out, cudaStream_t st) {
  if (dh < 1 || dw < 1 || dh > h || dw > w) return be_set_error("resize: only down-sampling is built");
  const double scale_x = 1.0 / (static_cast<double>(dw) / w), scale_y = 1.0 / (static_cast<double>(dh) / h);
  const int area2 = (h == 2 * dh && w == 2 * dw) ? 1 : 0;
  dim3 grid((dw + 127) / 128, dh, N);
  mk::resize_linear_u8_kernel<<<grid, 128, 0, st>>>(vol, stride_s, stride_y, stride_x, h, w, dh, dw, scale_y, scale_x, area2, out);
  return be_check_launch("resize_linear_u8_kernel");
}
int be_up4(const float* in, int planes, int h, int w, float* out, cudaStream_t st) {
  if (reinterpret_cast<uintptr_t>(out) & 15) return be_set_error("up4: output must be 16-byte aligned");
  dim3 grid((w + 127) / 128, 4 * h, planes);
  mk::up4_kernel<<<grid, 128, 0, st>>>(in, planes, h, w, out);
  return be_check_launch("up4_kernel");
}
int be_up2(const float* in, int B, int h, int w, float* out, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * 4 * h * w;
  mk::up2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, B, h, w, out);
  return be_check_launch("up2_kernel");
}
// idx_out [B][k]; state [B][8] u32, hist [B][3][2048] u32 scratch
int be_topk_uncertain(const float* x, int B, int n, int k, unsigned* state, unsigned* hist,
                      int* idx_out, cudaStream_t st) {
  if (k > n) return be_set_error("topk: k > n");
  mk::topk_init_kernel<<<(B * 3 * 2048 + 255) / 256, 256, 0, st>>>(state, hist, B, k);
  const int blocks = min(64, (n + 1023) / 1024);
  for (int pass = 0; pass < 3; ++pass) {
    mk::topk_hist_kernel<<<dim3(blocks, B), 1024, 0, st>>>(x, n, pass, state, hist);
    mk::topk_pick_kernel<<<B, 1024, 0, st>>>(pass, state, hist);
  }
  mk::topk_select_kernel<<<dim3(blocks, B), 1024, 0, st>>>(x, n, k, state, idx_out);
  return be_check_launch("topk kernels");
}
int be_pr_sample(const int* idx, int B, int k, int Hf, int Wf, const float* coarse,
                 const __nv_bfloat16* feat, int h4, int w4, int C, __nv_bfloat16* P,
                 __nv_bfloat16* P2, int ldp, float* coarse_pts, cudaStream_t st) {
  const int total = B * k;
  mk::pr_sample_kernel<<<(total + 7) / 8, 256, 0, st>>>(idx, k, Hf, Wf, coarse, feat, h4, w4, C, P, P2, ldp, coarse_pts, total);
  return be_check_launch("pr_sample_kernel");
}
int be_pr_predict(const __nv_bfloat16* X, int ldp, int C, const float* coarse_pts, const float* wp,
                  float bias, const int* idx, int B, int k, int HWf, float* sem, cudaStream_t st) {
  const int total = B * k;
  mk::pr_predict_kernel<<<(total + 7) / 8, 256, 0, st>>>(X, ldp, C, coarse_pts, wp, bias, idx, k, HWf, sem, total);
  return be_check_launch("pr_predict_kernel");
}

}  // extern "C"
