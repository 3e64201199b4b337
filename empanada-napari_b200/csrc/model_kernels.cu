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

namespace mk {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __bfloat162float(h[i].x); f[2 * i + 1] = __bfloat162float(h[i].y); }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// ------------------------------------------------------------------ stem
// One thread = one output pixel x 64 channels. 16x16 output tile per CTA, 37x37 input patch.
constexpr int ST = 16;
constexpr int SP = 2 * ST + 5;
__global__ void __launch_bounds__(ST * ST)
stem_kernel(const uint8_t* __restrict__ vol, long long stride_s, long long stride_y,
            long long stride_x, int s0, int h, int w, int H, int W, float mean255, float den,
            const float* __restrict__ wt /*[49][64]*/, const float* __restrict__ bias,
            bf16* __restrict__ out) {
  __shared__ float patch[SP][SP + 1];
  __shared__ __align__(16) float ws[49 * 64];
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * ST, ox0 = blockIdx.x * ST;
  const int Ho = H / 2, Wo = W / 2;
  const uint8_t* src = vol + static_cast<long long>(s0 + b) * stride_s;
  for (int i = threadIdx.x; i < 49 * 64; i += blockDim.x) ws[i] = wt[i];
  for (int i = threadIdx.x; i < SP * SP; i += blockDim.x) {
    const int py = i / SP, px = i - py * SP;
    const int y = 2 * oy0 - 3 + py, x = 2 * ox0 - 3 + px;
    float v = 0.0f;
    if (y >= 0 && y < h && x >= 0 && x < w)
      v = __fmul_rn(__fsub_rn(static_cast<float>(src[y * stride_y + x * stride_x]), mean255), den);
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
// CTA = 8 rows x 32 pixels x 32 channels. The (8+K-1) x (32+K-1) input patch is staged once in
// shared memory (halo re-read ~1.5x instead of K*K x from L2); a thread produces 4 consecutive
// pixels x 8 channels with the current filter row's taps held in registers.
// Optional fused producer: channels [0, Cup) of the input are the align_corners=True bilinear
// upsampling of a low-resolution NHWC tensor `up` (the decoder's `F.interpolate` + `torch.cat`,
// decoders/panoptic_deeplab.py:76-77), channels [Cup, C) come from `in`; the concatenated
// tensor is never materialised.
constexpr int DW_TY = 8, DW_TX = 32, DW_PX = 4, DW_CG = 4;  // DW_CG groups of 8 channels
constexpr int DW_PSTR = DW_CG + 1;                            // padded pixel stride (uint4 units)
template <int K>
__global__ void __launch_bounds__(256, 3)
dwconv_kernel(const bf16* __restrict__ in, long long in_ld, int B, int H, int W, int C,
              const float* __restrict__ wt /*[K*K][C]*/, bf16* __restrict__ out, long long out_ld,
              const bf16* __restrict__ up, int Cup, int Hu, int Wu) {
  constexpr int PAD = (K - 1) / 2;
  constexpr int PH = DW_TY + K - 1, PW = DW_TX + K - 1;
  __shared__ uint4 patch[PH * PW * DW_PSTR];
  __shared__ __align__(16) float wsm[K * K * 2 * DW_CG * 4];
  const int cgs = C / 8;
  const int cblocks = (cgs + DW_CG - 1) / DW_CG;
  const int b = blockIdx.z / cblocks;
  const int g0 = (blockIdx.z - b * cblocks) * DW_CG;
  const int y0 = blockIdx.y * DW_TY, x0 = blockIdx.x * DW_TX;
  for (int i = threadIdx.x; i < K * K * DW_CG * 8; i += blockDim.x) {
    const int tap = i / (DW_CG * 8), c = i - tap * (DW_CG * 8);
    const int gl = c >> 3, hf = (c >> 2) & 1, e = c & 3;
    const int ch = g0 * 8 + c;
    wsm[((tap * 2 + hf) * DW_CG + gl) * 4 + e] = (ch < C) ? wt[tap * C + ch] : 0.0f;
  }
  const int cup_g = Cup / 8;
  const float sy = (up != nullptr && H > 1) ? static_cast<float>(Hu - 1) / static_cast<float>(H - 1) : 0.0f;
  const float sx = (up != nullptr && W > 1) ? static_cast<float>(Wu - 1) / static_cast<float>(W - 1) : 0.0f;
  if (up == nullptr) {
    // plain path: all global loads of the thread are issued before the first shared store
    constexpr int NLD = (PH * PW * DW_CG + 255) / 256;
    uint4 v[NLD];
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = threadIdx.x + u * 256;
      const int gl = i % DW_CG, pix = i / DW_CG;
      const int px = pix % PW, py = pix / PW;
      const int y = y0 - PAD + py, x = x0 - PAD + px;
      v[u] = make_uint4(0, 0, 0, 0);
      if (i < PH * PW * DW_CG && y >= 0 && y < H && x >= 0 && x < W && g0 + gl < cgs)
        v[u] = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * H + y) * W + x) * in_ld + (g0 + gl) * 8));
    }
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = threadIdx.x + u * 256;
      if (i < PH * PW * DW_CG) patch[(i / DW_CG) * DW_PSTR + (i % DW_CG)] = v[u];
    }
  } else if (g0 < cup_g) {
    // upsampled channel block: stage the (few) low-resolution source pixels this patch needs,
    // then interpolate out of shared memory
    constexpr int SH = 8, SW = 14;
    __shared__ uint4 srcp[SH * SW * DW_CG];
    const int ya = max(y0 - PAD, 0), xa = max(x0 - PAD, 0);
    const int ys0 = static_cast<int>(sy * ya), xs0 = static_cast<int>(sx * xa);
    for (int i = threadIdx.x; i < SH * SW * DW_CG; i += blockDim.x) {
      const int gl = i % DW_CG, pix = i / DW_CG;
      const int yy = min(ys0 + pix / SW, Hu - 1), xx = min(xs0 + pix % SW, Wu - 1);
      srcp[i] = __ldg(reinterpret_cast<const uint4*>(up + ((static_cast<long long>(b) * Hu + yy) * Wu + xx) * Cup + (g0 + gl) * 8));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PH * PW * DW_CG; i += blockDim.x) {
      const int gl = i % DW_CG;
      const int pix = i / DW_CG;
      const int px = pix % PW, py = pix / PW;
      const int y = y0 - PAD + py, x = x0 - PAD + px;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const float fy = sy * y, fx = sx * x;
        const int yy0 = static_cast<int>(fy), xx0 = static_cast<int>(fx);
        const int yy1 = min(yy0 + 1, Hu - 1), xx1 = min(xx0 + 1, Wu - 1);
        const float ly = fy - yy0, lx = fx - xx0;
        float a[8], c[8], d[8], e[8], o[8];
        unpack8(srcp[((yy0 - ys0) * SW + (xx0 - xs0)) * DW_CG + gl], a);
        unpack8(srcp[((yy0 - ys0) * SW + (xx1 - xs0)) * DW_CG + gl], c);
        unpack8(srcp[((yy1 - ys0) * SW + (xx0 - xs0)) * DW_CG + gl], d);
        unpack8(srcp[((yy1 - ys0) * SW + (xx1 - xs0)) * DW_CG + gl], e);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = (1.0f - ly) * ((1.0f - lx) * a[j] + lx * c[j]) + ly * ((1.0f - lx) * d[j] + lx * e[j]);
        v = pack8(o);  // same bf16 rounding point as the materialised concat buffer
      }
      patch[pix * DW_PSTR + gl] = v;
    }
  } else {
    for (int i = threadIdx.x; i < PH * PW * DW_CG; i += blockDim.x) {
      const int gl = i % DW_CG;
      const int pix = i / DW_CG;
      const int px = pix % PW, py = pix / PW;
      const int y = y0 - PAD + py, x = x0 - PAD + px;
      const int g = g0 + gl;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (y >= 0 && y < H && x >= 0 && x < W && g < cgs)
        v = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(b) * H + y) * W + x) * in_ld + (g - cup_g) * 8));
      patch[pix * DW_PSTR + gl] = v;
    }
  }
  __syncthreads();
  const int ty = threadIdx.x >> 5;
  const int xq = (threadIdx.x & 31) >> 2;
  const int gl = threadIdx.x & 3;
  float acc[DW_PX][8];
#pragma unroll
  for (int px = 0; px < DW_PX; ++px)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[px][j] = 0.0f;
  const float4* w4 = reinterpret_cast<const float4*>(wsm);
#pragma unroll
  for (int ry = 0; ry < K; ++ry) {
    float wr[K][8];
#pragma unroll
    for (int sx2 = 0; sx2 < K; ++sx2) {
      const float4 wa = w4[((ry * K + sx2) * 2 + 0) * DW_CG + gl];
      const float4 wb = w4[((ry * K + sx2) * 2 + 1) * DW_CG + gl];
      wr[sx2][0] = wa.x; wr[sx2][1] = wa.y; wr[sx2][2] = wa.z; wr[sx2][3] = wa.w;
      wr[sx2][4] = wb.x; wr[sx2][5] = wb.y; wr[sx2][6] = wb.z; wr[sx2][7] = wb.w;
    }
    const uint4* prow = patch + ((ty + ry) * PW + xq * DW_PX) * DW_PSTR + gl;
#pragma unroll
    for (int c = 0; c < DW_PX + K - 1; ++c) {
      float f[8];
      unpack8(prow[c * DW_PSTR], f);
#pragma unroll
      for (int px = 0; px < DW_PX; ++px) {
        const int sx2 = c - px;
        if (sx2 < 0 || sx2 >= K) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[px][j] = fmaf(f[j], wr[sx2][j], acc[px][j]);
      }
    }
  }
  const int y = y0 + ty;
  if (y < H && g0 + gl < cgs) {
#pragma unroll
    for (int px = 0; px < DW_PX; ++px) {
      const int x = x0 + xq * DW_PX + px;
      if (x < W)
        *reinterpret_cast<uint4*>(out + ((static_cast<long long>(b) * H + y) * W + x) * out_ld + (g0 + gl) * 8) = pack8(acc[px]);
    }
  }
}

// ------------------------------------------------------------------ bilinear, align_corners=True
__global__ void bilinear_kernel(const bf16* __restrict__ in, long long in_ld, int B, int Hi, int Wi,
                                int C, bf16* __restrict__ out, long long out_ld, int out_coff,
                                int Ho, int Wo) {
  const int cg = C / 8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * Ho * Wo * cg;
  if (i >= total) return;
  const int g = static_cast<int>(i % cg);
  long long pix = i / cg;
  const int ox = static_cast<int>(pix % Wo); pix /= Wo;
  const int oy = static_cast<int>(pix % Ho);
  const int b = static_cast<int>(pix / Ho);
  const float sy = (Ho > 1) ? static_cast<float>(Hi - 1) / static_cast<float>(Ho - 1) : 0.0f;
  const float sx = (Wo > 1) ? static_cast<float>(Wi - 1) / static_cast<float>(Wo - 1) : 0.0f;
  const float fy = sy * oy, fx = sx * ox;
  const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
  const int y1 = min(y0 + 1, Hi - 1), x1 = min(x0 + 1, Wi - 1);
  const float ly = fy - y0, lx = fx - x0;
  const bf16* base = in + static_cast<long long>(b) * Hi * Wi * in_ld + g * 8;
  float a[8], c[8], d[8], e[8], o[8];
  unpack8(*reinterpret_cast<const uint4*>(base + (static_cast<long long>(y0) * Wi + x0) * in_ld), a);
  unpack8(*reinterpret_cast<const uint4*>(base + (static_cast<long long>(y0) * Wi + x1) * in_ld), c);
  unpack8(*reinterpret_cast<const uint4*>(base + (static_cast<long long>(y1) * Wi + x0) * in_ld), d);
  unpack8(*reinterpret_cast<const uint4*>(base + (static_cast<long long>(y1) * Wi + x1) * in_ld), e);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    o[j] = (1.0f - ly) * ((1.0f - lx) * a[j] + lx * c[j]) + ly * ((1.0f - lx) * d[j] + lx * e[j]);
  *reinterpret_cast<uint4*>(out + ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * out_ld + out_coff + g * 8) = pack8(o);
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

int be_stem(const uint8_t* vol, long long stride_s, long long stride_y, long long stride_x, int s0,
            int B, int h, int w, int H, int W, float mean255, float den, const float* wt,
            const float* bias, __nv_bfloat16* out, cudaStream_t st) {
  dim3 grid((W / 2 + mk::ST - 1) / mk::ST, (H / 2 + mk::ST - 1) / mk::ST, B);
  mk::stem_kernel<<<grid, mk::ST * mk::ST, 0, st>>>(vol, stride_s, stride_y, stride_x, s0, h, w, H, W,
                                                   mean255, den, wt, bias, out);
  return be_check_launch("stem_kernel");
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
  if (C % 8 || Cup % 8) return be_set_error("dwconv: C must be a multiple of 8");
  // the staged low-resolution source patch is 8 x 14 pixels: needs an upsampling factor >= ~3.3
  if (up != nullptr && (Cup % (8 * mk::DW_CG) != 0 || 100LL * (Wu - 1) > 31LL * (W - 1) ||
                        100LL * (Hu - 1) > 45LL * (H - 1)))
    return be_set_error("dwconv: fused upsampling needs a scale factor >= 3.3 and Cup % 32 == 0");
  const int cblocks = (C / 8 + mk::DW_CG - 1) / mk::DW_CG;
  dim3 grid((W + mk::DW_TX - 1) / mk::DW_TX, (H + mk::DW_TY - 1) / mk::DW_TY, B * cblocks);
  if (k == 5) mk::dwconv_kernel<5><<<grid, 256, 0, st>>>(in, in_ld, B, H, W, C, wt, out, out_ld, up, Cup, Hu, Wu);
  else if (k == 3) mk::dwconv_kernel<3><<<grid, 256, 0, st>>>(in, in_ld, B, H, W, C, wt, out, out_ld, up, Cup, Hu, Wu);
  else return be_set_error("dwconv: only 3x3 and 5x5 kernels are built");
  return be_check_launch("dwconv_kernel");
}
int be_bilinear(const __nv_bfloat16* in, long long in_ld, int B, int Hi, int Wi, int C,
                __nv_bfloat16* out, long long out_ld, int out_coff, int Ho, int Wo, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * Ho * Wo * (C / 8);
  mk::bilinear_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(in, in_ld, B, Hi, Wi, C, out, out_ld, out_coff, Ho, Wo);
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
