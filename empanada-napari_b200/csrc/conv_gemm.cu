// Implicit-GEMM convolution kernel: TMA -> smem (128B swizzle) -> tcgen05.mma -> TMEM ->
// fused epilogue (bias / residual / ReLU / SiLU) -> NHWC bf16 (and optional fp32) stores.
// See conv_gemm.cuh for the layout contract and the reference call sites this replaces.
#include "conv_gemm.cuh"
#include "sm100_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace convgemm {
using namespace sm100;

constexpr int MAX_RES_CHUNKS = MAX_N / 32;  // 16-column chunks per epilogue warp (half a tile)
constexpr int SBIAS_N = 2304;  // >= max Cout (2048) + one tile of zero padding
// barriers + tmem ptr + s_bias + s_head + s_hpart
constexpr int EPI_WARPS = (NUM_THREADS / 32) - 2;   // 16
constexpr int EPI_PARTS = EPI_WARPS / 4;              // column parts per lane quadrant
constexpr int TAIL_BYTES = 256 + 32 + (SBIAS_N + 2 * MAX_N + 2 * TILE_M * (EPI_PARTS - 1)) * 4 + 2 * TILE_M * 8;

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  if (act == ACT_SILU) return v / (1.0f + __expf(-v));
  return v;
}

__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Epilogue role (EPI_WARPS = 16 warps): TMEM -> registers -> bias / residual / activation -> bf16
// NHWC stores (or the fused 1x1 head). Warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32); the
// EPI_PARTS warps of a lane quadrant split the tile's columns. The chunk loop stays rolled so that the
// whole role fits the instruction cache; RES / HEAD are compile-time so unused paths vanish.
template <bool RES, bool HEAD>
__device__ __forceinline__ void epilogue_role(const Params& p, uint32_t s_bias, uint32_t s_head,
                                              float* s_hpart, uint64_t* tfull_bar,
                                              uint64_t* tempty_bar, uint32_t tmem_base, int warp,
                                              int lane, int n_iter, int per_nt, bool wstat) {
  const int ew = warp - 2;
  const int q = warp & 3;
  const int part = ew >> 2;            // column part handled by this warp (0 .. EPI_PARTS-1)
  const int m = q * 32 + lane;
  const int TW = p.TW, TH = p.TH, TB = p.TB, Wo = p.Wo, Ho = p.Ho, Bn = p.B, Cout = p.Cout;
  const int tiles_x = p.tiles_x, tiles_y = p.tiles_y, tiles_n = p.tiles_n, block_n = p.block_n;
  const int act = p.act;
  __nv_bfloat16* const outp = p.out;
  const long long out_ld = p.out_ld;
  const int out_coff = p.out_coff;
  const int tw = m % TW;
  const int th = (m / TW) % TH;
  const int tbi = m / (TW * TH);
  const int cols_part = (((block_n >> 4) + EPI_PARTS - 1) / EPI_PARTS) << 4;
  const int col_begin = min(part * cols_part, block_n);
  const int col_end = min(col_begin + cols_part, block_n);
  const long long hw = static_cast<long long>(Ho) * Wo;
  const bool per_img_bias = (p.bias != nullptr) && (p.bias_img_stride != 0);
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(outp) | (out_ld * 2) | (out_coff * 2)) & 15) == 0;
  const bool res_vec_ok = RES && (((reinterpret_cast<uintptr_t>(p.residual) | (p.res_ld * 2)) & 15) == 0);
  for (int t = 0; t < n_iter; ++t) {
    const int buf = t & 1;
    int nt, mt;
    if (wstat) { nt = t / per_nt; mt = blockIdx.x + (t - nt * per_nt) * gridDim.x; }
    else { const int tile_ = blockIdx.x + t * gridDim.x; nt = tile_ % tiles_n; mt = tile_ / tiles_n; }
    const int tx = mt % tiles_x; mt /= tiles_x;
    const int ty = mt % tiles_y;
    const int tb = mt / tiles_y;
    const int x = tx * TW + tw, y = ty * TH + th, b = tb * TB + tbi;
    const bool valid = (x < Wo) && (y < Ho) && (b < Bn);
    const long long pix = (static_cast<long long>(b) * Ho + y) * Wo + x;
    const int n0 = nt * block_n;
    const float* gbias = per_img_bias ? p.bias + static_cast<long long>(b) * p.bias_img_stride : nullptr;
    const __nv_bfloat16* rrow = RES ? p.residual + pix * p.res_ld + n0 : nullptr;
    uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
    if (RES && valid) {
      // pull this thread's residual segment towards L2 while the MMAs of the tile run, and
      // fetch the first chunk into registers
      for (int c = col_begin; c < col_end && n0 + c < Cout; c += 64) prefetch_l2(rrow + c);
      if (res_vec_ok && n0 + col_begin + 16 <= Cout) {
        r0 = __ldg(reinterpret_cast<const uint4*>(rrow + col_begin));
        r1 = __ldg(reinterpret_cast<const uint4*>(rrow + col_begin) + 1);
      }
    }
    mbar_wait(&tfull_bar[buf], (t >> 1) & 1);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * MAX_N;
    float hacc0 = 0.0f, hacc1 = 0.0f;
#pragma unroll 1
    for (int c0 = col_begin; c0 < col_end; c0 += 16) {
      uint32_t v[16];
      tmem_ld_32x16(taddr + c0, v);
      // residual of the NEXT chunk is requested before this chunk's math
      uint4 nr0 = make_uint4(0, 0, 0, 0), nr1 = nr0;
      const int n = n0 + c0;
      if (RES && valid && res_vec_ok && c0 + 16 < col_end && n + 32 <= Cout) {
        nr0 = __ldg(reinterpret_cast<const uint4*>(rrow + c0 + 16));
        nr1 = __ldg(reinterpret_cast<const uint4*>(rrow + c0 + 16) + 1);
      }
      tmem_ld_wait();
      if (valid && n < Cout) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        if (per_img_bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j) if (n + j < Cout) f[j] += __ldg(gbias + n + j);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // zero padded past Cout
            const float4 bb = lds_f32x4(s_bias + (n + 4 * g) * 4);
            f[4 * g] += bb.x; f[4 * g + 1] += bb.y; f[4 * g + 2] += bb.z; f[4 * g + 3] += bb.w;
          }
        }
        const bool full16 = (n + 16 <= Cout);
        if (RES) {
          if (res_vec_ok && full16) {
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rr[j]);
              f[2 * j] += __bfloat162float(h2.x);
              f[2 * j + 1] += __bfloat162float(h2.y);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n + j < Cout) f[j] += __bfloat162float(rrow[c0 + j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = apply_act(f[j], act);
        if (outp != nullptr) {
          __nv_bfloat16* op = outp + pix * out_ld + out_coff + n;
          if (full16 && vec_ok && ((n & 7) == 0)) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              pk[j] = *reinterpret_cast<uint32_t*>(&h2);
            }
            reinterpret_cast<uint4*>(op)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            reinterpret_cast<uint4*>(op)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n + j < Cout) op[j] = __float2bfloat16_rn(f[j]);
          }
        }
        if (!RES && !HEAD && p.out_f32 != nullptr) {
          if (p.out_f32_planar) {
            float* fp = p.out_f32 + (static_cast<long long>(b) * Cout + n) * hw + (static_cast<long long>(y) * Wo + x);
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n + j < Cout) fp[j * hw] = f[j];
          } else {
            float* fp = p.out_f32 + pix * p.out_f32_ld + n;
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n + j < Cout) fp[j] = f[j];
          }
        }
        if (HEAD) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // zero padded past Cout
            const float4 w0 = lds_f32x4(s_head + (n + 4 * g) * 4);
            const float4 w1 = lds_f32x4(s_head + (MAX_N + n + 4 * g) * 4);
            hacc0 = fmaf(f[4 * g], w0.x, hacc0); hacc0 = fmaf(f[4 * g + 1], w0.y, hacc0);
            hacc0 = fmaf(f[4 * g + 2], w0.z, hacc0); hacc0 = fmaf(f[4 * g + 3], w0.w, hacc0);
            hacc1 = fmaf(f[4 * g], w1.x, hacc1); hacc1 = fmaf(f[4 * g + 1], w1.y, hacc1);
            hacc1 = fmaf(f[4 * g + 2], w1.z, hacc1); hacc1 = fmaf(f[4 * g + 3], w1.w, hacc1);
          }
        }
      }
      r0 = nr0; r1 = nr1;
    }
    if (HEAD) {
      // combine the column parts of each row through shared memory
      if (part > 0) { s_hpart[((part - 1) * TILE_M + m) * 2] = hacc0; s_hpart[((part - 1) * TILE_M + m) * 2 + 1] = hacc1; }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      if (part == 0 && valid) {
#pragma unroll
        for (int pp = 0; pp < EPI_PARTS - 1; ++pp) { hacc0 += s_hpart[(pp * TILE_M + m) * 2]; hacc1 += s_hpart[(pp * TILE_M + m) * 2 + 1]; }
        float* hp = p.head_out + static_cast<long long>(b) * p.head_n * hw + (static_cast<long long>(y) * Wo + x);
        hp[0] = hacc0 + __ldg(p.head_b);
        if (p.head_n > 1) hp[hw] = hacc1 + __ldg(p.head_b + 1);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty_bar[buf]);
  }
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> the 16 destination bytes are zero filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void epi_bar() {
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Staged epilogue: rows live in threads after tcgen05.ld, but HBM wants whole pixel rows per
// warp. Each 64-channel sub-tile (128 px x 128 B) is therefore transposed through a 16 KiB
// XOR-swizzled shared buffer and leaves as fully coalesced 16-byte stores; the residual tile
// arrives the same way (cp.async issued BEFORE the accumulator barrier, so its HBM latency
// hides under the tile's MMAs). obuf: [2][16 KiB], rbuf: [block_n/64][16 KiB], rowpix: [128] i64.
template <bool RES>
__device__ __forceinline__ void epilogue_role_staged(const Params& p, uint32_t s_bias, uint32_t obuf,
                                                     uint32_t rbuf, long long* rowpix,
                                                     uint64_t* tfull_bar, uint64_t* tempty_bar,
                                                     uint32_t tmem_base, int warp, int lane,
                                                     int n_iter, int per_nt, bool wstat) {
  const int ew = warp - 2;
  const int q = warp & 3;
  const int part = ew >> 2;            // 16-column part of the 64-column sub-tile
  const int m = q * 32 + lane;
  const int etid = ew * 32 + lane;     // 0 .. EPI_WARPS*32-1
  const int TW = p.TW, TH = p.TH, TB = p.TB, Wo = p.Wo, Ho = p.Ho, Bn = p.B, Cout = p.Cout;
  const int tiles_x = p.tiles_x, tiles_y = p.tiles_y, tiles_n = p.tiles_n, block_n = p.block_n;
  const int act = p.act;
  __nv_bfloat16* const outp = p.out + p.out_coff;
  const long long out_ld = p.out_ld;
  const int nsub = (block_n + 63) >> 6;
  // tile t -> (n0, rowpix table). rowpix is double buffered: the table of tile t+1 is needed
  // while tile t is processed, to prefetch its residual rows.
  auto tile_n0 = [&](int t) {
    int nt;
    if (wstat) nt = t / per_nt; else nt = (blockIdx.x + t * gridDim.x) % tiles_n;
    return nt * block_n;
  };
  auto fill_rowpix = [&](int t, long long* tab) {
    if (etid < TILE_M) {
      int nt, mt;
      if (wstat) { nt = t / per_nt; mt = blockIdx.x + (t - nt * per_nt) * gridDim.x; }
      else { const int tile_ = blockIdx.x + t * gridDim.x; nt = tile_ % tiles_n; mt = tile_ / tiles_n; }
      const int tx = mt % tiles_x; mt /= tiles_x;
      const int ty = mt % tiles_y;
      const int tb = mt / tiles_y;
      const int tw = etid % TW, th = (etid / TW) % TH, tbi = etid / (TW * TH);
      const int x = tx * TW + tw, y = ty * TH + th, b = tb * TB + tbi;
      long long pix = (static_cast<long long>(b) * Ho + y) * Wo + x;
      if (p.shuffle2x2)  // output pixel (2y + dy, 2x + dx) of the 2Ho x 2Wo map, (dy, dx) = N tile
        pix = ((static_cast<long long>(b) * Ho + y) * 2 + (nt >> 1)) * (2LL * Wo) + 2 * x + (nt & 1);
      tab[etid] = ((x < Wo) && (y < Ho) && (b < Bn)) ? pix : -1;
    }
  };
  auto issue_residual = [&](int j, int n0, const long long* tab) {
    // sub-tile j of the residual: 128 rows x 8 chunks of 16 B, coalesced along channels
#pragma unroll
    for (int u = 0; u < (TILE_M * 8) / (EPI_WARPS * 32); ++u) {
      const int i = etid + u * (EPI_WARPS * 32);
      const int chunk = i & 7, row = i >> 3;
      const long long pix = tab[row];
      const int n = n0 + 64 * j + 8 * chunk;
      const bool ok = (pix >= 0) && (n < Cout);
      const __nv_bfloat16* src = p.residual + (ok ? pix * p.res_ld + n : 0);
      cp_async16(rbuf + j * 16384 + row * 128 + ((chunk ^ (row & 7)) << 4), src, ok);
    }
  };
  if (n_iter > 0) {
    fill_rowpix(0, rowpix);
    epi_bar();
    if (RES) for (int j = 0; j < nsub; ++j) issue_residual(j, tile_n0(0), rowpix);
  }
  for (int t = 0; t < n_iter; ++t) {
    const int buf = t & 1;
    long long* tab = rowpix + (t & 1) * TILE_M;
    long long* tab_next = rowpix + ((t + 1) & 1) * TILE_M;
    const int n0 = tile_n0(t);
    const bool has_next = (t + 1 < n_iter);
    if (has_next) fill_rowpix(t + 1, tab_next);
    if (RES) cp_async_wait_all();   // this tile's residual, requested one tile ago
    mbar_wait(&tfull_bar[buf], (t >> 1) & 1);
    tc_fence_after();
    epi_bar();                      // residual chunks + next rowpix table visible to everyone
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * MAX_N;
    for (int j = 0; j < nsub; ++j) {
      const int c0 = 64 * j + 16 * part;
      const uint32_t ob = obuf + (j & 1) * 16384;
      if (c0 < block_n) {
        uint32_t v[16];
        tmem_ld_32x16(taddr + c0, v);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        const int n = n0 + c0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // s_bias is zero padded past Cout
          const float4 bb = lds_f32x4(s_bias + (n + 4 * g) * 4);
          f[4 * g] += bb.x; f[4 * g + 1] += bb.y; f[4 * g + 2] += bb.z; f[4 * g + 3] += bb.w;
        }
        if (RES) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int chunk = 2 * part + hh;
            const uint4 r = lds_u4(rbuf + j * 16384 + m * 128 + ((chunk ^ (m & 7)) << 4));
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rr[i]);
              f[8 * hh + 2 * i] += __bfloat162float(h2.x);
              f[8 * hh + 2 * i + 1] += __bfloat162float(h2.y);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = apply_act(f[i], act);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * hh + 2 * i], f[8 * hh + 2 * i + 1]);
            pk[i] = *reinterpret_cast<uint32_t*>(&h2);
          }
          const int chunk = 2 * part + hh;
          sts_u4(ob + m * 128 + ((chunk ^ (m & 7)) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
        }
      }
      if (j == nsub - 1) {  // accumulator fully read: hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      }
      epi_bar();
      // rbuf[j] is free again: request the same sub-tile of the NEXT tile's residual now
      if (RES && has_next) issue_residual(j, tile_n0(t + 1), tab_next);
      // coalesced stores: 8 lanes cover one pixel row of the sub-tile (128 B)
#pragma unroll
      for (int u = 0; u < (TILE_M * 8) / (EPI_WARPS * 32); ++u) {
        const int i = etid + u * (EPI_WARPS * 32);
        const int chunk = i & 7, row = i >> 3;
        const long long pix = tab[row];
        const int n = n0 + 64 * j + 8 * chunk;
        if (pix >= 0 && n < Cout) {
          const uint4 val = lds_u4(ob + row * 128 + ((chunk ^ (row & 7)) << 4));
          *reinterpret_cast<uint4*>(outp + pix * out_ld + (p.shuffle2x2 ? n - n0 : n)) = val;
        }
      }
    }
    epi_bar();  // obuf / rowpix[t&1] are reused
  }
  if (RES) cp_async_wait_all();
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int a_bytes = TILE_M * KCHUNK * 2;          // 16 KiB
  const int b_bytes = p.block_n * KCHUNK * 2;       // block_n * 128 B
  const int taps = p.R * p.S;
  const int num_kb = taps * p.kchunks;
  // weight-stationary mode: all K blocks of the current N tile stay resident in shared memory
  // and only the activation tiles stream through the pipeline stages
  const bool wstat = p.b_stationary != 0;
  const int halo = p.halo;
  const int halo_bytes = (p.TH + 2) * HALO_PW * KCHUNK * 2;   // one (TH+2) x 16 px x 64 ch patch
  const int stage_bytes = halo ? halo_bytes : (wstat ? a_bytes : a_bytes + b_bytes);
  const int stages = p.stages;
  uint8_t* bres = smem;                                                    // [num_kb][b_bytes] (wstat)
  uint8_t* stage_base = smem + (wstat ? static_cast<size_t>(num_kb) * b_bytes : 0);

  const int nsub_k = (p.block_n + 63) >> 6;   // one output staging buffer is enough for a single sub-tile
  const int staging_bytes = p.staged ? ((nsub_k > 1 ? 2 : 1) + (p.residual != nullptr ? nsub_k : 0)) * 16384 : 0;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_base + static_cast<size_t>(stages) * stage_bytes + staging_bytes);
  uint64_t* empty_bar = full_bar + stages;
  uint64_t* tfull_bar = empty_bar + stages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint64_t* bfull_bar = tempty_bar + 2;       // weights landed      (wstat)
  uint64_t* bfree_bar = bfull_bar + 1;        // weights consumed    (wstat)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bfree_bar + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_ptr + 4);   // [SBIAS_N]   bias, zero padded
  float* s_head = s_bias + SBIAS_N;                         // [2][MAX_N]  fused head weights
  float* s_hpart = s_head + 2 * MAX_N;                      // [EPI_PARTS-1][128][2] head partial sums
  for (int i = threadIdx.x; i < SBIAS_N; i += blockDim.x)
    s_bias[i] = (p.bias != nullptr && p.bias_img_stride == 0 && i < p.Cout) ? p.bias[i] : 0.0f;
  for (int i = threadIdx.x; i < 2 * MAX_N; i += blockDim.x) {
    const int r = i / MAX_N, c = i - r * MAX_N;
    s_head[i] = (r < p.head_n && c < p.Cout) ? p.head_w[r * p.Cout + c] : 0.0f;
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_WARPS);
    }
    mbar_init(bfull_bar, 1);
    mbar_init(bfree_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_b;
  const int total_tiles = tiles_m * p.tiles_n;
  // Work of this CTA, identical for the three roles. Streaming: tiles blockIdx.x + i*grid with N
  // fastest. Weight-stationary: N tiles outermost, this CTA's M tiles (stride grid) innermost.
  const int per_nt = (static_cast<int>(blockIdx.x) < tiles_m)
                         ? (tiles_m - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
  const int n_iter = wstat ? per_nt * p.tiles_n
                           : ((static_cast<int>(blockIdx.x) < total_tiles)
                                  ? (total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0);
#define TILE_OF(it, nt, mt)                                                  \
  int nt, mt;                                                                \
  if (wstat) { nt = (it) / per_nt; mt = blockIdx.x + ((it) - nt * per_nt) * gridDim.x; } \
  else { const int tile_ = blockIdx.x + (it) * gridDim.x; nt = tile_ % p.tiles_n; mt = tile_ / p.tiles_n; }

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < n_iter; ++it) {
        TILE_OF(it, nt, mt_)
        int mt = mt_;
        const int tx = mt % p.tiles_x; mt /= p.tiles_x;
        const int ty = mt % p.tiles_y;
        const int tb = mt / p.tiles_y;
        const int x_in = tx * p.TW * p.stride - p.pad;
        const int y_in = ty * p.TH * p.stride - p.pad;
        const int b0 = tb * p.TB;
        const int n0 = nt * p.block_n;
        if (wstat && (it % per_nt) == 0) {
          // (re)load the resident weights of N tile nt once every MMA that read the previous
          // ones has retired
          mbar_wait(bfree_bar, (nt & 1) ^ 1);
          mbar_expect_tx(bfull_bar, num_kb * b_bytes);
          for (int kb = 0; kb < num_kb; ++kb) {
            const int tap = kb / p.kchunks, kc = kb - tap * p.kchunks;
            tma_load_2d(bres + static_cast<size_t>(kb) * b_bytes, &tmap_b, bfull_bar,
                        tap * p.Cin + kc * KCHUNK, n0);
          }
        }
        if (halo) {  // one patch per tile: rows y_in .. y_in+TH+1, pixels x_in .. x_in+15
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = stage_base + static_cast<size_t>(stage) * stage_bytes;
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          tma_load_4d(a_dst, &tmap_a, &full_bar[stage], 0, x_in, y_in, b0);
          if (++stage == stages) { stage = 0; phase ^= 1; }
          continue;
        }
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / p.S, s = tap - r * p.S;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* a_dst = stage_base + static_cast<size_t>(stage) * stage_bytes;
            mbar_expect_tx(&full_bar[stage], stage_bytes);
            tma_load_4d(a_dst, &tmap_a, &full_bar[stage], kc * KCHUNK, x_in + s * p.dil,
                        y_in + r * p.dil, b0);
            if (!wstat)
              tma_load_2d(a_dst + a_bytes, &tmap_b, &full_bar[stage], tap * p.Cin + kc * KCHUNK, n0);
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(TILE_M, p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < n_iter; ++t) {
        TILE_OF(t, nt, mt)
        (void)mt;
        const int buf = t & 1;
        if (wstat && (t % per_nt) == 0) {
          mbar_wait(bfull_bar, nt & 1);
          tc_fence_after();
        }
        mbar_wait(&tempty_bar[buf], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * MAX_N;
        if (halo) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t patch = smem_u32(stage_base + static_cast<size_t>(stage) * stage_bytes);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int r = tap / 3, sx = tap - 3 * r;
            // rows of tile row th are pixels sx .. sx+7 of patch row th + r: 8 consecutive 128 B rows;
            // consecutive tile rows are one patch row (16 px = 2 KiB) apart
            const uint64_t a_desc = make_sw128_kmajor_desc_ex(patch + (r * HALO_PW + sx) * 128, HALO_PW * 128,
                                                              halo == 2 ? sx : 0);
            const uint64_t b_desc = make_sw128_kmajor_desc(smem_u32(bres + static_cast<size_t>(tap) * b_bytes));
#pragma unroll
            for (int k = 0; k < KCHUNK / 16; ++k)
              umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (tap | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
          umma_commit(&tfull_bar[buf]);
          if (wstat && ((t + 1) % per_nt) == 0) umma_commit(bfree_bar);
          continue;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(stage_base + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t b_addr = wstat ? smem_u32(bres + static_cast<size_t>(kb) * b_bytes) : a_addr + a_bytes;
          const uint64_t a_desc = make_sw128_kmajor_desc(a_addr);
          const uint64_t b_desc = make_sw128_kmajor_desc(b_addr);
#pragma unroll
          for (int k = 0; k < KCHUNK / 16; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row (encoded >> 4)
            umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[buf]);
        if (wstat && ((t + 1) % per_nt) == 0) umma_commit(bfree_bar);
      }
    }
  } else if (p.staged) {
    const uint32_t obuf = smem_u32(stage_base + static_cast<size_t>(stages) * stage_bytes);
    const uint32_t rbuf = obuf + (nsub_k > 1 ? 2 : 1) * 16384;
    long long* rowpix = reinterpret_cast<long long*>(s_hpart + 2 * TILE_M * (EPI_PARTS - 1));
    if (p.residual != nullptr) epilogue_role_staged<true>(p, smem_u32(s_bias), obuf, rbuf, rowpix, tfull_bar, tempty_bar, tmem_base, warp, lane, n_iter, per_nt, wstat);
    else                       epilogue_role_staged<false>(p, smem_u32(s_bias), obuf, rbuf, rowpix, tfull_bar, tempty_bar, tmem_base, warp, lane, n_iter, per_nt, wstat);
  } else {
    if (p.residual != nullptr) epilogue_role<true, false>(p, smem_u32(s_bias), smem_u32(s_head), s_hpart, tfull_bar, tempty_bar, tmem_base, warp, lane, n_iter, per_nt, wstat);
    else if (p.head_n > 0)     epilogue_role<false, true>(p, smem_u32(s_bias), smem_u32(s_head), s_hpart, tfull_bar, tempty_bar, tmem_base, warp, lane, n_iter, per_nt, wstat);
    else                       epilogue_role<false, false>(p, smem_u32(s_bias), smem_u32(s_head), s_hpart, tfull_bar, tempty_bar, tmem_base, warp, lane, n_iter, per_nt, wstat);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------- host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) !=
            cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const unsigned long long* dims,
                     const unsigned long long* strides_bytes, const unsigned* box,
                     const unsigned* elem_strides, int swizzle_128b) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return -1;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides[i]; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                         const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzle_128b ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -100 - static_cast<int>(r);
}

static void pick_tile_geometry(int B, int Ho, int Wo, int* TW, int* TH, int* TB) {
  // minimise padded pixels; prefer wide rows on ties (longer contiguous TMA segments)
  long long best = -1;
  for (int tw = 128; tw >= 8; tw >>= 1) {
    for (int th = 128 / tw; th >= 1; th >>= 1) {
      const int tb = 128 / (tw * th);
      if (tb > 1 && (th < Ho || tw < Wo)) continue;  // only batch-tile maps that fit one tile
      const long long cost = 1LL * ((Wo + tw - 1) / tw) * tw * ((Ho + th - 1) / th) * th *
                             ((B + tb - 1) / tb) * tb;
      if (best < 0 || cost < best) {
        best = cost; *TW = tw; *TH = th; *TB = tb;
      }
    }
  }
}

int conv_gemm_plan_ex(Launch* L, const __nv_bfloat16* in, long long in_ld, int B, int Hi, int Wi,
                      int Cin, const __nv_bfloat16* w, int Cout, int R, int S, int stride, int dil,
                      int pad, int Ho, int Wo, __nv_bfloat16* out, long long out_ld, int out_coff,
                      float* out_f32, long long out_f32_ld, int out_f32_planar, const float* bias,
                      long long bias_img_stride, const __nv_bfloat16* residual, long long res_ld,
                      int act, const float* head_w, const float* head_b, float* head_out,
                      int head_n, int num_sms, int shuffle2x2);

// Describes one convolution call. Input: NHWC bf16 [B, Hi, Wi, >=Cin] with pixel stride in_ld.
// Weights: [Cout][R*S*Cin] bf16. Output map: Ho x Wo.
int conv_gemm_plan(Launch* L, const __nv_bfloat16* in, long long in_ld, int B, int Hi, int Wi,
                   int Cin, const __nv_bfloat16* w, int Cout, int R, int S, int stride, int dil,
                   int pad, int Ho, int Wo, __nv_bfloat16* out, long long out_ld, int out_coff,
                   float* out_f32, long long out_f32_ld, const float* bias,
                   const __nv_bfloat16* residual, long long res_ld, int act, int num_sms) {
  return conv_gemm_plan_ex(L, in, in_ld, B, Hi, Wi, Cin, w, Cout, R, S, stride, dil, pad, Ho, Wo, out,
                           out_ld, out_coff, out_f32, out_f32_ld, 0, bias, 0, residual, res_ld, act,
                           nullptr, nullptr, nullptr, 0, num_sms, 0);
}

int conv_gemm_plan_ex(Launch* L, const __nv_bfloat16* in, long long in_ld, int B, int Hi, int Wi,
                      int Cin, const __nv_bfloat16* w, int Cout, int R, int S, int stride, int dil,
                      int pad, int Ho, int Wo, __nv_bfloat16* out, long long out_ld, int out_coff,
                      float* out_f32, long long out_f32_ld, int out_f32_planar, const float* bias,
                      long long bias_img_stride, const __nv_bfloat16* residual, long long res_ld,
                      int act, const float* head_w, const float* head_b, float* head_out,
                      int head_n, int num_sms, int shuffle2x2) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return -1;
  if (Cin % 8 != 0 || in_ld % 8 != 0) return -2;
  if (R * S > 1 && Cin % KCHUNK != 0) return -3;
  Params& p = L->p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.Ho = Ho; p.Wo = Wo; p.Cin = Cin; p.Cout = Cout; p.R = R; p.S = S;
  p.stride = stride; p.dil = dil; p.pad = pad;
  pick_tile_geometry(B, Ho, Wo, &p.TW, &p.TH, &p.TB);
  if (p.TW * stride > 256 || p.TH * stride > 256) {
    // TMA box limit: fall back to narrower rows
    while (p.TW * stride > 256) { p.TW >>= 1; p.TH <<= 1; }
  }
  // halo schedule for the L2-bound 3x3 / Cin = 64 convolutions (ResNet layer1)
  int halo_mode = 0;
  {
    const char* e = getenv("B200_CONV_HALO");
    const int want = e ? atoi(e) : 1;   // on by default; B200_CONV_HALO=0 selects the nine-box schedule
    // (nine resident weight blocks + two patches + the output staging must fit: Cout <= 80)
    if (want > 0 && R == 3 && S == 3 && stride == 1 && dil == 1 && pad == 1 && Cin == KCHUNK && Cout <= 80 &&
        Cout % 16 == 0 && Wo >= HALO_TW && Ho >= HALO_TH && residual == nullptr && head_n == 0 && !shuffle2x2)
      halo_mode = want;
  }
  if (halo_mode) { p.TW = HALO_TW; p.TH = HALO_TH; p.TB = 1; }
  p.halo = halo_mode;
  p.tiles_x = (Wo + p.TW - 1) / p.TW;
  p.tiles_y = (Ho + p.TH - 1) / p.TH;
  p.tiles_b = (B + p.TB - 1) / p.TB;
  int bn = ((Cout + 15) / 16) * 16;
  if (bn > MAX_N) bn = (Cout % 256 == 0) ? 256 : ((Cout % 192 == 0) ? 192 : ((Cout % 128 == 0) ? 128 : 256));
  if (shuffle2x2) {
    if (Cout % 4 != 0 || (Cout / 4) % 16 != 0 || Cout / 4 > MAX_N || R * S != 1 || stride != 1 ||
        residual != nullptr || head_n != 0 || out_f32 != nullptr)
      return -7;
    bn = Cout / 4;
  }
  p.shuffle2x2 = shuffle2x2 ? 1 : 0;
  p.block_n = bn;
  p.tiles_n = (Cout + bn - 1) / bn;
  p.kchunks = (Cin + KCHUNK - 1) / KCHUNK;
  if (Cout > SBIAS_N - 256) return -6;
  const int a_bytes_h = TILE_M * KCHUNK * 2, b_bytes_h = bn * KCHUNK * 2;
  const long long num_kb_h = 1LL * R * S * p.kchunks;
  const long long tiles_m_h = 1LL * p.tiles_x * p.tiles_y * p.tiles_b;
  // staged (smem-transposed, coalesced) epilogue: bf16 NHWC output only, 16-byte aligned rows
  const bool can_stage = out != nullptr && out_f32 == nullptr && head_n == 0 && bias_img_stride == 0 &&
                         Cout % 8 == 0 && out_ld % 8 == 0 && out_coff % 8 == 0 &&
                         (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                         (residual == nullptr || (res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0));
  const long long staging_h = can_stage ? (((bn + 63) / 64 > 1 ? 2 : 1) + (residual != nullptr ? (bn + 63) / 64 : 0)) * 16384LL : 0;
  long long smem_budget = 227 * 1024 - 1024 - TAIL_BYTES - staging_h;
  p.staged = can_stage ? 1 : 0;
  {
    // deep-K convolutions are MMA bound and want >= 4 operand stages more than a coalesced
    // epilogue; shallow ones (few K blocks per tile) are epilogue bound
    const long long kb_total = 1LL * R * S * p.kchunks;
    const long long stage_full = TILE_M * KCHUNK * 2 + bn * KCHUNK * 2;
    const bool deep = kb_total > 16;
    if (can_stage && (smem_budget < 2 * stage_full || (deep && smem_budget < 4 * stage_full))) {
      p.staged = 0;
      smem_budget += staging_h;
    }
  }
  if (shuffle2x2 && !p.staged) return -8;
  const long long staging_used = p.staged ? staging_h : 0;
  // weight-stationary when the whole K extent of one N tile fits beside >= 3 activation stages
  // and every CTA reuses it for at least two M tiles
  p.b_stationary = (num_kb_h * b_bytes_h + 3 * a_bytes_h <= smem_budget && tiles_m_h >= 2LL * num_sms) ? 1 : 0;
  const int halo_bytes_h = (p.TH + 2) * HALO_PW * KCHUNK * 2;
  if (halo_mode) {
    if (num_kb_h * b_bytes_h + 2LL * halo_bytes_h > smem_budget) return -9;
    p.b_stationary = 1;
  }
  const int stage_bytes = halo_mode ? halo_bytes_h : (p.b_stationary ? a_bytes_h : a_bytes_h + b_bytes_h);
  int stages = static_cast<int>((smem_budget - (p.b_stationary ? num_kb_h * b_bytes_h : 0)) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return -4;
  p.stages = stages;
  p.out = out; p.out_ld = out_ld; p.out_coff = out_coff;
  p.out_f32 = out_f32; p.out_f32_ld = out_f32_ld;
  p.bias = bias; p.residual = residual; p.res_ld = res_ld; p.act = act;
  p.bias_img_stride = bias_img_stride; p.out_f32_planar = out_f32_planar;
  p.head_w = head_w; p.head_b = head_b; p.head_out = head_out; p.head_n = head_n;
  if (head_n > 0 && (p.tiles_n != 1 || head_n > 2 || Cout % 16 != 0)) return -5;
  L->smem = static_cast<size_t>(stages) * stage_bytes + (p.b_stationary ? num_kb_h * b_bytes_h : 0) + staging_used + 1024 + TAIL_BYTES;
  const long long total = p.b_stationary ? tiles_m_h : tiles_m_h * p.tiles_n;
  L->grid = static_cast<int>(total < num_sms ? total : num_sms);

  {
    cuuint64_t gdim[4] = {static_cast<cuuint64_t>(Cin), static_cast<cuuint64_t>(Wi),
                          static_cast<cuuint64_t>(Hi), static_cast<cuuint64_t>(B)};
    cuuint64_t gstr[3] = {static_cast<cuuint64_t>(in_ld) * 2,
                          static_cast<cuuint64_t>(in_ld) * 2 * Wi,
                          static_cast<cuuint64_t>(in_ld) * 2 * Wi * Hi};
    cuuint32_t box[4] = {KCHUNK, static_cast<cuuint32_t>(p.TW * stride),
                         static_cast<cuuint32_t>(p.TH * stride), static_cast<cuuint32_t>(p.TB)};
    if (halo_mode) { box[1] = HALO_PW; box[2] = static_cast<cuuint32_t>(p.TH + 2); box[3] = 1; }
    cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
    CUresult r = enc(&L->tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                     const_cast<__nv_bfloat16*>(in), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -100 - static_cast<int>(r);
  }
  {
    const long long ktot = 1LL * R * S * Cin;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ktot), static_cast<cuuint64_t>(Cout)};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ktot) * 2};
    cuuint32_t box[2] = {KCHUNK, static_cast<cuuint32_t>(bn)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&L->tmap_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<__nv_bfloat16*>(w), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -200 - static_cast<int>(r);
  }
  return 0;
}

int conv_gemm_launch(const Launch* L, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  conv_gemm_kernel<<<L->grid, NUM_THREADS, L->smem, stream>>>(L->tmap_a, L->tmap_b, L->p);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace convgemm
