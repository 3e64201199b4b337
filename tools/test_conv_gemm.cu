// Standalone self-check of the tcgen05 implicit-GEMM conv kernel against a CPU loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/test_conv_gemm \
//        tools/test_conv_gemm.cu empanada-napari_b200/csrc/conv_gemm.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../empanada-napari_b200/csrc/conv_gemm.cuh"

namespace convgemm {
int conv_gemm_plan(Launch* L, const __nv_bfloat16* in, long long in_ld, int B, int Hi, int Wi,
                   int Cin, const __nv_bfloat16* w, int Cout, int R, int S, int stride, int dil,
                   int pad, int Ho, int Wo, __nv_bfloat16* out, long long out_ld, int out_coff,
                   float* out_f32, long long out_f32_ld, const float* bias,
                   const __nv_bfloat16* residual, long long res_ld, int act, int num_sms);
int conv_gemm_launch(const Launch* L, cudaStream_t stream);
}
using namespace convgemm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)

static uint32_t rng_state = 12345;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f; }

struct Case { const char* name; int B, Hi, Wi, Cin, Cout, R, stride, dil, pad; bool bias, res; int act; bool f32; int full_check; };

static int run_case(const Case& c, int num_sms, bool timing) {
  const int Ho = (c.Hi + 2 * c.pad - c.dil * (c.R - 1) - 1) / c.stride + 1;
  const int Wo = (c.Wi + 2 * c.pad - c.dil * (c.R - 1) - 1) / c.stride + 1;
  const size_t n_in = (size_t)c.B * c.Hi * c.Wi * c.Cin, n_w = (size_t)c.Cout * c.R * c.R * c.Cin;
  const size_t n_out = (size_t)c.B * Ho * Wo * c.Cout;
  std::vector<__nv_bfloat16> h_in(n_in), h_w(n_w), h_res(n_out);
  std::vector<float> f_in(n_in), f_w(n_w), f_res(n_out), h_bias(c.Cout);
  for (size_t i = 0; i < n_in; ++i) { h_in[i] = __float2bfloat16(frand()); f_in[i] = __bfloat162float(h_in[i]); }
  for (size_t i = 0; i < n_w; ++i) { h_w[i] = __float2bfloat16(frand() * 0.25f); f_w[i] = __bfloat162float(h_w[i]); }
  for (size_t i = 0; i < n_out; ++i) { h_res[i] = __float2bfloat16(frand()); f_res[i] = __bfloat162float(h_res[i]); }
  for (int i = 0; i < c.Cout; ++i) h_bias[i] = frand();
  __nv_bfloat16 *d_in, *d_w, *d_out, *d_res; float *d_bias, *d_f32;
  CK(cudaMalloc(&d_in, n_in * 2)); CK(cudaMalloc(&d_w, n_w * 2)); CK(cudaMalloc(&d_out, n_out * 2));
  CK(cudaMalloc(&d_res, n_out * 2)); CK(cudaMalloc(&d_bias, c.Cout * 4)); CK(cudaMalloc(&d_f32, n_out * 4));
  CK(cudaMemcpy(d_in, h_in.data(), n_in * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_w, h_w.data(), n_w * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_res, h_res.data(), n_out * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bias, h_bias.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_out, 0xFF, n_out * 2)); CK(cudaMemset(d_f32, 0xFF, n_out * 4));
  Launch L;
  int rc = conv_gemm_plan(&L, d_in, c.Cin, c.B, c.Hi, c.Wi, c.Cin, d_w, c.Cout, c.R, c.R, c.stride, c.dil, c.pad,
                          Ho, Wo, d_out, c.Cout, 0, c.f32 ? d_f32 : nullptr, c.Cout, c.bias ? d_bias : nullptr,
                          c.res ? d_res : nullptr, c.Cout, c.act, num_sms);
  if (rc != 0) { printf("[%s] plan failed rc=%d\n", c.name, rc); return 1; }
  printf("[%s] Ho=%d Wo=%d tile TW=%d TH=%d TB=%d bn=%d stages=%d wstat=%d halo=%d grid=%d smem=%zu\n", c.name, Ho, Wo, L.p.TW, L.p.TH, L.p.TB, L.p.block_n, L.p.stages, L.p.b_stationary, L.p.halo, L.grid, L.smem);
  rc = conv_gemm_launch(&L, 0);
  if (rc != 0) { printf("[%s] launch failed rc=%d\n", c.name, rc); return 1; }
  CK(cudaDeviceSynchronize());
  if (timing) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) conv_gemm_launch(&L, 0);
    CK(cudaEventRecord(e0));
    const int iters = 10;
    for (int i = 0; i < iters; ++i) conv_gemm_launch(&L, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
    double flops = 2.0 * c.B * Ho * Wo * (double)c.Cout * c.R * c.R * c.Cin;
    printf("[%s] %.3f ms  %.1f TFLOP/s\n", c.name, ms, flops / ms * 1e-9);
  }
  std::vector<__nv_bfloat16> h_out(n_out); std::vector<float> h_f32(n_out);
  CK(cudaMemcpy(h_out.data(), d_out, n_out * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_f32.data(), d_f32, n_out * 4, cudaMemcpyDeviceToHost));
  // CPU check (all outputs or a strided sample)
  size_t step = c.full_check ? 1 : 9973;
  double max_err = 0; size_t bad = 0, checked = 0;
  for (size_t o = 0; o < n_out; o += step) {
    int n = o % c.Cout; size_t pix = o / c.Cout; int x = pix % Wo; int y = (pix / Wo) % Ho; int b = pix / ((size_t)Wo * Ho);
    double acc = 0;
    for (int r = 0; r < c.R; ++r) for (int s = 0; s < c.R; ++s) {
      int yi = y * c.stride - c.pad + r * c.dil, xi = x * c.stride - c.pad + s * c.dil;
      if (yi < 0 || yi >= c.Hi || xi < 0 || xi >= c.Wi) continue;
      const float* ip = &f_in[(((size_t)b * c.Hi + yi) * c.Wi + xi) * c.Cin];
      const float* wp = &f_w[((size_t)n * c.R * c.R + r * c.R + s) * c.Cin];
      for (int k = 0; k < c.Cin; ++k) acc += (double)ip[k] * wp[k];
    }
    if (c.bias) acc += h_bias[n];
    if (c.res) acc += f_res[o];
    if (c.act == ACT_RELU) acc = acc > 0 ? acc : 0;
    if (c.act == ACT_SILU) acc = acc / (1 + exp(-acc));
    double got = __bfloat162float(h_out[o]);
    double err = fabs(got - acc), tol = 0.02 + 0.01 * fabs(acc);
    if (c.f32) { double e2 = fabs(h_f32[o] - acc); if (e2 > 1e-3 + 1e-3 * fabs(acc)) { if (bad < 5) printf("  f32 mismatch o=%zu got %f want %f\n", o, h_f32[o], acc); ++bad; } }
    if (!(err <= tol)) { if (bad < 5) printf("  mismatch o=%zu (b%d y%d x%d n%d) got %f want %f\n", o, b, y, x, n, got, acc); ++bad; }
    if (err > max_err) max_err = err; ++checked;
  }
  printf("[%s] checked %zu outputs, max_err %.4f, bad %zu -> %s\n", c.name, checked, max_err, bad, bad ? "FAIL" : "ok");
  cudaFree(d_in); cudaFree(d_w); cudaFree(d_out); cudaFree(d_res); cudaFree(d_bias); cudaFree(d_f32);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  const int sms = prop.multiProcessorCount;
  Case cases[] = {
    {"1x1 gemm", 1, 64, 64, 256, 256, 1, 1, 1, 0, false, false, ACT_NONE, false, 1},
    {"1x1 bias relu f32", 1, 32, 32, 64, 64, 1, 1, 1, 0, true, false, ACT_RELU, true, 1},
    {"3x3 d2", 2, 64, 64, 128, 256, 3, 1, 2, 2, true, false, ACT_RELU, false, 1},
    {"3x3 s2", 1, 64, 64, 64, 128, 3, 2, 1, 1, true, false, ACT_RELU, false, 1},
    {"1x1 s2", 1, 64, 64, 256, 512, 1, 2, 1, 0, true, false, ACT_NONE, false, 1},
    {"3x3 d6 odd 65", 1, 65, 65, 64, 64, 3, 1, 6, 6, true, true, ACT_RELU, false, 1},
    {"1x1 k288 res", 1, 48, 40, 288, 256, 1, 1, 1, 0, true, true, ACT_RELU, false, 1},
    {"1x1 cout2 f32", 1, 32, 32, 256, 2, 1, 1, 1, 0, true, false, ACT_NONE, true, 1},
    {"3x3 tiny 8x8 B4 silu", 4, 8, 8, 128, 128, 3, 1, 1, 1, true, false, ACT_SILU, false, 1},
    {"3x3 s2 odd 33", 1, 33, 33, 64, 64, 3, 2, 1, 1, false, false, ACT_NONE, false, 1},
    {"1x1 wide 256x256", 1, 256, 256, 64, 256, 1, 1, 1, 0, true, false, ACT_RELU, false, 0},
    {"wstat 1x1 res 64->256", 3, 128, 128, 64, 256, 1, 1, 1, 0, true, true, ACT_RELU, false, 1},
    {"wstat 1x1 128->512 (2 n-tiles)", 3, 128, 128, 128, 512, 1, 1, 1, 0, true, true, ACT_RELU, false, 1},
    {"wstat 3x3 64->64", 3, 128, 128, 64, 64, 3, 1, 1, 1, true, false, ACT_RELU, false, 1},
    {"3x3 64->64 odd 50x37", 2, 50, 37, 64, 64, 3, 1, 1, 1, true, false, ACT_RELU, false, 1},
    {"wstat 1x1 k288 odd 250x130", 3, 250, 130, 288, 256, 1, 1, 1, 0, true, false, ACT_RELU, false, 1},
  };
  int fails = 0;
  const int only = (argc > 2) ? atoi(argv[2]) : -1;
  if (only < 0 || argc > 3) for (const Case& c : cases) fails += run_case(c, sms, false);
  if (argc > 1) {
    Case perf[] = {
      {"ASPP 3x3 d6 2048->512 B8", 8, 64, 64, 2048, 512, 3, 1, 6, 6, true, false, ACT_RELU, false, 0},
      {"ASPP 3x3 d2 2048->256 B8", 8, 64, 64, 2048, 256, 3, 1, 2, 2, true, false, ACT_RELU, false, 0},
      {"l4 3x3 d2 512->512 B8", 8, 64, 64, 512, 512, 3, 1, 2, 2, true, false, ACT_RELU, false, 0},
      {"l3 1x1 1024->256 B8", 8, 64, 64, 1024, 256, 1, 1, 1, 0, true, false, ACT_RELU, false, 0},
      {"l1 1x1 64->256 B8 256^2", 8, 256, 256, 64, 256, 1, 1, 1, 0, true, true, ACT_RELU, false, 0},
      {"head 1x1 256->256 B8 256^2", 8, 256, 256, 256, 256, 1, 1, 1, 0, true, false, ACT_RELU, false, 0},
      {"l1 3x3 64->64 B16 256^2", 16, 256, 256, 64, 64, 3, 1, 1, 1, true, false, ACT_RELU, false, 0},
    };
    int idx = 0;
    for (const Case& c : perf) { if (only < 0 || only == idx) fails += run_case(c, sms, true); ++idx; }
  }
  printf("TOTAL FAILS %d\n", fails);
  return fails ? 1 : 0;
}
