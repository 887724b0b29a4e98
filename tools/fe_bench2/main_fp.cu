// integer field (csrc/fe25519.h) vs FP64-pipe field (csrc/fed25519.h): throughput and byte equality of the outputs
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
extern "C" void new_run(const uint8_t *, int, int, float *, uint8_t *);
extern "C" void fp_run(const uint8_t *, int, int, float *, uint8_t *);
extern "C" void fpi_run(const uint8_t *, int, int, float *, uint8_t *);
extern "C" void inl_run(const uint8_t *, int, int, float *, uint8_t *);
int main() {
  const int it_mul = 1000, it_madd = 256;
  for (int wps : {8, 16, 32}) {
    int n = 148 * wps * 32;
    std::vector<uint8_t> uni(64 * 256);
    srand(12345);
    for (auto &b : uni) b = rand() & 0xff;
    uint8_t *d_uni; cudaMalloc(&d_uni, uni.size()); cudaMemcpy(d_uni, uni.data(), uni.size(), cudaMemcpyHostToDevice);
    std::vector<uint8_t> o_int((size_t)4 * 64 * n), o_fp((size_t)4 * 64 * n), o_fpi((size_t)4 * 64 * n), o_inl((size_t)4 * 64 * n);
    float ms_int[4], ms_fp[4], ms_fpi[4], ms_inl[4];
    new_run(d_uni, n, wps, ms_int, o_int.data());
    inl_run(d_uni, n, wps, ms_inl, o_inl.data());
    fp_run(d_uni, n, wps, ms_fp, o_fp.data());
    fpi_run(d_uni, n, wps, ms_fpi, o_fpi.data());
    const char *names[3] = {"fe_mul", "fe_sq", "ge_madd"};
    const double ops[3] = {2.0 * it_mul, 2.0 * it_mul, (double)it_madd};
    for (int t = 0; t < 3; t++) {
      bool same = memcmp(o_int.data() + (size_t)t * 64 * n, o_fp.data() + (size_t)t * 64 * n, (size_t)64 * n) == 0 &&
                  memcmp(o_int.data() + (size_t)t * 64 * n, o_fpi.data() + (size_t)t * 64 * n, (size_t)64 * n) == 0 &&
                  memcmp(o_int.data() + (size_t)t * 64 * n, o_inl.data() + (size_t)t * 64 * n, (size_t)64 * n) == 0;
      printf("{\"op\": \"%s\", \"warps_per_sm\": %d, \"int8x32_call_Gops\": %.2f, \"int8x32_inline_Gops\": %.2f, \"fp64_5x51_call_Gops\": %.2f, \"fp64_5x51_inline_Gops\": %.2f, \"outputs_equal\": %s}\n", names[t], wps,
             ops[t] * n / ms_int[t] / 1e6, ops[t] * n / ms_inl[t] / 1e6, ops[t] * n / ms_fp[t] / 1e6, ops[t] * n / ms_fpi[t] / 1e6, same ? "true" : "false");
    }
    cudaFree(d_uni);
  }
  return 0;
}
