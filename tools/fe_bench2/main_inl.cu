// fe_mul / fe_sq as real device functions (BP_FE_CALL=1, the product build) vs fully inlined: same kernels, outputs compared
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
extern "C" void new_run(const uint8_t *, int, int, float *, uint8_t *);
extern "C" void inl_run(const uint8_t *, int, int, float *, uint8_t *);
int main() {
  const int it_mul = 1000, it_madd = 256, it_dbl = 512;
  for (int wps : {8, 16, 32}) {
    int n = 148 * wps * 32;
    std::vector<uint8_t> uni(64 * 256);
    srand(12345);
    for (auto &b : uni) b = rand() & 0xff;
    uint8_t *d_uni; cudaMalloc(&d_uni, uni.size()); cudaMemcpy(d_uni, uni.data(), uni.size(), cudaMemcpyHostToDevice);
    std::vector<uint8_t> o_a((size_t)4 * 64 * n), o_b((size_t)4 * 64 * n);
    float ms_a[4], ms_b[4];
    new_run(d_uni, n, wps, ms_a, o_a.data());
    inl_run(d_uni, n, wps, ms_b, o_b.data());
    const char *names[4] = {"fe_mul", "fe_sq", "ge_madd", "ge_dbl"};
    const double ops[4] = {2.0 * it_mul, 2.0 * it_mul, (double)it_madd, (double)it_dbl};
    for (int t = 0; t < 4; t++) {
      bool same = memcmp(o_a.data() + (size_t)t * 64 * n, o_b.data() + (size_t)t * 64 * n, (size_t)64 * n) == 0;
      printf("{\"op\": \"%s\", \"warps_per_sm\": %d, \"call_Gops\": %.2f, \"inline_Gops\": %.2f, \"speedup\": %.3f, \"outputs_equal\": %s}\n", names[t], wps,
             ops[t] * n / ms_a[t] / 1e6, ops[t] * n / ms_b[t] / 1e6, ms_a[t] / ms_b[t], same ? "true" : "false");
    }
    cudaFree(d_uni);
  }
  return 0;
}
