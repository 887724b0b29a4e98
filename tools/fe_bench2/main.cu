// Field/point microbenchmark: legacy radix-2^25.5 field (tools/fe_bench2/old) vs the current 8x32 saturated field,
// same kernels, outputs compared byte for byte.  Developer tool; results are copied to profiles/.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
extern "C" void new_run(const uint8_t *, int, int, float *, uint8_t *);
extern "C" void old_run(const uint8_t *, int, int, float *, uint8_t *);
extern "C" void kara_run(const uint8_t *, int, int, float *, uint8_t *);
int main(int argc, char **argv) {
  const int it_mul = 1000, it_madd = 256, it_dbl = 512;
  for (int wps : {8, 16, 32}) {
    int n = 148 * wps * 32;
    std::vector<uint8_t> uni(64 * 256);
    srand(12345);
    for (auto &b : uni) b = rand() & 0xff;
    uint8_t *d_uni; cudaMalloc(&d_uni, uni.size()); cudaMemcpy(d_uni, uni.data(), uni.size(), cudaMemcpyHostToDevice);
    std::vector<uint8_t> o_new((size_t)4 * 64 * n), o_old((size_t)4 * 64 * n), o_kara((size_t)4 * 64 * n);
    float ms_new[4], ms_old[4], ms_kara[4];
    old_run(d_uni, n, wps, ms_old, o_old.data());
    new_run(d_uni, n, wps, ms_new, o_new.data());
    kara_run(d_uni, n, wps, ms_kara, o_kara.data());
    const char *names[4] = {"fe_mul", "fe_sq", "ge_madd", "ge_dbl"};
    const double ops[4] = {2.0 * it_mul, 2.0 * it_mul, (double)it_madd, (double)it_dbl};
    for (int t = 0; t < 4; t++) {
      bool same = memcmp(o_new.data() + (size_t)t * 64 * n, o_old.data() + (size_t)t * 64 * n, (size_t)64 * n) == 0;
      bool samek = memcmp(o_kara.data() + (size_t)t * 64 * n, o_old.data() + (size_t)t * 64 * n, (size_t)64 * n) == 0;
      printf("{\"op\": \"%s\", \"warps_per_sm\": %d, \"legacy_Gops\": %.2f, \"sat8x32_Gops\": %.2f, \"sat8x32_karatsuba_Gops\": %.2f, \"outputs_equal\": %s, \"karatsuba_outputs_equal\": %s}\n", names[t], wps,
             ops[t] * n / ms_old[t] / 1e6, ops[t] * n / ms_new[t] / 1e6, ops[t] * n / ms_kara[t] / 1e6, same ? "true" : "false", samek ? "true" : "false");
    }
    cudaFree(d_uni);
  }
  return 0;
}
