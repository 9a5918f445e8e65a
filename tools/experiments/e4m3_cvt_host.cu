#include <cuda_fp8.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f) / 4; fseek(f, 0, SEEK_SET);
  std::vector<float> x(n); fread(x.data(), 4, n, f); fclose(f);
  std::vector<unsigned char> y(n);
  for (long i = 0; i < n; ++i) y[i] = (unsigned char)__nv_cvt_float_to_fp8(x[i], __NV_SATFINITE, __NV_E4M3);
  f = fopen(argv[2], "wb"); fwrite(y.data(), 1, n, f); fclose(f);
  return 0;
}
