// development probe: dfmir_conv_umma_wgrad on a 1x1 convolution, compared with a CPU loop
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "dfmir_b200.h"
int main(int argc, char** argv) {
  const int N = 1, H = argc > 1 ? atoi(argv[1]) : 1, W = argc > 2 ? atoi(argv[2]) : 32, Cin = 128, Cout = argc > 3 ? atoi(argv[3]) : 128;
  const int P = N * H * W;
  std::vector<float> x(P * Cin), dy(P * Cout), dw(Cin * Cout, 0.f), ref(Cin * Cout, 0.f);
  for (int i = 0; i < P * Cin; ++i) x[i] = (float)((i * 7 + 3) % 11 - 5);
  for (int i = 0; i < P * Cout; ++i) dy[i] = (float)((i * 5 + 1) % 7 - 3);
  for (int p = 0; p < P; ++p) for (int ci = 0; ci < Cin; ++ci) for (int co = 0; co < Cout; ++co) ref[ci * Cout + co] += x[p * Cin + ci] * dy[p * Cout + co];
  float *dx_, *ddy, *ddw;
  cudaMalloc(&dx_, x.size() * 4); cudaMalloc(&ddy, dy.size() * 4); cudaMalloc(&ddw, dw.size() * 4);
  cudaMemcpy(dx_, x.data(), x.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(ddy, dy.data(), dy.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(ddw, 0, dw.size() * 4);
  dfmir_conv_desc d{};
  d.nd = 2; d.N = N; d.Cin = Cin; d.Cout = Cout; d.in_shape[0] = H; d.in_shape[1] = W; d.out_shape[0] = H; d.out_shape[1] = W;
  d.kernel[0] = d.kernel[1] = 1; d.pad[0] = d.pad[1] = 0; d.stride = 1; d.act = 0;
  d.x_strides[0] = (long long)H * W * Cin; d.x_strides[1] = (long long)W * Cin; d.x_strides[2] = Cin; d.x_strides[3] = 1;
  d.y_strides[0] = (long long)H * W * Cout; d.y_strides[1] = (long long)W * Cout; d.y_strides[2] = Cout; d.y_strides[3] = 1;
  printf("supported=%d\n", dfmir_conv_umma_wgrad_supported(&d));
  int rc = dfmir_conv_umma_wgrad(dx_, ddy, ddw, nullptr, &d, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  printf("rc=%d err=%s sync=%s\n", rc, dfmir_last_error(), cudaGetErrorString(e));
  cudaMemcpy(dw.data(), ddw, dw.size() * 4, cudaMemcpyDeviceToHost);
  int nz = 0, bad = 0; double maxe = 0;
  for (int i = 0; i < Cin * Cout; ++i) { if (dw[i] != 0) ++nz; double er = fabs(dw[i] - ref[i]); if (er > maxe) maxe = er; if (er > 1e-3) ++bad; }
  printf("nonzero=%d bad=%d/%d maxerr=%g\n", nz, bad, Cin * Cout, maxe);
  for (int ci = 0; ci < 3; ++ci) { for (int co = 0; co < 8; ++co) printf("%8.1f/%8.1f ", dw[ci * Cout + co], ref[ci * Cout + co]); printf("\n"); }
  return 0;
}
