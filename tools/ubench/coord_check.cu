// Does the packed coordinate sequence of lean_tile_packed equal the scalar one for a borderline input? (debug aid)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float yf, float v, float i2y, float Hf, float* out) {
  // scalar
  const float ay = __fadd_rn(yf, v);
  const float ty = __fadd_rn(__fadd_rn(__fmul_rn(ay, i2y), -1.0f), 1.0f);
  const float iy = __fmul_rn(__fmaf_rn(ty, Hf, -1.0f), 0.5f);
  // packed
  const float2 ay2 = __fadd2_rn(make_float2(yf, yf + 2.0f), make_float2(v, v));
  const float2 ty2 = __fadd2_rn(__fadd2_rn(__fmul2_rn(ay2, make_float2(i2y, i2y)), make_float2(-1.0f, -1.0f)), make_float2(1.0f, 1.0f));
  const float2 r2 = __ffma2_rn(ty2, make_float2(Hf, Hf), make_float2(-1.0f, -1.0f));
  const float2 iy2 = __fmul2_rn(r2, make_float2(0.5f, 0.5f));
  out[0] = ay; out[1] = ty; out[2] = iy; out[3] = ay2.x; out[4] = ty2.x; out[5] = r2.x; out[6] = iy2.x; out[7] = floorf(iy2.x);
  out[8] = __fmaf_rn(ty, Hf, -1.0f);
}
int main() {
  float* d; cudaMalloc(&d, 64);
  const float inv = 1.0f / 1079.0f;
  k<<<1, 1>>>(544.0f, 50.94862747192383f, 2.0f * inv, 1080.0f, d);
  float h[9]; cudaMemcpy(h, d, 36, cudaMemcpyDeviceToHost);
  printf("scalar ay %.9g ty %.9g r %.9g iy %.9g | packed ay %.9g ty %.9g r %.9g iy %.9g floor %.9g\n", h[0], h[1], h[8], h[2], h[3], h[4], h[5], h[6], h[7]);
  return 0;
}
