// Throughput of the compare / min-max instruction forms the kNN selector is built from (per SM sub-partition).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float thr) {
  float a[8], r[8];
  int ri[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) { a[u] = threadIdx.x * 0.001f + u; r[u] = 0.f; ri[u] = 0; }
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {        // FSETP + predicated FADD
        asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; @p add.f32 %0, %0, 0f3F800000;}" : "+f"(r[u]) : "f"(a[u]), "f"(thr));
      } else if (MODE == 1) { // ISETP + predicated IADD
        asm volatile("{.reg .pred p; setp.gt.s32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(ri[u]) : "r"(__float_as_int(a[u])), "r"(__float_as_int(thr)));
      } else if (MODE == 2) { // FMNMX
        asm volatile("max.f32 %0, %0, %1;" : "+f"(r[u]) : "f"(a[u]));
      } else if (MODE == 3) { // FADD
        asm volatile("add.f32 %0, %0, %1;" : "+f"(r[u]) : "f"(a[u]));
      } else if (MODE == 4) { // FSETP alone (predicate consumed by a selp outside the count)
        asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; @p mov.b32 %0, %1;}" : "+f"(r[u]) : "f"(a[u]), "f"(thr));
      } else if (MODE == 5) { // FADD + ISETP on the sign + predicated IADD
        float d; asm volatile("sub.f32 %0, %1, %2;" : "=f"(d) : "f"(a[u]), "f"(thr));
        asm volatile("{.reg .pred p; setp.ge.s32 p, %1, 0; @p add.s32 %0, %0, 1;}" : "+r"(ri[u]) : "r"(__float_as_int(d)));
      } else if (MODE == 6) { // IMNMX
        asm volatile("max.s32 %0, %0, %1;" : "+r"(ri[u]) : "r"(__float_as_int(a[u])));
      } else if (MODE == 7) { // predicated-off STS
        asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; @p st.shared.b32 [%3], %1;}" :: "f"(r[u]), "f"(a[u]), "f"(thr), "r"(threadIdx.x * 4));
      }
    }
    thr += 1e-9f;
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int u = 0; u < 8; ++u) s += r[u] + ri[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(t1 - t0);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
template <int MODE>
void run(const char* name, float* d) {
  for (int warps : {4, 8, 16}) {
    k<MODE><<<1, warps * 32>>>(d, 2000, 1e30f);
    cudaDeviceSynchronize();
    k<MODE><<<1, warps * 32>>>(d, 2000, 1e30f);
    float cyc; cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost);
    printf("%-44s warps/SM %2d: %.2f cycles per warp-instr-group per SMSP\n", name, warps, cyc / (2000.0 * 8) / (warps / 4.0));
  }
}
int main() {
  float* d; cudaMalloc(&d, 1 << 20);
  run<0>("FSETP + @p FADD", d);
  run<1>("ISETP + @p IADD", d);
  run<2>("FMNMX", d);
  run<3>("FADD", d);
  run<4>("FSETP + @p MOV", d);
  run<5>("FADD(sub) + ISETP(sign) + @p IADD", d);
  run<6>("IMNMX", d);
  run<7>("FSETP + @!p STS (never taken)", d);
  return 0;
}
