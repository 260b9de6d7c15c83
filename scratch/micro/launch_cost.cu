// host cost of a kernel launch by parameter size, of cudaFuncSetAttribute, of event create/record/destroy (B200 box)
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
template <int N> struct P { char b[N]; };
template <int N> __global__ void k(const __grid_constant__ P<N> p) { if (p.b[0] == 77 && threadIdx.x == 999) printf("x"); }
template <int N> double run(int smem, bool attr, bool event) {
  P<N> p{}; cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int i = 0; i < 200; ++i) k<N><<<2, 512, smem, st>>>(p);
  cudaStreamSynchronize(st);
  const int n = 800;
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; ++i) {
    if (attr) cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<N><<<2, 512, smem, st>>>(p);
    if (event) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); cudaEventRecord(e, st); cudaEventDestroy(e); }
  }
  auto t1 = std::chrono::steady_clock::now();
  cudaStreamSynchronize(st);
  return std::chrono::duration<double, std::micro>(t1 - t0).count() / n;
}
int main() {
  printf("params 256 B:   %.2f us/launch\n", run<256>(0, false, false));
  printf("params 2 KB:    %.2f us/launch\n", run<2048>(0, false, false));
  printf("params 4 KB:    %.2f us/launch\n", run<4096>(0, false, false));
  printf("params 11 KB:   %.2f us/launch\n", run<11264>(0, false, false));
  printf("params 11 KB + 160 KB smem:              %.2f us/launch\n", run<11264>(160 * 1024, false, false));
  printf("params 11 KB + 160 KB smem + attr:       %.2f us/launch\n", run<11264>(160 * 1024, true, false));
  printf("params 11 KB + 160 KB smem + attr + ev:  %.2f us/launch\n", run<11264>(160 * 1024, true, true));
  printf("params 2 KB + 160 KB smem:               %.2f us/launch\n", run<2048>(160 * 1024, false, false));
  return 0;
}
