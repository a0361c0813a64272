// Diagnostic: how many spinning kernels on distinct non-blocking streams of ONE
// process actually run concurrently on this GPU (relevant only to the
// in-process multi-rank TESTS; deployment runs one rank per process per GPU).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <thread>
#include <unistd.h>
#include <vector>

__global__ void spin(volatile int* started, volatile int* go, unsigned long long timeout_ns) {
  if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd((int*)started, 1);
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!*go) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) break;
  }
}

int main(int argc, char** argv) {
  int K = argc > 1 ? atoi(argv[1]) : 8;
  int threaded = argc > 2 ? atoi(argv[2]) : 0;
  int with_events = argc > 3 ? atoi(argv[3]) : 0;
  int *started, *go;
  cudaHostAlloc(&started, sizeof(int), cudaHostAllocMapped);
  cudaHostAlloc(&go, sizeof(int), cudaHostAllocMapped);
  *started = 0; *go = 0;
  std::vector<cudaStream_t> st(K);
  std::vector<cudaEvent_t> ev(2 * K);
  for (int i = 0; i < K; ++i) {
    cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    cudaEventCreate(&ev[2 * i]); cudaEventCreate(&ev[2 * i + 1]);
  }
  spin<<<1, 32, 0, st[0]>>>(started, go, 1000); cudaDeviceSynchronize(); *started = 0;  // load module
  auto launch = [&](int i) {
    if (with_events) cudaEventRecord(ev[2 * i], st[i]);
    spin<<<4, 128, 0, st[i]>>>(started, go, 3000000000ull);
    if (with_events) cudaEventRecord(ev[2 * i + 1], st[i]);
  };
  if (threaded) {
    std::vector<std::thread> th;
    for (int i = 0; i < K; ++i) th.emplace_back(launch, i);
    for (auto& t : th) t.join();
  } else {
    for (int i = 0; i < K; ++i) launch(i);
  }
  usleep(500000);
  int s = *started;
  *go = 1;
  cudaDeviceSynchronize();
  const char* e = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
  printf("K=%d threaded=%d events=%d MAX_CONNECTIONS=%s -> %d running concurrently after 0.5 s\n", K, threaded,
         with_events, e ? e : "(unset)", s);
  return 0;
}
