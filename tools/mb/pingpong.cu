// Round trip GPU thread -> host thread -> GPU thread through mapped pinned host memory, the exchange a
// Fiat-Shamir round would need if the transcript ran on the host while the sumcheck kernel stays resident.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pingpong pingpong.cu -lpthread ; ./pingpong
// One device thread posts a 48-byte request + sequence number (system-scope release), a host thread
// polls, does `work` SHA-256-like dummy iterations, answers with 16 bytes + sequence number; the device
// thread polls the answer over PCIe.  Both spins are bounded (no hang if the other side dies).
// Variants: answer polled in HOST memory (the device reads across PCIe), or the host writes the answer
// into DEVICE memory through a second mapped window is not available without GPUDirect, so only the first.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <cuda_runtime.h>

struct Slot {
  volatile uint32_t req_seq;
  uint32_t pad0[3];
  volatile uint32_t req[12];
  uint32_t pad1[16];
  volatile uint32_t rsp_seq;
  uint32_t pad2[3];
  volatile uint32_t rsp[4];
};

__global__ void k_ping(Slot* s, int n, long long* cycles, int* ok) {
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int i = 1; i <= n; ++i) {
    for (int k = 0; k < 12; ++k) s->req[k] = acc + k + i;
    __threadfence_system();
    s->req_seq = i;
    long long spin0 = clock64();
    while (s->rsp_seq != (uint32_t)i) {
      if (clock64() - spin0 > 400000000LL) {  // ~0.2 s: give up
        *ok = -i;
        return;
      }
    }
    __threadfence_system();
    acc += s->rsp[0] + s->rsp[3];
  }
  *cycles = clock64() - t0;
  *ok = (int)(acc | 1);
}

int main() {
  Slot* h;
  cudaHostAlloc(&h, sizeof(Slot), cudaHostAllocMapped);
  memset((void*)h, 0, sizeof(Slot));
  long long* d_cyc;
  int* d_ok;
  cudaMalloc(&d_cyc, 8);
  cudaMalloc(&d_ok, 4);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int work : {0, 200, 1000}) {
    const int n = 2000;
    memset((void*)h, 0, sizeof(Slot));
    std::atomic<bool> stop{false};
    std::thread srv([&] {
      uint32_t x = 1;
      for (int i = 1; i <= n && !stop; ++i) {
        auto t0 = std::chrono::steady_clock::now();
        while (h->req_seq != (uint32_t)i) {
          if (stop) return;
          if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) return;
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        for (int k = 0; k < 12; ++k) x += h->req[k];
        for (int w = 0; w < work; ++w) x = (x << 7 | x >> 25) * 2654435761u + w;
        for (int k = 0; k < 4; ++k) h->rsp[k] = x + k;
        std::atomic_thread_fence(std::memory_order_release);
        h->rsp_seq = i;
      }
    });
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_ping<<<1, 1>>>(h, n, d_cyc, d_ok);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    stop = true;
    srv.join();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int ok = 0;
    long long cyc = 0;
    cudaMemcpy(&ok, d_ok, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("{\"host_work_iters\": %d, \"round_trips\": %d, \"us_per_round_trip\": %.3f, \"ok\": %d, \"cuda\": \"%s\"}\n", work,
           n, 1e3 * ms / n, ok, cudaGetErrorString(err));
  }
  return 0;
}
