// Stand-alone micro-benchmark behind the "persistent batch-1 decode step" decision (not part of libdsocr.so; results:
// profiles/r02_microbench_step.log):
// what does one *dependent phase* of a batch-1 decode step cost as
//   (a) a kernel node of a CUDA graph (76 dependent launches per token today, ~7 us each),
//   (b) the same with programmatic dependent launch,
//   (c) a phase of ONE persistent cooperative kernel separated by a grid-wide barrier,
// and how much of (a) is instruction fetch: (d) the phases alternate between 8 different large kernels (cold
// instruction cache at every launch, as in the real step) or reuse one small kernel.
// Every phase does the same small amount of work: each of 148 x 256 threads reads 64 B of a 2.4 MB buffer that the
// previous phase wrote, and writes 4 B - so any time beyond ~2 us is launch / barrier / fetch overhead.
// Build + run on a B200:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mb microbench_step.cu && ./mb
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kBlocks = 148, kThreads = 256, kPhases = 76;

__device__ __forceinline__ float phase_work(const float* __restrict__ in, int phase) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const float4* p = reinterpret_cast<const float4*>(in) + (size_t)((gid * 4 + phase) % (kBlocks * kThreads * 4));
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float4 v = __ldcg(p + i * kBlocks * kThreads); s += v.x + v.y + v.z + v.w; }
  return s;
}

__global__ void small_phase(const float* in, float* out, int phase, int pdl) {
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  out[blockIdx.x * blockDim.x + threadIdx.x] = phase_work(in, phase);
}

// same work behind 5000 straight-line FFMAs (80 KB of code) that every warp executes once (what an unrolled GEMV body looks like)
template <int ID>
__global__ void fat_phase(const float* in, float* out, int phase) {
  float s = phase_work(in, phase);
  float a = s * 1e-30f + (float)ID;
#pragma unroll
  for (int i = 0; i < 5000; ++i) a = fmaf(a, 1.0000001f, 1e-9f * (float)(i + ID));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + a * 1e-30f;
}

__global__ void persistent_step(float* a, float* b, unsigned* bar, int phases) {
  cg::grid_group grid = cg::this_grid();
  for (int p = 0; p < phases; ++p) {
    const float* in = (p & 1) ? b : a;
    float* out = (p & 1) ? a : b;
    out[blockIdx.x * blockDim.x + threadIdx.x] = phase_work(in, p);
    grid.sync();
  }
  (void)bar;
}

// hand-rolled barrier: one arrive per block, generation flag
__global__ void persistent_step_flag(float* a, float* b, unsigned* bar, int phases) {
  __shared__ unsigned gen_s;
  unsigned gen = 0;
  for (int p = 0; p < phases; ++p) {
    const float* in = (p & 1) ? b : a;
    float* out = (p & 1) ? a : b;
    out[blockIdx.x * blockDim.x + threadIdx.x] = phase_work(in, p);
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      ++gen;
      if (atomicAdd(&bar[0], 1u) == gen * gridDim.x - 1) {
        atomicExch(&bar[1], gen);  // last block of this generation releases everyone
      } else {
        while (atomicAdd(&bar[1], 0u) < gen) { }
      }
      __threadfence();
      gen_s = gen;
    }
    __syncthreads();
    gen = gen_s;
  }
}

typedef void (*FatFn)(const float*, float*, int);

static float time_graph(cudaStream_t s, cudaGraphExec_t g, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 5; ++i) CK(cudaGraphLaunch(g, s));
  CK(cudaEventRecord(e0, s));
  for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(g, s));
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms * 1e3f / reps / kPhases;  // us per phase
}

int main() {
  const size_t n = (size_t)kBlocks * kThreads * 16 + 64;
  float *a, *b;
  unsigned* bar;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&bar, 64));
  CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  const FatFn fat[8] = {fat_phase<0>, fat_phase<1>, fat_phase<2>, fat_phase<3>, fat_phase<4>, fat_phase<5>, fat_phase<6>, fat_phase<7>};

  for (int mode = 0; mode < 4; ++mode) {  // 0 small kernels, 1 small + PDL, 2 eight alternating fat kernels, 3 one fat kernel
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int p = 0; p < kPhases; ++p) {
      const float* in = (p & 1) ? b : a;
      float* out = (p & 1) ? a : b;
      if (mode == 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kBlocks); cfg.blockDim = dim3(kThreads); cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, small_phase, in, out, p, 1));
      } else if (mode == 0) {
        small_phase<<<kBlocks, kThreads, 0, s>>>(in, out, p, 0);
      } else {
        fat[mode == 2 ? (p % 8) : 0]<<<kBlocks, kThreads, 0, s>>>(in, out, p);
      }
    }
    CK(cudaStreamEndCapture(s, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    const char* names[4] = {"graph, 76 small kernels", "graph, 76 small kernels, PDL", "graph, 8 alternating 80 KB straight-line kernels",
                            "graph, one 80 KB straight-line kernel"};
    printf("%-48s %6.2f us per phase\n", names[mode], time_graph(s, ge, 200));
    CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
  }

  for (int mode = 0; mode < 2; ++mode) {
    int phases = kPhases;
    void* args[4] = {&a, &b, &bar, &phases};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const void* fn = mode == 0 ? (const void*)persistent_step : (const void*)persistent_step_flag;
    for (int i = 0; i < 3; ++i) { CK(cudaMemsetAsync(bar, 0, 64, s)); CK(cudaLaunchCooperativeKernel(fn, dim3(kBlocks), dim3(kThreads), args, 0, s)); }
    CK(cudaStreamSynchronize(s));
    float total = 0;
    const int reps = 100;
    for (int i = 0; i < reps; ++i) {
      CK(cudaMemsetAsync(bar, 0, 64, s));
      CK(cudaEventRecord(e0, s));
      CK(cudaLaunchCooperativeKernel(fn, dim3(kBlocks), dim3(kThreads), args, 0, s));
      CK(cudaEventRecord(e1, s));
      CK(cudaStreamSynchronize(s));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      total += ms;
    }
    printf("%-48s %6.2f us per phase\n", mode == 0 ? "persistent kernel, cooperative grid.sync()" : "persistent kernel, atomic counter + flag barrier",
           total * 1e3f / reps / kPhases);
  }
  return 0;
}
