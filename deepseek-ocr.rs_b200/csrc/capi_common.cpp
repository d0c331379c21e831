// Error plumbing shared by every extern "C" entry point.
#include "dsocr.h"
#include "util.h"

#include <atomic>
#include <map>
#include <thread>
#include <stdio.h>

namespace dsocr {
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
std::atomic<long long>& launch_counter() {
  static std::atomic<long long> n{0};
  return n;
}

namespace {
struct TimingState {  // a diagnostic of ONE engine at a time: only launches of the thread that enabled it are recorded
  std::atomic<bool> on{false};
  std::thread::id owner;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> pool;
  std::vector<std::pair<const char*, const char*>> names;  // (phase, kernel) that ended at event i+1
  const char* phase = "";
  size_t used = 0;
  cudaEvent_t next() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cuda_check(cudaEventCreate(&e), "cudaEventCreate");
      pool.push_back(e);
    }
    return pool[used++];
  }
} g_timing;
}  // namespace

void launch_check(const char* what) {
  ++launch_counter();
  cuda_check(cudaGetLastError(), what);
  if (g_timing.on.load(std::memory_order_relaxed) && g_timing.owner == std::this_thread::get_id()) {
    cuda_check(cudaEventRecord(g_timing.next(), g_timing.stream), "timing event");
    g_timing.names.push_back({g_timing.phase, what});
  }
}
void kernel_timing_phase(const char* phase) { g_timing.phase = phase; }
bool kernel_timing_enabled() { return g_timing.on.load() && g_timing.owner == std::this_thread::get_id(); }
void kernel_timing_begin(cudaStream_t stream) {
  g_timing.owner = std::this_thread::get_id();
  g_timing.on = true;
  g_timing.stream = stream;
  g_timing.used = 0;
  g_timing.names.clear();
  cuda_check(cudaEventRecord(g_timing.next(), stream), "timing event");
}
std::string kernel_timing_end_json() {
  g_timing.on = false;
  if (g_timing.used == 0) return "[]";
  cuda_check(cudaEventSynchronize(g_timing.pool[g_timing.used - 1]), "timing sync");
  std::vector<std::string> order;
  std::map<std::string, std::pair<long long, double>> agg;
  for (size_t i = 0; i < g_timing.names.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_timing.pool[i], g_timing.pool[i + 1]);
    const std::string key = std::string(g_timing.names[i].first) + g_timing.names[i].second;
    auto it = agg.find(key);
    if (it == agg.end()) { order.push_back(key); it = agg.emplace(key, std::make_pair(0LL, 0.0)).first; }
    it->second.first += 1;
    it->second.second += ms;
  }
  std::string out = "[";
  for (size_t i = 0; i < order.size(); ++i) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s{\"name\": \"%s\", \"launches\": %lld, \"ms\": %.6f}", i ? ", " : "", order[i].c_str(),
             agg[order[i]].first, agg[order[i]].second);
    out += buf;
  }
  return out + "]";
}
}  // namespace dsocr

extern "C" const char* dsocr_last_error(void) { return dsocr::g_last_error.c_str(); }
extern "C" const char* dsocr_version(void) { return "dsocr-b200 0.1 (sm_100a)"; }
