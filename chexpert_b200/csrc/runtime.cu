// Host-side runtime of libaaconv_b200: error text, launch accounting, per-launch event profile.
#include <atomic>
#include <mutex>
#include <vector>
#include <cstring>
#include <nvtx3/nvToolsExt.h>
#include "common.cuh"

namespace aaconv {

std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

namespace {
std::atomic<long long> g_launches{0};
std::mutex g_prof_mu;
std::atomic<bool> g_prof_on{false};
struct Mark { const char* name; cudaEvent_t start, stop; };
std::vector<Mark> g_marks;
thread_local cudaEvent_t t_pending_start = nullptr;   // start event of the launch being issued by this thread
}  // namespace

cudaStream_t pre_launch(cudaStream_t st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return st;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return st;
  cudaEventRecord(ev, st);
  if (t_pending_start) cudaEventDestroy(t_pending_start);
  t_pending_start = ev;
  return st;
}

void note_launch(const char* name, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  cudaEvent_t start = t_pending_start;
  t_pending_start = nullptr;
  if (!start) return;                          // launch was issued before the profile began
  cudaEvent_t stop;
  if (cudaEventCreate(&stop) != cudaSuccess) { cudaEventDestroy(start); return; }
  cudaEventRecord(stop, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_marks.push_back({name, start, stop});
}

}  // namespace aaconv

using namespace aaconv;

extern "C" {

long long aaconv_launch_count(void) { return g_launches.load(); }

// Start recording a (start, stop) CUDA event pair around every kernel launch this library makes (any thread, any stream).
int aaconv_profile_begin(void* stream) {
  (void)stream;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& m : g_marks) { cudaEventDestroy(m.start); cudaEventDestroy(m.stop); }
  g_marks.clear();
  g_prof_on = true;
  return 0;
}

// Stop recording; synchronises, then writes up to `max_entries` (name, ms) pairs: names '\n'-joined into names_buf.
// Each ms is the device time between the launch's own start and stop events (host gaps between launches are excluded).
// Returns the number of entries (or < 0).
int aaconv_profile_end(char* names_buf, size_t names_len, float* ms, int max_entries) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = false;
  int n = 0;
  size_t off = 0;
  if (names_buf && names_len) names_buf[0] = 0;
  for (auto& m : g_marks) {
    if (cudaEventSynchronize(m.stop) != cudaSuccess) return fail(AACONV_E_CUDA, "profile: event sync failed");
    if (n < max_entries) {
      float t = 0.f;
      cudaEventElapsedTime(&t, m.start, m.stop);
      ms[n] = t;
      const size_t len = strlen(m.name);
      if (names_buf && off + len + 2 < names_len) {
        memcpy(names_buf + off, m.name, len);
        off += len;
        names_buf[off++] = '\n';
        names_buf[off] = 0;
      }
      ++n;
    }
  }
  for (auto& m : g_marks) { cudaEventDestroy(m.start); cudaEventDestroy(m.stop); }
  g_marks.clear();
  return n;
}

}  // extern "C"
