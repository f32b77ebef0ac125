// Host-side runtime of libaaconv_b200: error text, launch accounting, per-launch event profile.
#include <atomic>
#include <mutex>
#include <vector>
#include <cstring>
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

namespace {
std::atomic<long long> g_launches{0};
std::mutex g_prof_mu;
bool g_prof_on = false;
struct Mark { const char* name; cudaEvent_t ev; };
std::vector<Mark> g_marks;
cudaEvent_t g_prof_start = nullptr;
cudaStream_t g_prof_stream = nullptr;
}  // namespace

void note_launch(const char* name, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_prof_on) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  if (g_marks.empty()) {                       // the first launch after begin(): its start mark
    // start event was recorded at begin() on the default-captured stream; if the stream differs we
    // cannot order against it, so re-record here (the first kernel's own time is then not measured).
    if (st != g_prof_stream) { cudaEventRecord(g_prof_start, st); g_prof_stream = st; }
  }
  cudaEventRecord(ev, st);
  g_marks.push_back({name, ev});
}

}  // namespace aaconv

using namespace aaconv;

extern "C" {

long long aaconv_launch_count(void) { return g_launches.load(); }

// Start recording one CUDA event after every kernel launch this library makes (any thread) on `stream`.
int aaconv_profile_begin(void* stream) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& m : g_marks) cudaEventDestroy(m.ev);
  g_marks.clear();
  if (!g_prof_start) AACONV_CUDA_OK(cudaEventCreate(&g_prof_start));
  g_prof_stream = static_cast<cudaStream_t>(stream);
  AACONV_CUDA_OK(cudaEventRecord(g_prof_start, g_prof_stream));
  g_prof_on = true;
  return 0;
}

// Stop recording; synchronises, then writes up to `max_entries` (name, ms) pairs: names '\n'-joined into
// names_buf.  Each ms is the time between the previous mark and this launch's mark on the stream, i.e.
// the kernel's duration when launches are back to back.  Returns the number of entries (or < 0).
int aaconv_profile_end(char* names_buf, size_t names_len, float* ms, int max_entries) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = false;
  int n = 0;
  size_t off = 0;
  if (names_buf && names_len) names_buf[0] = 0;
  cudaEvent_t prev = g_prof_start;
  for (auto& m : g_marks) {
    if (cudaEventSynchronize(m.ev) != cudaSuccess) return fail(AACONV_E_CUDA, "profile: event sync failed");
    if (n < max_entries) {
      float t = 0.f;
      cudaEventElapsedTime(&t, prev, m.ev);
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
    prev = m.ev;
  }
  for (auto& m : g_marks) cudaEventDestroy(m.ev);
  g_marks.clear();
  return n;
}

}  // extern "C"
