// Library-wide plumbing of the C ABI: error string, launch counter, device properties.
#include <stdarg.h>
#include <string.h>

#include "gf_common.cuh"

namespace gf {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches += n; }

static thread_local cudaEvent_t g_stage_ev[ST_COUNT] = {nullptr, nullptr, nullptr, nullptr, nullptr};
static thread_local bool g_stage_armed = false;

void stage_mark(int stage, cudaStream_t st) {
  if (!g_stage_armed || stage < 0 || stage >= ST_COUNT) return;
  if (g_stage_ev[stage]) {
    // inside a stream capture the event becomes an EXTERNAL record node of the graph: every replay stamps it,
    // and cudaEventElapsedTime between two of them is valid after the replay has finished
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive)
      cudaEventRecordWithFlags(g_stage_ev[stage], st, cudaEventRecordExternal);
    else
      cudaEventRecord(g_stage_ev[stage], st);
  }
  if (stage == ST_GEO_DONE) g_stage_armed = false;  // one call only
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace gf

extern "C" const char *gf_last_error(void) { return gf::g_err; }
extern "C" int gf_version(void) { return 100; }
extern "C" int64_t gf_launch_count(void) { return gf::g_launches; }
extern "C" void gf_reset_launch_count(void) { gf::g_launches = 0; }
extern "C" int gf_set_stage_events(void **events, int n) {
  if (events == nullptr || n <= 0) {
    gf::g_stage_armed = false;
    return GF_OK;
  }
  for (int i = 0; i < gf::ST_COUNT; ++i) gf::g_stage_ev[i] = i < n ? (cudaEvent_t)events[i] : nullptr;
  gf::g_stage_armed = true;
  return GF_OK;
}

// tiny event helpers so that a ctypes host can time stages without a CUDA binding of its own
extern "C" void *gf_event_create(void) {
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) {
    gf::set_error("cudaEventCreate failed");
    return nullptr;
  }
  return (void *)e;
}
extern "C" void gf_event_destroy(void *e) {
  if (e) cudaEventDestroy((cudaEvent_t)e);
}
extern "C" float gf_event_elapsed_ms(void *a, void *b) {
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) {
    (void)cudaGetLastError();
    return -1.f;
  }
  return ms;
}

// ---- peer-visible memory (seed-sharded scenes, include/geoformer_b200.h) ---------------------------
extern "C" int gf_peer_alloc(void **dev_ptr, size_t bytes) {
  GF_CHECK_ARG(dev_ptr && bytes > 0, "peer_alloc: null pointer or empty allocation");
  GF_CUDA(cudaMalloc(dev_ptr, bytes));
  return GF_OK;
}
extern "C" int gf_peer_free(void *dev_ptr) {
  if (dev_ptr) GF_CUDA(cudaFree(dev_ptr));
  return GF_OK;
}
extern "C" int gf_peer_export(void *dev_ptr, void *handle_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == GF_PEER_HANDLE_BYTES, "CUDA IPC handle size");
  GF_CHECK_ARG(dev_ptr && handle_out, "peer_export: null pointer");
  cudaIpcMemHandle_t h;
  GF_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(handle_out, &h, sizeof(h));
  return GF_OK;
}
extern "C" int gf_peer_open(const void *handle, void **dev_ptr) {
  GF_CHECK_ARG(handle && dev_ptr, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  GF_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GF_OK;
}
extern "C" int gf_peer_close(void *dev_ptr) {
  if (dev_ptr) GF_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return GF_OK;
}
