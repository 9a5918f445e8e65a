// Error plumbing, device queries and bookkeeping shared by every entry point of libsdb200.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace sdb {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) {
    g_launches.fetch_add(strstr(what, "cudaGetLastError") ? 1 : 0, std::memory_order_relaxed);
    return 0;
  }
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return SDB_ERR_CUDA;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* v = getenv("SDB_PDL");
    on = (v && *v == '0') ? 0 : 1;
  }
  return on != 0;
}

static int g_host_pack_mode = SDB_FMT_F16X2;
int host_pack_mode() { return g_host_pack_mode; }
void set_host_pack_mode(int m) { g_host_pack_mode = m; }

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace sdb

extern "C" int sdb_version(void) { return 1; }
extern "C" const char* sdb_last_error(void) { return sdb::g_err; }
extern "C" int64_t sdb_launch_count(void) { return sdb::g_launches.load(); }

extern "C" int sdb_set_pack_mode(int fmt, void* stream) {
  SDB_REQUIRE(fmt == SDB_FMT_F16X2 || fmt == SDB_FMT_F8C, "sdb_set_pack_mode: bad format %d", fmt);
  // one flag per translation unit that writes packed operands (no relocatable device code)
  int rc = sdb::set_pack_mode_elementwise(fmt, sdb::as_stream(stream));
  if (!rc) rc = sdb::set_pack_mode_gemm(fmt, sdb::as_stream(stream));
  if (!rc) rc = sdb::set_pack_mode_attention(fmt, sdb::as_stream(stream));
  if (!rc) rc = sdb::set_pack_mode_attention_tc(fmt, sdb::as_stream(stream));
  if (!rc) sdb::set_host_pack_mode(fmt);
  return rc;
}
