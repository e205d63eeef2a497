// Error state, launch counter and the stream-ordered workspace arena.
#include "common.cuh"

#include <cstdlib>
#include <vector>

namespace hsidm {

thread_local std::string g_last_error;
int64_t g_launches = 0;
bool g_pdl_on = std::getenv("HSIDM_NO_PDL") == nullptr;
bool g_pdl_pass = true;

void set_last_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

// ---- profiling ------------------------------------------------------------------------------------------------
bool g_prof_on = false;
namespace {
struct ProfRec {
  cudaEvent_t a, b;
  int kind;
  double work;
  std::string tag;
};
std::vector<ProfRec> g_prof_recs;
}  // namespace

int prof_start(int kind, double work, cudaStream_t stream, const char* tag) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return -1;
  }
  ProfRec r;
  r.kind = kind, r.work = work;
  if (tag) r.tag = tag;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
  cudaEventRecord(r.a, stream);
  g_prof_recs.push_back(r);
  return (int)g_prof_recs.size() - 1;
}

void prof_stop(int token, cudaStream_t stream) { cudaEventRecord(g_prof_recs[token].b, stream); }

void prof_reset() {
  for (auto& r : g_prof_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof_recs.clear();
}

int prof_read(int kind, double* ms, double* work, int64_t* launches) {
  *ms = 0, *work = 0, *launches = 0;
  HSIDM_CUDA(cudaDeviceSynchronize());
  for (auto& r : g_prof_recs) {
    if (r.kind != kind) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) *ms += t, *work += r.work, *launches += 1;
  }
  return HSIDM_OK;
}

int prof_dump(const char* path) {
  HSIDM_CUDA(cudaDeviceSynchronize());
  FILE* f = fopen(path, "w");
  if (!f) HSIDM_FAIL(HSIDM_BAD_ARG, "cannot open %s", path);
  fprintf(f, "kind,tag,work,ms\n");
  for (auto& r : g_prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) fprintf(f, "%d,%s,%.6g,%.6f\n", r.kind, r.tag.c_str(), r.work, t);
  }
  fclose(f);
  return HSIDM_OK;
}

static constexpr int64_t kAlign = 1024;
static char* const kDryBase = reinterpret_cast<char*>(uintptr_t(1) << 40);  // never dereferenced

Arena::~Arena() {
  if (base_) cudaFree(base_);
}

int Arena::reserve(int64_t bytes) {
  if (bytes <= cap_) return HSIDM_OK;
  if (base_) {
    HSIDM_CUDA(cudaDeviceSynchronize());
    HSIDM_CUDA(cudaFree(base_));
    base_ = nullptr;
    cap_ = 0;
  }
  bytes = round_up(bytes, 1 << 20);
  cudaError_t e = cudaMalloc(&base_, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    base_ = nullptr;
    HSIDM_FAIL(HSIDM_OOM_WORKSPACE, "workspace of %lld bytes could not be allocated: %s", (long long)bytes,
               cudaGetErrorString(e));
  }
  cap_ = bytes;
  return HSIDM_OK;
}

void Arena::begin(bool dry) {
  dry_ = dry;
  failed_ = false;
  nblocks_ = 0;
  top_ = 0;
  if (dry) peak_ = 0;
}

void* Arena::alloc(int64_t bytes) {
  bytes = round_up(bytes < 1 ? 1 : bytes, kAlign);
  char* base = dry_ ? kDryBase : base_;
  int best = -1;
  for (int i = 0; i < nblocks_; ++i)
    if (!blocks_[i].used && blocks_[i].size >= bytes && (best < 0 || blocks_[i].size < blocks_[best].size)) best = i;
  if (best >= 0) {
    Block& b = blocks_[best];
    if (b.size > bytes && nblocks_ < kMaxBlocks) {  // split, keep the list sorted by offset
      for (int j = nblocks_; j > best + 1; --j) blocks_[j] = blocks_[j - 1];
      blocks_[best + 1] = Block{b.off + bytes, b.size - bytes, false};
      ++nblocks_;
      blocks_[best].size = bytes;
    }
    blocks_[best].used = true;
    return base + blocks_[best].off;
  }
  if (nblocks_ >= kMaxBlocks) {
    failed_ = true;
    return nullptr;
  }
  // grow at the top (absorbing a trailing free block)
  int64_t off = top_;
  if (nblocks_ > 0 && !blocks_[nblocks_ - 1].used) {
    off = blocks_[nblocks_ - 1].off;
    --nblocks_;
  }
  blocks_[nblocks_++] = Block{off, bytes, true};
  top_ = off + bytes;
  if (top_ > peak_) peak_ = top_;
  if (!dry_ && top_ > cap_) {
    failed_ = true;
    return nullptr;
  }
  return base + off;
}

void Arena::free(void* p) {
  if (!p) return;
  char* base = dry_ ? kDryBase : base_;
  int64_t off = static_cast<char*>(p) - base;
  for (int i = 0; i < nblocks_; ++i) {
    if (blocks_[i].off != off || !blocks_[i].used) continue;
    blocks_[i].used = false;
    if (i + 1 < nblocks_ && !blocks_[i + 1].used) {  // merge right
      blocks_[i].size += blocks_[i + 1].size;
      for (int j = i + 1; j + 1 < nblocks_; ++j) blocks_[j] = blocks_[j + 1];
      --nblocks_;
    }
    if (i > 0 && !blocks_[i - 1].used) {  // merge left
      blocks_[i - 1].size += blocks_[i].size;
      for (int j = i; j + 1 < nblocks_; ++j) blocks_[j] = blocks_[j + 1];
      --nblocks_;
    }
    if (nblocks_ > 0 && !blocks_[nblocks_ - 1].used) {  // give the tail back
      top_ = blocks_[nblocks_ - 1].off;
      --nblocks_;
    }
    return;
  }
}

}  // namespace hsidm
