// io.cu — streamed transfer between the native on-disk store (tokens.bf16.bin) and device memory.
//
// Replaces the persistence half of JinaColBERTRetriever.index / load (local_rag_complete.py:742-753: torch.save /
// torch.load of ONE dense fp32 tensor through host memory).  A 10M-passage corpus is 327.7 GB, 41 GB per GPU on eight
// GPUs: a rank must never hold its shard in host memory.  A rank reads only ITS byte range of the token file, through
// two pinned staging buffers: while chunk i is on its way to the device (one cudaMemcpyAsync per chunk), pread fills
// the other buffer with chunk i + 1.  Host memory in use: 2 x chunk_bytes, whatever the shard size.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cstring>

#include "hrc_common.cuh"

namespace hrc {
namespace {

struct Staging {
  uint8_t* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int fd = -1;
  ~Staging() {
    for (int i = 0; i < 2; ++i) {
      if (buf[i]) cudaFreeHost(buf[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
    }
    if (fd >= 0) close(fd);
  }
  int init(size_t chunk) {
    for (int i = 0; i < 2; ++i) {
      HRC_CHECK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&buf[i]), chunk, cudaHostAllocDefault));
      HRC_CHECK_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    return 0;
  }
};

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace
}  // namespace hrc

using namespace hrc;

extern "C" {

int hrc_store_read_file(const char* path, int64_t file_offset, int64_t n_bytes, void* d_dst, size_t chunk_bytes,
                        void* stream, double* seconds_out) {
  HRC_REQUIRE(path != nullptr && file_offset >= 0 && n_bytes >= 0 && (n_bytes == 0 || d_dst != nullptr),
              "store_read_file: bad argument");
  if (chunk_bytes == 0) chunk_bytes = size_t(256) << 20;
  if (size_t(n_bytes) < chunk_bytes) chunk_bytes = n_bytes > 0 ? size_t(n_bytes) : 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double t0 = now_s();
  Staging s;
  s.fd = open(path, O_RDONLY);
  HRC_REQUIRE(s.fd >= 0, "store_read_file: cannot open %s: %s", path, strerror(errno));
  struct stat sb;
  HRC_REQUIRE(fstat(s.fd, &sb) == 0 && int64_t(sb.st_size) >= file_offset + n_bytes,
              "store_read_file: %s is shorter than offset %lld + %lld bytes", path, (long long)file_offset, (long long)n_bytes);
  if (int rc = s.init(chunk_bytes)) return rc;
  uint8_t* dst = static_cast<uint8_t*>(d_dst);
  int64_t done = 0;
  for (int i = 0; done < n_bytes; ++i) {
    const int b = i & 1;
    if (i >= 2) HRC_CHECK_CUDA(cudaEventSynchronize(s.ev[b]));          // the copy that last used this buffer has finished
    const size_t n = size_t(n_bytes - done) < chunk_bytes ? size_t(n_bytes - done) : chunk_bytes;
    size_t got = 0;
    while (got < n) {
      const ssize_t r = pread(s.fd, s.buf[b] + got, n - got, file_offset + done + int64_t(got));
      HRC_REQUIRE(r > 0, "store_read_file: read error in %s at %lld: %s", path, (long long)(file_offset + done + int64_t(got)),
                  r == 0 ? "unexpected end of file" : strerror(errno));
      got += size_t(r);
    }
    HRC_CHECK_CUDA(cudaMemcpyAsync(dst + done, s.buf[b], n, cudaMemcpyHostToDevice, st));   // ONE copy per chunk
    HRC_CHECK_CUDA(cudaEventRecord(s.ev[b], st));
    done += int64_t(n);
  }
  HRC_CHECK_CUDA(cudaStreamSynchronize(st));
  if (seconds_out != nullptr) *seconds_out = now_s() - t0;
  return 0;
}

int hrc_store_write_file(const char* path, int64_t file_offset, int64_t n_bytes, const void* d_src, size_t chunk_bytes,
                         void* stream, double* seconds_out) {
  HRC_REQUIRE(path != nullptr && file_offset >= 0 && n_bytes >= 0 && (n_bytes == 0 || d_src != nullptr),
              "store_write_file: bad argument");
  if (chunk_bytes == 0) chunk_bytes = size_t(256) << 20;
  if (size_t(n_bytes) < chunk_bytes) chunk_bytes = n_bytes > 0 ? size_t(n_bytes) : 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double t0 = now_s();
  Staging s;
  s.fd = open(path, O_WRONLY | O_CREAT, 0644);
  HRC_REQUIRE(s.fd >= 0, "store_write_file: cannot open %s: %s", path, strerror(errno));
  if (int rc = s.init(chunk_bytes)) return rc;
  const uint8_t* src = static_cast<const uint8_t*>(d_src);
  const int64_t n_chunks = (n_bytes + int64_t(chunk_bytes) - 1) / int64_t(chunk_bytes);
  auto chunk_len = [&](int64_t i) { return size_t(n_bytes - i * int64_t(chunk_bytes) < int64_t(chunk_bytes) ? n_bytes - i * int64_t(chunk_bytes) : int64_t(chunk_bytes)); };
  if (n_chunks > 0) {
    HRC_CHECK_CUDA(cudaMemcpyAsync(s.buf[0], src, chunk_len(0), cudaMemcpyDeviceToHost, st));
    HRC_CHECK_CUDA(cudaEventRecord(s.ev[0], st));
  }
  for (int64_t i = 0; i < n_chunks; ++i) {
    const int b = int(i & 1);
    if (i + 1 < n_chunks) {                                               // chunk i + 1 travels while chunk i is written
      HRC_CHECK_CUDA(cudaMemcpyAsync(s.buf[b ^ 1], src + (i + 1) * int64_t(chunk_bytes), chunk_len(i + 1), cudaMemcpyDeviceToHost, st));
      HRC_CHECK_CUDA(cudaEventRecord(s.ev[b ^ 1], st));
    }
    HRC_CHECK_CUDA(cudaEventSynchronize(s.ev[b]));
    const size_t n = chunk_len(i);
    size_t put = 0;
    while (put < n) {
      const ssize_t r = pwrite(s.fd, s.buf[b] + put, n - put, file_offset + i * int64_t(chunk_bytes) + int64_t(put));
      HRC_REQUIRE(r > 0, "store_write_file: write error in %s: %s", path, strerror(errno));
      put += size_t(r);
    }
  }
  if (seconds_out != nullptr) *seconds_out = now_s() - t0;
  return 0;
}

}  // extern "C"
