// Asynchronous file IO for checkpoint writing and NVMe optimizer-state offload.
//
// Capability parity: TensorNVMe's AsyncFileWriter / DiskOffloader pthread backend used by the reference
// (colossalai/utils/safetensors.py:11-205, nn/optimizer/nvme_optimizer.py:32).  Design: one small worker pool per
// handle executing positional pwrite/pread requests from a queue; callers enqueue (pointer, bytes, offset) and later
// `synchronize()`.  Large requests are split into 8 MiB slices so several workers stream one tensor in parallel.
#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <fcntl.h>
#include <mutex>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

struct Req {
  bool write;
  char* ptr;
  int64_t bytes;
  int64_t offset;
};

class AioFile {
 public:
  AioFile(const char* path, int flags, int n_threads) : fd_(::open(path, flags, 0644)), stop_(false), pending_(0), err_(0) {
    if (fd_ < 0) { err_ = errno; return; }
    for (int i = 0; i < n_threads; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~AioFile() {
    {
      std::unique_lock<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
    if (fd_ >= 0) ::close(fd_);
  }
  int error() const { return err_; }
  void submit(bool write, void* ptr, int64_t bytes, int64_t offset) {
    const int64_t SLICE = 8ll << 20;
    std::unique_lock<std::mutex> lk(mu_);
    for (int64_t o = 0; o < bytes; o += SLICE) {
      q_.push_back(Req{write, (char*)ptr + o, bytes - o < SLICE ? bytes - o : SLICE, offset + o});
      ++pending_;
    }
    lk.unlock();
    cv_.notify_all();
  }
  int synchronize() {
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [this] { return pending_ == 0; });
    return err_;
  }
  int fsync_file() { return ::fsync(fd_); }

 private:
  void loop() {
    for (;;) {
      Req r;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
        if (q_.empty()) { if (stop_) return; continue; }
        r = q_.front();
        q_.pop_front();
      }
      int64_t done = 0;
      while (done < r.bytes) {
        const ssize_t n = r.write ? ::pwrite(fd_, r.ptr + done, r.bytes - done, r.offset + done)
                                  : ::pread(fd_, r.ptr + done, r.bytes - done, r.offset + done);
        if (n < 0) { if (errno == EINTR) continue; err_ = errno; break; }
        if (n == 0) { if (!r.write) err_ = EIO; break; }
        done += n;
      }
      {
        std::unique_lock<std::mutex> lk(mu_);
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }
  int fd_;
  bool stop_;
  int64_t pending_;
  std::atomic<int> err_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::deque<Req> q_;
  std::vector<std::thread> workers_;
};

}  // namespace

extern "C" {

void* cb_aio_open(const char* path, int for_write, int truncate, int n_threads) {
  int flags = for_write ? (O_RDWR | O_CREAT | (truncate ? O_TRUNC : 0)) : O_RDONLY;
  AioFile* f = new AioFile(path, flags, n_threads > 0 ? n_threads : 4);
  if (f->error()) { delete f; return nullptr; }
  return f;
}
void cb_aio_write(void* h, void* ptr, int64_t bytes, int64_t offset) { ((AioFile*)h)->submit(true, ptr, bytes, offset); }
void cb_aio_read(void* h, void* ptr, int64_t bytes, int64_t offset) { ((AioFile*)h)->submit(false, ptr, bytes, offset); }
int cb_aio_synchronize(void* h) { return ((AioFile*)h)->synchronize(); }
int cb_aio_fsync(void* h) { return ((AioFile*)h)->fsync_file(); }
void cb_aio_close(void* h) { delete (AioFile*)h; }

}  // extern "C"
