// run_file.cu -- file range -> operator -> file, streamed through a bounded ring of pinned buffers.
//
//   reader side    worker.PlainFile(path, partitions, delim) + ReadFixer   bigseqkit/helper.go:148-195,
//                                                                          bigseqkit-lib/helper.go:41-66
//   writer side    FileStore, element + '\n' in input order                bigseqkit-lib/helper.go:378-460
//
// The reference reads a partition, maps it, and stores it, one after the other.  Here a range of any size (larger than
// host memory included) moves through five overlapped stages:
//
//   reader thread    pread of the next record-aligned block (several threads per block) into one of K pinned slots
//   copy-in stream   H2D of block i+1
//   ctx stream       kernels of block i (the same process_block chain as bsk_run_buffer: state such as the stats
//                    histogram or the rmdup key table carries from block to block)
//   copy-out stream  D2H of block i-1 into one of K pinned output slots
//   writer thread    the finished slot goes to the output file at its running offset
//
// Host memory is bounded by 2 K slots of one block each, whatever the length of the range.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "engine.h"

namespace bsk {

static size_t stream_block_bytes() {
  const char *e = getenv("BSK_BLOCK_BYTES");
  size_t v = e ? strtoull(e, nullptr, 10) : 0;
  if (v < 4096) v = 64ull << 20;
  if (v > kMaxBlockBytes / 2) v = kMaxBlockBytes / 2;
  return v;
}

static unsigned io_threads() {
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  if (const char *e = getenv("BSK_IO_THREADS")) { const int v = atoi(e); if (v > 0) hw = (unsigned)v; }
  return hw > 16 ? 16 : hw;
}

// pread / pwrite of one byte range by several threads: a single stream reads the page cache at memcpy speed of one
// core, far below what the H2D pipeline takes
bool par_io(int fd, unsigned char *buf, uint64_t len, uint64_t off, bool write) {
  if (len == 0) return true;
  const uint64_t kMinChunk = 8ull << 20;
  uint64_t nt = (len + kMinChunk - 1) / kMinChunk;
  if (nt > io_threads()) nt = io_threads();
  const uint64_t chunk = ((len + nt - 1) / nt + 4095) & ~4095ull;
  std::atomic<bool> ok{true};
  auto work = [&](uint64_t t) {
    uint64_t p = t * chunk;
    const uint64_t end = p + chunk < len ? p + chunk : len;
    while (p < end) {
      const ssize_t r = write ? pwrite(fd, buf + p, end - p, (off_t)(off + p)) : pread(fd, buf + p, end - p, (off_t)(off + p));
      if (r <= 0) { ok = false; return; }
      p += (uint64_t)r;
    }
  };
  std::vector<std::thread> th;
  for (uint64_t t = 1; t < nt; t++) th.emplace_back(work, t);
  work(0);
  for (auto &x : th) x.join();
  return ok;
}

// Output side: pwrite()s to one file serialise on the inode lock, so the range is mapped and filled by several
// threads instead (page allocation then runs in parallel); plain pwrite when the file cannot be mapped.
bool par_write(int fd, const unsigned char *buf, uint64_t len, uint64_t off) {
  if (len == 0) return true;
  struct stat sb;
  const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE);
  if (getenv("BSK_NO_MMAP_WRITE") == nullptr && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode)) {
    const uint64_t need = off + len;
    // grow by writing this range's own last byte: never shrinks a file that another rank has already extended
    if ((uint64_t)sb.st_size >= need || pwrite(fd, buf + len - 1, 1, (off_t)(need - 1)) == 1) {
      const uint64_t m0 = off & ~(page - 1);
      void *m = mmap(nullptr, need - m0, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)m0);
      if (m != MAP_FAILED) {
        unsigned char *dst = (unsigned char *)m + (off - m0);
        uint64_t nt = (len + (8ull << 20) - 1) / (8ull << 20);
        if (nt > io_threads()) nt = io_threads();
        const uint64_t chunk = ((len + nt - 1) / nt + page - 1) & ~(page - 1);
        auto work = [&](uint64_t t) {
          const uint64_t p = t * chunk;
          if (p < len) memcpy(dst + p, buf + p, p + chunk < len ? chunk : len - p);
        };
        std::vector<std::thread> th;
        for (uint64_t t = 1; t < nt; t++) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
        munmap(m, need - m0);
        return true;
      }
    }
  }
  return par_io(fd, const_cast<unsigned char *>(buf), len, off, true);
}

// last record start in (0, n): the cut of a block read from the middle of a file (SURVEY C.1 record-start rule);
// 0 when there is none
static size_t last_record_start(const u8 *d, size_t n, bool fq) {
  const u8 marker = fq ? '@' : '>';
  for (size_t L = n; L-- > 1;) {
    if (d[L - 1] != '\n' || d[L] != marker) continue;
    if (fq && L >= 3 && d[L - 3] == '\n' && d[L - 2] == '+') continue;
    return L;
  }
  return 0;
}

namespace {
struct Sem {
  std::mutex m;
  std::condition_variable cv;
  long v;
  explicit Sem(long v0) : v(v0) {}
  void post() {
    { std::lock_guard<std::mutex> g(m); v++; }
    cv.notify_one();
  }
  void wait() {
    std::unique_lock<std::mutex> g(m);
    cv.wait(g, [&] { return v > 0; });
    v--;
  }
};
struct InBlock { int slot; size_t n; bool last; bool ok; };
struct OutJob { int slot; uint64_t n, off; bool stop; };
template <class T>
struct Chan {
  std::mutex m;
  std::condition_variable cv;
  std::deque<T> q;
  void push(const T &x) {
    { std::lock_guard<std::mutex> g(m); q.push_back(x); }
    cv.notify_one();
  }
  T pop() {
    std::unique_lock<std::mutex> g(m);
    cv.wait(g, [&] { return !q.empty(); });
    T x = q.front();
    q.pop_front();
    return x;
  }
};
}  // namespace

int Engine::run_stream(int fd, u64 off, u64 len, int64_t pid, int out_fd, u64 out_off, u64 *out_bytes, u64 *n_records, u64 *n_elem) {
  if (device_ >= 0) BSK_CUDA(cudaSetDevice(device_));
  launches_ = 0;
  timings = bsk_timings{};
  alphabet_ = o_.alphabet;
  alphabet_known_ = false;
  first_block_ = true;
  if (!s_in_) {
    BSK_CUDA(cudaStreamCreateWithFlags(&s_in_, cudaStreamNonBlocking));
    BSK_CUDA(cudaStreamCreateWithFlags(&s_out_, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      BSK_CUDA(cudaEventCreateWithFlags(&ev_in_done_[i], cudaEventDisableTiming));
      BSK_CUDA(cudaEventCreateWithFlags(&ev_in_free_[i], cudaEventDisableTiming));
      BSK_CUDA(cudaEventCreateWithFlags(&ev_out_ready_[i], cudaEventDisableTiming));
      BSK_CUDA(cudaEventCreateWithFlags(&ev_out_free_[i], cudaEventDisableTiming));
    }
  }
  constexpr int K = kStreamSlots;
  const size_t blk = stream_block_bytes();
  const bool want_out = out_fd >= 0;
  // first byte of the range decides the format (every range starts on a record)
  u8 first = 0;
  if (len && pread(fd, &first, 1, (off_t)off) != 1) { err = "short read"; return BSK_ERR_DATA; }
  const bool fq = first == '@';

  // ---- reader: record-aligned blocks into the pinned input slots
  Sem in_free(K);
  Chan<InBlock> in_ready;
  std::atomic<bool> stop{false};
  std::string read_err;
  for (int s = 0; s < K; s++) s_in_slot_[s].reserve(std::min<u64>(len, blk) + 64);
  std::thread reader([&] {
    if (device_ >= 0) cudaSetDevice(device_);  // a slot that has to grow is page-locked from this thread
    u64 pos = off;
    const u64 end = off + len;
    int j = 0;
    for (;;) {
      in_free.wait();
      if (stop) return;
      const int slot = j % K;
      size_t got = 0, cut = 0;
      size_t want = (size_t)std::min<u64>(blk, end - pos);
      bool ok = true;
      for (;;) {  // grows only when a single record is longer than the block
        PinnedBuf &b = s_in_slot_[slot];
        try {
          b.reserve(want + 64, got > 0, got);
        } catch (const std::exception &e) { read_err = e.what(); ok = false; break; }
        if (!par_io(fd, b.as<u8>() + got, want - got, pos + got, false)) { read_err = "short read"; ok = false; break; }
        got = want;
        if (pos + got >= end) { cut = got; break; }
        cut = last_record_start(b.as<u8>(), got, fq);
        if (cut) break;
        if (got >= kMaxBlockBytes / 2) { read_err = "a single record exceeds the 4 GiB block limit"; ok = false; break; }
        want = (size_t)std::min<u64>((u64)got * 2, end - pos);
        if (want > kMaxBlockBytes - 64) want = kMaxBlockBytes - 64;
      }
      pos += cut;
      const bool last = !ok || pos >= end;
      in_ready.push(InBlock{slot, cut, last, ok});
      j++;
      if (last) return;
    }
  });

  // ---- writer: finished output slots to the file
  Sem out_free(K);
  Chan<OutJob> out_jobs;
  std::atomic<bool> write_ok{true};
  std::thread writer([&] {
    for (;;) {
      const OutJob jb = out_jobs.pop();
      if (jb.stop) return;
      if (jb.n && write_ok && !par_write(out_fd, s_out_slot_[jb.slot].as<u8>(), jb.n, jb.off)) write_ok = false;
      out_free.post();
    }
  });
  auto shutdown = [&]() {
    stop = true;
    for (int s = 0; s < K + 1; s++) in_free.post();
    out_jobs.push(OutJob{0, 0, 0, true});
    reader.join();
    writer.join();
  };

  int rc = BSK_OK;
  u64 out_used = 0, elem_used = 0, n_rec_total = 0, n_jobs = 0;  // jobs finish in the order they are handed out: job j uses slot j % K
  try {
    InBlock cur = in_ready.pop();
    if (!cur.ok) { err = read_err; shutdown(); return BSK_ERR_DATA; }
    u8 *d_in[2] = {nullptr, nullptr};
    auto upload = [&](const InBlock &b, size_t i) {
      d_in[i & 1] = (i & 1) ? b_in2_.get<u8>(b.n + 64) : b_in_.get<u8>(b.n + 64);  // (grows only while no copy of that buffer is in flight)
      if (i >= 2) BSK_CUDA(cudaStreamWaitEvent(s_in_, ev_in_free_[i & 1], 0));
      if (b.n) BSK_CUDA(cudaMemcpyAsync(d_in[i & 1], s_in_slot_[b.slot].as<u8>(), b.n, cudaMemcpyHostToDevice, s_in_));
      BSK_CUDA(cudaEventRecord(ev_in_done_[i & 1], s_in_));
    };
    bool out_busy[2] = {false, false};
    int ob = 0;
    struct Pending { bool any; int ob, slot; u64 n, off; } pend{false, 0, 0, 0, 0};
    auto flush_pending = [&]() {  // the D2H of the previous block is done: its slot goes to the writer
      if (!pend.any) return;
      BSK_CUDA(cudaEventSynchronize(ev_out_free_[pend.ob]));
      out_jobs.push(OutJob{pend.slot, pend.n, pend.off, false});
      pend.any = false;
    };
    upload(cur, 0);
    for (size_t i = 0;; i++) {
      InBlock nxt{0, 0, true, true};
      if (!cur.last) {
        nxt = in_ready.pop();  // (the kernels of block i-1 are done: its device buffer may be refilled, or grow)
        if (!nxt.ok) { err = read_err; rc = BSK_ERR_DATA; break; }
        upload(nxt, i + 1);
      }
      BSK_CUDA(cudaStreamWaitEvent(stream, ev_in_done_[i & 1], 0));
      if (out_busy[ob]) BSK_CUDA(cudaStreamWaitEvent(stream, ev_out_free_[ob], 0));
      BlockOut bo;
      rc = process_block(d_in[i & 1], (u32)cur.n, pid, bo);  // synchronises the ctx stream (and so the H2D of this block)
      if (rc != BSK_OK) break;
      BSK_CUDA(cudaEventRecord(ev_in_free_[i & 1], stream));
      in_free.post();  // the pinned slot of block i has been copied
      n_rec_total += bo.n_rec;
      elem_used += bo.n_elem;
      if (want_out && bo.n) {
        out_free.wait();
        const int slot = (int)(n_jobs++ % K);
        s_out_slot_[slot].reserve(bo.n + 64);
        BSK_CUDA(cudaEventRecord(ev_out_ready_[ob], stream));
        BSK_CUDA(cudaStreamWaitEvent(s_out_, ev_out_ready_[ob], 0));
        BSK_CUDA(cudaMemcpyAsync(s_out_slot_[slot].as<u8>(), bo.d_data, bo.n, cudaMemcpyDeviceToHost, s_out_));
        BSK_CUDA(cudaEventRecord(ev_out_free_[ob], s_out_));
        flush_pending();
        pend = Pending{true, ob, slot, bo.n, out_off + out_used};
        const bool swappable = bo.d_data == b_out_.p;
        if (swappable && !cur.last) {
          std::swap(b_out_.p, b_out2_.p);
          std::swap(b_out_.cap, b_out2_.cap);
          std::swap(b_elem_.p, b_elem2_.p);
          std::swap(b_elem_.cap, b_elem2_.cap);
          out_busy[ob] = true;
          ob ^= 1;
        } else {
          flush_pending();
          out_busy[ob] = false;
        }
      }
      out_used += bo.n;
      if (cur.last) break;
      cur = nxt;
    }
    if (rc == BSK_OK) flush_pending();
    BSK_CUDA(cudaStreamSynchronize(s_out_));
    BSK_CUDA(cudaStreamSynchronize(s_in_));
    if (rc == BSK_OK && op_ == OP_GREP && o_.Count) {
      BlockOut bo;
      rc = finish_grep_count(bo);
      if (rc == BSK_OK) {
        if (want_out && bo.n) {
          out_free.wait();
          const int slot = (int)(n_jobs++ % K);
          s_out_slot_[slot].reserve(bo.n + 64);
          BSK_CUDA(cudaMemcpy(s_out_slot_[slot].as<u8>(), bo.d_data, bo.n, cudaMemcpyDeviceToHost));
          out_jobs.push(OutJob{slot, bo.n, out_off, false});
        }
        out_used = bo.n;
        elem_used = 1;
      }
    }
  } catch (...) {
    shutdown();
    throw;
  }
  shutdown();
  if (rc != BSK_OK) return rc;
  if (!write_ok) { err = "short write"; return BSK_ERR_DATA; }
  timings.kernel_launches = launches_;
  timings.in_bytes = len;
  timings.out_bytes = out_used;
  if (out_bytes) *out_bytes = out_used;
  if (n_records) *n_records = n_rec_total;
  if (n_elem) *n_elem = elem_used;
  return BSK_OK;
}

}  // namespace bsk
