// Shuffled activation loader with the shuffle pool in HBM (replaces saev's ShuffledDataLoader + ReservoirBuffer,
// src/saev/data/shuffled.py:133-363, 380-699 and src/saev/data/buffers.py:91-231, and the pageable H2D copy of
// src/saev/framework/train.py:333).
//
// saev shuffles on the host: I/O threads push rows one by one into a shared-memory reservoir (Python loops, ~21 k
// rows/s), the consumer pops random rows into a pageable batch, the training loop copies it to the GPU.  Here the
// reservoir lives in device memory (a B200 has 180 GB; the default 64-batch pool is 4.3 GB at c3):
//
//   I/O threads  : pread() whole examples (content tokens of the requested layer are contiguous on disk:
//                  shard layout [example, layer, token, d_model] fp32, shards.py:168-180) into PINNED staging
//                  chunks, write the (example_idx, token_idx) pairs next to them, optionally drop rows by label.
//                  ZERO-COPY mode (shards on tmpfs, or forced): the shard file is mmap()ed and registered with the
//                  driver (cudaHostRegister), and the feeder's copies DMA straight out of the page cache -- no
//                  pread -> staging memcpy on the host.  With 8 ranks on one host that memcpy was what bounded the
//                  end-to-end rate (round 1: 11 M of 26 M activations/s).
//   feeder thread: owns the loader's CUDA stream.  Appends ready chunks to the tail of the device pool with
//                  cudaMemcpyAsync, and prepares batches ahead of the consumer: draws `need` distinct random pool
//                  positions (host RNG), launches a gather kernel pool -> batch buffer, then a move kernel that
//                  back-fills the holes from the pool tail, so the live part of the pool stays [0, fill).
//   consumer     : saev_b200_loader_next() hands out the next prepared batch as DEVICE pointers and makes the
//                  caller's stream wait for the gather (event, no host sync).
//
// Every (example, content-token) row of the rank's shards is delivered exactly once per epoch, in an order that
// is uniform over the pool contents at draw time — the same contract tests/test_reservoir_buffer.py pins for the
// reference's reservoir.  The pure-host pieces (chunk reader, draw/compaction planner) are exported separately so
// that they are testable without a GPU.
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/vfs.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/saev_b200.h"
#include "kernels.h"

namespace {

thread_local char g_loader_err[512] = "";

// ---------------------------------------------------------------------------------------------
// RNG (splitmix64 seeded xoshiro256**), host only
// ---------------------------------------------------------------------------------------------
struct Rng {
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) {
    for (auto& v : s) v = splitmix(seed);
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return r;
  }
  // uniform in [0, n), n > 0 (Lemire's multiply-shift with rejection)
  uint64_t below(uint64_t n) {
    uint64_t x = next();
    __uint128_t m = static_cast<__uint128_t>(x) * n;
    uint64_t l = static_cast<uint64_t>(m);
    if (l < n) {
      const uint64_t t = (0 - n) % n;
      while (l < t) {
        x = next();
        m = static_cast<__uint128_t>(x) * n;
        l = static_cast<uint64_t>(m);
      }
    }
    return static_cast<uint64_t>(m >> 64);
  }
};

// Draw `need` distinct positions of [0, fill) (Floyd's algorithm, then a Fisher-Yates pass so the order inside
// the batch is random too) and the moves that make [0, fill - need) the live region again: every drawn position
// below the new fill level (a hole) receives one un-drawn row from the tail.
void plan_draw(Rng& rng, int64_t fill, int32_t need, int32_t* sel, int32_t* mv_src, int32_t* mv_dst, int32_t* n_moves) {
  // membership as a byte map over the pool positions (a hash set here cost ~2 ms per 16 k-row batch: with one feeder
  // per rank and 8 ranks on a 32-vCPU host that was a visible share of the end-to-end step)
  static thread_local std::vector<uint8_t> chosen;
  chosen.assign(static_cast<size_t>(fill), 0);
  int32_t n = 0;
  for (int64_t j = fill - need; j < fill; ++j) {
    const int64_t t = static_cast<int64_t>(rng.below(static_cast<uint64_t>(j + 1)));
    const int64_t pick = chosen[static_cast<size_t>(t)] ? j : t;
    chosen[static_cast<size_t>(pick)] = 1;
    sel[n++] = static_cast<int32_t>(pick);
  }
  for (int32_t i = need - 1; i > 0; --i) {
    const int32_t j = static_cast<int32_t>(rng.below(static_cast<uint64_t>(i + 1)));
    const int32_t tmp = sel[i];
    sel[i] = sel[j];
    sel[j] = tmp;
  }
  const int64_t new_fill = fill - need;
  std::vector<uint8_t> tail_taken(static_cast<size_t>(need), 0);
  int32_t nm = 0;
  for (int32_t i = 0; i < need; ++i) {
    if (sel[i] >= new_fill) tail_taken[static_cast<size_t>(sel[i] - new_fill)] = 1;
  }
  int64_t cursor = 0;  // next candidate filler in the tail
  for (int32_t i = 0; i < need; ++i) {
    if (sel[i] < new_fill) {
      while (tail_taken[static_cast<size_t>(cursor)]) ++cursor;
      mv_dst[nm] = sel[i];
      mv_src[nm] = static_cast<int32_t>(new_fill + cursor);
      ++cursor;
      ++nm;
    }
  }
  *n_moves = nm;
}

// ---------------------------------------------------------------------------------------------
// shard geometry + chunk reader (host only)
// ---------------------------------------------------------------------------------------------
struct Geometry {
  std::string dir;
  int examples_per_shard = 0, n_layers = 0, tokens_per_example = 0, d_model = 0;
  int layer_index = 0, cls_token = 0, content_tokens = 0;
  const uint8_t* labels = nullptr;  // [n_examples_total, content_tokens] or null
  uint8_t ignore[256] = {0};
  bool filter = false;
};

bool pread_full(int fd, void* dst, size_t bytes, off_t off) {
  char* p = static_cast<char*>(dst);
  while (bytes > 0) {
    const ssize_t r = pread(fd, p, bytes, off);
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    if (r == 0) return false;  // short file
    p += r;
    off += r;
    bytes -= static_cast<size_t>(r);
  }
  return true;
}

// Reads the content tokens of examples [ex_begin, ex_begin + n_ex) of shard `shard` (layer geo.layer_index) into
// act[rows, d_model] / meta[rows, 2] = (global example index, content-token index); returns rows kept (after the
// optional label filter) or -1 on I/O error.
int64_t read_chunk(const Geometry& g, int fd, int shard, int ex_begin, int n_ex, float* act, int32_t* meta) {
  const size_t D = g.d_model, T = g.tokens_per_example, L = g.n_layers, C = g.content_tokens;
  const size_t row_bytes = D * 4;
  const bool contiguous = (L == 1 && g.cls_token == 0 && C == T);
  if (contiguous) {
    const off_t off = static_cast<off_t>(ex_begin) * T * row_bytes;
    if (!pread_full(fd, act, static_cast<size_t>(n_ex) * C * row_bytes, off)) return -1;
  } else {
    for (int e = 0; e < n_ex; ++e) {
      const off_t off =
          (((static_cast<off_t>(ex_begin) + e) * L + g.layer_index) * T + g.cls_token) * static_cast<off_t>(row_bytes);
      if (!pread_full(fd, act + static_cast<size_t>(e) * C * D, C * row_bytes, off)) return -1;
    }
  }
  const int64_t ex0 = static_cast<int64_t>(shard) * g.examples_per_shard + ex_begin;
  int64_t kept = 0;
  for (int e = 0; e < n_ex; ++e) {
    for (size_t t = 0; t < C; ++t) {
      const int64_t row = static_cast<int64_t>(e) * C + t;
      if (g.filter && g.ignore[g.labels[(ex0 + e) * C + t]]) continue;
      if (kept != row) memmove(act + kept * D, act + row * D, row_bytes);
      meta[2 * kept] = static_cast<int32_t>(ex0 + e);
      meta[2 * kept + 1] = static_cast<int32_t>(t);
      ++kept;
    }
  }
  return kept;
}

// ---------------------------------------------------------------------------------------------
// device kernels: one warp per row, float4 lanes
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loader_gather_kernel(const float* __restrict__ pool_act,
                                                            const int2* __restrict__ pool_meta,
                                                            const int* __restrict__ sel, int n, int D,
                                                            float* __restrict__ out_act, int* __restrict__ out_ex,
                                                            int* __restrict__ out_tok) {
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    const long long src = sel[i];
    const float4* s = reinterpret_cast<const float4*>(pool_act + src * D);
    float4* d = reinterpret_cast<float4*>(out_act + static_cast<long long>(i) * D);
    for (int v = lane; v < (D >> 2); v += 32) d[v] = s[v];
    if (lane == 0) {
      const int2 m = pool_meta[src];
      out_ex[i] = m.x;
      out_tok[i] = m.y;
    }
  }
}

__global__ void __launch_bounds__(256) loader_move_kernel(float* __restrict__ pool_act, int2* __restrict__ pool_meta,
                                                          const int* __restrict__ src, const int* __restrict__ dst,
                                                          int n, int D) {
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
    const long long a = src[i], b = dst[i];
    const float4* s = reinterpret_cast<const float4*>(pool_act + a * D);
    float4* d = reinterpret_cast<float4*>(pool_act + b * D);
    for (int v = lane; v < (D >> 2); v += 32) d[v] = s[v];
    if (lane == 0) pool_meta[b] = pool_meta[a];
  }
}

// a shard file mapped and registered for DMA (zero-copy mode); unmapped once every chunk cut from it has been copied
struct Mapping {
  char* base = nullptr;
  size_t len = 0;
  int pending = 0;         // chunks handed to the feeder whose copies have not been enqueued yet (under L->mu)
  bool closed = false;     // the I/O thread has cut its last chunk
  bool resident = false;   // kept mapped + registered for the loader's lifetime (within the pin budget)
  cudaEvent_t last = nullptr;  // recorded after the newest copy out of this mapping
};

struct Chunk {
  int slot = -1;
  int64_t rows = 0;
  Mapping* map = nullptr;  // zero-copy: the activations are `n_seg` segments of `seg_bytes` at map->base + off0, `stride` apart
  size_t off0 = 0, seg_bytes = 0, stride = 0;
  int n_seg = 0;
};

}  // namespace

// =================================================================================================
struct saev_b200_loader {
  Geometry geo;
  std::vector<int32_t> shard_order, shard_examples;
  int batch = 0, n_threads = 1, n_out = 3, chunk_examples = 1, device = 0;
  int64_t capacity = 0, chunk_rows_max = 0, rows_limit = -1;
  uint64_t seed = 0;
  float min_fill = 0.f;

  // device
  cudaStream_t stream = nullptr;
  float* pool_act = nullptr;
  int2* pool_meta = nullptr;
  std::vector<float*> out_act;
  std::vector<int*> out_ex, out_tok;
  std::vector<cudaEvent_t> out_ready, out_released, stage_done;
  std::vector<int*> idx_dev;   // [n_out][3 * batch] device copies of sel / src / dst
  std::vector<int*> idx_host;  // pinned
  std::vector<char*> stage;    // pinned staging chunks: [chunk_rows_max * D floats][chunk_rows_max * 2 int32]

  // state shared between threads
  std::mutex mu;
  std::condition_variable cv;
  std::deque<int> free_stage;    // staging slots available to I/O threads
  std::deque<Chunk> ready;       // filled chunks waiting for the feeder
  std::deque<int> prepared;      // out slots ready for the consumer (FIFO)
  std::vector<int> prepared_rows;
  std::vector<int> out_state;    // 0 free, 1 prepared, 2 handed out, 3 released (event recorded)
  size_t next_work = 0;          // next index into shard_order
  int io_running = 0;
  bool stop = false, feeder_done = false, epoch_active = false;
  int64_t rows_appended = 0, rows_drawn = 0, rows_expected = 0, pool_fill = 0;
  int handed_out = -1;           // slot currently owned by the consumer
  std::string error;
  int zero_copy = 0;             // 0: pread into pinned staging; 1: mmap + cudaHostRegister
  std::vector<Mapping*> mappings;  // live transient zero-copy mappings (under mu)
  // Shard files that stay mapped + registered for the loader's lifetime, up to `pin_budget` bytes (default 16 GiB,
  // SAEV_B200_LOADER_PIN_GB): registering pins every page, ~25 us per MB on this class of host, and an epoch over a
  // tmpfs-resident data set would otherwise pay it again for every shard it revisits -- or, with several ranks on one
  // host, queue behind the other ranks' registrations.  They are registered up front in saev_b200_loader_create.
  std::unordered_map<int, Mapping*> resident;
  size_t pin_budget = 0, pinned_bytes = 0;
  std::vector<std::thread> io_threads;
  std::thread feeder;
  std::atomic<long long> bytes_read{0};

  size_t stage_bytes() const { return static_cast<size_t>(chunk_rows_max) * (geo.d_model * 4 + 8); }
};

namespace {

int lfail(int code, const char* msg, const char* detail = "") {
  snprintf(g_loader_err, sizeof(g_loader_err), "%s%s", msg, detail);
  return code;
}

void set_error(saev_b200_loader* L, const std::string& e) {
  std::lock_guard<std::mutex> lk(L->mu);
  if (L->error.empty()) L->error = e;
  L->stop = true;
  L->cv.notify_all();
}

// mmap + cudaHostRegister one shard file (read-only); nullptr if the file cannot be mapped or registered
Mapping* map_shard(const saev_b200_loader* L, int shard) {
  char name[64];
  snprintf(name, sizeof(name), "/acts%06d.bin", shard);
  const std::string path = L->geo.dir + name;
  const int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return nullptr;
  Mapping* map = nullptr;
  struct stat stt;
  if (fstat(fd, &stt) == 0 && stt.st_size > 0) {
    void* base = mmap(nullptr, static_cast<size_t>(stt.st_size), PROT_READ, MAP_SHARED | MAP_POPULATE, fd, 0);
    if (base != MAP_FAILED) {
      cudaError_t e = cudaHostRegister(base, static_cast<size_t>(stt.st_size), cudaHostRegisterReadOnly);
      if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaHostRegister(base, static_cast<size_t>(stt.st_size), cudaHostRegisterDefault);
      }
      if (e == cudaSuccess) {
        map = new Mapping();
        map->base = static_cast<char*>(base);
        map->len = static_cast<size_t>(stt.st_size);
        cudaEventCreateWithFlags(&map->last, cudaEventDisableTiming);
      } else {
        cudaGetLastError();
        munmap(base, static_cast<size_t>(stt.st_size));
      }
    }
  }
  close(fd);
  return map;
}

void free_mapping(Mapping* m) {
  cudaHostUnregister(m->base);
  munmap(m->base, m->len);
  cudaEventDestroy(m->last);
  delete m;
}

// Unmap the zero-copy mappings whose chunks have all been copied (`wait`: block on the copies; else only the finished
// ones).  Called by the I/O threads between shards and when the threads are joined.
void reap_mappings(saev_b200_loader* L, bool wait) {
  std::vector<Mapping*> done;
  {
    std::lock_guard<std::mutex> lk(L->mu);
    for (size_t i = 0; i < L->mappings.size();) {
      Mapping* m = L->mappings[i];
      if (m->closed && m->pending == 0 && (wait || cudaEventQuery(m->last) == cudaSuccess)) {
        done.push_back(m);
        L->mappings[i] = L->mappings.back();
        L->mappings.pop_back();
      } else {
        ++i;
      }
    }
  }
  cudaGetLastError();  // (cudaErrorNotReady from the queries is not an error)
  for (Mapping* m : done) {
    cudaEventSynchronize(m->last);
    free_mapping(m);
  }
}

void io_main(saev_b200_loader* L) {
  cudaSetDevice(L->device);
  const Geometry& g = L->geo;
  for (;;) {
    size_t w;
    {
      std::lock_guard<std::mutex> lk(L->mu);
      if (L->stop || L->next_work >= L->shard_order.size()) break;
      w = L->next_work++;
    }
    const int shard = L->shard_order[w], n_examples = L->shard_examples[w];
    char name[64];
    snprintf(name, sizeof(name), "/acts%06d.bin", shard);
    const std::string path = g.dir + name;
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) {
      set_error(L, "cannot open " + path + ": " + strerror(errno));
      break;
    }
    bool ok = true;
    // zero-copy: map + register the whole shard file once; every chunk is then a (strided) window of it
    Mapping* map = nullptr;
    if (L->zero_copy && !g.filter) {
      reap_mappings(L, false);
      {
        std::lock_guard<std::mutex> lk(L->mu);
        auto it = L->resident.find(shard);
        if (it != L->resident.end()) map = it->second;
      }
      if (map == nullptr) {
        map = map_shard(L, shard);  // (nullptr: this file cannot be registered -> pread path for it)
        if (map != nullptr) {
          std::lock_guard<std::mutex> lk(L->mu);
          if (L->pinned_bytes + map->len <= L->pin_budget) {
            map->resident = true;
            L->pinned_bytes += map->len;
            L->resident[shard] = map;
          } else {
            L->mappings.push_back(map);
          }
        }
      }
    }
    for (int ex = 0; ex < n_examples && ok; ex += L->chunk_examples) {
      const int n_ex = std::min(L->chunk_examples, n_examples - ex);
      int slot;
      {
        std::unique_lock<std::mutex> lk(L->mu);
        L->cv.wait(lk, [&] { return L->stop || !L->free_stage.empty(); });
        if (L->stop) {
          ok = false;
          break;
        }
        slot = L->free_stage.front();
        L->free_stage.pop_front();
      }
      // the previous H2D copy out of this slot must have finished before it is overwritten
      cudaEventSynchronize(L->stage_done[slot]);
      float* act = reinterpret_cast<float*>(L->stage[slot]);
      int32_t* meta = reinterpret_cast<int32_t*>(L->stage[slot] + static_cast<size_t>(L->chunk_rows_max) * g.d_model * 4);
      Chunk ch;
      ch.slot = slot;
      if (map != nullptr) {
        // only the (example, token) pairs go through the staging slot; the activations stay where they are
        const size_t row_bytes = static_cast<size_t>(g.d_model) * 4, T = g.tokens_per_example, Lr = g.n_layers, C = g.content_tokens;
        const bool contiguous = (Lr == 1 && g.cls_token == 0 && C == T);
        ch.map = map;
        ch.off0 = ((static_cast<size_t>(ex) * Lr + g.layer_index) * T + g.cls_token) * row_bytes;
        ch.seg_bytes = contiguous ? static_cast<size_t>(n_ex) * C * row_bytes : C * row_bytes;
        ch.stride = Lr * T * row_bytes;
        ch.n_seg = contiguous ? 1 : n_ex;
        const size_t last_end = ch.off0 + static_cast<size_t>(ch.n_seg - 1) * ch.stride + ch.seg_bytes;
        if (last_end > map->len) {
          set_error(L, "short file " + path);
          ok = false;
          break;
        }
        const int64_t ex0 = static_cast<int64_t>(shard) * g.examples_per_shard + ex;
        int64_t r = 0;
        for (int e = 0; e < n_ex; ++e)
          for (size_t t = 0; t < C; ++t, ++r) {
            meta[2 * r] = static_cast<int32_t>(ex0 + e);
            meta[2 * r + 1] = static_cast<int32_t>(t);
          }
        ch.rows = r;
      } else {
        ch.rows = read_chunk(g, fd, shard, ex, n_ex, act, meta);
        if (ch.rows < 0) {
          set_error(L, "short read / I/O error in " + path);
          ok = false;
          break;
        }
      }
      L->bytes_read += static_cast<long long>(n_ex) * g.content_tokens * g.d_model * 4;
      {
        std::lock_guard<std::mutex> lk(L->mu);
        if (ch.rows > 0) {
          if (map != nullptr && !map->resident) ++map->pending;
          L->ready.push_back(ch);
        } else {
          L->free_stage.push_back(slot);
        }
        L->cv.notify_all();
      }
    }
    if (map != nullptr && !map->resident) {
      std::lock_guard<std::mutex> lk(L->mu);
      map->closed = true;
    }
    close(fd);
    if (!ok) break;
  }
  std::lock_guard<std::mutex> lk(L->mu);
  --L->io_running;
  L->cv.notify_all();
}

void feeder_main(saev_b200_loader* L) {
  cudaSetDevice(L->device);
  const int D = L->geo.d_model;
  Rng rng(L->seed);
  int next_slot = 0;
  for (;;) {
    Chunk ch;
    bool have_chunk = false, do_draw = false;
    int need = 0, slot = -1;
    bool wait_release = false;
    {
      std::unique_lock<std::mutex> lk(L->mu);
      for (;;) {
        if (L->stop) {
          L->feeder_done = true;
          L->cv.notify_all();
          return;
        }
        const int64_t remaining = L->rows_expected - L->rows_drawn;
        if (remaining <= 0) {
          L->feeder_done = true;
          L->cv.notify_all();
          return;
        }
        // 1. append a ready chunk if it fits
        if (!L->ready.empty() && L->pool_fill + L->ready.front().rows <= L->capacity) {
          ch = L->ready.front();
          L->ready.pop_front();
          have_chunk = true;
          break;
        }
        // 2. prepare a batch if enough rows are pooled and an output slot is free
        need = static_cast<int>(std::min<int64_t>(L->batch, remaining));
        const bool producers_done = (L->rows_appended >= L->rows_expected);
        const int64_t min_rows = producers_done ? 0 : static_cast<int64_t>(L->min_fill * static_cast<float>(std::min<int64_t>(L->capacity, remaining)));
        const int st = L->out_state[next_slot];
        if (L->pool_fill >= need && L->pool_fill >= min_rows && (st == 0 || st == 3)) {
          slot = next_slot;
          wait_release = (st == 3);
          L->out_state[slot] = 1;
          do_draw = true;
          break;
        }
        if (L->io_running == 0 && L->ready.empty() && L->pool_fill < need && L->rows_appended < L->rows_expected) {
          if (L->error.empty()) L->error = "loader: I/O threads finished before the expected number of rows arrived";
          L->stop = true;
          continue;
        }
        L->cv.wait(lk);
      }
    }
    if (have_chunk) {
      const float* act = reinterpret_cast<const float*>(L->stage[ch.slot]);
      const char* meta = L->stage[ch.slot] + static_cast<size_t>(L->chunk_rows_max) * D * 4;
      if (ch.map != nullptr) {  // zero-copy: DMA out of the registered file mapping (strided when examples interleave layers)
        if (ch.n_seg == 1)
          cudaMemcpyAsync(L->pool_act + L->pool_fill * D, ch.map->base + ch.off0, ch.seg_bytes, cudaMemcpyHostToDevice,
                          L->stream);
        else
          cudaMemcpy2DAsync(L->pool_act + L->pool_fill * D, ch.seg_bytes, ch.map->base + ch.off0, ch.stride, ch.seg_bytes,
                            static_cast<size_t>(ch.n_seg), cudaMemcpyHostToDevice, L->stream);
        cudaEventRecord(ch.map->last, L->stream);
      } else {
        cudaMemcpyAsync(L->pool_act + L->pool_fill * D, act, static_cast<size_t>(ch.rows) * D * 4, cudaMemcpyHostToDevice,
                        L->stream);
      }
      cudaMemcpyAsync(L->pool_meta + L->pool_fill, meta, static_cast<size_t>(ch.rows) * 8, cudaMemcpyHostToDevice,
                      L->stream);
      cudaEventRecord(L->stage_done[ch.slot], L->stream);
      std::lock_guard<std::mutex> lk(L->mu);
      if (ch.map != nullptr && !ch.map->resident) --ch.map->pending;
      L->pool_fill += ch.rows;
      L->rows_appended += ch.rows;
      L->free_stage.push_back(ch.slot);
      L->cv.notify_all();
      continue;
    }
    if (do_draw) {
      // the pinned index buffer of this slot was last read by the copy recorded in out_ready[slot]
      cudaEventSynchronize(L->out_ready[slot]);
      int* sel = L->idx_host[slot];
      int* src = sel + L->batch;
      int* dst = src + L->batch;
      int n_moves = 0;
      plan_draw(rng, L->pool_fill, need, sel, src, dst, &n_moves);
      if (wait_release) cudaStreamWaitEvent(L->stream, L->out_released[slot], 0);
      cudaMemcpyAsync(L->idx_dev[slot], sel, static_cast<size_t>(3) * L->batch * 4, cudaMemcpyHostToDevice, L->stream);
      const int blocks = std::min((need + 7) / 8, 148 * 8);
      loader_gather_kernel<<<blocks, 256, 0, L->stream>>>(L->pool_act, L->pool_meta, L->idx_dev[slot], need, D,
                                                         L->out_act[slot], L->out_ex[slot], L->out_tok[slot]);
      ++sb::g_launch_count;
      if (n_moves > 0) {
        loader_move_kernel<<<std::min((n_moves + 7) / 8, 148 * 8), 256, 0, L->stream>>>(
            L->pool_act, L->pool_meta, L->idx_dev[slot] + L->batch, L->idx_dev[slot] + 2 * L->batch, n_moves, D);
        ++sb::g_launch_count;
      }
      cudaEventRecord(L->out_ready[slot], L->stream);
      if (cudaGetLastError() != cudaSuccess) {
        set_error(L, "loader: CUDA error while preparing a batch");
        continue;
      }
      std::lock_guard<std::mutex> lk(L->mu);
      L->pool_fill -= need;
      L->rows_drawn += need;
      L->prepared.push_back(slot);
      L->prepared_rows[slot] = need;
      next_slot = (next_slot + 1) % L->n_out;
      L->cv.notify_all();
    }
  }
}

void join_threads(saev_b200_loader* L) {
  {
    std::lock_guard<std::mutex> lk(L->mu);
    L->stop = true;
    L->cv.notify_all();
  }
  for (auto& t : L->io_threads)
    if (t.joinable()) t.join();
  L->io_threads.clear();
  if (L->feeder.joinable()) L->feeder.join();
  if (L->stream) cudaStreamSynchronize(L->stream);
  {  // chunks that were cut but never copied (early stop) no longer pin their mapping
    std::lock_guard<std::mutex> lk(L->mu);
    for (Mapping* m : L->mappings) {
      m->pending = 0;
      m->closed = true;
    }
  }
  reap_mappings(L, true);
}

}  // namespace

extern "C" {

const char* saev_b200_loader_last_error(const saev_b200_loader* L) {
  if (L) {
    std::lock_guard<std::mutex> lk(const_cast<saev_b200_loader*>(L)->mu);
    if (!L->error.empty()) {
      snprintf(g_loader_err, sizeof(g_loader_err), "%s", L->error.c_str());
    }
  }
  return g_loader_err;
}

int saev_b200_loader_plan_draw(uint64_t seed, int64_t fill, int32_t need, int32_t* sel, int32_t* mv_src,
                               int32_t* mv_dst, int32_t* n_moves) {
  if (need <= 0 || fill < need || !sel || !mv_src || !mv_dst || !n_moves)
    return lfail(120, "loader_plan_draw: need 0 < need <= fill and non-null outputs");
  Rng rng(seed);
  plan_draw(rng, fill, need, sel, mv_src, mv_dst, n_moves);
  return 0;
}

int64_t saev_b200_loader_read_chunk(const saev_b200_loader_cfg* c, int32_t shard, int32_t ex_begin, int32_t n_ex,
                                    float* host_act, int32_t* host_meta) {
  if (!c || !c->shards_dir || !host_act || !host_meta) return lfail(-121, "loader_read_chunk: null argument");
  Geometry g;
  g.dir = c->shards_dir;
  g.examples_per_shard = c->examples_per_shard;
  g.n_layers = c->n_layers;
  g.tokens_per_example = c->tokens_per_example;
  g.d_model = c->d_model;
  g.layer_index = c->layer_index;
  g.cls_token = c->cls_token;
  g.content_tokens = c->content_tokens;
  g.labels = c->labels;
  g.filter = c->labels != nullptr && c->ignore_lut != nullptr;
  if (g.filter) memcpy(g.ignore, c->ignore_lut, 256);
  char name[64];
  snprintf(name, sizeof(name), "/acts%06d.bin", shard);
  const std::string path = g.dir + name;
  const int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return lfail(-122, "loader_read_chunk: cannot open ", path.c_str());
  const int64_t rows = read_chunk(g, fd, shard, ex_begin, n_ex, host_act, host_meta);
  close(fd);
  if (rows < 0) return lfail(-123, "loader_read_chunk: short read in ", path.c_str());
  return rows;
}

int saev_b200_loader_create(const saev_b200_loader_cfg* c, saev_b200_loader** out) {
  if (!c || !out) return lfail(124, "loader_create: null argument");
  *out = nullptr;
  if (!c->shards_dir || c->d_model <= 0 || c->d_model % 4 || c->batch_size <= 0 || c->pool_batches <= 0 ||
      c->content_tokens <= 0 || c->n_order < 0 || c->examples_per_shard <= 0 || c->n_layers <= 0 ||
      c->layer_index < 0 || c->layer_index >= c->n_layers || c->cls_token + c->content_tokens > c->tokens_per_example)
    return lfail(125, "loader_create: invalid configuration");
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return lfail(126, "loader_create: no CUDA device (the shuffle pool lives in HBM; there is no host-only mode)");
  }
  saev_b200_loader* L = new (std::nothrow) saev_b200_loader();
  if (!L) return lfail(127, "loader_create: out of host memory");
  L->device = dev;
  L->geo.dir = c->shards_dir;
  L->geo.examples_per_shard = c->examples_per_shard;
  L->geo.n_layers = c->n_layers;
  L->geo.tokens_per_example = c->tokens_per_example;
  L->geo.d_model = c->d_model;
  L->geo.layer_index = c->layer_index;
  L->geo.cls_token = c->cls_token;
  L->geo.content_tokens = c->content_tokens;
  L->geo.labels = c->labels;
  L->geo.filter = c->labels != nullptr && c->ignore_lut != nullptr;
  if (L->geo.filter) memcpy(L->geo.ignore, c->ignore_lut, 256);
  L->shard_order.assign(c->shard_order, c->shard_order + c->n_order);
  L->shard_examples.assign(c->shard_examples, c->shard_examples + c->n_order);
  L->batch = c->batch_size;
  L->n_threads = c->n_threads > 0 ? c->n_threads : 1;
  L->n_out = c->n_out_slots >= 2 ? c->n_out_slots : 3;
  L->seed = c->seed;
  L->min_fill = c->min_buffer_fill;
  L->rows_limit = c->n_rows_limit;
  // zero-copy mode (cfg.reserved: 0 = automatic -- on when the shards live on tmpfs, i.e. already in RAM --, 1 = off,
  // 2 = on for any file system; SAEV_B200_LOADER_ZERO_COPY=0|1 overrides).  Label filtering drops rows on the host
  // and therefore always takes the staging path.
  {
    int mode = c->reserved;
    if (const char* e = getenv("SAEV_B200_LOADER_ZERO_COPY")) mode = (e[0] == '0') ? 1 : 2;
    if (mode == 0) {
      struct statfs sfs;
      mode = (statfs(c->shards_dir, &sfs) == 0 && static_cast<unsigned long>(sfs.f_type) == 0x01021994UL) ? 2 : 1;
    }
    L->zero_copy = (mode == 2) ? 1 : 0;
    double gb = 16.0;
    if (const char* e = getenv("SAEV_B200_LOADER_PIN_GB")) gb = atof(e);
    L->pin_budget = gb > 0 ? static_cast<size_t>(gb * (1ull << 30)) : 0;
  }
  // chunk: whole examples, about 8 MB of activations unless the caller fixed it
  const int64_t ex_bytes = static_cast<int64_t>(c->content_tokens) * c->d_model * 4;
  int ce = c->chunk_examples > 0 ? c->chunk_examples : static_cast<int>(std::max<int64_t>(1, (8LL << 20) / ex_bytes));
  ce = std::min(ce, c->examples_per_shard);
  L->chunk_examples = ce;
  L->chunk_rows_max = static_cast<int64_t>(ce) * c->content_tokens;
  // the pool must be able to take one more chunk while it still holds a full batch
  L->capacity = std::max<int64_t>(static_cast<int64_t>(c->pool_batches) * c->batch_size, c->batch_size + L->chunk_rows_max);
  const int D = c->d_model;
  bool ok = cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&L->pool_act, static_cast<size_t>(L->capacity) * D * 4) == cudaSuccess;
  ok = ok && cudaMalloc(&L->pool_meta, static_cast<size_t>(L->capacity) * 8) == cudaSuccess;
  const int n_stage = L->n_threads + 2;
  L->out_act.assign(L->n_out, nullptr);
  L->out_ex.assign(L->n_out, nullptr);
  L->out_tok.assign(L->n_out, nullptr);
  L->idx_dev.assign(L->n_out, nullptr);
  L->idx_host.assign(L->n_out, nullptr);
  L->out_ready.assign(L->n_out, nullptr);
  L->out_released.assign(L->n_out, nullptr);
  L->stage.assign(n_stage, nullptr);
  L->stage_done.assign(n_stage, nullptr);
  L->out_state.assign(L->n_out, 0);
  L->prepared_rows.assign(L->n_out, 0);
  for (int i = 0; ok && i < L->n_out; ++i) {
    ok = ok && cudaMalloc(&L->out_act[i], static_cast<size_t>(L->batch) * D * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&L->out_ex[i], static_cast<size_t>(L->batch) * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&L->out_tok[i], static_cast<size_t>(L->batch) * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&L->idx_dev[i], static_cast<size_t>(3) * L->batch * 4) == cudaSuccess;
    ok = ok && cudaHostAlloc(&L->idx_host[i], static_cast<size_t>(3) * L->batch * 4, cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&L->out_ready[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&L->out_released[i], cudaEventDisableTiming) == cudaSuccess;
  }
  for (int i = 0; ok && i < n_stage; ++i) {
    ok = ok && cudaHostAlloc(&L->stage[i], L->stage_bytes(), cudaHostAllocDefault) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&L->stage_done[i], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    cudaGetLastError();
    saev_b200_loader_destroy(L);
    return lfail(128, "loader_create: CUDA allocation failed (pool / pinned staging)");
  }
  if (L->zero_copy && !L->geo.filter && L->pin_budget > 0) {
    // register this rank's shard files now, in visiting order, until the pin budget is spent (n_threads at a time)
    std::atomic<size_t> next{0};
    auto work = [&]() {
      cudaSetDevice(L->device);
      for (;;) {
        const size_t w = next.fetch_add(1);
        if (w >= L->shard_order.size()) return;
        {
          std::lock_guard<std::mutex> lk(L->mu);
          if (L->pinned_bytes >= L->pin_budget) return;
        }
        Mapping* m = map_shard(L, L->shard_order[w]);
        if (m == nullptr) continue;
        bool keep = false;
        {
          std::lock_guard<std::mutex> lk(L->mu);
          if (L->pinned_bytes + m->len <= L->pin_budget && L->resident.find(L->shard_order[w]) == L->resident.end()) {
            m->resident = keep = true;
            L->pinned_bytes += m->len;
            L->resident[L->shard_order[w]] = m;
          }
        }
        if (!keep) free_mapping(m);
      }
    };
    std::vector<std::thread> ts;
    for (int i = 0; i < L->n_threads; ++i) ts.emplace_back(work);
    for (auto& t : ts) t.join();
  }
  *out = L;
  return 0;
}

int saev_b200_loader_start_epoch(saev_b200_loader* L, int64_t rows_expected, uint64_t seed, void* consumer_stream) {
  if (!L) return lfail(124, "loader_start_epoch: null loader");
  join_threads(L);
  {
    std::lock_guard<std::mutex> lk(L->mu);
    // batches of the previous epoch may still be in use on the consumer's stream: their buffers must not be
    // overwritten before that work has finished
    if (L->handed_out >= 0) {
      cudaEventRecord(L->out_released[L->handed_out], static_cast<cudaStream_t>(consumer_stream));
      L->out_state[L->handed_out] = 3;
    }
    for (int i = 0; i < L->n_out; ++i)
      if (L->out_state[i] == 3) cudaStreamWaitEvent(L->stream, L->out_released[i], 0);
    L->stop = false;
    L->feeder_done = false;
    L->error.clear();
    L->free_stage.clear();
    for (size_t i = 0; i < L->stage.size(); ++i) L->free_stage.push_back(static_cast<int>(i));
    L->ready.clear();
    L->prepared.clear();
    std::fill(L->out_state.begin(), L->out_state.end(), 0);
    L->next_work = 0;
    L->rows_appended = L->rows_drawn = L->pool_fill = 0;
    L->rows_expected = rows_expected;
    L->handed_out = -1;
    L->seed = seed;
    L->io_running = L->n_threads;
    L->epoch_active = true;
  }
  for (int i = 0; i < L->n_threads; ++i) L->io_threads.emplace_back(io_main, L);
  L->feeder = std::thread(feeder_main, L);
  return 0;
}

int saev_b200_loader_next(saev_b200_loader* L, void* consumer_stream, double timeout_s, float** act,
                          int32_t** example_idx, int32_t** token_idx, int32_t* n_rows) {
  if (!L || !act || !example_idx || !token_idx || !n_rows) return lfail(124, "loader_next: null argument");
  cudaStream_t cs = static_cast<cudaStream_t>(consumer_stream);
  std::unique_lock<std::mutex> lk(L->mu);
  // the batch handed out by the previous call is released: everything the caller enqueued on its stream so far
  // has to finish before the feeder may overwrite that slot
  if (L->handed_out >= 0) {
    cudaEventRecord(L->out_released[L->handed_out], cs);
    L->out_state[L->handed_out] = 3;
    L->handed_out = -1;
    L->cv.notify_all();
  }
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(timeout_s > 0 ? timeout_s : 1e9);
  while (L->prepared.empty()) {
    if (!L->error.empty()) return 130;
    if (L->feeder_done || !L->epoch_active) {
      *n_rows = 0;  // epoch exhausted
      L->epoch_active = false;
      return 0;
    }
    if (L->cv.wait_until(lk, deadline) == std::cv_status::timeout && L->prepared.empty()) {
      snprintf(g_loader_err, sizeof(g_loader_err), "loader_next: no batch within %.1f s", timeout_s);
      return 131;  // TimeoutError; state is untouched, the call can be repeated
    }
  }
  const int slot = L->prepared.front();
  L->prepared.pop_front();
  L->out_state[slot] = 2;
  L->handed_out = slot;
  *n_rows = L->prepared_rows[slot];
  *act = L->out_act[slot];
  *example_idx = L->out_ex[slot];
  *token_idx = L->out_tok[slot];
  lk.unlock();
  if (cudaStreamWaitEvent(cs, L->out_ready[slot], 0) != cudaSuccess) return lfail(132, "loader_next: cudaStreamWaitEvent failed");
  return 0;
}

int saev_b200_loader_stats(saev_b200_loader* L, int64_t* pool_rows, int64_t* pool_capacity, int64_t* rows_delivered,
                           int64_t* bytes_read) {
  if (!L) return lfail(124, "loader_stats: null loader");
  std::lock_guard<std::mutex> lk(L->mu);
  if (pool_rows) *pool_rows = L->pool_fill;
  if (pool_capacity) *pool_capacity = L->capacity;
  if (rows_delivered) *rows_delivered = L->rows_drawn;
  if (bytes_read) *bytes_read = L->bytes_read.load();
  return 0;
}

int saev_b200_loader_zero_copy(const saev_b200_loader* L) { return L ? L->zero_copy : 0; }

int saev_b200_loader_stop(saev_b200_loader* L) {
  if (!L) return 0;
  join_threads(L);
  std::lock_guard<std::mutex> lk(L->mu);
  L->epoch_active = false;
  return 0;
}

int saev_b200_loader_destroy(saev_b200_loader* L) {
  if (!L) return 0;
  join_threads(L);
  cudaSetDevice(L->device);
  if (L->pool_act) cudaFree(L->pool_act);
  if (L->pool_meta) cudaFree(L->pool_meta);
  for (auto p : L->out_act) if (p) cudaFree(p);
  for (auto p : L->out_ex) if (p) cudaFree(p);
  for (auto p : L->out_tok) if (p) cudaFree(p);
  for (auto p : L->idx_dev) if (p) cudaFree(p);
  for (auto p : L->idx_host) if (p) cudaFreeHost(p);
  for (auto p : L->stage) if (p) cudaFreeHost(p);
  for (auto& kv : L->resident) free_mapping(kv.second);
  L->resident.clear();
  for (auto e : L->out_ready) if (e) cudaEventDestroy(e);
  for (auto e : L->out_released) if (e) cudaEventDestroy(e);
  for (auto e : L->stage_done) if (e) cudaEventDestroy(e);
  if (L->stream) cudaStreamDestroy(L->stream);
  delete L;
  return 0;
}

}  // extern "C"
