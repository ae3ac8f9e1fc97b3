// Host-side C++ runtime for bigslice_b200 (pybind11 module `_bsnative`).
//
//  * EvalCore — the evaluator's dependency bookkeeping (phase-aware waitlist:
//    consumers wait on whole producer groups).  Python drives task state and
//    executor calls; this core decides what becomes runnable.  Mirrors
//    bigslice_b200/exec/eval.py:_PyCore; role of exec/eval.go:272-451 `state`.
//  * crc32        — slice-by-8 CRC32 (IEEE) for the frame codec.
//  * murmur3_*    — seeded MurmurHash3_x86_32 over strings / integer buffers
//    (the host half of the bit-exact partitioner).
//  * merge_bound  — helper for batched k-way merges.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace py = pybind11;

// ---------------------------------------------------------------------------
class EvalCore {
 public:
  void set_count(int64_t task, int n) {
    counts_[task] = n;
    if (n == 0) todo_.push_back(task);
  }
  bool has_group(int64_t g) const { return group_missing_.count(g) != 0; }
  void add_group(int64_t g, const std::vector<int64_t>& missing) {
    auto& s = group_missing_[g];
    for (int64_t t : missing) {
      s.insert(t);
      task_groups_[t].push_back(g);
    }
  }
  void wait_on(int64_t task, int64_t g) { group_waiters_[g].push_back(task); }
  void schedule(int64_t task) { todo_.push_back(task); }
  void done_ok(int64_t task) {
    auto it = task_groups_.find(task);
    if (it == task_groups_.end()) return;
    std::vector<int64_t> groups = std::move(it->second);
    task_groups_.erase(it);
    for (int64_t g : groups) {
      auto gm = group_missing_.find(g);
      if (gm == group_missing_.end()) continue;
      gm->second.erase(task);
      if (!gm->second.empty()) continue;
      group_missing_.erase(gm);
      auto gw = group_waiters_.find(g);
      if (gw == group_waiters_.end()) continue;
      std::vector<int64_t> waiters = std::move(gw->second);
      group_waiters_.erase(gw);
      for (int64_t c : waiters) {
        int& n = counts_[c];
        if (--n == 0) todo_.push_back(c);
      }
    }
  }
  std::vector<int64_t> take_todo() {
    std::vector<int64_t> t;
    t.swap(todo_);
    return t;
  }
  size_t num_waiting_groups() const { return group_missing_.size(); }

 private:
  std::unordered_map<int64_t, int> counts_;
  std::unordered_map<int64_t, std::unordered_set<int64_t>> group_missing_;
  std::unordered_map<int64_t, std::vector<int64_t>> group_waiters_;
  std::unordered_map<int64_t, std::vector<int64_t>> task_groups_;
  std::vector<int64_t> todo_;
};

// ---------------------------------------------------------------------------
static uint32_t crc_tab[8][256];
static bool crc_ready = false;
static void crc_init() {
  if (crc_ready) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    crc_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) crc_tab[t][i] = (crc_tab[t - 1][i] >> 8) ^ crc_tab[0][crc_tab[t - 1][i] & 0xff];
  crc_ready = true;
}
static uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
  crc_init();
  crc = ~crc;
  while (n && (reinterpret_cast<uintptr_t>(p) & 7)) {
    crc = crc_tab[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
    --n;
  }
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    uint32_t lo = (uint32_t)v ^ crc, hi = (uint32_t)(v >> 32);
    crc = crc_tab[7][lo & 0xff] ^ crc_tab[6][(lo >> 8) & 0xff] ^ crc_tab[5][(lo >> 16) & 0xff] ^
          crc_tab[4][lo >> 24] ^ crc_tab[3][hi & 0xff] ^ crc_tab[2][(hi >> 8) & 0xff] ^
          crc_tab[1][(hi >> 16) & 0xff] ^ crc_tab[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) crc = crc_tab[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
  return ~crc;
}
static uint32_t py_crc32(py::buffer b, uint32_t value) {
  py::buffer_info info = b.request();
  size_t n = (size_t)info.size * (size_t)info.itemsize;
  return crc32_update(value, static_cast<const uint8_t*>(info.ptr), n);
}

// GF(2) crc32_combine (zlib algorithm) for per-block device CRCs.
static uint32_t gf2_times(const uint32_t* mat, uint32_t vec) {
  uint32_t sum = 0;
  while (vec) {
    if (vec & 1) sum ^= *mat;
    vec >>= 1;
    ++mat;
  }
  return sum;
}
static void gf2_square(uint32_t* sq, const uint32_t* mat) {
  for (int n = 0; n < 32; ++n) sq[n] = gf2_times(mat, mat[n]);
}
static uint32_t crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
  if (len2 == 0) return crc1;
  uint32_t even[32], odd[32];
  odd[0] = 0xEDB88320u;
  uint32_t row = 1;
  for (int n = 1; n < 32; ++n) {
    odd[n] = row;
    row <<= 1;
  }
  gf2_square(even, odd);
  gf2_square(odd, even);
  do {
    gf2_square(even, odd);
    if (len2 & 1) crc1 = gf2_times(even, crc1);
    len2 >>= 1;
    if (!len2) break;
    gf2_square(odd, even);
    if (len2 & 1) crc1 = gf2_times(odd, crc1);
    len2 >>= 1;
  } while (len2);
  return crc1 ^ crc2;
}
static uint32_t py_crc32_combine_blocks(py::array_t<uint32_t> crcs, uint64_t total, uint64_t block) {
  auto r = crcs.unchecked<1>();
  uint32_t crc = 0;
  uint64_t left = total;
  for (py::ssize_t i = 0; i < r.shape(0); ++i) {
    uint64_t len = left < block ? left : block;
    crc = (i == 0) ? r(i) : crc32_combine(crc, r(i), len);
    left -= len;
  }
  return crc;
}

// ---------------------------------------------------------------------------
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static uint32_t murmur3_32(const uint8_t* data, size_t len, uint32_t seed) {
  uint32_t h = seed;
  const size_t nblocks = len / 4;
  for (size_t i = 0; i < nblocks; ++i) {
    uint32_t k;
    memcpy(&k, data + 4 * i, 4);
    k *= 0xcc9e2d51u;
    k = rotl32(k, 15);
    k *= 0x1b873593u;
    h ^= k;
    h = rotl32(h, 13);
    h = h * 5 + 0xe6546b64u;
  }
  const uint8_t* tail = data + 4 * nblocks;
  uint32_t k = 0;
  switch (len & 3) {
    case 3: k ^= (uint32_t)tail[2] << 16; [[fallthrough]];
    case 2: k ^= (uint32_t)tail[1] << 8; [[fallthrough]];
    case 1:
      k ^= tail[0];
      k *= 0xcc9e2d51u;
      k = rotl32(k, 15);
      k *= 0x1b873593u;
      h ^= k;
  }
  h ^= (uint32_t)len;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
static py::array_t<uint32_t> py_murmur3_strings(const std::vector<py::bytes>& items, uint32_t seed) {
  py::array_t<uint32_t> out((py::ssize_t)items.size());
  auto w = out.mutable_unchecked<1>();
  for (size_t i = 0; i < items.size(); ++i) {
    char* p;
    py::ssize_t n;
    PyBytes_AsStringAndSize(items[i].ptr(), &p, &n);
    w((py::ssize_t)i) = murmur3_32(reinterpret_cast<const uint8_t*>(p), (size_t)n, seed);
  }
  return out;
}
static py::array_t<uint32_t> py_murmur3_fixed(py::buffer b, int width, uint32_t seed) {
  py::buffer_info info = b.request();
  size_t nbytes = (size_t)info.size * (size_t)info.itemsize;
  size_t n = nbytes / (size_t)width;
  py::array_t<uint32_t> out((py::ssize_t)n);
  auto w = out.mutable_unchecked<1>();
  const uint8_t* p = static_cast<const uint8_t*>(info.ptr);
  for (size_t i = 0; i < n; ++i) w((py::ssize_t)i) = murmur3_32(p + i * width, (size_t)width, seed);
  return out;
}
static uint32_t py_murmur3_bytes(py::bytes b, uint32_t seed) {
  char* p;
  py::ssize_t n;
  PyBytes_AsStringAndSize(b.ptr(), &p, &n);
  return murmur3_32(reinterpret_cast<const uint8_t*>(p), (size_t)n, seed);
}

PYBIND11_MODULE(_bsnative, m) {
  m.doc() = "bigslice_b200 host runtime: evaluator core, crc32, murmur3";
  py::class_<EvalCore>(m, "EvalCore")
      .def(py::init<>())
      .def("set_count", &EvalCore::set_count)
      .def("has_group", &EvalCore::has_group)
      .def("add_group", &EvalCore::add_group)
      .def("wait_on", &EvalCore::wait_on)
      .def("schedule", &EvalCore::schedule)
      .def("done_ok", &EvalCore::done_ok)
      .def("take_todo", &EvalCore::take_todo)
      .def("num_waiting_groups", &EvalCore::num_waiting_groups);
  m.def("crc32", &py_crc32, py::arg("data"), py::arg("value") = 0);
  m.def("crc32_combine", &crc32_combine);
  m.def("crc32_combine_blocks", &py_crc32_combine_blocks);
  m.def("murmur3_strings", &py_murmur3_strings);
  m.def("murmur3_fixed", &py_murmur3_fixed);
  m.def("murmur3_bytes", &py_murmur3_bytes);
}
